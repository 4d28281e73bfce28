#!/usr/bin/env python
"""tools/bench_pipeline.py — end-to-end timing of the two small reservoir configurations of BASELINE.json
(SURVEY.md §8d): C1 (12 qubits, T = 100, per-qubit <Z>) and C5 (LSM front end driving a 20-qubit reservoir with the
QInfo partial-trace / entropy readouts per step).  Both are loops of a short circuit followed by host readouts, so
what is measured is the whole call path — gate-list construction, planning (with the structure-keyed plan cache),
launches, readout kernels, device-to-host copies and the host-side entropies — in steps per second, next to the CPU
oracle port doing the same steps (oracle/qsim_oracle.c, all host threads) on a bounded prefix."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def run_c5(pkg, n, T, k, cpu_steps):
    if cpu_steps:
        from oracle import literal as LIT
        from oracle import oracle as O
    R = pkg.reservoir
    rng = np.random.default_rng(2001)
    lsm = R.LSMNet(rng.uniform(-1, 1, size=(50, 3)), 0.05 * rng.normal(size=(50, 50)), leak_rate=0.3)
    drive = lsm.collect_states(rng.uniform(0, 1, size=(3, T)))[:n]
    qr = R.QuantumReservoir(n, entropy_k=k)
    qr.collect_states(drive[:, :3])                      # warm-up: module load, allocations, plan cache
    qr.state.counters_reset()
    t0 = time.perf_counter()
    feats = qr.collect_states(drive)
    wall = time.perf_counter() - t0
    c = qr.state.counters()
    # CPU port on a prefix of the same steps (cpu_steps = 0: GPU side only)
    ref = O.statevector(n, 0) if cpu_steps else None
    t1 = time.perf_counter()
    err = 0.0
    for t in range(cpu_steps):
        O.apply_circuit(ref, qr.step_gates(drive[:, t]))
        z = O.expect_z_all(ref, 1e-10)
        SB = LIT.von_neumann_entropy(O.reduced_density(ref, True, k))
        SA = LIT.von_neumann_entropy(O.reduced_density(ref, False, k))
        err = max(err, float(np.max(np.abs(feats[:n, t] - (z[:, 0] - z[:, 1])))), abs(feats[n, t] - SB), abs(feats[n + 1, t] - SA))
    cpu = time.perf_counter() - t1
    return {"workload": "C5: LSM (50 units) -> %d-qubit reservoir, <Z> all qubits + entropies of rho_B / rho_A (k = %d) per step" % (n, k),
            "steps": T, "gates_per_step": len(qr.step_gates(drive[:, 0])), "steps_per_s": T / wall, "ms_per_step": 1e3 * wall / T,
            "launches_per_step": c["launches"] / T, "passes_per_step": c["passes"] / T,
            "cpu_port_steps_per_s": (cpu_steps / cpu) if cpu_steps else None, "cpu_steps_timed": cpu_steps,
            "cpu_threads": O.num_threads() if cpu_steps else None,
            "max_abs_diff_vs_cpu_port": err if cpu_steps else None}


def run_c1(pkg, n, T, cpu_steps):
    if cpu_steps:
        from oracle import oracle as O
    Q, C = pkg.qsim, pkg.circuits
    u, w, layer, gamma = C.c1_inputs(n, T)

    def step_gates(t):
        gl = C.GateList(n, "reference")
        for q in range(1, n + 1):
            gl.ry(q, np.pi * u[t] * w[q - 1])
        for q in range(1, n + 1):
            gl.rx(q, layer[q - 1][0]); gl.rz(q, layer[q - 1][1])
        for q in range(1, n):
            gl.cnot(q, q + 1)
        gl.crz(n, 1, gamma)
        return gl.array()

    st = pkg.B200State(n, 0)
    for t in range(3):
        st.apply_circuit(step_gates(t)); st.expect_z_all()
    st.reset(0); st.counters_reset()
    out = np.empty((n, T))
    t0 = time.perf_counter()
    for t in range(T):
        st.apply_circuit(step_gates(t))
        z = st.expect_z_all()
        out[:, t] = z[:, 0] - z[:, 1]
    wall = time.perf_counter() - t0
    c = st.counters()
    ref = O.statevector(n, 0) if cpu_steps else None
    t1 = time.perf_counter()
    err = 0.0
    for t in range(cpu_steps):
        O.apply_circuit(ref, step_gates(t))
        z = O.expect_z_all(ref, 1e-10)
        err = max(err, float(np.max(np.abs(out[:, t] - (z[:, 0] - z[:, 1])))))
    cpu = time.perf_counter() - t1
    st.close()
    return {"workload": "C1: %d-qubit reservoir, %d gates per step (incl. the non-adjacent crz!), <Z> of every qubit per step" % (n, len(step_gates(0))),
            "steps": T, "steps_per_s": T / wall, "ms_per_step": 1e3 * wall / T, "launches_per_step": c["launches"] / T,
            "cpu_port_steps_per_s": (cpu_steps / cpu) if cpu_steps else None, "cpu_steps_timed": cpu_steps,
            "cpu_threads": O.num_threads() if cpu_steps else None,
            "max_abs_diff_vs_cpu_port": err if cpu_steps else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--c5-qubits", type=int, default=20)
    ap.add_argument("--cpu-steps", type=int, default=20)
    a = ap.parse_args()
    pkg = ge.load_package()
    print(json.dumps(run_c1(pkg, 12, a.steps, min(a.cpu_steps * 5, a.steps))), flush=True)
    print(json.dumps(run_c5(pkg, a.c5_qubits, a.steps, 4, min(a.cpu_steps, a.steps))), flush=True)


if __name__ == "__main__":
    main()
