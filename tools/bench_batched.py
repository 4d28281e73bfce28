#!/usr/bin/env python
"""tools/bench_batched.py — config C2: 4096 independent 16-qubit registers (4 GiB) evolved in
lockstep, per-state ry encodings + shared reservoir layer + CNOT chain (63 gates per step per register), all-qubit
<Z> readout per step.  Prints gate applications per second and the achieved HBM rate of the fused passes."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def run_c2(pkg, steps=8, batch=4096, n=16, max_cost=-1):
    C, L = pkg.circuits, pkg._lib
    B = batch
    gates = C.c2_step_gates(n, seed=1603)
    w = [C.SplitMix64(7).uniform() for _ in range(n)]
    st = pkg.B200State(n, 0, batch=B)
    if max_cost >= 0:
        st.set_option(L.OPT_MAX_COST, max_cost)
    encs = [C.c2_encoding(B, n, t, w) for t in range(3)]
    for t in range(2):
        st.apply_batched(gates, encs[t]); st.expect_z_all()
    st.sync(); st.counters_reset()
    dev_ms, t0 = 0.0, time.perf_counter()
    for t in range(steps):
        st.apply_batched(gates, encs[t % 3])
        st.expect_z_all()
        dev_ms += st.counters()["ms_device"]
    wall = time.perf_counter() - t0
    c = st.counters()
    # the same loop as a pipeline (split readout): the readout of step t is only ENQUEUED, the host goes straight on to
    # classifying and uploading the encodings of step t + 1 while the device still runs step t, and fetches z(t) afterwards
    st.sync()
    for t in range(2):          # untimed: the first split readouts allocate their two pinned result buffers and events
        st.apply_batched(gates, encs[t]); st.expect_z_all_begin()
    st.expect_z_all_end(); st.expect_z_all_end()
    st.sync()
    zs = []
    t1 = time.perf_counter()
    for t in range(steps):
        st.apply_batched(gates, encs[t % 3])
        st.expect_z_all_begin()
        if t > 0:
            zs.append(st.expect_z_all_end())
    zs.append(st.expect_z_all_end())
    wall_pipe = time.perf_counter() - t1
    assert len(zs) == steps and abs(float(zs[-1][0].sum()) - n) < 1e-3
    per_pass = dev_ms / max(c["passes"], 1)
    gbs = 32.0 * B * (1 << n) / (per_pass * 1e-3) / 1e9
    out = {"workload": "C2: %d independent %d-qubit registers in lockstep, per-state ry encodings + shared reservoir layer + "
                       "CNOT chain, all-qubit <Z> per register per step" % (B, n),
           "batch": B, "n_qubits": n, "steps": steps, "gates_per_step_per_register": len(gates),
           "passes_per_step": c["passes"] / steps, "ms_circuit_per_step_device": dev_ms / steps, "ms_per_pass": per_pass,
           "fused_pass_gbs": gbs, "ms_per_step_wall_incl_readout_and_encoding_upload": 1e3 * wall / steps,
           "ms_per_step_wall_pipelined_split_readout": 1e3 * wall_pipe / steps,
           "gate_applications_per_s_wall_pipelined": B * len(gates) * steps / wall_pipe,
           "gate_applications_per_s_device": B * len(gates) * steps / (dev_ms * 1e-3),
           "gate_applications_per_s_wall": B * len(gates) * steps / wall,
           "compiled_passes_per_step": c["compiled_passes"] / steps,
           "h2d_bytes_per_step": c["bytes_h2d"] / steps, "d2h_bytes_per_step": c["bytes_d2h"] / steps}
    st.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--qubits", type=int, default=16)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--max-cost", type=int, default=-1)
    a = ap.parse_args()
    print(json.dumps(run_c2(ge.load_package(), a.steps, a.batch, a.qubits, a.max_cost)))


if __name__ == "__main__":
    main()
