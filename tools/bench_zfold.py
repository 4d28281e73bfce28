"""tools/bench_zfold.py — the reservoir step (config C4: encoding layer + reservoir layer + <Z> of every qubit) on ONE GPU with the
readout folded into the last pass (qm_apply_circuit_expect_z, QM_OPT_ZFOLD 1) against circuit + separate readout (QM_OPT_ZFOLD 0),
same process, same register, same step structures, interleaved A/B/A/B so that clock drift under the power cap hits both.

    python tools/bench_zfold.py [--qubits 33] [--steps 6] [--rounds 2] [--out gpurun_out/zfold.json]

Also checks, at full size, that the folded sums equal the separate readout of the same state (1e-13) and that the norm is 1."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=33)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    pkg = ge.load_package()
    L = pkg._lib
    n = args.qubits
    st = pkg.B200State(n, 0)
    st.set_option(L.OPT_JIT, 1)
    gates = [pkg.circuits.c4_step(n, s) for s in range(6 + args.steps)]
    res = {"n_qubits": n, "steps_per_leg": args.steps, "gates_per_step": int(len(gates[0]))}
    # warm-up: every step structure (period 6) once per flavour, so that both legs launch cached kernels
    for fold in (1, 0):
        st.set_option(L.OPT_ZFOLD, fold)
        for s in range(6):
            st.apply_circuit_expect_z(gates[s])
    # correctness at full size: folded sums against the separate readout of the very same state
    st.set_option(L.OPT_ZFOLD, 1)
    st.counters_reset()
    z = st.apply_circuit_expect_z(gates[0])
    res["folded"] = int(st.counters()["folded_readouts"])
    res["fold_vs_separate"] = float(np.max(np.abs(z - st.expect_z_all())))
    z0 = st.apply_circuit_expect_z(gates[1], -1.0)
    res["fold_vs_separate_nothr"] = float(np.max(np.abs(z0 - st.expect_z_all(-1.0))))
    res["norm2"] = st.norm2()
    res["norm2_from_fold"] = float(z0[0].sum())
    legs = {0: [], 1: []}
    for r in range(args.rounds):
        for fold in (1, 0):
            st.set_option(L.OPT_ZFOLD, fold)
            st.sync()
            st.counters_reset()
            t0 = time.perf_counter()
            for s in range(args.steps):
                st.apply_circuit_expect_z(gates[6 + s])
            dt = time.perf_counter() - t0          # every step ends with a synchronous readout: wall time is device time + launch
            c = st.counters()
            legs[fold].append({"ms_per_step": 1e3 * dt / args.steps, "passes_per_step": c["passes"] / args.steps,
                               "launches_per_step": c["launches"] / args.steps, "folded_per_step": c["folded_readouts"] / args.steps,
                               "compiled_passes_per_step": c["compiled_passes"] / args.steps})
    res["folded_readout"] = legs[1]
    res["separate_readout"] = legs[0]
    a = min(x["ms_per_step"] for x in legs[1]); b = min(x["ms_per_step"] for x in legs[0])
    res["best_ms_per_step"] = {"folded": a, "separate": b, "saved_ms": b - a, "readout_bytes_not_read": 16 * (1 << n)}
    res["jit"] = L.jit_stats()
    ok = res["folded"] == 1 and res["fold_vs_separate"] < 1e-13 and res["fold_vs_separate_nothr"] < 1e-13 and abs(res["norm2"] - 1.0) < 1e-12
    res["ok"] = bool(ok)
    s = json.dumps(res)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")
    st.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
