#!/usr/bin/env python
"""tools/pass_model.py — per-pass times of a C3 prefix against what the planner knew about each pass (groups,
gates, cost): the data behind the planner's cost cap."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--costs", default="32,48,64")
    ap.add_argument("--jit", type=int, default=0, help="QM_OPT_JIT (1: compiled pass kernels)")
    ap.add_argument("--dump", default="", help="write the raw rows [budget, ms, groups, gates, planner cost, full, half, xor, thr] here (.npy)")
    a = ap.parse_args()
    pkg = ge.load_package()
    L = pkg._lib
    gates = pkg.circuits.c3_circuit(n=a.qubits, depth=a.depth)
    st = pkg.B200State(a.qubits, 0)
    st.set_option(L.OPT_JIT, a.jit)
    st.apply_circuit(gates); st.sync()                # dense state
    st.set_option(L.OPT_TIME_PASSES, 1)
    allrows = []
    for cost in [int(x) for x in a.costs.split(",")]:
        st.set_option(L.OPT_MAX_COST, cost)
        st.apply_circuit(gates); st.pass_times()      # warm
        st.apply_circuit(gates)
        t = st.pass_times()
        allrows.append(t.astype(np.float64))
        if a.dump:
            np.save(a.dump, np.concatenate([np.concatenate([np.full((len(r), 1), float(c)), r], axis=1)
                                            for c, r in zip([int(x) for x in a.costs.split(",")], allrows)]))
        ms, groups, ng, pc = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
        A = np.stack([np.ones_like(ms), groups, pc - ng], axis=1)      # ms ~ c0 + c1 groups + c2 (FP64 ops per amplitude)
        coef, *_ = np.linalg.lstsq(A, ms, rcond=None)
        desc = L.plan_describe(a.qubits, gates, max_cost=float(cost))
        if len(desc) == len(ms):
            for i in np.argsort(-ms)[:3]:
                print("   slow pass %d: %.2f ms" % (i, ms[i]), desc[i], flush=True)
            for i in np.argsort(ms)[:2]:
                print("   fast pass %d: %.2f ms" % (i, ms[i]), desc[i], flush=True)
        print(json.dumps({"max_cost": cost, "passes": len(ms), "ms_mean": float(ms.mean()), "ms_min": float(ms.min()), "ms_max": float(ms.max()),
                          "ms_p10": float(np.percentile(ms, 10)), "ms_p90": float(np.percentile(ms, 90)),
                          "fit_ms = c0 + c1*groups + c2*fp64_per_amp": [float(x) for x in coef],
                          "by_groups": {int(g): [int((groups == g).sum()), float(ms[groups == g].mean())] for g in np.unique(groups)}}), flush=True)
    # one model over all budgets, on the passes that are clearly not HBM-bound: ms ~ a + g*groups + s*gates + f*fp64
    if allrows:
        R = np.concatenate(allrows)
        R = R[R[:, 0] > 6.6]
        A = np.stack([np.ones(len(R)), R[:, 1], R[:, 2], R[:, 3] - R[:, 2]], axis=1)
        coef, res, *_ = np.linalg.lstsq(A, R[:, 0], rcond=None)
        pred = A @ coef
        print(json.dumps({"compute_bound_passes": int(len(R)), "fit_ms = a + g*groups + s*gates + f*fp64_per_amp": [float(x) for x in coef],
                          "rms_residual_ms": float(np.sqrt(np.mean((pred - R[:, 0]) ** 2)))}), flush=True)
        A = np.stack([np.ones(len(R)), R[:, 1], R[:, 4], R[:, 5], R[:, 6], R[:, 7]], axis=1)
        coef, res, *_ = np.linalg.lstsq(A, R[:, 0], rcond=None)
        pred = A @ coef
        print(json.dumps({"fit_ms = a + g*groups + full + half + xor + thr (per record)": [float(x) for x in coef],
                          "rms_residual_ms": float(np.sqrt(np.mean((pred - R[:, 0]) ** 2)))}), flush=True)
    st.close()


if __name__ == "__main__":
    main()
