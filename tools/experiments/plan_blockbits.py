"""tools/experiments/plan_blockbits.py — planning study (CPU only): gates per fused pass with 4-qubit (today) and 5-qubit
register-block groups, by the engine's own planner and the compiled kernels' cost model, and what the fitted pass model
(ms at 30 qubits = max(HBM 5.8, 1.2 + 1.45 per group + 0.05 per budget unit)) makes of it.  The 5-qubit rows are a
PROJECTION: there is no kernel for them; the model's per-unit compute cost is the one fitted on 16-amplitude threads.

    python tools/experiments/plan_blockbits.py [--qubits 30] [--depth 200]"""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=200)
    a = ap.parse_args()
    so = "/tmp/libplan_blockbits.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-attributes", "-I/usr/local/cuda/include", "-x", "c++",
                           "-o", so, os.path.join(HERE, "plan_blockbits.cpp")])
    lib = C.CDLL(so)
    lib.plan_blockbits.restype = C.c_int32
    lib.plan_blockbits.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
    pkg = ge.load_package()
    c3 = pkg.circuits.c3_circuit(n=a.qubits, depth=a.depth, seed=3001)
    # a reservoir step ends with a readout: every step is planned on its own (short-circuit planning, as the engine does)
    workloads = [("C3 %dq depth %d" % (a.qubits, a.depth), [c3], a.qubits),
                 ("C4 steps 0-5, 33q", [pkg.circuits.c4_step(33, s) for s in range(6)], 33)]
    print("%-22s %5s %6s %7s %10s %9s %10s %11s %13s" % ("workload", "block", "budget", "passes", "gates/pass", "groups/p", "units/pass", "model ms/p", "model gates/s"))
    for name, circuits, n in workloads:
        for bb in (4, 5):
            for budget in (80, 100, 120, 140, 170, 200):
                allrows = []
                for gates in circuits:
                    rows = np.zeros((65536, 4), dtype=np.int32)
                    npass = lib.plan_blockbits(n, 5 if len(circuits) == 1 else 4, float(budget), bb, 0.0, gates.ctypes.data, len(gates), rows.ctypes.data, 65536)
                    if npass <= 0:
                        raise SystemExit("planner error %d" % npass)
                    allrows.append(rows[:npass].copy())
                r = np.concatenate(allrows)
                ngates = sum(len(g) for g in circuits)
                groups = r[:, 1]
                units = r[:, 2] / 16.0                       # PlanPass::cost: gate costs of the pass in budget units (0.05 ms at 30 qubits), groups not included
                scale = 2.0 ** (n - 30)
                ms = np.maximum(5.8, 1.2 + 1.45 * groups + 0.05 * units)           # at 30 qubits
                print("%-22s %5d %6d %7d %10.1f %9.2f %10.1f %11.2f %13.0f" % (name, bb, budget, len(r), r[:, 0].mean(), groups.mean(), units.mean(),
                                                                              ms.mean(), ngates / (ms.sum() * 1e-3)))


if __name__ == "__main__":
    main()
