#!/usr/bin/env python
"""tools/debug_passes.py — applies a circuit pass by pass (the planner's own pass boundaries) on the GPU and
compares with the oracle after every pass: localises a wrong pass program."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=22)
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--tile", type=int, default=11)
    ap.add_argument("--cost", type=int, default=24)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    pkg = ge.load_package()
    L = pkg._lib
    n = a.qubits
    gates = pkg.circuits.c3_circuit(n=n, depth=a.depth)
    pg, po = L.plan_gates(n, gates, tile_bits=a.tile, max_cost=float(a.cost))
    npass = int(po.max()) + 1
    print("passes", npass, flush=True)
    for rep in range(a.reps):
        ref = O.random_state(n, 3 + rep)
        st = pkg.B200State.from_numpy(ref)
        st.set_option(L.OPT_TILE_BITS, a.tile)
        st.set_option(L.OPT_MAX_COST, 0)
        bad = 0
        for p in range(npass):
            g = pg[po == p]
            st.apply_circuit(g)
            got = st.to_numpy()
            ref = O.apply_circuit(ref, g)
            e = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
            if e > 1e-12:
                bad += 1
                d = st.plan_describe(g) if hasattr(st, "plan_describe") else L.plan_describe(n, g, tile_bits=a.tile)
                print("rep", rep, "pass", p, "err %.3e" % e, "gates", len(g), d, flush=True)
                idx = np.nonzero(np.abs(got - ref) > 1e-9)[0]
                print("   wrong amplitudes:", len(idx), "first", idx[:8], "last", idx[-4:], flush=True)
                st.close()
                st = pkg.B200State.from_numpy(ref)
                st.set_option(L.OPT_TILE_BITS, a.tile)
                st.set_option(L.OPT_MAX_COST, 0)
        print("rep", rep, "bad passes", bad, flush=True)
        st.close()


if __name__ == "__main__":
    main()
