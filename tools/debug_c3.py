#!/usr/bin/env python
"""tools/debug_c3.py — a C3 prefix applied in one call, compared with the oracle; repeated to expose races."""
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
    ap.add_argument("--qubits", type=int, default=24)
    ap.add_argument("--depth", type=int, default=6)
    ap.add_argument("--tile", type=int, default=11)
    ap.add_argument("--cost", type=int, default=24)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    pkg = ge.load_package()
    L = pkg._lib
    n = a.qubits
    gates = pkg.circuits.c3_circuit(n=n, depth=a.depth)
    ref = O.apply_circuit(O.statevector(n, 0), gates)
    for rep in range(a.reps):
        st = pkg.B200State(n, 0)
        st.set_option(L.OPT_TILE_BITS, a.tile)
        st.set_option(L.OPT_MAX_COST, a.cost)
        st.apply_circuit(gates)
        got = st.to_numpy()
        e = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        idx = np.nonzero(np.abs(got - ref) > 1e-9)[0]
        print("rep", rep, "err %.3e" % e, "passes", st.counters()["passes"], "wrong", len(idx), idx[:6], flush=True)
        st.close()


if __name__ == "__main__":
    main()
