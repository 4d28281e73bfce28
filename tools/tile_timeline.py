#!/usr/bin/env python
"""tools/tile_timeline.py — where a fused pass spends its time, from the kernel's own phase trace (QM_OPT_TRACE:
CTA 0 stamps the SM clock per tile: team wait start, tile landed, after each register-block group, arrival; the
producer's `computed` seen and next load issued).  Prints per-team phase means in cycles, how often the two teams
are inside their gate bodies at the same time, and the producer's turnaround."""
import argparse
import ctypes as C
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
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--costs", default="80,110,150,250")
    ap.add_argument("--tile", type=int, default=12)
    a = ap.parse_args()
    pkg = ge.load_package()
    L = pkg._lib
    n = a.qubits
    gates = pkg.circuits.c3_circuit(n=n, depth=a.depth, seed=3001)
    for cost in [int(x) for x in a.costs.split(",")]:
        st = pkg.B200State(n, 0)
        st.set_option(L.OPT_MAX_COST, cost)
        st.set_option(L.OPT_TILE_BITS, a.tile)
        st.apply_circuit(gates); st.sync()                 # dense state, warm
        st.set_option(L.OPT_TRACE, 1)
        st.set_option(L.OPT_TIME_PASSES, 1)
        st.apply_circuit(gates); st.sync()
        pt = st.pass_times()
        ntiles_cta = ((1 << (n - a.tile)) + 147) // 148
        buf = np.zeros(ntiles_cta * 8, dtype=np.uint64)
        L.check(L.lib().qm_trace_read(st._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), len(buf)))
        t = buf.reshape(-1, 8).astype(np.int64)
        last = pt[-1]
        ngroups = int(last[1])
        teams = 2 if a.tile == 12 else 4
        t = t[8:-8]                                            # steady state
        end_body = t[:, 2 + min(ngroups, 3) - 1]
        out = {"tile": a.tile, "budget": cost, "last_pass_ms": float(last[0]), "groups": ngroups, "gates": int(last[2]),
               "wait_full": float(np.mean(t[:, 1] - t[:, 0])), "groups_total": float(np.mean(end_body - t[:, 1])),
               "arrive": float(np.mean(t[:, 5] - end_body)),
               "tile_period_per_team": float(np.mean(np.diff(t[::teams, 0]))),
               "producer_store_to_load": float(np.mean((t[:, 7] - t[:, 6])[t[:, 7] > 0])),
               "computed_seen_after_arrive": float(np.mean(t[:, 6] - t[:, 5]))}
        # overlap of body intervals [landed, end_body] between consecutive tiles (which belong to different teams)
        a0, a1 = t[:-1, 1], end_body[:-1]
        b0, b1 = t[1:, 1], end_body[1:]
        ov = np.maximum(0, np.minimum(a1, b1) - np.maximum(a0, b0))
        out["body_overlap_frac_of_body"] = float(np.mean(ov) / max(np.mean(a1 - a0), 1))
        print(json.dumps(out), flush=True)
        st.close()


if __name__ == "__main__":
    main()
