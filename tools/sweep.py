#!/usr/bin/env python
"""tools/sweep.py — times the fused-pass kernel over planner/tile settings in ONE process (one GPU
call).  Prints one JSON line per configuration: pass count, mean pass time, achieved HBM GB/s."""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--tiles", default="11,12,13")
    ap.add_argument("--lows", default="5")
    ap.add_argument("--costs", default="0,32,64,128")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--pipes", default="1")
    ap.add_argument("--jits", default="0", help="QM_OPT_JIT values: 0 interpreter, 1 compiled pass kernels (compiled in the first repetition)")
    ap.add_argument("--clocks", action="store_true", help="sample nvidia-smi clocks / power while a configuration runs")
    ap.add_argument("--reset", action="store_true", help="reset to |0..0> before every repetition")
    args = ap.parse_args()
    pkg = ge.load_package()
    L = pkg._lib
    n = args.qubits
    gates = pkg.circuits.c3_circuit(n=n, depth=args.depth)
    st = pkg.B200State(n, 0)
    for jit, pipe, tile, low, cost in itertools.product([int(x) for x in args.jits.split(",")], [int(x) for x in args.pipes.split(",")],
                                                   [int(x) for x in args.tiles.split(",")],
                                                   [int(x) for x in args.lows.split(",")],
                                                   [int(x) for x in args.costs.split(",")]):
        st.set_option(L.OPT_JIT, jit)
        st.set_option(L.OPT_PIPELINE, pipe)
        st.set_option(L.OPT_TILE_BITS, tile)
        st.set_option(L.OPT_LOW_BITS, low)
        st.set_option(L.OPT_MAX_COST, cost)
        best = None
        sampler = None
        if args.clocks:
            from bench import ClockSampler
            sampler = ClockSampler(0)
            sampler.start()
        for r in range(args.reps + 1):
            if r == 0 or args.reset:      # by default the state keeps evolving: dense amplitudes, realistic power draw
                st.reset(0)
            st.counters_reset()
            st.apply_circuit(gates)
            c = st.counters()
            if best is None or c["ms_device"] < best["ms_device"]:
                best = c
        clk = sampler.stop() if sampler else {}
        ms = best["ms_device"]
        per = ms / best["passes"]
        gbs = 32.0 * (1 << n) / (per * 1e-3) / 1e9
        js = L.jit_stats()
        print(json.dumps({"jit": jit, "compiled_passes": best["compiled_passes"], "jit_compile_ms": round(js["compile_ms"], 1), "jit_backend": js["backend"], "pipe": pipe, "tile": tile, "low": low, "max_cost": cost, "passes": best["passes"], "ms_total": round(ms, 2),
                          "ms_per_pass": round(per, 3), "hbm_gbs": round(gbs, 1), "gates_per_s": round(len(gates) / ms * 1e3, 1),
                          "plan_ms": round(best["ms_plan"], 2), "sm_mhz": clk.get("sm_mhz"), "power_w": clk.get("power_w"),
                          "reasons": clk.get("reasons")}), flush=True)
    st.close()


if __name__ == "__main__":
    main()
