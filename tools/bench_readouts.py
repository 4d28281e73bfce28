#!/usr/bin/env python
"""tools/bench_readouts.py — the readout kernels against their rooflines on one B200 (CUDA events on the engine's
stream): all-qubit <Z> (one read), all-qubit <X>/<Y> (one read per window of qubits; FP64-bound: 24 FP64 ops per
amplitude per qubit) against the per-qubit loop it replaces, and the reduced density K6 at k = 4, 6, 8, 10
(8 * 2^(n+k) real FMAs on top of one read of the state: HBM-bound for small k, FP64-bound above)."""
import argparse
import ctypes as C
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
    ap.add_argument("--qubits", type=int, default=28)
    a = ap.parse_args()
    import torch
    pkg = ge.load_package()
    L = pkg._lib
    n = a.qubits
    st = pkg.B200State(n, 0)
    st.apply_circuit(pkg.circuits.c3_circuit(n=n, depth=4, seed=9))
    st.sync()
    amps, stream = C.c_void_p(), C.c_void_p()
    L.check(L.lib().qm_device_ptr(st._h, C.byref(amps), C.byref(stream)))
    ext = torch.cuda.ExternalStream(stream.value)

    def timed(fn, reps=3):
        fn()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext); fn(); e1.record(ext); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    state_gb = 16.0 * (1 << n) / 1e9
    hbm = 6554.9
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    fp64_peak = 59 * 148 * 1.965e9          # DFMA lanes per second at the maximum SM clock (tools/microbench/dfma_bench.cu)
    ms = timed(lambda: st.expect_z_all())
    print(json.dumps({"readout": "all-qubit Z (k_zpartial + k_zfinal)", "n": n, "ms": ms, "gbs": state_gb / ms * 1e3, "frac_hbm": state_gb / ms * 1e3 / hbm}), flush=True)
    for basis in ("x", "y"):
        ms = timed(lambda: st.expect_all(basis))
        windows = 1 if n <= 12 else 1 + -(-(n - 12) // 9)
        flops = 24.0 * n * (1 << n) / 2           # per pair and qubit: rotation 16 + abs2 6 + 2 adds, i.e. 12 per amplitude
        one = timed(lambda: pkg.qsim.measure_x(st, n // 2) if basis == "x" else pkg.qsim.measure_y(st, n // 2))
        print(json.dumps({"readout": "all-qubit %s (k_pauli_tile x %d reads)" % (basis.upper(), windows), "n": n, "ms": ms,
                          "reads_gbs": windows * state_gb / ms * 1e3, "fp64_ops_per_s": flops / (ms * 1e-3),
                          "frac_fp64_peak": flops / (ms * 1e-3) / fp64_peak,
                          "per_qubit_loop_ms_estimate": one * n, "one_qubit_ms": one}), flush=True)
    for k in (4, 6, 8, 10):
        for keep_low in (True, False):
            ms = timed(lambda: st.reduced_density(keep_low, k), reps=2)
            fma = 8.0 * (1 << (n + k)) / 2          # Hermitian half is not exploited by the kernel: 4 real FMAs per product, 2^(n+k) products
            fma = 4.0 * (1 << (n + k))
            print(json.dumps({"readout": "reduced density K6", "n": n, "k": k, "keep_low": keep_low, "ms": ms,
                              "read_gbs": state_gb / ms * 1e3, "frac_hbm": state_gb / ms * 1e3 / hbm,
                              "fma_per_s": fma / (ms * 1e-3), "frac_fp64_peak": fma / (ms * 1e-3) / fp64_peak}), flush=True)
    st.close()


if __name__ == "__main__":
    main()
