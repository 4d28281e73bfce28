#!/usr/bin/env python
"""bench.py — the reservoir hot path on B200, measured as BASELINE.json asks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 : config C3 — 30-qubit ComplexF64 random reservoir circuit, depth 200 (8,900 gates), fused
        gate passes on one B200; one step = the whole circuit from |0...0> plus the all-qubit <Z>
        readout.
N > 1 : config C4 — sharded register, 2^33 amplitudes (128 GiB) per GPU (34 q @ 2, 35 q @ 4,
        36 q @ 8), one step = encoding layer + one C3-style layer + all-qubit <Z> readout.

`value`  = gate applications per second in units of one gate over 2^30 amplitudes (at N = 1 this is
           plain gates/s of the 30-qubit circuit; at 36 q one gate counts 64), device-timed with the
           state resident in HBM.
`e2e`    = the same metric through the public C-ABI calls with HOST buffers, wall clock: qm_reset,
           qm_apply_circuit(host gate array) (planning + pass-program upload + kernels) and
           qm_expect_z_all (device -> host readout) all inside the timed region.
`roofline` is for the dominant kernel k_fused_tma: 32 * 2^n bytes per launch (16 read + 16 written
           per amplitude) over its mean launch duration from CUDA events on the
           launching stream, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` is oracle/qsim_oracle.c (matrix-free OpenMP restatement of the reference's
           arithmetic; the Julia reference itself cannot run here) on this box's host cores over a
           bounded sample of the same circuit.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3_QUBITS, C3_DEPTH, C3_SEED = 30, 200, 3001
SHARD_QUBITS = 33


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


KERNEL_SOURCES = ("jit_codegen.hpp", "encode_v4.hpp")


def kernel_source_hash():
    """Identifies the code of the dominant kernel (the compiled fused pass): the complete PTX text the library hands to the
    assembler for a fixed reference pass — the kernel template embedded in the library with generated groups spliced in,
    mangled entry name normalised — plus the sources of the generator and the encoder.  Taken from the loaded library, so it
    describes the binary that runs, not the files next to it."""
    import hashlib
    import re
    import __graft_entry__ as ge
    pkg = ge.load_package()
    gates = pkg.circuits.c3_circuit(n=14, depth=3, seed=3001)
    ptx = pkg._lib.jit_check(14, gates, max_cost=80.0, compile=False, want_ptx=True)["ptx"]
    ptx = re.sub(r"k_fused_tmaI(?:Li\d+E)+(?:Lb[01]E)+EEv", "k_fused_tmaI_EEv", ptx)     # template arguments: a defaulted one may be appended
    h = hashlib.sha256(ptx.encode())
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "quromorphic.jl_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def static_traffic(n):
    """DRAM bytes of one k_fused_tma launch.  It cannot be measured inside a plain run (it needs the profiler's
    counters), so it is a STATIC figure: the dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
    capture committed under profiles/, accepted only when that capture was taken from the very kernel text this run's
    library assembles (kernel_source_hash) and at this register size; otherwise null."""
    try:
        with open(os.path.join(ROOT, "profiles", "r3_traffic.json")) as f:
            t = json.load(f)["k_fused_tma"]
        if t["n_qubits"] == n and t["kernel_source_hash"] == kernel_source_hash():
            return t["dram_bytes_per_launch"], "static: %s (ncu --set full, one launch; same kernel sources)" % t["capture"]
    except Exception:
        pass
    return None, "not captured for this binary"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            try:
                pw.append(float(parts[3]))
            except ValueError:
                pass
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """every host core this process may use (torchrun exports OMP_NUM_THREADS=1: the oracle is told explicitly)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample(n, gates, budget_s, threads=None, keep_state=False):
    """time the oracle on a prefix of `gates` at n qubits: returns (gates/s, ngates, threads, seconds[, state])"""
    import numpy as np
    from oracle import oracle as O
    O.set_threads(threads or host_threads())
    s = O.statevector(n, 0)
    O.apply_circuit(s, gates[:1].copy())            # touch pages / warm up
    s = O.statevector(n, 0)
    t0 = time.perf_counter()
    done = 0
    chunk = 4
    while done < len(gates):
        O.apply_circuit(s, np.ascontiguousarray(gates[done:done + chunk]))
        done += min(chunk, len(gates) - done)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    if keep_state:
        return done / dt, done, O.num_threads(), dt, s
    return done / dt, done, O.num_threads(), dt


def parity_block(pkg, n, gates, ref_state, tol=1e-12):
    """the GPU path on the very gate prefix the CPU leg just ran (same size, same circuit), compared with the
    oracle's state: all-qubit <Z> with the reference's 1e-10 threshold and without one, the norm, both 4-qubit
    reduced density matrices (every amplitude enters them) and 64 seeded shots (bit for bit)."""
    import numpy as np
    from oracle import oracle as O
    L = pkg._lib
    g = np.ascontiguousarray(gates)
    # the timed steps run compiled pass kernels (QM_OPT_JIT): so does the check; the interpreter kernel of the shipped binary
    # runs the same gates too (its planner prices gates differently, so the passes differ and the two agree to rounding only)
    st0 = pkg.B200State(n, 0)
    st0.set_option(L.OPT_JIT, 0)
    st0.apply_circuit(g)
    z_interp = st0.expect_z_all(threshold=-1.0)
    norm_interp = st0.norm2()
    st0.close()
    st = pkg.B200State(n, 0)
    st.set_option(L.OPT_JIT, 1)
    st.apply_circuit(g)
    cc = st.counters()
    errs = {}
    for name, thr in (("z_thr1e-10", 1e-10), ("z_nothr", -1.0)):
        errs[name] = float(np.max(np.abs(st.expect_z_all(threshold=thr) - O.expect_z_all(ref_state, thr))))
    errs["norm2"] = abs(st.norm2() - O.norm2(ref_state))
    k = min(4, n - 1)
    errs["rho_low%d" % k] = float(np.max(np.abs(st.reduced_density(True, k) - O.reduced_density(ref_state, True, k))))
    errs["rho_high%d" % k] = float(np.max(np.abs(st.reduced_density(False, k) - O.reduced_density(ref_state, False, k))))
    shots_equal = bool(np.array_equal(st.sample(20261020, 64), O.sample(ref_state, 20261020, 64)))
    norm2 = st.norm2()
    errs["compiled_vs_interpreter_z"] = float(np.max(np.abs(st.expect_z_all(threshold=-1.0) - z_interp)))
    errs["compiled_vs_interpreter_norm2"] = abs(norm2 - norm_interp)
    st.close()
    worst = max(errs.values())
    compiled = bool(cc["compiled_passes"] == cc["passes"] and cc["passes"] > 0)
    return {"n_qubits": n, "gates": int(len(gates)), "against": "oracle/qsim_oracle.c on the same gates from |0..0>",
            "path": "compiled pass kernels (%d of %d passes)" % (cc["compiled_passes"], cc["passes"]),
            "max_abs_err": worst, "tol": tol, "errors": errs, "sampler_bit_exact": shots_equal, "norm2": norm2,
            "ok": bool(worst < tol and shots_equal and compiled)}


def pick_cpu_qubits(n):
    """largest register <= n whose state (16 * 2^n bytes) fits comfortably in host memory"""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    while n > 20 and (16 << n) * 1.5 > avail:
        n -= 1
    return n


def run_reference(args, rank):
    """--impl reference: the CPU restatement of the reference's own arithmetic on ALL host cores (set explicitly:
    torchrun exports OMP_NUM_THREADS=1).  N = 1: the C3 circuit itself.  N > 1: the C4 reservoir step, timed at the
    largest register that fits host memory and converted per amplitude (one gate over 2^30 amplitudes = one unit) —
    the config says which size was timed."""
    if rank != 0:
        return
    import __graft_entry__ as ge
    pkg = ge.load_package()
    n_full = C3_QUBITS if args.gpus == 1 else SHARD_QUBITS + (args.gpus.bit_length() - 1)
    n_cpu = pick_cpu_qubits(min(n_full, 30))
    if args.gpus == 1:
        gates = pkg.circuits.c3_circuit(n=n_cpu, depth=2 if n_cpu != n_full else C3_DEPTH, seed=C3_SEED)
        what = "C3 circuit"
    else:
        gates = pkg.circuits.c4_step(n_cpu, 0)
        what = "C4 reservoir step (encoding layer + C3-style layer)"
    budget = 6.0
    vals = []
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        gps, ng, thr, dt = cpu_sample(n_cpu, gates, budget)
        if i >= args.warmup:
            vals.append((gps, ng, dt))
    gps = sum(v[1] for v in vals) / sum(v[2] for v in vals)
    scale = 2.0 ** (n_cpu - 30)                    # convert to units of one gate over 2^30 amplitudes
    value = gps * scale
    sample = "first %d gates of the %s at %d qubits per step (state 16*2^%d B), %d OpenMP threads" % (
        vals[0][1], what, n_cpu, n_cpu, thr)
    cfg = workload_config(args.gpus, n_full, c4=args.gpus > 1)
    cfg["cpu_timed_qubits"] = n_cpu
    if n_cpu != n_full:
        cfg["cpu_extrapolation"] = ("a %d-qubit state does not fit host memory: the same gate sequence is timed at %d qubits and "
                                    "counted per amplitude (one gate over 2^30 amplitudes = one unit), i.e. the CPU cost per "
                                    "amplitude is assumed flat up to 2^%d" % (n_full, n_cpu, n_full))
    line = {
        "impl": "reference", "metric": "gates_per_sec", "value": value, "unit": "gates/s (one gate over 2^30 amplitudes)",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(v[2] for v in vals) / len(vals), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "gates/s (one gate over 2^30 amplitudes)", "cores": thr,
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "gates/s (one gate over 2^30 amplitudes)", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


ALT_MAX_COST = 160         # planner budget of the secondary "deep fusion" measurement (engine default: 80; two register-block groups per pass)


def st_max_cost(args):
    return int(args.max_cost) if args.max_cost >= 0 else 80


def workload_config(ngpus, n, c4=False):
    if ngpus == 1 and c4:
        return {"workload": "C4 base: %d-qubit reservoir step on one B200 (2^%d amplitudes), no exchange" % (n, n),
                "n_qubits": n, "amps_per_gpu_log2": n, "seed": 3601, "l2": "state is far larger than L2"}
    if ngpus == 1:
        return {"workload": "C3: %d-qubit ComplexF64 random reservoir circuit, depth %d, fused gate passes, "
                            "all-qubit <Z> readout" % (n, C3_DEPTH),
                "n_qubits": n, "depth": C3_DEPTH, "seed": C3_SEED, "state_bytes": 16 << n,
                "l2": "state (16 GiB) is far larger than the 126 MB L2; no flush needed"}
    return {"workload": "C4: %d-qubit reservoir step sharded over %d B200 (2^%d amplitudes per GPU), global qubits "
                        "remapped by in-place all-to-all over NVLink" % (n, ngpus, SHARD_QUBITS),
            "n_qubits": n, "amps_per_gpu_log2": SHARD_QUBITS, "seed": 3601,
            "l2": "shard (128 GiB) is far larger than L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the register size (testing)")
    ap.add_argument("--depth", type=int, default=C3_DEPTH)
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--low-bits", type=int, default=-1)
    ap.add_argument("--max-cost", type=float, default=-1)
    ap.add_argument("--short-pct", type=int, default=-1, help="QM_OPT_SHORT_PCT (default 180; 100 = off)")
    ap.add_argument("--jit", type=int, default=-1, help="QM_OPT_JIT (compiled pass kernels): -1 = default (C3: engine default 2; C4: 1, with --structure-warmup)")
    ap.add_argument("--structure-warmup", type=int, default=12,
                    help="C4: extra untimed reservoir steps in front of the W warm-up steps (the pass programs of a step repeat with period 6; 12 = every structure twice)")
    ap.add_argument("--defer-pass", type=int, default=-1, help="sharded: QM_OPT_DEFER_PASS (default: engine default, 1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-alt", action="store_true", help="skip the deep-fusion secondary measurement")
    ap.add_argument("--overlap", type=int, default=-1, help="sharded: QM_OPT_OVERLAP (default: engine default, 1)")
    ap.add_argument("--xsm", type=int, default=-1, help="sharded: SMs of the overlapped exchange kernel")
    ap.add_argument("--xun", type=int, default=-1, help="sharded: pairs in flight per exchange thread (4 or 8)")
    ap.add_argument("--slab-bits", type=int, default=-1, help="sharded: log2 of the number of slabs")
    ap.add_argument("--overlap-depth", type=int, default=-1, help="sharded: passes in front of an exchange that run slab by slab beside it")
    ap.add_argument("--zfold", type=int, default=1, help="C4: 1 = one call per step (qm_apply_circuit_expect_z, <Z> sums inside the last pass), 0 = circuit then readout")
    ap.add_argument("--workload", default="auto", choices=["auto", "c3", "c4"],
                    help="auto: C3 at N = 1, C4 at N > 1; c4 at N = 1 runs the 33-qubit weak-scaling base")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    L = pkg._lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    if world > 1 or args.workload == "c4":
        from bench_dist import run_dist            # sharded path (C4)
        run_dist(args, pkg, rank, world, local_rank)
        return

    torch.cuda.set_device(0)
    n = args.qubits or C3_QUBITS
    gates = pkg.circuits.c3_circuit(n=n, depth=args.depth, seed=C3_SEED)
    ngates = len(gates)
    st = pkg.B200State(n, 0)
    if args.tile_bits:
        st.set_option(L.OPT_TILE_BITS, args.tile_bits)
    if args.low_bits >= 0:
        st.set_option(L.OPT_LOW_BITS, args.low_bits)
    if args.max_cost >= 0:
        st.set_option(L.OPT_MAX_COST, int(args.max_cost))
    if args.short_pct >= 0:
        st.set_option(L.OPT_SHORT_PCT, args.short_pct)
    if args.jit >= 0:
        st.set_option(L.OPT_JIT, args.jit)
    import ctypes as C
    amps, stream = C.c_void_p(), C.c_void_p()
    L.check(L.lib().qm_device_ptr(st._h, C.byref(amps), C.byref(stream)))
    ext = torch.cuda.ExternalStream(stream.value)
    zout = np.empty((n, 2))

    def step():
        st.reset(0)
        st.apply_circuit(gates)
        return st.expect_z_all()

    for _ in range(args.warmup):
        step()
    st.sync()

    # ---- device-timed region: K steps, events on the library's own stream ----
    st.counters_reset()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    pass_ms, npasses = 0.0, 0
    t_wall0 = time.perf_counter()
    e0.record(ext)
    for _ in range(args.steps):
        z = step()
        c = st.counters()
        pass_ms += c["ms_device"]
    e1.record(ext)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    c = st.counters()
    npasses = c["passes"]
    unit = "gates/s (one gate over 2^30 amplitudes)"
    scale = 2.0 ** (n - 30)
    value = args.steps * ngates / (dev_ms * 1e-3) * scale
    e2e = args.steps * ngates / wall * scale

    hbm_peak, peak_kind = peaks()
    traffic, traffic_from = static_traffic(n)
    bytes_per_launch = 32.0 * (1 << n)
    avg_pass_ms = pass_ms / max(npasses, 1)
    achieved = bytes_per_launch / (avg_pass_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "k_fused_tma<12,2,3,JIT> (compiled pass kernels)" if c["compiled_passes"] else "k_fused_tma", "achieved": achieved, "peak": hbm_peak, "peak_kind": peak_kind,
            "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_from": traffic_from,
            "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_pass_ms, "launches": npasses,
            "gates_per_pass": args.steps * ngates / max(npasses, 1),
            "frac_of_nominal_8TBs": achieved / 8000.0}

    line = {
        "metric": "gates_per_sec", "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1, n),
        "roofline": roof, "clocks": clocks,
        "e2e": {"value": e2e, "unit": unit, "h2d_bytes_per_step": c["bytes_h2d"] / args.steps,
                "d2h_bytes_per_step": c["bytes_d2h"] / args.steps, "ms_per_step": 1e3 * wall / args.steps},
        "gpu_launches": int(c["launches"]), "gates_per_step": ngates,
        "readout_z0": [float(z[0][0]), float(z[0][1])],
    }
    js = L.jit_stats()
    line["compiled_passes"] = {"passes_compiled": int(c["compiled_passes"]), "passes": int(npasses), "kernels_in_cache": int(js["cached"]),
                               "failed": int(js["failed"]), "backend": js["backend"], "compile_wall_ms_total": js["compile_ms"],
                               "note": "compiled once, in the warm-up step that saw the circuit's structure for the second time; "
                                       "the timed steps launch cached kernels"}
    if args.depth != C3_DEPTH or n != C3_QUBITS:
        line["config"]["workload"] += " [NON-DEFAULT size: depth %d, %d qubits]" % (args.depth, n)
    line["config"]["planner_max_cost"] = st_max_cost(args)

    # ---- same workload with deeper fusion (two register-block groups per pass: fewer, FP64-bound passes): the other end of the
    # trade-off.  TWO untimed steps first: the first sees the new plan's opcode sequences once (and compiles those that repeat
    # inside the circuit), the second compiles the rest from the cached plan.  Until the last session of round 2 there was one,
    # so the leftover compilation (~1.9 s) sat inside this block's two timed steps and its `value` read 3,3xx gates/s while
    # its own per-pass times said 5,1xx (profiles/r3i, r3p, r4b: 46.8 gates per 9.07-9.25 ms pass) ----
    if not args.no_alt:
        st.set_option(L.OPT_MAX_COST, ALT_MAX_COST)
        step(); st.sync()
        step(); st.sync()
        st.counters_reset()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        tw = time.perf_counter()
        a0.record(ext)
        alt_pass_ms = 0.0
        for _ in range(2):
            step()
            alt_pass_ms += st.counters()["ms_device"]
        a1.record(ext)
        torch.cuda.synchronize()
        tw = time.perf_counter() - tw
        ca = st.counters()
        alt_ms = a0.elapsed_time(a1)
        alt_ach = bytes_per_launch / (alt_pass_ms / max(ca["passes"], 1) * 1e-3) / 1e9
        line["deep_fusion"] = {"planner_max_cost": ALT_MAX_COST, "value": 2 * ngates / (alt_ms * 1e-3) * scale,
                               "e2e": 2 * ngates / tw * scale, "unit": unit, "steps": 2,
                               "gates_per_pass": 2 * ngates / max(ca["passes"], 1),
                               "avg_launch_ms": alt_pass_ms / max(ca["passes"], 1),
                               "roofline_achieved_gbs": alt_ach, "roofline_frac": alt_ach / hbm_peak,
                               "steady_state_from_pass_times": (2 * ngates / max(ca["passes"], 1)) / (alt_pass_ms / max(ca["passes"], 1) * 1e-3) * scale,
                               "compiled_passes": int(ca["compiled_passes"]), "passes": int(ca["passes"]),
                               "note": "FP64-bound passes (two register-block groups); the default budget keeps passes HBM-bound"}
    # ---- the same workload through the interpreter kernel of the shipped binary (QM_OPT_JIT 0, its own planner budget) ----
    if not args.no_alt:
        st.set_option(L.OPT_JIT, 0)
        st.set_option(L.OPT_MAX_COST, 80)
        step(); st.sync()
        st.counters_reset()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a0.record(ext)
        step()
        ip_ms = st.counters()["ms_device"]
        a1.record(ext)
        torch.cuda.synchronize()
        ci = st.counters()
        ip_ach = bytes_per_launch / (ip_ms / max(ci["passes"], 1) * 1e-3) / 1e9
        line["interpreter"] = {"planner_max_cost": 80, "value": ngates / (a0.elapsed_time(a1) * 1e-3) * scale, "unit": unit, "steps": 1,
                               "gates_per_pass": ngates / max(ci["passes"], 1), "avg_launch_ms": ip_ms / max(ci["passes"], 1),
                               "roofline_achieved_gbs": ip_ach, "roofline_frac": ip_ach / hbm_peak,
                               "compiled_passes": int(ci["compiled_passes"])}
    st.close()
    del st
    torch.cuda.empty_cache()

    # ---- the weak-scaling base of the N > 1 lines (config C4 on ONE GPU: 2^33 amplitudes, same reservoir step, no
    # exchange), so that the 2/4/8-GPU values have their own-workload reference in the N = 1 line ----
    if not args.no_alt and n == C3_QUBITS and args.depth == C3_DEPTH:
        import copy
        a4 = copy.copy(args)
        a4.qubits, a4.max_cost, a4.tile_bits, a4.short_pct = 0, -1, 0, -1
        try:
            from bench_dist import run_dist
            b = run_dist(a4, pkg, 0, 1, 0, emit=False)
            line["c4_base"] = {"workload": b["config"]["workload"], "n_qubits": b["n_qubits"], "value": b["value"], "unit": unit,
                               "ms_per_step": b["ms_per_step"], "gates_per_step": b["gates_per_step"],
                               "passes_per_step": b["passes_per_step"], "fused_pass_gbs": b["roofline"]["achieved"],
                               "step_call": b.get("step_call"), "folded_readouts_per_step": b.get("folded_readouts_per_step"),
                               "step_readout_vs_separate_readout": b.get("step_readout_vs_separate_readout"), "norm2": b.get("norm2"),
                               "note": "weak-scaling efficiency of the N-GPU C4 lines = (ms_per_step / gates_per_step here) / (theirs)"}
        except Exception as e:            # the base is informational: never lose the headline line over it
            line["c4_base"] = {"error": str(e)[:200]}

    parity_ok = True
    if not args.no_cpu:
        n_cpu = pick_cpu_qubits(n)
        cg = gates if n_cpu == n else pkg.circuits.c3_circuit(n=n_cpu, depth=2, seed=C3_SEED)
        gps, ng, thr, dt, ref_state = cpu_sample(n_cpu, cg, 15.0, keep_state=True)
        line["cpu_baseline"] = {"value": gps * 2.0 ** (n_cpu - 30), "unit": unit, "cores": thr, "kind": "port",
                                "sample": "first %d gates of the same circuit at %d qubits from |0..0> "
                                          "(%.1f s of oracle/qsim_oracle.c, OpenMP)" % (ng, n_cpu, dt)}
        # parity at the headline size: the GPU path on the same gate prefix against the state the CPU leg just made
        try:
            line["parity"] = parity_block(pkg, n_cpu, cg[:ng], ref_state)
        except Exception as e:
            line["parity"] = {"ok": False, "error": str(e)[:300]}
        parity_ok = bool(line["parity"].get("ok"))
        del ref_state

    # ---- the other BASELINE.json configurations, so that the driver's record carries a number for each ----
    if not args.no_alt and n == C3_QUBITS and args.depth == C3_DEPTH:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        try:
            import bench_batched
            import bench_pipeline
            cpu_steps = 0 if args.no_cpu else 1
            line["c2_batched"] = bench_batched.run_c2(pkg, steps=8)
            line["c1"] = bench_pipeline.run_c1(pkg, 12, 100, 25 * cpu_steps)
            line["c5"] = bench_pipeline.run_c5(pkg, 20, 100, 4, 10 * cpu_steps)
        except Exception as e:
            line["other_configs_error"] = str(e)[:300]
    print(json.dumps(line), flush=True)
    if not parity_ok:
        sys.stderr.write("bench.py: PARITY FAILED against the oracle: %s\n" % json.dumps(line.get("parity")))
        sys.exit(1)


if __name__ == "__main__":
    main()
