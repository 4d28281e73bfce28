"""bench_dist.py — the N > 1 leg of bench.py: config C4, a register sharded over N B200s (one process per
GPU, launched by torchrun), 2^33 amplitudes (128 GiB) per GPU unless --qubits says otherwise.

One step = encoding layer (ry on every qubit) + one C3-style layer (rotation per qubit + brickwork
cnot!/crz!) + the all-qubit <Z> readout.  The global qubits are exchanged with the top local ones by the
engine's in-place chunked all-to-all (grouped ncclSend/ncclRecv over NVLink) about once per step.
Timing: barrier + synchronize on both sides, CUDA events on the engine's stream, MAX over ranks.
`value` counts one gate over 2^30 amplitudes as one unit, so the sizes are comparable: at 36 qubits a
gate counts 64."""
import ctypes as C
import json
import time


def run_dist(args, pkg, rank, world, local_rank, emit=True):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bench import SHARD_QUBITS, ClockSampler, peaks, workload_config
    L = pkg._lib
    torch.cuda.set_device(local_rank)
    single = world == 1                      # the 33-qubit weak-scaling base: same step, one GPU, no exchange
    if not single:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = world.bit_length() - 1
    n = args.qubits or (SHARD_QUBITS + g)
    parity, base = None, None
    if not single:
        parity = dist_parity(pkg, rank, world, g, args)
        # the weak-scaling base of this line: the same reservoir step on ONE GPU (2^33 amplitudes, no exchange), timed
        # by rank 0 on its own GPU right here, in the same run, while the other ranks wait
        if n == SHARD_QUBITS + g and not getattr(args, "no_alt", False):
            if rank == 0:
                import copy
                a1 = copy.copy(args)
                a1.qubits = 0
                try:
                    b = run_dist(a1, pkg, 0, 1, local_rank, emit=False)
                    base = {"n_qubits": b["n_qubits"], "ms_per_step": b["ms_per_step"], "gates_per_step": b["gates_per_step"],
                            "passes_per_step": b["passes_per_step"], "value": b["value"]}
                except Exception as e:
                    base = {"error": str(e)[:200]}
                torch.cuda.empty_cache()
            dist.barrier()
    st = pkg.B200State(n, 0) if single else pkg.dist.sharded_state(n, 0)
    for key, name in ((L.OPT_OVERLAP, "overlap"), (L.OPT_XSM, "xsm"), (L.OPT_XUN, "xun"), (L.OPT_SLAB_BITS, "slab_bits"),
                      (L.OPT_OVERLAP_DEPTH, "overlap_depth")):
        v = getattr(args, name, -1)
        if not single and v is not None and v >= 0:
            st.set_option(key, v)
    if args.tile_bits:
        st.set_option(L.OPT_TILE_BITS, args.tile_bits)
    if getattr(args, "low_bits", -1) >= 0:
        st.set_option(L.OPT_LOW_BITS, args.low_bits)
    if getattr(args, "defer_pass", -1) >= 0:
        st.set_option(L.OPT_DEFER_PASS, args.defer_pass)
    if args.max_cost >= 0:
        st.set_option(L.OPT_MAX_COST, int(args.max_cost))
    if args.short_pct >= 0:
        st.set_option(L.OPT_SHORT_PCT, args.short_pct)
    amps, stream = C.c_void_p(), C.c_void_p()
    L.check(L.lib().qm_device_ptr(st._h, C.byref(amps), C.byref(stream)))
    ext = torch.cuda.ExternalStream(stream.value)
    # Compiled pass kernels (QM_OPT_JIT): a reservoir step's pass programs repeat with period 6 (gate kind (q + l) mod 3,
    # cnot / crz brickwork l mod 2; the angles are new every step and are data of the launch, not part of the code).  A
    # reservoir runs for thousands of steps; the bench runs `structure_warmup` extra untimed steps first (default 12: every
    # structure twice) so that the W warm-up steps and the K timed steps launch cached kernels, as every later step would.
    jit = getattr(args, "jit", -1)
    st.set_option(L.OPT_JIT, 1 if jit < 0 else jit)
    sw = max(0, getattr(args, "structure_warmup", 12))
    pre = [pkg.circuits.c4_step(n, s) for s in range(sw)]
    steps = [pkg.circuits.c4_step(n, sw + s) for s in range(args.warmup + args.steps)]
    ngates = sum(len(x) for x in steps[args.warmup:])

    # one reservoir step = circuit + all-qubit <Z>.  Default: ONE call (qm_apply_circuit_expect_z), which takes the <Z> sums inside
    # the last fused pass when that pass is a compiled kernel over the whole register (QM_OPT_ZFOLD) — the separate 16 * 2^n
    # byte read of the readout kernels disappears; --zfold 0 makes the two calls one after the other as before.
    zfold = getattr(args, "zfold", 1) != 0

    def run_step(g):
        if zfold:
            return st.apply_circuit_expect_z(g)
        st.apply_circuit(g)
        return st.expect_z_all()

    def step(i):
        return run_step(steps[i])

    for gates_pre in pre:
        run_step(gates_pre)
    for i in range(args.warmup):
        step(i)
    st.sync()
    st.set_option(L.OPT_TIME_PASSES, 1)      # CUDA events around every whole-register pass (slab launches are not timed one by one)
    st.pass_times()
    st.counters_reset()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if not single:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(ext)
    pass_ms = 0.0
    for i in range(args.warmup, args.warmup + args.steps):
        z = step(i)
        pass_ms += st.counters()["ms_device"]
    e1.record(ext)
    torch.cuda.synchronize()
    if not single:
        dist.barrier()
    wall = time.perf_counter() - t0
    dev_ms = e0.elapsed_time(e1)
    c = st.counters()
    pt = st.pass_times()
    full_ms = float(pt[:, 0].mean()) if len(pt) else 0.0
    t = torch.tensor([dev_ms, wall, c["ms_remap"], pass_ms, full_ms], dtype=torch.float64, device="cuda")
    if not single:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall, ms_remap, pass_ms, full_ms = (float(x) for x in t.tolist())
    clocks = sampler.stop() if rank == 0 else None
    # after the timed region: the literal (unthresholded) norm and readout of the final state — at 36 qubits the
    # reference's 1e-10 threshold drops most of the mass, so the thresholded readout alone proves nothing
    norm2 = st.norm2()
    z_nothr = st.expect_z_all(threshold=-1.0)
    # ... and the last timed step's readout (taken inside its last pass when the fold applies) against the ordinary readout
    # kernels on the very same final state, at full size (collective calls: every rank makes them)
    fold_diff = float(np.max(np.abs(np.asarray(z) - np.asarray(st.expect_z_all())))) if args.steps > 0 else 0.0
    if rank == 0:
        unit = "gates/s (one gate over 2^30 amplitudes)"
        scale = 2.0 ** (n - 30)
        hbm_peak, peak_kind = peaks()
        nl = n - g
        npasses = c["passes"]
        # the fused pass on its own: mean CUDA-event time of the whole-register launches (the two passes that run slab by
        # slab beside an exchange cannot be separated from it and are left out of this figure)
        achieved = 32.0 * (1 << nl) / (max(full_ms, 1e-9) * 1e-3) / 1e9
        link_gbs = (c["bytes_link"] / (ms_remap * 1e-3) / 1e9) if ms_remap > 0 else None
        line = {
            "metric": "gates_per_sec", "value": ngates / (dev_ms * 1e-3) * scale, "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, n, c4=True),
            "roofline": {"bound": "hbm", "kernel": "k_fused_tma", "achieved": achieved, "peak": hbm_peak, "peak_kind": peak_kind,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "bytes_per_launch": 32.0 * (1 << nl), "launches": int(len(pt)), "avg_launch_ms": full_ms, "per": "GPU",
                         "note": "whole-register launches only; passes cut into slabs beside an exchange are not in this mean"},
            "link": {"bound": "nvlink", "achieved": link_gbs, "peak": 770.0, "unit": "GB/s per GPU per direction",
                     "frac": (link_gbs / 770.0) if link_gbs else None, "ms_remap_per_step": ms_remap / args.steps,
                     "note": "span of the exchange kernels on their own stream; with the overlapped path the passes of the neighbouring slabs run inside this span",
                     "bytes_sent_per_gpu_per_step": c["bytes_link"] / args.steps,
                     "peak_kind": "measured peer copy on this pool (B200_PROFILING.md)"},
            "clocks": clocks,
            "e2e": {"value": ngates / wall * scale, "unit": unit, "h2d_bytes_per_step": c["bytes_h2d"] / args.steps,
                    "d2h_bytes_per_step": c["bytes_d2h"] / args.steps, "ms_per_step": 1e3 * wall / args.steps},
            "gpu_launches": int(c["launches"]) * world, "gates_per_step": ngates / args.steps,
            "gates_per_sec_absolute": ngates / (dev_ms * 1e-3), "n_qubits": n,
            "readout_z0": [float(z[0][0]), float(z[0][1])],
            "readout_z0_unthresholded": [float(z_nothr[0][0]), float(z_nothr[0][1])],
            "norm2": norm2, "norm2_ok": bool(abs(norm2 - 1.0) < 1e-12),
            "passes_per_step": npasses / args.steps,
            "overlapped_remaps_per_step": c.get("overlapped_remaps", 0) / args.steps,
            "overlapped_passes_per_step": c.get("overlapped_passes", 0) / args.steps,
            "compiled_passes_per_step": c.get("compiled_passes", 0) / args.steps,
            "deferred_passes_per_step": c.get("deferred_passes", 0) / args.steps,
            "folded_readouts_per_step": c.get("folded_readouts", 0) / args.steps,
            "step_call": "qm_apply_circuit_expect_z" if zfold else "qm_apply_circuit + qm_expect_z_all",
            "step_readout_vs_separate_readout": fold_diff, "step_readout_ok": bool(fold_diff < 1e-12),
            "structure_warmup_steps": sw,
            "compiled_pass_kernels": dict(L.jit_stats(), note="rank 0's process; compiled during the untimed structure warm-up"),
        }
        if parity is not None:
            line["parity"] = parity
        if base is not None:
            line["c4_base_same_run"] = base
            if "ms_per_step" in base:
                line["weak_scaling_eff"] = (base["ms_per_step"] / base["gates_per_step"]) / ((dev_ms / args.steps) / (ngates / args.steps))
        if n != SHARD_QUBITS + g:
            line["config"]["workload"] += " [NON-DEFAULT size: %d qubits]" % n
        if emit:
            print(json.dumps(line), flush=True)
    if not single:
        dist.barrier()
    st.close()
    if not single:
        dist.destroy_process_group()
    if rank == 0 and emit and parity is not None and not (parity.get("ok") and line["norm2_ok"] and line["step_readout_ok"]):
        import sys
        sys.stderr.write("bench_dist.py: PARITY FAILED: %s norm2=%r\n" % (json.dumps(parity), norm2))
        sys.exit(1)
    return line if rank == 0 else None


def dist_parity(pkg, rank, world, g, args=None):
    """Before anything is timed: a (21+g)-qubit register sharded over the same ranks runs a 260-gate random circuit
    (several global<->local exchanges, the overlapped path included) and every collective readout — amplitudes, <Z>
    with and without the threshold, X/Y/Z marginals of local and global qubits, both reduced densities, the seeded
    sampler bit for bit, a reset after an odd number of layout changes — is compared with the CPU oracle on every rank
    (tests/gpu_dist_worker.py check_sharded: the same body the GPU tests run)."""
    import os
    import sys
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(root, "tests"))
    import gpu_dist_worker as W
    from oracle import oracle as O
    from bench import host_threads
    O.set_threads(max(1, host_threads() // world))
    n = 21 + g
    R = W.reference(pkg, n)
    st = pkg.dist.sharded_state(n, R["m"])
    L = pkg._lib
    for key, name in ((L.OPT_OVERLAP, "overlap"), (L.OPT_XSM, "xsm"), (L.OPT_XUN, "xun"), (L.OPT_SLAB_BITS, "slab_bits"),
                      (L.OPT_OVERLAP_DEPTH, "overlap_depth")):        # the exchange settings under test are the ones checked
        v = getattr(args, name, -1) if args is not None else -1
        if v is not None and v >= 0:
            st.set_option(key, v)
    out = W.check_sharded(pkg, st, rank, world, R)
    dist.barrier()                       # nobody frees a shard another rank may still be writing flags into
    st.close()
    worst = max(out[k] for k in ("err", "zerr", "z0err", "nerr", "merr", "rerr", "err_after_reset"))
    t = torch.tensor([worst, 0.0 if out["ok"] else 1.0, 0.0 if out["sample_bitexact"] else 1.0, float(out["overlapped_remaps"])],
                     dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst, bad, sbad, ovl = (float(x) for x in t.tolist())
    return {"n_qubits": n, "ranks": world, "gates": int(len(R["gates"])), "against": "oracle/qsim_oracle.c, every rank's shard",
            "max_abs_err": worst, "tol": 1e-12, "sampler_bit_exact": sbad == 0.0, "remap_bytes_rank0": out["link_bytes"],
            "overlapped_remaps": int(ovl), "ok": bad == 0.0, "rank0": {k: out[k] for k in ("err", "zerr", "z0err", "nerr", "merr", "rerr", "err_after_reset")}}
