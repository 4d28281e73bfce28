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
    st = pkg.B200State(n, 0) if single else pkg.dist.sharded_state(n, 0)
    if args.tile_bits:
        st.set_option(L.OPT_TILE_BITS, args.tile_bits)
    if args.max_cost >= 0:
        st.set_option(L.OPT_MAX_COST, int(args.max_cost))
    if args.short_pct >= 0:
        st.set_option(L.OPT_SHORT_PCT, args.short_pct)
    amps, stream = C.c_void_p(), C.c_void_p()
    L.check(L.lib().qm_device_ptr(st._h, C.byref(amps), C.byref(stream)))
    ext = torch.cuda.ExternalStream(stream.value)
    steps = [pkg.circuits.c4_step(n, s) for s in range(args.warmup + args.steps)]
    ngates = sum(len(x) for x in steps[args.warmup:])

    def step(i):
        st.apply_circuit(steps[i])
        return st.expect_z_all()

    for i in range(args.warmup):
        step(i)
    st.sync()
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
    t = torch.tensor([dev_ms, wall, c["ms_remap"], pass_ms], dtype=torch.float64, device="cuda")
    if not single:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall, ms_remap, pass_ms = (float(x) for x in t.tolist())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        unit = "gates/s (one gate over 2^30 amplitudes)"
        scale = 2.0 ** (n - 30)
        hbm_peak, peak_kind = peaks()
        nl = n - g
        npasses = c["passes"]
        pure_pass_ms = max(pass_ms - ms_remap, 1e-9)
        achieved = 32.0 * (1 << nl) * npasses / (pure_pass_ms * 1e-3) / 1e9
        link_gbs = (c["bytes_link"] / (ms_remap * 1e-3) / 1e9) if ms_remap > 0 else None
        line = {
            "metric": "gates_per_sec", "value": ngates / (dev_ms * 1e-3) * scale, "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, n, c4=True),
            "roofline": {"bound": "hbm", "kernel": "k_fused_tma", "achieved": achieved, "peak": hbm_peak, "peak_kind": peak_kind,
                         "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "bytes_per_launch": 32.0 * (1 << nl), "launches": npasses, "per": "GPU"},
            "link": {"bound": "nvlink", "achieved": link_gbs, "peak": 770.0, "unit": "GB/s per GPU per direction",
                     "frac": (link_gbs / 770.0) if link_gbs else None, "ms_remap_per_step": ms_remap / args.steps,
                     "bytes_sent_per_gpu_per_step": c["bytes_link"] / args.steps,
                     "peak_kind": "measured peer copy on this pool (B200_PROFILING.md)"},
            "clocks": clocks,
            "e2e": {"value": ngates / wall * scale, "unit": unit, "h2d_bytes_per_step": c["bytes_h2d"] / args.steps,
                    "d2h_bytes_per_step": c["bytes_d2h"] / args.steps, "ms_per_step": 1e3 * wall / args.steps},
            "gpu_launches": int(c["launches"]) * world, "gates_per_step": ngates / args.steps,
            "gates_per_sec_absolute": ngates / (dev_ms * 1e-3), "n_qubits": n,
            "readout_z0": [float(z[0][0]), float(z[0][1])],
            "passes_per_step": npasses / args.steps,
        }
        if n != SHARD_QUBITS + g:
            line["config"]["workload"] += " [NON-DEFAULT size: %d qubits]" % n
        if emit:
            print(json.dumps(line), flush=True)
    st.close()
    if not single:
        dist.destroy_process_group()
    return line if rank == 0 else None
