"""GPU check of the sharded engine: gates, remaps (overlapped with the passes when the plan allows) and every collective
readout must match the single-process oracle.  check_sharded() is the body; it runs
  * one process per GPU (spawned by tests/test_gpu_parity.py::test_sharded_two_ranks, torchrun-style environment), and
  * all ranks as threads of one process on ONE GPU (tests/test_gpu_parity.py::test_sharded_same_process_world), and
  * as the parity block of bench_dist.py at every N."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def reference(pkg, n, seed_base=4100, ngates=260, m=3):
    """the circuit and the oracle's answers (computed once, shared by all ranks)"""
    from oracle import oracle as O
    from test_abi_cpu import random_circuit
    gates = random_circuit(pkg, n, ngates, seed_base + n)
    ref = O.apply_circuit(O.statevector(n, m), gates)
    gates2 = random_circuit(pkg, n, 60, seed_base + 500 + n)
    ref2 = O.apply_circuit(O.statevector(n, 1), gates2)
    ks = ((True, 3), (False, 2), (False, min(4, n // 2)), (True, 5))
    return {"n": n, "m": m, "gates": gates, "ref": ref, "z": O.expect_z_all(ref, 1e-10), "z0": O.expect_z_all(ref, -1.0),
            "norm2": O.norm2(ref),
            "meas": {(b, t): f(ref, t) for t in (1, n // 2, n)
                     for b, f in (("z", O.measure_z), ("x", O.measure_x), ("y", O.measure_y))},
            "rho": {kk: O.reduced_density(ref, kk[0], kk[1]) for kk in ks},
            "shots": O.sample(ref, 20261020 + n, 257), "gates2": gates2, "ref2": ref2}


def check_sharded(pkg, st, rank, world, R):
    """every rank calls this with its shard handle `st` (created at basis state R["m"]); returns a dict of errors"""
    n = R["n"]
    g = world.bit_length() - 1
    nl = n - g
    Q = pkg.qsim
    st.apply_circuit(R["gates"])
    z = st.expect_z_all()
    out = {"zerr": float(np.max(np.abs(z - R["z"]))),
           "z0err": float(np.max(np.abs(st.expect_z_all(threshold=-1.0) - R["z0"]))),
           "nerr": abs(st.norm2() - R["norm2"])}
    merr = 0.0
    for (b, t), want in R["meas"].items():
        got = getattr(Q, "measure_" + b)(st, t)
        merr = max(merr, float(np.max(np.abs(np.array(got) - np.array(want)))))
    out["merr"] = merr
    # reduced densities (QInfo.partial_trace of psi psi') and the seeded sampler on the sharded register: rho_B needs
    # one all-reduce, rho_A brings the kept qubits home by exact swaps and puts them back, the sampler must agree with
    # the single-device definition bit for bit
    rerr = 0.0
    for (keep_low, k), want in R["rho"].items():
        rerr = max(rerr, float(np.max(np.abs(st.reduced_density(keep_low, k) - want))))
    out["rerr"] = rerr
    out["sample_bitexact"] = bool(np.array_equal(st.sample(20261020 + n, 257), R["shots"]))
    c = st.counters()
    shard = st.to_numpy()                        # canonical layout: this rank's contiguous slice (the swaps undone)
    ref = R["ref"]
    out["err"] = float(np.max(np.abs(shard - ref[rank << nl:(rank + 1) << nl])) / np.max(np.abs(ref)))
    out["link_bytes"], out["passes"], out["overlapped_remaps"] = int(c["bytes_link"]), int(c["passes"]), int(c["overlapped_remaps"])
    # reset after an odd number of layout changes, then a second circuit: the layout flag must follow the reset
    st.reset(1)
    st.apply_circuit(R["gates2"])
    shard2 = st.to_numpy()
    ref2 = R["ref2"]
    out["err_after_reset"] = float(np.max(np.abs(shard2 - ref2[rank << nl:(rank + 1) << nl])) / np.max(np.abs(ref2)))
    out["ok"] = bool(max(out["err"], out["zerr"], out["z0err"], out["nerr"], out["merr"], out["rerr"], out["err_after_reset"]) < 1e-12
                     and out["sample_bitexact"] and out["link_bytes"] > 0)
    return out


def main():
    import __graft_entry__ as ge
    n = int(sys.argv[1])
    pkg = ge.load_package()
    dist = pkg.dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    R = reference(pkg, n)
    st = pkg.dist.sharded_state(n, R["m"])
    out = check_sharded(pkg, st, rank, world, R)
    print("%s rank=%d %s" % ("GPU_DIST_OK" if out["ok"] else "GPU_DIST_BAD", rank,
                             " ".join("%s=%s" % (k, ("%.2e" % v) if isinstance(v, float) else v) for k, v in out.items())), flush=True)
    st.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
