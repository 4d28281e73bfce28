"""2-rank GPU check of the sharded engine (spawned by tests/test_gpu_parity.py::test_sharded_two_gpus):
each rank owns one GPU; gates, remaps and readouts must match the single-process oracle."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import __graft_entry__ as ge  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_abi_cpu import random_circuit  # noqa: E402


def main():
    n = int(sys.argv[1])
    pkg = ge.load_package()
    dist = pkg.dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = world.bit_length() - 1
    nl = n - g
    st = pkg.dist.sharded_state(n, 3)
    gates = random_circuit(pkg, n, 260, 4100 + n)
    st.apply_circuit(gates)
    ref = O.apply_circuit(O.statevector(n, 3), gates)
    z = st.expect_z_all()
    zerr = float(np.max(np.abs(z - O.expect_z_all(ref, 1e-10))))
    nerr = abs(st.norm2() - O.norm2(ref))
    Q = pkg.qsim
    merr = 0.0
    for t in (1, n // 2, n):
        for f, fo in ((Q.measure_z, O.measure_z), (Q.measure_x, O.measure_x), (Q.measure_y, O.measure_y)):
            merr = max(merr, float(np.max(np.abs(np.array(f(st, t)) - np.array(fo(ref, t))))))
    # reduced densities (QInfo.partial_trace of psi psi') and the seeded sampler on the sharded register: rho_B needs
    # one all-reduce, rho_A brings the kept qubits home by exact swaps and puts them back, the sampler must agree with
    # the single-device definition bit for bit
    rerr = 0.0
    for keep_low, k in ((True, 3), (False, 2), (False, min(4, n // 2)), (True, 5)):
        got = st.reduced_density(keep_low, k)
        want = O.reduced_density(ref, keep_low, k)
        rerr = max(rerr, float(np.max(np.abs(got - want))))
    shots = st.sample(20261020 + n, 257)
    sample_ok = bool(np.array_equal(shots, O.sample(ref, 20261020 + n, 257)))
    c = st.counters()
    shard = st.to_numpy()                        # canonical layout: this rank's contiguous slice (the swaps undone)
    err = float(np.max(np.abs(shard - ref[rank << nl:(rank + 1) << nl])) / np.max(np.abs(ref)))
    ok = (err < 1e-12 and zerr < 1e-12 and nerr < 1e-12 and merr < 1e-12 and rerr < 1e-12 and sample_ok
          and c["bytes_link"] > 0)
    print("%s rank=%d err=%.2e zerr=%.2e nerr=%.2e merr=%.2e rerr=%.2e sample_bitexact=%s link_bytes=%d passes=%d"
          % ("GPU_DIST_OK" if ok else "GPU_DIST_BAD", rank, err, zerr, nerr, merr, rerr, sample_ok, c["bytes_link"],
             c["passes"]), flush=True)
    st.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
