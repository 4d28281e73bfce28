"""World-size-N emulation of the sharded path on CPU (gloo): the LIBRARY's planner (qm_plan_gates with
global_bits = log2(world)) decides what runs locally and where the global<->local remaps go; the data
path is emulated with numpy shards (rank predicates for controls / diagonal targets on global qubits)
and torch.distributed all_to_all for the remap.  Rank 0 gathers the shards and compares with the
single-process oracle.  Launched by tests/test_dist_cpu.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import __graft_entry__ as ge  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_abi_cpu import random_circuit  # noqa: E402


def apply_local(shard, nl, rank, g, L):
    """one wire-format gate (physical bits, textbook) on this rank's shard"""
    kind, q0, q1 = int(g["kind"]), int(g["q0"]), int(g["q1"])
    U = g["m"].view(np.complex128).reshape(2, 2).T
    ctrl, tgt = (q0, q1) if kind == L.GATE_CU2 else (-1, q0)
    if ctrl >= nl:                               # control on a global qubit: rank predicate
        if not (rank >> (ctrl - nl)) & 1:
            return
        ctrl = -1
    if tgt >= nl:                                # diagonal target on a global qubit: rank-dependent phase
        assert U[0, 1] == 0 and U[1, 0] == 0, "planner placed a non-diagonal gate on a global qubit"
        d = U[1, 1] if (rank >> (tgt - nl)) & 1 else U[0, 0]
        if ctrl < 0:
            shard *= d
        else:
            idx = np.arange(len(shard))
            shard[(idx >> ctrl) & 1 == 1] *= d
        return
    if ctrl < 0:
        O.u2(shard, tgt + 1, U)
    else:
        O.controlled(shard, ctrl + 1, tgt + 1, U, quirk=False)


def remap(shard, nl, g, world, rank):
    chunk = 1 << (nl - g)
    send = [torch.from_numpy(shard[j * chunk:(j + 1) * chunk].view(np.float64).copy()) for j in range(world)]
    recv = [torch.empty_like(t) for t in send]
    dist.all_to_all(recv, send) if dist.get_backend() != "gloo" else _a2a_gloo(recv, send, rank, world)
    for j in range(world):
        shard[j * chunk:(j + 1) * chunk] = recv[j].numpy().view(np.complex128)


def _a2a_gloo(recv, send, rank, world):
    reqs = []
    for j in range(world):
        if j == rank:
            recv[j].copy_(send[j])
        else:
            reqs.append(dist.isend(send[j], j))
            reqs.append(dist.irecv(recv[j], j))
    for r in reqs:
        r.wait()


def main():
    n, tile, low, seed = (int(x) for x in sys.argv[1:5])
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    g = world.bit_length() - 1
    nl = n - g
    pkg = ge.load_package()
    L = pkg._lib
    gates = random_circuit(pkg, n, 220, seed)
    order, _ = L.plan_gates(n, gates, global_bits=g, tile_bits=tile, low_bits=low)
    psi0 = O.random_state(n, seed + 1)
    shard = psi0[rank << nl:(rank + 1) << nl].copy()
    swapped = False
    nremap = 0
    for gt in order:
        if int(gt["kind"]) == -1:
            remap(shard, nl, g, world, rank)
            swapped = not swapped
            nremap += 1
        else:
            apply_local(shard, nl, rank, gt, L)
    if swapped:
        remap(shard, nl, g, world, rank)
    parts = [torch.empty(2 << nl, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(torch.from_numpy(shard.view(np.float64).copy()), parts, dst=0)
    if rank == 0:
        got = np.concatenate([p.numpy().view(np.complex128) for p in parts])
        ref = O.apply_circuit(psi0.copy(), gates)
        err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        print("DIST_OK" if err < 1e-12 else "DIST_BAD", "err=%.2e" % err, "remaps=%d" % nremap, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
