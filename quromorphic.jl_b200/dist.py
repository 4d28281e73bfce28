"""dist.py — torch.distributed plumbing for sharded registers (one process per GPU).

The engine does its own NCCL traffic (grouped ncclSend/ncclRecv all-to-all inside libqm_b200.so); torch is
only used to agree on the NCCL unique id and for the bench's barriers.  Launch with torchrun
(`--master-addr 127.0.0.1`); RANK / LOCAL_RANK / WORLD_SIZE come from the environment."""
import os

from .qsim import B200State


def init_process_group(backend="gloo"):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=int(os.environ.get("RANK", "0")),
                                world_size=int(os.environ.get("WORLD_SIZE", "1")))
    return dist


def sharded_state(n, m=0, **kw):
    """collective: every rank calls it; returns this rank's shard handle"""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    box = [B200State.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    st = B200State.create_dist(n, rank, world, box[0], m=m, device=local_rank, **kw)
    if os.environ.get("QM_DIST_P2P", "1") != "0":
        # peer-memory exchange: all-gather the CUDA IPC handles of the shards, map them
        import ctypes as C
        from . import _lib as L
        mine = (C.c_uint8 * 64)()
        L.check(L.lib().qm_dist_ipc_export(st._h, mine))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(mine))
        blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
        rc = L.lib().qm_dist_ipc_import(st._h, blob)
        st.p2p = (rc == 0)
        oks = [None] * world
        dist.all_gather_object(oks, st.p2p)
        if not all(oks):
            raise RuntimeError("CUDA IPC mapping failed on some rank (%s): set QM_DIST_P2P=0 to use NCCL" % L.last_error())
    return st
