"""dist.py — host plumbing for sharded registers: torch.distributed (one process per GPU), or threads (all ranks in one process).

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


_local_key = [0x51D0]


def run_local_world(n, world, body, m=0, devices=None, options=None, timeout=600.0):
    """All `world` ranks of an n-qubit sharded register in THIS process, one host thread per rank (ctypes releases the
    GIL inside the engine, so collective calls of different ranks overlap as they would between processes).
    devices[r] is rank r's CUDA device (default: all on device 0 — how the sharded kernels, flag fences and mailbox
    collectives are exercised on a 1-GPU box).  body(rank, state) runs on every rank; returns the list of results."""
    import threading
    devices = devices or [0] * world
    _local_key[0] += 1
    key = _local_key[0]
    states, results, errors = [None] * world, [None] * world, [None] * world
    barrier = threading.Barrier(world)

    def worker(r):
        try:
            st = B200State.create_dist_local(n, r, world, key, m=m, device=devices[r])
            states[r] = st
            for k, v in (options or {}).items():
                st.set_option(k, v)
            barrier.wait(timeout)
            st.connect_local()
            barrier.wait(timeout)
            results[r] = body(r, st)
            barrier.wait(timeout)
        except BaseException as e:          # noqa: BLE001 — reported to the caller, the other ranks are released
            errors[r] = e
            barrier.abort()
        finally:
            try:
                if states[r] is not None:
                    states[r].sync()
            except Exception:
                pass

    threads = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    alive = [t.is_alive() for t in threads]
    for st in states:
        if st is not None and not any(alive):
            st.close()
    real = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    if any(alive):
        raise TimeoutError("a rank of the same-process world did not finish")
    if any(e is not None for e in errors):
        raise RuntimeError("a rank aborted")
    return results
