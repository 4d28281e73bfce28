"""The N > 1 path on CPU: world_size-2 and -4 gloo runs of tests/dist_worker.py (the library's sharded
planner + an emulated data path) must reproduce the single-process oracle."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n,tile,low,seed", [(2, 9, 5, 2, 11), (2, 10, 6, 3, 12), (4, 10, 5, 2, 13)])
def test_sharded_planner_with_gloo_ranks(world, n, tile, low, seed):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(n), str(tile), str(low), str(seed)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DIST_OK" in outs[0], outs[0]
    assert "remaps=0" not in outs[0], "the test circuit must exercise at least one remap: " + outs[0]
