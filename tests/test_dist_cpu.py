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


import numpy as np

import emu  # noqa: E402  (tests/emu.py: the host emulator library)
import pytest


@pytest.mark.parametrize("n,g,kind,seed", [(14, 1, "random", 1), (15, 2, "random", 2), (16, 3, "random", 3), (14, 2, "c4", 4),
                                           (15, 1, "c4", 5), (16, 2, "c3deep", 6), (15, 3, "perm", 7)])
def test_sharded_encoded_programs_emulated_on_cpu(pkg, emu_mode, n, g, kind, seed):
    """every rank's ENCODED pass programs (the bytes the TMA kernel is launched with) walked on the host over the
    rank's shard — rank predicates for controls on global qubits, rank-dependent phases for diagonal targets on them,
    permutation tails folded into the stores, the short-circuit planner across exchanges — with the chunk exchange
    and the relabelling done as the engine does them; the whole register must equal the oracle's."""
    from oracle import oracle as O
    from test_abi_cpu import random_circuit
    C = pkg.circuits
    if kind == "random":
        chunks = [random_circuit(pkg, n, 260, 5200 + seed)]
    elif kind == "c4":
        chunks = [C.c4_step(n, s) for s in range(3)]              # short circuits, consecutive steps on one register
    elif kind == "c3deep":
        chunks = [C.c3_circuit(n=n, depth=36, seed=77)]           # > 768 gates: budget-limited passes
    else:
        rng = np.random.default_rng(seed)
        gl = C.GateList(n, "corrected")
        for _ in range(200):
            t, c = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
            if c == t: gl.x(t)
            elif rng.uniform() < 0.7: gl.cnot(c, t)
            else: gl.swap(c, t)
        chunks = [gl.array()]
    psi = O.random_state(n, 40 + seed)
    ref = psi.copy()
    remaps = 0
    for gates in chunks:
        psi, npass, nrem = emu.plan_emulate_sharded(n, g, gates, psi, max_cost=80.0)
        ref = O.apply_circuit(ref, gates)
        remaps += nrem
        assert npass > 0
    assert remaps > 0                                             # the global qubits did get non-diagonal gates
    if kind == "perm":
        assert np.array_equal(psi, ref)
    else:
        assert np.max(np.abs(psi - ref)) / np.max(np.abs(ref)) < 1e-12


@pytest.mark.parametrize("n,g,low", [(19, 1, 4), (20, 2, 3), (19, 1, 5)])
def test_last_pass_before_an_exchange_is_postponed_when_that_saves_a_pass(pkg, emu_mode, n, g, low):
    """a reservoir layer on a sharded register: the last pass before the exchange does not target an outgoing qubit, and
    after the exchange it fits one tile together with the gates of the ex-global qubits (planner.hpp try_defer_last_pass):
    one pass per step less, same amplitudes"""
    from oracle import oracle as O
    C = pkg.circuits
    res = {}
    for defer in (0, 1):
        emu.set_defer(defer)
        try:
            psi = O.random_state(n, 90 + n)
            ref = psi.copy()
            passes = deferred = 0
            for s in range(2):
                gates = C.c4_step(n, s)
                psi, npass, nrem = emu.plan_emulate_sharded(n, g, gates, psi, low_bits=low, max_cost=80.0)
                ref = O.apply_circuit(ref, gates)
                passes += npass
                deferred += emu.last_deferred()
                assert nrem >= 1
            assert np.max(np.abs(psi - ref)) / np.max(np.abs(ref)) < 1e-12
            res[defer] = (passes, deferred)
        finally:
            emu.set_defer(1)
    assert res[0][1] == 0 and res[1][1] > 0, res
    assert res[1][0] < res[0][0], res


def test_reservoir_step_on_a_sharded_register_takes_as_many_passes_as_the_local_base(pkg):
    """planning only, at the bench's sizes (config C4): 33 local qubits per GPU.  With the outgoing qubits scheduled first and
    the last pass postponed across the exchange a step takes 4 passes at 34 and 35 qubits (like the 33-qubit base) and 4.5 at
    36, instead of 5 everywhere."""
    C = pkg.circuits
    emu.set_defer(1)
    for n, g, bound in ((33, 0, 4.0), (34, 1, 4.2), (35, 2, 4.2), (36, 3, 4.5)):
        tot = 0
        for s in range(12, 24):
            passes, remaps, deferred, seq = emu.plan_count_sharded(n, g, C.c4_step(n, s), low_bits=4, max_cost=80.0, cost_model=1)
            assert remaps == (1 if g else 0)
            tot += passes
        assert tot / 12.0 <= bound, (n, g, tot / 12.0)
