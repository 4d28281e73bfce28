"""GPU parity tests of the FOLDED readout (qm_apply_circuit_expect_z, QM_OPT_ZFOLD): the all-qubit measure_z sums
(QSim.jl:297-318) are taken inside the last fused pass of the circuit — the ZFOLD instantiation of the compiled pass kernel,
csrc/fused_tma_kernel.cuh — instead of by the readout kernels.  Bar: 1e-12 against the CPU oracle (threshold rule included),
1e-14 against the separate readout of the same state, and the state the pass leaves behind must be the same as without the fold."""
import numpy as np
import pytest

from oracle import oracle as O
from test_abi_cpu import random_circuit

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("n,cost,seed", [(12, 80, 1), (13, 80, 2), (14, 24, 3), (16, 0, 4), (20, 80, 5), (22, 120, 6), (21, 250, 7)])
def test_folded_readout_matches_oracle_and_separate_readout(pkg, n, cost, seed):
    L = pkg._lib
    gates = random_circuit(pkg, n, 200, 7300 + seed)
    psi0 = O.random_state(n, 60 + seed)
    ref = O.apply_circuit(psi0.copy(), gates)
    st = pkg.B200State.from_numpy(psi0)
    st.set_option(L.OPT_JIT, 1)
    st.set_option(L.OPT_MAX_COST, cost)
    # a threshold that really drops amplitudes (about the median probability), the reference's own 1e-10, and none
    thr_mid = float(np.median(np.abs(ref) ** 2))
    z = st.apply_circuit_expect_z(gates, 1e-10)
    c = st.counters()
    assert c["folded_readouts"] == 1 and c["compiled_passes"] == c["passes"] > 0, c
    assert float(np.max(np.abs(z - O.expect_z_all(ref, 1e-10)))) < RTOL
    assert float(np.max(np.abs(z - st.expect_z_all(1e-10)))) < 1e-14
    assert rel_err(st.to_numpy(), ref) < RTOL                     # the pass itself is not disturbed by the fold
    # second and third step on top (a reservoir loop): other thresholds, the kernels come from the cache now
    for thr in (thr_mid, -1.0):
        ref = O.apply_circuit(ref, gates)
        z = st.apply_circuit_expect_z(gates, thr)
        assert float(np.max(np.abs(z - O.expect_z_all(ref, thr)))) < RTOL, thr
        assert float(np.max(np.abs(z - st.expect_z_all(thr)))) < 1e-14, thr
        if thr == thr_mid:
            assert 0.05 < float(z[0].sum()) < 0.999              # the threshold did drop mass, and the sums say so
    assert st.counters()["folded_readouts"] == 3
    assert rel_err(st.to_numpy(), ref) < RTOL
    assert L.jit_stats()["failed"] == 0
    st.close()


def test_folded_readout_is_deterministic_and_optional(pkg):
    """same circuit twice from the same state: the folded sums agree to the last bit (fixed summation order); with QM_OPT_ZFOLD 0,
    with the interpreter kernel, and for a batch the call falls back to circuit + ordinary readout with the same results"""
    L = pkg._lib
    n = 17
    gates = random_circuit(pkg, n, 150, 8100)
    psi0 = O.random_state(n, 5)
    ref = O.apply_circuit(psi0.copy(), gates)
    want = O.expect_z_all(ref, 1e-10)
    runs = []
    for rep in range(2):
        st = pkg.B200State.from_numpy(psi0)
        st.set_option(L.OPT_JIT, 1)
        runs.append(st.apply_circuit_expect_z(gates, 1e-10))
        assert st.counters()["folded_readouts"] == 1
        st.close()
    assert np.array_equal(runs[0], runs[1])
    assert float(np.max(np.abs(runs[0] - want))) < RTOL
    for opts in ({L.OPT_JIT: 1, L.OPT_ZFOLD: 0}, {L.OPT_JIT: 0}, {}):          # {}: default JIT gate (small register: interpreter)
        st = pkg.B200State.from_numpy(psi0)
        for k, v in opts.items():
            st.set_option(k, v)
        z = st.apply_circuit_expect_z(gates, 1e-10)
        assert st.counters()["folded_readouts"] == 0
        assert float(np.max(np.abs(z - want))) < RTOL
        assert rel_err(st.to_numpy(), ref) < RTOL
        st.close()
    st = pkg.B200State(n, 3, batch=4)
    st.set_option(L.OPT_JIT, 1)
    zb = st.apply_circuit_expect_z(gates, 1e-10)
    refb = O.apply_circuit(O.statevector(n, 3), gates)
    assert zb.shape == (4, n, 2) and st.counters()["folded_readouts"] == 0
    for b in range(4):
        assert float(np.max(np.abs(zb[b] - O.expect_z_all(refb, 1e-10)))) < RTOL
    st.close()


def test_folded_readout_c4_steps(pkg):
    """the reservoir step the fold is for (config C4: encoding layer + reservoir layer + <Z> of every qubit), 24 qubits, six
    steps = every step structure once; every step's readout against the oracle"""
    L = pkg._lib
    n = 24
    st = pkg.B200State(n, 0)
    st.set_option(L.OPT_JIT, 1)
    ref = O.statevector(n, 0)
    for s in range(6):
        g = pkg.circuits.c4_step(n, s)
        ref = O.apply_circuit(ref, g)
        z = st.apply_circuit_expect_z(g)
        assert float(np.max(np.abs(z - O.expect_z_all(ref, 1e-10)))) < RTOL, s
    c = st.counters()
    assert c["folded_readouts"] == 6, c
    assert abs(st.norm2() - 1.0) < 1e-12
    st.close()


@pytest.mark.timeout(150, method="thread")
@pytest.mark.parametrize("n,world", [(23, 2), (22, 4)])
def test_folded_readout_on_a_sharded_register(pkg, n, world):
    """sharded register (all ranks in this process, one GPU): every rank folds its shard's sums into its last pass when that pass
    is a whole-register launch; the collective finish is the one of qm_expect_z_all"""
    L = pkg._lib
    gates = random_circuit(pkg, n, 220, 8800 + n)
    ref = O.apply_circuit(O.statevector(n, 2), gates)
    want = O.expect_z_all(ref, 1e-10)
    want0 = O.expect_z_all(ref, -1.0)

    def body(r, st):
        z = st.apply_circuit_expect_z(gates, 1e-10)
        z2 = st.expect_z_all(1e-10)
        z0 = st.expect_z_all(-1.0)
        c = st.counters()
        return {"err": float(np.max(np.abs(z - want))), "err_sep": float(np.max(np.abs(z - z2))), "err0": float(np.max(np.abs(z0 - want0))),
                "folded": int(c["folded_readouts"]), "remaps": int(c["bytes_link"] > 0)}

    res = pkg.dist.run_local_world(n, world, body, m=2, options={L.OPT_JIT: 1}, timeout=120.0)
    for r, out in enumerate(res):
        assert out["err"] < RTOL and out["err_sep"] < 1e-14 and out["err0"] < RTOL, (r, out)
    assert len({out["folded"] for out in res}) == 1, res             # all ranks took the same path


def test_precompiled_circuit_folds_at_first_sight(pkg):
    """qm_precompile_circuit also compiles the fold flavour of the last pass: on a register large enough for the default
    QM_OPT_JIT (2^26 amplitudes) the FIRST qm_apply_circuit_expect_z already runs compiled and folded"""
    n = 26
    gates = pkg.circuits.c4_step(n, 1)
    st = pkg.B200State(n, 0)
    k = st.precompile_circuit(gates)
    assert k >= 2, k                                  # at least one pass kernel and the fold flavour of the last one
    st.counters_reset()
    z = st.apply_circuit_expect_z(gates)
    c = st.counters()
    assert c["folded_readouts"] == 1 and c["compiled_passes"] == c["passes"] > 0, c
    assert float(np.max(np.abs(z - st.expect_z_all()))) < 1e-14
    z0 = st.apply_circuit_expect_z(pkg.circuits.c4_step(n, 7), -1.0)                  # same structure (period 6), new angles, no threshold
    assert st.counters()["folded_readouts"] == 2
    assert abs(float(z0[0].sum()) - 1.0) < 1e-12 and abs(st.norm2() - 1.0) < 1e-12
    st.close()
