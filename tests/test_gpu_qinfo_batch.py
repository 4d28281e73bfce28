"""GPU parity of the batched QInfo measures: one Jacobi warp per matrix on the device against the
host mirrors in quromorphic.jl_b200/qinfo.py (numpy eigh, the reference's formulas QInfo.jl:24-50, :199-204, :289-357).
Tolerance 1e-9 on the measures (they pass through log2 / pow of eigenvalues), 1e-12 on eigenvalues."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def random_density(rng, d, rank=None):
    r = rank or d
    M = rng.normal(size=(d, r)) + 1j * rng.normal(size=(d, r))
    rho = M @ M.conj().T
    return rho / np.trace(rho).real


@pytest.mark.parametrize("d", [2, 4, 8, 16])
def test_measures_of_many_matrices_match_host_qinfo(pkg, d):
    QI = pkg.qinfo
    rng = np.random.default_rng(100 + d)
    count = 257
    full = np.array([random_density(rng, d) for _ in range(count)])
    defic = np.array([random_density(rng, d, rank=max(1, d // 2)) for _ in range(count)])       # rank-deficient: zero eigenvalues
    sigmas = np.array([random_density(rng, d) for _ in range(count)])
    for rhos in (full, defic):
        ev = QI.eigvalsh_batch(rhos)
        assert np.max(np.abs(ev - np.array([np.linalg.eigvalsh(r) for r in rhos]))) < 1e-12

    def check(name, dev, host, tol=1e-9):
        assert np.max(np.abs(dev - np.array(host))) < tol, name
    for tag, rhos in (("full rank", full), ("rank-deficient", defic)):
        check(tag + " von_neumann", QI.von_neumann_entropy_batch(rhos), [QI.von_neumann_entropy(r) for r in rhos])
        check(tag + " renyi 2", QI.renyi_entropy_batch(rhos, 2.0), [QI.renyi_entropy(r, 2.0) for r in rhos])
        check(tag + " renyi 3", QI.renyi_entropy_batch(rhos, 3.0), [QI.renyi_entropy(r, 3.0) for r in rhos])
        check(tag + " min", QI.min_entropy_batch(rhos), [QI.min_entropy(r) for r in rhos])
        check(tag + " trace_distance", QI.trace_distance_batch(rhos, sigmas), [QI.trace_distance(r, s) for r, s in zip(rhos, sigmas)])
    # measures that take a root of the eigenvalues are only conditioned to ~sqrt(eps) = 1e-8 where eigenvalues vanish (a zero
    # eigenvalue comes out as +-1e-17 from any solver, and its square root is 3e-9): 1e-9 on full-rank matrices, 1e-6 otherwise
    for tag, rhos, tol in (("full rank", full, 1e-9), ("rank-deficient", defic, 1e-6)):
        check(tag + " renyi 0.5", QI.renyi_entropy_batch(rhos, 0.5), [QI.renyi_entropy(r, 0.5) for r in rhos], tol)
        check(tag + " fidelity", QI.fidelity_batch(rhos, sigmas), [QI.fidelity(r, s) for r, s in zip(rhos, sigmas)], tol)
    # max_entropy counts eigenvalues above 1e-10: exact on full-rank matrices; on rank-deficient ones the count must be the rank
    assert np.array_equal(QI.max_entropy_batch(full), np.array([QI.max_entropy(r) for r in full]))
    assert np.array_equal(QI.max_entropy_batch(defic), np.full(count, np.log2(max(1, d // 2))))
    # fidelity of a state with itself, trace distance to itself
    assert np.max(np.abs(QI.fidelity_batch(full, full) - 1.0)) < 1e-9
    assert np.max(np.abs(QI.trace_distance_batch(full, full))) < 1e-12


def test_quantum_mutual_information_batch(pkg):
    QI = pkg.qinfo
    rng = np.random.default_rng(7)
    rhos = np.array([random_density(rng, 16, rank=3) for _ in range(40)])
    got = QI.quantum_mutual_information_batch(rhos, 4, 4)
    want = np.array([QI.quantum_mutual_information(r, 4, 4) for r in rhos])
    assert np.max(np.abs(got - want)) < 1e-9


@pytest.mark.parametrize("n,B", [(10, 64), (16, 512)])
def test_entropy_of_every_register_of_a_batch(pkg, n, B):
    """reduced matrix + entropy for a whole batch on the device (entanglement_entropy per register, QInfo.jl:178-182)
    against per-register qm_reduced_density + host eigh"""
    QI = pkg.qinfo
    st = pkg.B200State(n, 0, batch=B)
    C = pkg.circuits
    gates = C.c2_step_gates(n, seed=1603) if n == 16 else C.c3_circuit(n=n, depth=3, seed=5)
    if n == 16:
        w = [C.SplitMix64(7).uniform() for _ in range(n)]
        st.apply_batched(gates, C.c2_encoding(B, n, 0, w))
    else:
        for b in range(B):
            st.upload(O.random_state(n, 300 + b), b)
        st.apply_circuit(gates)
    sample = sorted(set([0, 1, B // 2, B - 1]))
    for keep_low in (True, False):
        for k in (1, 3, 4):
            rb = st.reduced_density_batch(keep_low, k)
            assert rb.shape == (B, 1 << k, 1 << k)
            for b in sample:
                assert np.max(np.abs(rb[b] - st.reduced_density(keep_low, k, b))) < 1e-14
            for kind, host in (("von_neumann", QI.von_neumann_entropy), ("min", QI.min_entropy),
                               ("renyi", lambda r: QI.renyi_entropy(r, 2.0))):
                got = st.entropy_batch(keep_low, k, kind, alpha=2.0)
                want = np.array([host(rb[b]) for b in range(B)])
                assert np.max(np.abs(got - want)) < 1e-9, (keep_low, k, kind)
    dA = 1 << 4
    ee = QI.entanglement_entropy_batch(st, dA, (1 << n) // dA)
    assert abs(ee[sample[-1]] - QI.entanglement_entropy(st, dA, (1 << n) // dA, sample[-1])) < 1e-9
    st.close()
