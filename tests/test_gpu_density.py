"""Density-matrix (Matrix{ComplexF64}) methods of QSim.jl on the engine: rho lives on the
GPU as vec(rho), a 2q-qubit register; every method must match the literal restatement of the reference's
`op * rho * op'` (oracle/literal.py, QSim.jl:67-79, :209-232, :277-290, :295, :332-365) and the reference's
own assertions on density matrices (test/test_QSim.jl, cited per check)."""
import io

import numpy as np
import pytest

from oracle import literal as LIT

pytestmark = pytest.mark.gpu
RTOL = 1e-12
APPROX = dict(rel=1.5e-8, abs=1e-10)


def ap(v):
    return pytest.approx(v, **APPROX)


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def is_hermitian(M):
    return np.max(np.abs(M - M.conj().T)) < 1e-12


def test_reference_assertions_on_density_matrices(pkg):
    Q = pkg.qsim
    rho = Q.density_matrix(2, 0).to_numpy()                                   # TQ:44-48
    assert rho.shape == (4, 4) and rho[0, 0] == ap(1.0) and np.trace(rho).real == ap(1.0) and is_hermitian(rho)
    s = Q.statevector(2, 0); Q.h(s, 1)                                         # TQ:50-55
    rho2 = Q.statevector_to_density_matrix(s).to_numpy()
    assert rho2.shape == (4, 4) and np.trace(rho2).real == ap(1.0) and is_hermitian(rho2)
    r3 = Q.density_matrix(3, 0)                                                # TQ:59-64
    assert Q.nb(r3) == 3
    p = Q.mp(r3)
    assert len(p) == 8 and p[0] == ap(1.0) and p.sum() == ap(1.0)
    r = Q.density_matrix(1, 0); Q.h(r, 1); M = r.to_numpy()                    # TQ:126-131
    assert np.trace(M).real == ap(1.0) and is_hermitian(M) and M[0, 0].real == ap(0.5) and M[1, 1].real == ap(0.5)
    r = Q.density_matrix(1, 0); Q.x(r, 1); M = r.to_numpy()                    # TQ:136-140
    assert M[0, 0].real == ap(0.0) and M[1, 1].real == ap(1.0) and np.trace(M).real == ap(1.0)
    r = Q.density_matrix(1, 0); Q.y(r, 1); M = r.to_numpy()                    # TQ:143-146
    assert np.trace(M).real == ap(1.0) and is_hermitian(M)
    r = Q.density_matrix(1, 1); Q.z(r, 1); M = r.to_numpy()                    # TQ:149-152
    assert M[1, 1].real == ap(1.0) and np.trace(M).real == ap(1.0)
    for f, th in ((Q.rx, np.pi / 4), (Q.ry, np.pi / 3), (Q.rz, np.pi / 6)):    # TQ:156-169
        r = Q.density_matrix(1, 0); f(r, 1, th); M = r.to_numpy()
        assert np.trace(M).real == ap(1.0) and is_hermitian(M)
    r = Q.density_matrix(2, 0); Q.cnot(r, 1, 2); M = r.to_numpy()              # TQ:195-199
    assert M[0, 0].real == ap(1.0) and np.trace(M).real == ap(1.0) and is_hermitian(M)
    r = Q.density_matrix(2, 2); Q.cnot(r, 2, 1); M = r.to_numpy()              # TQ:201-204
    assert M[3, 3].real == ap(1.0) and np.trace(M).real == ap(1.0)
    for f in (Q.crx, Q.cry, Q.crz):                                            # TQ:244-259
        r = Q.density_matrix(2, 2); f(r, 1, 2, np.pi / 4); M = r.to_numpy()
        assert np.trace(M).real == ap(1.0) and is_hermitian(M)
    r = Q.density_matrix(1, 0); Q.h(r, 1)                                      # TQ:285-289
    p0, p1 = Q.measure_z(r, 1)
    assert p0 == ap(0.5) and p1 == ap(0.5)
    p0, p1 = Q.measure_x(Q.density_matrix(1, 0), 1)                            # TQ:298-301
    assert p0 + p1 == ap(1.0)
    p0, p1 = Q.measure_y(Q.density_matrix(1, 0), 1)                            # TQ:310-313
    assert p0 + p1 == ap(1.0)
    buf = io.StringIO()                                                        # TQ:323-324
    Q.prstate(Q.density_matrix(2, 0), file=buf)
    assert "|00⟩ : 1.0" in buf.getvalue()
    r = Q.density_matrix(2, 0)                                                 # TQ:364-372
    with pytest.raises(pkg.QSimError):
        Q.h(r, 3)
    with pytest.raises(pkg.QSimError):
        Q.cnot(r, 1, 3)
    with pytest.raises(pkg.QSimError):
        Q.measure_z(r, 3)
    with pytest.raises(pkg.QSimError):
        Q.cnot(r, 1, 1)


@pytest.mark.parametrize("q,seed", [(3, 1), (5, 2), (6, 3)])
def test_random_gate_sequences_match_literal_op_rho_opT(pkg, q, seed):
    """every method against op * rho * op' built with kron (the reference's literal algorithm), including the
    non-adjacent controlled pairs, starting from a mixed state"""
    Q = pkg.qsim
    rng = np.random.default_rng(seed)
    d = 1 << q
    A = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    ref = A @ A.conj().T
    ref /= np.trace(ref).real                                   # a full-rank mixed state
    rho = pkg.B200Density.from_numpy(ref)
    for _ in range(60):
        k = int(rng.integers(0, 8))
        t, c = int(rng.integers(1, q + 1)), int(rng.integers(1, q + 1))
        th = float(rng.uniform(-np.pi, np.pi))
        if k == 0:
            Q.h(rho, t); ref = LIT.u2_dm(ref, t, LIT.H)
        elif k == 1:
            Q.rx(rho, t, th); ref = LIT.u2_dm(ref, t, LIT.rx_gate(th))
        elif k == 2:
            Q.ry(rho, t, th); ref = LIT.u2_dm(ref, t, LIT.ry_gate(th))
        elif k == 3:
            Q.rz(rho, t, th); ref = LIT.u2_dm(ref, t, LIT.rz_gate(th))
        elif k == 4:
            Q.y(rho, t); ref = LIT.u2_dm(ref, t, LIT.Y)
        elif c != t:
            if k == 5:
                Q.cnot(rho, c, t); ref = LIT.controlled_dm(ref, c, t, LIT.X)
            elif k == 6:
                Q.crz(rho, c, t, th); ref = LIT.controlled_dm(ref, c, t, LIT.rz_gate(th))
            else:
                Q.swap(rho, c, t); ref = LIT.swap_dm(ref, c, t)
    assert rel_err(rho.to_numpy(), ref) < RTOL
    assert np.max(np.abs(Q.mp(rho) - LIT.mp_dm(ref))) < RTOL
    for t in range(1, q + 1):
        assert Q.measure_z(rho, t) == pytest.approx(LIT.measure_z_dm(ref, t), rel=RTOL, abs=1e-14)
        assert Q.measure_x(rho, t) == pytest.approx(LIT.measure_x_dm(ref, t), rel=RTOL, abs=1e-14)
        assert Q.measure_y(rho, t) == pytest.approx(LIT.measure_y_dm(ref, t), rel=RTOL, abs=1e-14)


def test_density_of_evolved_statevector_equals_evolved_density(pkg):
    """statevector_to_density_matrix on the device, and U (s s') U' == (U s)(U s)' at 10 qubits (a 20-qubit register)"""
    Q = pkg.qsim
    q = 10
    s = Q.statevector(q, 3)
    rho = Q.density_matrix(q, 3)
    rng = np.random.default_rng(9)
    for _ in range(40):
        t = int(rng.integers(1, q + 1)); th = float(rng.uniform(-3, 3))
        for obj in (s, rho):
            Q.ry(obj, t, th); Q.rz(obj, t, 0.5 * th)
            if t < q:
                Q.cnot(obj, t, t + 1)
    A = Q.statevector_to_density_matrix(s).to_numpy()
    B = rho.to_numpy()
    v = s.to_numpy()
    assert rel_err(A, np.outer(v, v.conj())) < RTOL
    assert rel_err(B, A) < RTOL
    assert abs(np.trace(B).real - 1.0) < 1e-12
