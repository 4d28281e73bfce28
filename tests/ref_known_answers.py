"""The reference's own known-answer assertions for the hot path, restated once and run against
every implementation (literal numpy restatement, C oracle, and — under -m gpu — the B200 engine
through its QSim-shaped host API).

Each check cites the assertion it restates in /root/reference/test/test_QSim.jl (TQ) or
/root/reference/test/test_QInfo.jl (TI).  A backend is any object with the reference's function
names (bang dropped): statevector, nb, mp, h, x, y, z, rx, ry, rz, cnot, crx, cry, crz, swap,
measure_z, measure_x, measure_y, to_numpy(state) and Error (the exception type standing in for
Julia's ErrorException).
"""
import numpy as np
import pytest

S2 = 1 / np.sqrt(2)
APPROX = dict(rel=1.5e-8, abs=1e-10)   # Julia's default isapprox rtol = sqrt(eps); tests use atol 1e-10


def ap(v):
    return pytest.approx(v, **APPROX)


def check_state_creation(B):
    s1 = B.to_numpy(B.statevector(2, 0))                       # TQ:12-15
    assert len(s1) == 4 and s1[0] == ap(1.0) and np.all(np.abs(s1[1:]) < 1e-12)
    s2 = B.to_numpy(B.statevector(3, 5))                       # TQ:17-20
    assert len(s2) == 8 and s2[5] == ap(1.0) and np.sum(np.abs(s2) ** 2) == ap(1.0)
    assert B.nb(B.statevector(3, 0)) == 3                      # TQ:25-26
    assert B.nb(B.statevector(5, 0)) == 5                      # TQ:28-29
    p = np.asarray(B.mp(B.statevector(2, 1)))                  # TQ:34-38
    assert len(p) == 4 and p[1] == ap(1.0) and p.sum() == ap(1.0)


def check_single_qubit_gates(B):
    s = B.statevector(1, 0); B.h(s, 1); v = B.to_numpy(s)      # TQ:71-75
    assert abs(v[0]) == ap(S2) and abs(v[1]) == ap(S2) and np.sum(np.abs(v) ** 2) == ap(1.0)
    s = B.statevector(2, 0); B.h(s, 1); v = B.to_numpy(s)      # TQ:78-81
    assert abs(v[0]) == ap(S2) and abs(v[1]) == ap(S2)
    s = B.statevector(1, 0); B.x(s, 1); v = B.to_numpy(s)      # TQ:86-89
    assert abs(v[0]) < 1e-12 and v[1] == ap(1.0)
    s = B.statevector(1, 0); B.y(s, 1); v = B.to_numpy(s)      # TQ:92-95
    assert abs(v[0]) < 1e-12 and abs(v[1]) == ap(1.0)
    assert v[1] == ap(1j)                                      # Y|0> = i|1>  (QSim.jl:14; phase the reference test does not pin)
    s = B.statevector(1, 1); B.z(s, 1); v = B.to_numpy(s)      # TQ:98-101
    assert abs(v[0]) < 1e-12 and v[1] == ap(-1.0)
    s = B.statevector(1, 0); B.rx(s, 1, np.pi); v = B.to_numpy(s)        # TQ:106-109
    assert abs(v[0]) < 1e-10 and abs(v[1]) == ap(1.0)
    s = B.statevector(1, 0); B.ry(s, 1, np.pi / 2); v = B.to_numpy(s)    # TQ:112-115
    assert abs(v[0]) == ap(S2) and abs(v[1]) == ap(S2)
    s = B.statevector(1, 0); B.rz(s, 1, np.pi / 2); v = B.to_numpy(s)    # TQ:118-120
    assert np.sum(np.abs(v) ** 2) == ap(1.0)
    assert v[0] == ap(np.exp(-1j * np.pi / 4))                 # RZ = diag(c - is, c + is), QSim.jl:143-145


def check_two_qubit_gates(B):
    s = B.statevector(2, 0); B.cnot(s, 2, 1); v = B.to_numpy(s)          # TQ:176-179
    assert v[0] == ap(1.0) and np.all(np.abs(v[1:]) < 1e-12)
    s = B.statevector(2, 2); B.cnot(s, 2, 1); v = B.to_numpy(s)          # TQ:182-185
    assert abs(v[2]) < 1e-12 and v[3] == ap(1.0)
    s = B.statevector(2, 1); B.cnot(s, 1, 2); v = B.to_numpy(s)          # TQ:188-191
    assert abs(v[1]) < 1e-12 and v[3] == ap(1.0)
    s = B.statevector(2, 1); B.swap(s, 1, 2); v = B.to_numpy(s)          # TQ:209-212
    assert abs(v[1]) < 1e-12 and v[2] == ap(1.0)
    s = B.statevector(2, 1); before = B.to_numpy(s).copy(); B.swap(s, 1, 1)   # TQ:221-224
    assert np.array_equal(B.to_numpy(s), before)
    for f, th in ((B.crx, np.pi), (B.cry, np.pi / 2), (B.crz, np.pi / 4)):    # TQ:229-241
        s = B.statevector(2, 2); f(s, 1, 2, th)
        assert np.sum(np.abs(B.to_numpy(s)) ** 2) == ap(1.0)


def check_measurements(B):
    s = B.statevector(2, 0); p0, p1 = B.measure_z(s, 1)        # TQ:267-270
    assert p0 == ap(1.0) and p1 == ap(0.0)
    s = B.statevector(2, 2); p0, p1 = B.measure_z(s, 2)        # TQ:272-275
    assert p0 == ap(0.0) and p1 == ap(1.0)
    s = B.statevector(1, 0); B.h(s, 1); p0, p1 = B.measure_z(s, 1)   # TQ:278-282
    assert p0 == ap(0.5) and p1 == ap(0.5)
    s = B.statevector(1, 0); p0, p1 = B.measure_x(s, 1)        # TQ:293-296
    assert p0 == ap(0.5) and p1 == ap(0.5)
    s = B.statevector(1, 0); p0, p1 = B.measure_y(s, 1)        # TQ:305-308
    assert p0 == ap(0.5) and p1 == ap(0.5)


def check_errors(B):
    s = B.statevector(2, 0)                                    # TQ:347-360
    with pytest.raises(B.Error): B.h(s, 3)
    with pytest.raises(B.Error): B.x(s, 6)
    with pytest.raises(B.Error): B.measure_z(s, 5)
    with pytest.raises(B.Error): B.cnot(s, 1, 1)
    with pytest.raises(B.Error): B.cnot(s, 1, 3)
    with pytest.raises(B.Error): B.swap(s, 1, 7)


def check_cache_consistency(B):
    s1 = B.statevector(1, 0); s2 = B.statevector(1, 0)         # TQ:379-390
    B.rx(s1, 1, np.pi / 4); B.rx(s2, 1, np.pi / 4)
    assert np.array_equal(B.to_numpy(s1), B.to_numpy(s2))
    s3 = B.statevector(1, 0); B.rx(s3, 1, np.pi / 3)
    assert not np.allclose(B.to_numpy(s1), B.to_numpy(s3))


def check_circuits(B):
    s = B.statevector(2, 0); B.h(s, 1); B.cnot(s, 1, 2); v = B.to_numpy(s)       # TQ:397-408
    assert abs(v[0]) == ap(S2) and abs(v[3]) == ap(S2) and np.sum(np.abs(v) ** 2) == ap(1.0)
    p0, p1 = B.measure_z(s, 1)
    assert p0 == ap(0.5) and p1 == ap(0.5)
    s = B.statevector(3, 0); B.h(s, 1); B.cnot(s, 1, 2); B.cnot(s, 1, 3); v = B.to_numpy(s)   # TQ:413-420
    assert abs(v[0]) == ap(S2) and abs(v[7]) == ap(S2) and np.sum(np.abs(v) ** 2) == ap(1.0)
    s = B.statevector(2, 0); B.h(s, 1); B.h(s, 2)              # TQ:426-437
    for _ in range(10):
        B.rx(s, 1, 0.1); B.ry(s, 2, 0.1); B.cnot(s, 1, 2)
    assert abs(np.sum(np.abs(B.to_numpy(s)) ** 2) - 1.0) < 1e-10


def check_partial_trace(B):
    """B.reduced(psi, keep_low, k) must equal partial_trace(psi psi', dA, dB, 1 if keep_low else 2)."""
    bell = np.array([1, 0, 0, 1], dtype=np.complex128) / np.sqrt(2)      # TI:24-35
    mixed = np.array([[0.5, 0], [0, 0.5]], dtype=np.complex128)
    assert np.allclose(B.reduced(bell, False, 1), mixed, atol=1e-10)
    assert np.allclose(B.reduced(bell, True, 1), mixed, atol=1e-10)
    prod = np.array([1, 0, 0, 0], dtype=np.complex128)                   # TI:37-44
    pure = np.array([[1, 0], [0, 0]], dtype=np.complex128)
    assert np.allclose(B.reduced(prod, False, 1), pure, atol=1e-10)
    assert np.allclose(B.reduced(prod, True, 1), pure, atol=1e-10)
    ghz = np.zeros(8, dtype=np.complex128); ghz[0] = ghz[7] = 1 / np.sqrt(2)   # TI:298-308
    rB = B.reduced(ghz, True, 2)            # partial_trace(rho, 2, 4, 1): keep the 4-dim low factor
    assert rB.shape == (4, 4)
    assert B.entropy(rB) == ap(1.0)         # TI:308 only asserts > 0; the analytic value is 1 bit
    assert B.entropy(B.reduced(bell, False, 1)) == ap(1.0)               # TI:49-51
    assert B.entropy(B.reduced(prod, False, 1)) == ap(0.0)               # TI:54-56


ALL_QSIM = [check_state_creation, check_single_qubit_gates, check_two_qubit_gates,
            check_measurements, check_errors, check_cache_consistency, check_circuits]
