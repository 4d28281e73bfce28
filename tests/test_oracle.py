"""Pins the CPU oracle (oracle/qsim_oracle.c):
  (1) the reference's own known-answer assertions (tests/ref_known_answers.py), and
  (2) the literal dense-kron numpy transcription of QSim.jl (oracle/literal.py) for every gate
      kind and EVERY ordered (c, t) pair — which is what establishes the reference's
      non-adjacent-control behaviour (QSim.jl:181-197) in an executable form.
CPU only."""
import itertools

import numpy as np
import pytest

import ref_known_answers as KA
from backends import literal_backend, oracle_backend
from oracle import literal as L
from oracle import oracle as O

RTOL = 1e-12


def rel_err(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("check", KA.ALL_QSIM, ids=lambda f: f.__name__)
@pytest.mark.parametrize("backend", ["literal", "oracle_seq", "oracle_tree"])
def test_reference_known_answers(check, backend):
    B = {"literal": literal_backend, "oracle_seq": lambda: oracle_backend(True),
         "oracle_tree": lambda: oracle_backend(False)}[backend]()
    check(B)


@pytest.mark.parametrize("backend", ["literal", "oracle"])
def test_reference_partial_trace_known_answers(backend):
    KA.check_partial_trace(literal_backend() if backend == "literal" else oracle_backend())


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7])
def test_u2_matches_literal_every_target(n):
    rng = np.random.default_rng(100 + n)
    for t in range(1, n + 1):
        for U in (L.H, L.X, L.Y, L.Z, L.rx_gate(0.37), L.ry_gate(-1.2), L.rz_gate(2.9),
                  rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))):
            s = O.random_state(n, 7 * n + t)
            ref = L.u2(s.copy(), t, U)
            got = O.u2(s.copy(), t, U)
            assert rel_err(got, ref) < RTOL


@pytest.mark.parametrize("n", [2, 3, 5, 6])
def test_controlled_matches_literal_every_pair_including_non_adjacent(n):
    rng = np.random.default_rng(n)
    V = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    for c, t in itertools.permutations(range(1, n + 1), 2):
        for M in (L.X, L.rz_gate(0.81), V):
            s = O.random_state(n, 1000 * c + t)
            ref = L.controlled(s.copy(), c, t, M)
            got = O.controlled(s.copy(), c, t, M, quirk=True)
            assert rel_err(got, ref) < RTOL, (c, t)
            # the textbook gate agrees with the reference only for adjacent pairs
            txt = O.controlled(s.copy(), c, t, M, quirk=False)
            same = rel_err(txt, ref) < 1e-9
            assert same == (abs(c - t) == 1), (c, t, same)


def test_controlled_permutation_is_bit_exact():
    n = 6
    for c, t in itertools.permutations(range(1, n + 1), 2):
        s = O.random_state(n, 5 + c * 10 + t)
        ref = L.cnot(s.copy(), c, t)
        got = O.cnot(s.copy(), c, t)
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("n", [2, 4, 6])
def test_swap_matches_literal(n):
    for q1 in range(1, n + 1):
        for q2 in range(1, n + 1):
            s = O.random_state(n, 31 * q1 + q2)
            assert np.array_equal(O.swap(s.copy(), q1, q2), L.swap(s.copy(), q1, q2))


@pytest.mark.parametrize("n", [1, 3, 6, 9])
def test_measurements_match_literal(n):
    s = O.random_state(n, 77 + n)
    s[0] = 1e-6 + 0j          # |a|^2 = 1e-12 < 1e-10: dropped by the threshold (QSim.jl:308)
    for t in range(1, n + 1):
        for fo, fl in ((O.measure_z, L.measure_z), (O.measure_x, L.measure_x), (O.measure_y, L.measure_y)):
            ref = fl(s, t)
            assert fo(s, t, 1e-10, True) == pytest.approx(ref, rel=1e-13, abs=1e-15)
            assert fo(s, t, 1e-10, False) == pytest.approx(ref, rel=1e-12, abs=1e-15)
    ez = O.expect_z_all(s, 1e-10)
    for t in range(1, n + 1):
        assert tuple(ez[t - 1]) == pytest.approx(L.measure_z(s, t), rel=1e-12, abs=1e-15)
    assert np.array_equal(O.mp(s), L.mp(s))


def test_threshold_is_strict_and_drops_small_probabilities():
    s = np.zeros(4, dtype=np.complex128)
    s[0] = np.sqrt(1e-10)       # abs2 rounds to <= 1e-10 or just above: use exact values instead
    s[0] = 1e-5                 # abs2 == 1.0000000000000002e-10 > 1e-10  -> kept
    s[1] = 0.9e-5               # abs2 == 8.1e-11 -> dropped
    s[3] = 1.0
    p0, p1 = O.measure_z(s, 1, 1e-10, True)
    assert p0 == (1e-5 * 1e-5) and p1 == 1.0
    assert O.measure_z(s, 1, -1.0, True) == (1e-5 * 1e-5, 0.9e-5 * 0.9e-5 + 1.0)


@pytest.mark.parametrize("n,k", [(2, 1), (4, 1), (4, 2), (6, 3), (7, 2)])
def test_reduced_density_matches_literal_partial_trace(n, k):
    s = O.random_state(n, 9 * n + k)
    rho = L.statevector_to_density_matrix(s)
    for keep_low in (True, False):
        dB = (1 << k) if keep_low else (1 << (n - k))
        dA = (1 << n) // dB
        ref = L.partial_trace(rho, dA, dB, 1 if keep_low else 2)
        got = O.reduced_density(s, keep_low, k)
        assert np.max(np.abs(got - ref)) < 1e-14


def test_sampler_definition_small():
    """the sampler has no reference counterpart; pin its definition against a numpy restatement"""
    n = 14
    s = O.random_state(n, 4242)
    p = O.mp(s)
    leaves = np.array([np.add.reduce(p[i:i + 4096]) for i in range(0, len(p), 4096)])
    got = O.sample(s, 99, 64)
    assert got.max() < (1 << n)
    # statistical sanity + determinism
    assert np.array_equal(got, O.sample(s, 99, 64))
    assert not np.array_equal(got, O.sample(s, 100, 64))
    # each sampled index lies in a leaf with non-zero mass
    assert np.all(leaves[(got // 4096).astype(int)] > 0)


def test_circuit_runner_equals_call_sequence():
    n = 8
    gates = np.zeros(5, dtype=O.GATE_DTYPE)
    gates[0] = (0, 2, 0, 0, O._mat(L.H))
    gates[1] = (1, 2, 5, 0, O._mat(L.X))                # non-adjacent, textbook on the wire
    gates[2] = (0, 7, 0, 0, O._mat(L.rx_gate(0.3)))
    gates[3] = (2, 1, 6, 0, np.zeros(8))
    gates[4] = (1, 6, 0, 0, O._mat(L.rz_gate(1.1)))
    s = O.random_state(n, 3)
    got = O.apply_circuit(s.copy(), gates)
    ref = s.copy()
    O.h(ref, 3); O.cnot(ref, 3, 6, quirk=False); O.rx(ref, 8, 0.3); O.swap(ref, 2, 7, quirk=False)
    O.crz(ref, 7, 1, 1.1, quirk=False)
    assert np.array_equal(got, ref)


def test_literal_density_matrix_methods_agree_with_vector_methods():
    """the literal Matrix methods (op rho op', QSim.jl:67-79, :209-232) on a pure state equal the outer product of
    the literal vector methods' result — for every gate kind and every ordered (c, t), non-adjacent pairs included"""
    from oracle import literal as LIT
    q = 4
    rng = np.random.default_rng(3)
    v = rng.normal(size=1 << q) + 1j * rng.normal(size=1 << q)
    v /= np.linalg.norm(v)
    for t in range(1, q + 1):
        for U in (LIT.H, LIT.X, LIT.Y, LIT.Z, LIT.rx_gate(0.7), LIT.ry_gate(-1.1), LIT.rz_gate(2.3)):
            w = LIT.u2(v.copy(), t, U)
            assert np.max(np.abs(LIT.u2_dm(np.outer(v, v.conj()), t, U) - np.outer(w, w.conj()))) < 1e-14
    for c in range(1, q + 1):
        for t in range(1, q + 1):
            if c == t:
                continue
            for V in (LIT.X, LIT.rz_gate(0.9), LIT.ry_gate(0.4)):
                w = LIT.controlled(v.copy(), c, t, V)
                assert np.max(np.abs(LIT.controlled_dm(np.outer(v, v.conj()), c, t, V) - np.outer(w, w.conj()))) < 1e-14
            w = LIT.swap(v.copy(), c, t)
            assert np.max(np.abs(LIT.swap_dm(np.outer(v, v.conj()), c, t) - np.outer(w, w.conj()))) < 1e-14
    rho = np.outer(v, v.conj())
    for t in range(1, q + 1):
        assert LIT.measure_z_dm(rho, t) == pytest.approx(LIT.measure_z(v, t), abs=1e-14)
        assert LIT.measure_x_dm(rho, t) == pytest.approx(LIT.measure_x(v, t), abs=1e-14)
        assert LIT.measure_y_dm(rho, t) == pytest.approx(LIT.measure_y(v, t), abs=1e-14)
    assert np.max(np.abs(LIT.mp_dm(rho) - LIT.mp(v))) < 1e-15
