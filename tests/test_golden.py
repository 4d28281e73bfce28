"""Committed golden vectors replayed through (CPU) the C oracle and the host emulator of the encoded pass programs and (GPU)
the CUDA path behind the reference-named Python mirror.  Two files, same cases and call sequences:
  * tests/golden/ref_exec_golden.json — made by EXECUTING the reference's own src/Quantum/QSim.jl and QInfo.jl with the
    Julia-subset interpreter oracle/jl_subset.py (tests/golden/make_ref_exec_golden.py; see tests/test_ref_exec.py), and
  * tests/golden/qsim_golden.json — made from oracle/literal.py, the hand transcription of the same algorithm.
1e-12 relative for amplitudes and readouts; the permutation-only case (x!, cnot!, swap!) bit for bit."""
import json
import os

import numpy as np

import emu  # noqa: E402  (tests/emu.py: the host emulator library)
import pytest

from oracle import literal as LIT
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "qsim_golden.json")) as f:
    GOLDEN = json.load(f)
CASES = GOLDEN["cases"]
with open(os.path.join(HERE, "golden", "ref_exec_golden.json")) as f:
    REF_EXEC_CASES = json.load(f)["cases"]
ALL_CASES = [("refexec", c) for c in REF_EXEC_CASES] + [("literal", c) for c in CASES]
ALL_IDS = ["%s-%s" % (src, c["name"]) for src, c in ALL_CASES]
RTOL = 1e-12


def unhex(v):
    return np.array([float.fromhex(x) for x in v], dtype=np.float64)


def unhexc(v):
    a = unhex(v).reshape(-1, 2)
    return a[:, 0] + 1j * a[:, 1]


def replay(B, s, ops):
    for op in ops:
        name = op[0]
        if isinstance(op[-1], str):
            getattr(B, name)(s, *op[1:-1], float.fromhex(op[-1]))
        else:
            getattr(B, name)(s, *op[1:])
    return s


def check_case(case, amps, mp, mz, mx, my, reduced, entropy, all_qubits=None):
    n = case["n"]
    # the all-qubit readouts (one call for every qubit: no reference counterpart of their own) against the per-qubit goldens
    for basis, fn in (all_qubits or {}).items():
        want = np.array([unhex(v) for v in case["measure_" + basis]])
        got = np.asarray(fn())
        assert got.shape == (n, 2) and np.max(np.abs(got - want)) <= RTOL, (case["name"], "all-qubit " + basis)
    want = unhexc(case["amps"])
    if case["name"].endswith("permutations_only"):
        assert np.array_equal(amps, want)
    assert np.max(np.abs(amps - want)) <= RTOL * np.max(np.abs(want))
    assert np.max(np.abs(mp - unhex(case["mp"]))) <= RTOL
    for t in range(1, n + 1):
        for got, key in ((mz, "measure_z"), (mx, "measure_x"), (my, "measure_y")):
            assert np.max(np.abs(np.array(got(t)) - unhex(case[key][t - 1]))) <= RTOL, (case["name"], key, t)
    for r in case["reduced"]:
        k = r["k"]
        rB, rA = reduced(True, k), reduced(False, k)
        assert np.max(np.abs(rB.ravel() - unhexc(r["rho_B_low"]))) <= RTOL
        assert np.max(np.abs(rA.ravel() - unhexc(r["rho_A_high"]))) <= RTOL
        assert entropy(rB) == pytest.approx(float.fromhex(r["S_B"]), abs=1e-9)
        assert entropy(rA) == pytest.approx(float.fromhex(r["S_A"]), abs=1e-9)


def test_golden_file_matches_its_generator():
    """the committed file is what tests/golden/make_golden.py writes today (the literal restatement has not drifted)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    assert len(CASES) == len(mg.CASES) >= 9
    for case in CASES[:3]:
        s = LIT.statevector(case["n"], case["m"])
        for op in case["ops"]:
            mg.apply_op(s, op)
        assert mg.hxc(s) == case["amps"]


@pytest.mark.parametrize("src_case", ALL_CASES, ids=ALL_IDS)
def test_c_oracle_reproduces_golden_vectors(src_case):
    case = src_case[1]
    n = case["n"]
    s = replay(O, O.statevector(n, case["m"]), case["ops"])
    check_case(case, s, O.mp(s),
               lambda t: O.measure_z(s, t, 1e-10, True), lambda t: O.measure_x(s, t, 1e-10, True),
               lambda t: O.measure_y(s, t, 1e-10, True),
               lambda keep_low, k: O.reduced_density(np.ascontiguousarray(s), keep_low, k), LIT.von_neumann_entropy,
               all_qubits={"z": lambda: O.expect_z_all(s, 1e-10)})


def test_encoded_tma_programs_reproduce_the_12_qubit_golden_case(pkg, emu_mode):
    """planner + encoder of the fused TMA kernel against the literal algorithm, without a GPU: the 12-qubit golden
    case as a wire-format gate list, walked by the host emulator of the encoded pass programs (csrc/emulate.hpp)"""
    cases12 = [c for c in REF_EXEC_CASES if c["n"] == 12]
    assert len(cases12) == 2                              # the mixed case and three time steps of config C1
    for case in cases12:
        gl = pkg.circuits.GateList(12, "reference")
        replay_gl = lambda op: getattr(gl, op[0])(*op[1:-1], float.fromhex(op[-1])) if isinstance(op[-1], str) else getattr(gl, op[0])(*op[1:])
        for op in case["ops"]:
            replay_gl(op)
        got, npass = emu.plan_emulate(12, gl.array(), O.statevector(12, case["m"]), max_cost=75.0)
        want = unhexc(case["amps"])
        assert npass >= 1 and np.max(np.abs(got - want)) <= RTOL * np.max(np.abs(want)), case["name"]


@pytest.mark.gpu
@pytest.mark.parametrize("src_case", ALL_CASES, ids=ALL_IDS)
def test_cuda_path_reproduces_golden_vectors(pkg, src_case):
    """the reference-named mirror over the C ABI (gates queue and fuse; readouts flush)"""
    case = src_case[1]
    Q, QI = pkg.qsim, pkg.qinfo
    n = case["n"]
    s = replay(Q, Q.statevector(n, case["m"]), case["ops"])

    def reduced(keep_low, k):
        if keep_low:
            return QI.partial_trace(s, 1 << (n - k), 1 << k, 1)
        return QI.partial_trace(s, 1 << k, 1 << (n - k), 2)
    check_case(case, s.to_numpy(), Q.mp(s), lambda t: Q.measure_z(s, t), lambda t: Q.measure_x(s, t),
               lambda t: Q.measure_y(s, t), reduced, QI.von_neumann_entropy,
               all_qubits={"z": lambda: s.expect_z_all(), "x": lambda: s.expect_all("x"), "y": lambda: s.expect_all("y")})
