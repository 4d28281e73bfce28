"""How the oracle is pinned to the reference ITSELF although no Julia runtime exists here (CPU only).

oracle/jl_subset.py is a generic interpreter for the subset of Julia the reference's src/Quantum/*.jl and test/*.jl are written
in.  With it
  1. the reference's own test files run against the reference's own source (test/test_QSim.jl: 126 assertions, test/test_QInfo.jl:
     67) — every assertion must pass, which is what qualifies the interpreter;
  2. tests/golden/ref_exec_golden.json holds outputs of the reference's source (u2!, controlled!, swap!, measure_z/x/y, mp,
     partial_trace, von_neumann_entropy ... as its authors wrote them) on seeded call sequences; the small cases are regenerated
     here and must come out the same;
  3. the hand transcription (oracle/literal.py -> qsim_golden.json) must agree with the executed source entry by entry (5e-15),
     and tests/test_golden.py holds the C oracle, the emulator of the pass programs and the CUDA path to the executed-source file.
Items 1 and 2 need /root/reference and are skipped where it does not exist (the GPU box); item 3 and the interpreter's own
syntax tests run everywhere."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from oracle import jl_subset as J  # noqa: E402

REF = os.environ.get("QM_REFERENCE_ROOT", "/root/reference")
HAVE_REF = os.path.exists(os.path.join(REF, "src", "Quantum", "QSim.jl"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="the reference sources are not on this machine")


def flat(x):
    out = []
    if isinstance(x, list):
        for y in x:
            out += flat(y)
    elif isinstance(x, str):
        out.append(float.fromhex(x))
    return out


# ---- the interpreter on its own: Julia semantics the reference relies on --------------------------------------------------
def run(src):
    it = J.Interp()
    it.run(src)
    return it.globals.vars


def test_matrix_literals_follow_julias_whitespace_rule():
    g = run("""
    a = 2
    A = [0 -a; a 0]
    B = [0 - a  a; a-1 0]
    C = [1 2 3]
    D = [a;;]
    v = [1.0, 2.0, 3.0]
    E = ComplexF64[0.5 0.0; 0.0 0.5]
    """)
    assert g["A"].tolist() == [[0, -2], [2, 0]]            # "0 -a" is two elements ...
    assert g["B"].tolist() == [[-2, 2], [1, 0]]            # ... "0 - a" and "a-1" are one each
    assert g["A"].shape == (2, 2) and g["C"].shape == (1, 3) and g["D"].shape == (1, 1) and g["v"].shape == (3,)
    assert g["E"].dtype == np.complex128 and g["E"][0, 0] == 0.5


def test_one_based_indexing_ranges_end_and_column_major_order():
    g = run("""
    v = zeros(ComplexF64, 8)
    v[6] = ComplexF64(0, 1)
    w = v[2:end]
    n = length(w)
    M = reshape([1, 2, 3, 4, 5, 6], 2, 3)
    x = M[2, 3]
    y = M[4]
    s = 0
    for i in 0:(length(v)-1)
        s += (i & 4) == 0 ? 0 : 1
    end
    t = 0
    for i in 1:2, j in 1:3
        t += M[i, j] * (i == j ? 1 : 0)
    end
    """)
    assert g["v"][5] == 1j and g["n"] == 7 and g["w"][4] == 1j
    assert g["M"].tolist() == [[1, 3, 5], [2, 4, 6]] and g["x"] == 6 and g["y"] == 4
    assert g["s"] == 4 and g["t"] == 1 + 4


def test_dispatch_on_annotated_methods_and_scoping():
    g = run("""
    const CACHE = Dict{Float64, Matrix{ComplexF64}}()
    f(s::Vector{ComplexF64}) = "vector"
    f(rho::Matrix{ComplexF64}) = "matrix"
    f(x::Real) = "real"
    f(x::Int) = "int"
    function gate(theta::Real)
        t = Float64(theta)
        if haskey(CACHE, t)
            return CACHE[t]
        end
        m = ComplexF64(cos(t/2)) * Matrix{ComplexF64}([1 0; 0 1])
        CACHE[t] = m
        return m
    end
    function count_up(n)
        acc = 0.0
        for i in 1:n
            acc += i
        end
        acc
    end
    a = f(zeros(ComplexF64, 2)); b = f(zeros(ComplexF64, 2, 2)); c = f(1.5); d = f(3)
    g1 = gate(1); g2 = gate(1.0)
    same = g1 === g2
    k = length(CACHE)
    u = count_up(4)
    q = 3 <= 4 && 2 <= 4 || error("no")
    """)
    assert (g["a"], g["b"], g["c"], g["d"]) == ("vector", "matrix", "real", "int")
    assert g["k"] == 1 and g["u"] == 10.0 and g["q"] is True
    it = J.Interp()
    it.run("h(x::Vector{ComplexF64}) = 1")
    with pytest.raises(J.JuliaError, match="MethodError"):
        it.run("h([1.0, 2.0])")                       # a Vector{Float64} is not a Vector{ComplexF64}
    with pytest.raises(J.JuliaError, match="out of range"):
        it.run('1 <= 0 || error("Qubit index out of range")')


def test_adjoint_outer_product_broadcast_and_linear_algebra():
    g = run("""
    s = [ComplexF64(1), ComplexF64(0, 1)] / √2
    rho = s * s'
    p = abs2.(s)
    d = real.(diag(rho))
    K = kron(Diagonal(ones(ComplexF64, 2)), rho)
    tmp = similar(s)
    mul!(tmp, rho, s)
    vals, vecs = eigen(Hermitian(rho))
    big = sum(vals .> 1e-10)
    L2 = vecs * Diagonal(log2.(max.(vals, 1e-10))) * vecs'
    tr1 = real(tr(rho))
    """)
    assert np.allclose(g["rho"], [[0.5, -0.5j], [0.5j, 0.5]]) and np.allclose(g["p"], [0.5, 0.5]) and np.allclose(g["d"], [0.5, 0.5])
    assert g["K"].shape == (4, 4) and np.allclose(g["tmp"], g["s"]) and g["big"] == 1 and abs(g["tr1"] - 1.0) < 1e-15
    assert np.allclose(g["vals"], [0.0, 1.0])


def test_structs_refs_keywords_and_broadcast_assignment():
    """what the shim and LSM.jl need on top of QSim.jl: (mutable) structs with outer constructors, Ref boxes, ccall through a hook,
    optional positional against keyword parameters, kw... forwarding, `.=` / `.*=` into an existing array with a Vector spreading
    along columns, I as UniformScaling"""
    it = J.Interp()
    log = []
    it.ccall_handler = lambda name, args: (log.append((name, [a.value if isinstance(a, J.JRef) else a for a in args])),
                                            [setattr(a, "value", 7) for a in args if isinstance(a, J.JRef)], 0)[2]
    it.run("""
    const LIB = "libsomething.so"
    mutable struct Box
        h::Ptr{Cvoid}
        n::Int
        mode::Symbol
    end
    function Box(n::Int, m::Int = 0; mode::Symbol = :reference, scale = 1)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        ccall((:make, LIB), Int32, (Int32, Int32, Ptr{Ptr{Cvoid}}), n * scale, m, r)
        Box(r[], n, mode)
    end
    wrap(n; kw...) = Box(n, 1; kw...)
    pick(c::Int, t::Int) = (b = max(c, t); c < t ? (b - 1, b) : (b, b - 1))
    a = Box(3)
    b = wrap(4; mode = :corrected, scale = 2)
    b.n = 9
    same = a.mode === :reference && b.mode === :corrected
    p = pick(1, 4); q = pick(4, 1)
    M = zeros(2, 3)
    v = [1.0, 2.0]
    M .= v
    M .*= 2
    M[:, 2] .= [5.0, 6.0]
    A = [1.0 2.0; 3.0 4.0] + 0.5 * I
    early(x) = (x > 0 && return "positive"; "other")
    e1 = early(1); e2 = early(-1)
    """, it.globals)
    g = it.globals.vars
    assert log == [("make", [3, 0, 0]), ("make", [8, 1, 0])]
    assert g["a"].fields == {"h": 7, "n": 3, "mode": "reference"} and g["b"].fields["n"] == 9 and g["same"] is True
    assert g["p"] == (3, 4) and g["q"] == (4, 3)
    assert g["M"].tolist() == [[2.0, 5.0, 2.0], [4.0, 6.0, 4.0]] and g["A"].tolist() == [[1.5, 2.0], [3.0, 4.5]]
    assert (g["e1"], g["e2"]) == ("positive", "other")
    with pytest.raises(J.JuliaError, match="MethodError"):
        it.run('Box("x")')


# ---- the reference executed ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref():
    import make_ref_exec_golden as MR
    return MR, MR.load_reference()


@needs_ref
def test_the_references_own_tests_pass_on_its_own_source_under_the_interpreter(ref):
    """test/test_QSim.jl and test/test_QInfo.jl, unmodified, against src/Quantum/QSim.jl and QInfo.jl, unmodified"""
    MR, it = ref
    counts = {}
    for name in ("test_QSim.jl", "test_QInfo.jl"):
        p0, f0 = it.tests.passed, len(it.tests.failed)
        it.load_file(os.path.join(REF, "test", name))
        counts[name] = (it.tests.passed - p0, it.tests.failed[f0:])
    assert counts["test_QSim.jl"][1] == [] and counts["test_QInfo.jl"][1] == [], counts
    assert counts["test_QSim.jl"][0] >= 126 and counts["test_QInfo.jl"][0] >= 67, counts


@needs_ref
def test_executed_source_golden_file_is_reproducible(ref):
    MR, it = ref
    with open(os.path.join(HERE, "golden", "ref_exec_golden.json")) as f:
        cases = {c["name"]: c for c in json.load(f)["cases"]}
    for name, n, m, nops, seed, kinds in MR.MG.CASES:
        if n > 8:
            continue                                  # the 12-qubit case builds 4096 x 4096 operators: generator only
        ops = MR.build_ops(name, n, nops, seed, kinds)
        assert ops == cases[name]["ops"]
        s = MR.ref_call(it, "statevector", n, m)
        for op in ops:
            MR.apply_op(it, s, op)
        # bit for bit on this machine; 1e-14 is allowed so that another BLAS build (another summation order inside its
        # matrix products) does not fail the suite
        assert np.max(np.abs(np.array(flat(MR.MG.hxc(s))) - np.array(flat(cases[name]["amps"])))) < 1e-14, name
        my = [MR.MG.hx(MR.ref_call(it, "measure_y", s, t)) for t in range(1, n + 1)]
        assert np.max(np.abs(np.array(flat(my)) - np.array(flat(cases[name]["measure_y"])))) < 1e-14, name


@needs_ref
def test_reference_quirks_in_executable_form(ref):
    """what SURVEY.md calls Q1 and Q2, observed by running the reference: controlled! between non-adjacent qubits acts on the pair
    (b-1, b), b = max(c, t); measure_z drops probabilities that are not > 1e-10"""
    MR, it = ref
    R = lambda name, *a: MR.ref_call(it, name, *a)
    s = R("statevector", 3, 0)
    R("x!", s, 1)                       # |001>: qubit 1 set
    R("cnot!", s, 1, 3)                 # textbook: qubit 3 flips -> |101>; the reference acts on qubits (2, 3): control 2 is clear
    assert abs(s[1]) == 1.0 and np.count_nonzero(s) == 1
    t = R("statevector", 3, 0)
    R("x!", t, 2)
    R("cnot!", t, 1, 3)                 # ... and with qubit 2 set it does flip qubit 3
    assert abs(t[0b110]) == 1.0
    u = R("statevector", 2, 0)
    R("ry!", u, 1, 2e-5)                # P(qubit 1 = 1) = sin^2(1e-5) = 1e-10: exactly at the cut, dropped (strict >)
    p0, p1 = R("measure_z", u, 1)
    assert p1 == 0.0 and 0.0 < 1.0 - p0 < 2e-10


# ---- the two golden files against each other ----------------------------------------------------------------------------
def test_hand_transcription_agrees_with_the_executed_reference_source():
    with open(os.path.join(HERE, "golden", "qsim_golden.json")) as f:
        lit = json.load(f)["cases"]
    with open(os.path.join(HERE, "golden", "ref_exec_golden.json")) as f:
        rex = json.load(f)
    assert "jl_subset" in rex["generator"] and len(lit) == len(rex["cases"]) >= 10
    worst = 0.0
    for a, b in zip(lit, rex["cases"]):
        assert a["name"] == b["name"] and a["ops"] == b["ops"] and a["n"] == b["n"] and a["m"] == b["m"]
        for key in ("amps", "mp", "measure_z", "measure_x", "measure_y"):
            x, y = np.array(flat(a[key])), np.array(flat(b[key]))
            assert x.shape == y.shape
            worst = max(worst, float(np.max(np.abs(x - y))))
        assert len(a["reduced"]) == len(b["reduced"])
        for ra, rb in zip(a["reduced"], b["reduced"]):
            for key in ("rho_B_low", "rho_A_high"):
                worst = max(worst, float(np.max(np.abs(np.array(flat(ra[key])) - np.array(flat(rb[key]))))))
            for key in ("S_B", "S_A"):
                worst = max(worst, abs(float.fromhex(ra[key]) - float.fromhex(rb[key])))
        if a["name"].endswith("permutations_only"):
            assert a["amps"] == b["amps"]
    assert worst < 5e-15, worst


# ---- the other transcriptions, against the executed source ---------------------------------------------------------------
@needs_ref
def test_density_matrix_transcription_agrees_with_the_executed_reference(ref):
    """oracle/literal.py's Matrix{ComplexF64} methods (what tests/test_gpu_density.py holds qm_dm_* to) against the reference's own
    u2!(rho, ..), controlled!(rho, ..), swap!(rho, ..), mp(rho), measure_z/x/y(rho, ..) (QSim.jl:67-79, :209-232, :277-290, :295, :332-365)"""
    from oracle import literal as L
    MR, it = ref
    R = lambda name, *a: MR.ref_call(it, name, *a)
    rng = np.random.default_rng(2024)
    for n in (1, 2, 3, 4):
        rho_ref = R("density_matrix", n, int(rng.integers(0, 1 << n)))
        rho_lit = rho_ref.copy()
        for step in range(25):
            kind = int(rng.integers(0, 6)) if n > 1 else int(rng.integers(0, 3))
            t = int(rng.integers(1, n + 1))
            th = float(rng.uniform(-6.3, 6.3))
            if kind == 0:
                R("h!", rho_ref, t); rho_lit = L.u2_dm(rho_lit, t, L.H)
            elif kind == 1:
                R("rx!", rho_ref, t, th); rho_lit = L.u2_dm(rho_lit, t, L.rx_gate(th))
            elif kind == 2:
                R("rz!", rho_ref, t, th); rho_lit = L.u2_dm(rho_lit, t, L.rz_gate(th))
            else:
                c = int(rng.integers(1, n)); c = c + 1 if c >= t else c
                if kind == 3:
                    R("cnot!", rho_ref, c, t); rho_lit = L.controlled_dm(rho_lit, c, t, L.X)
                elif kind == 4:
                    R("cry!", rho_ref, c, t, th); rho_lit = L.controlled_dm(rho_lit, c, t, L.ry_gate(th))
                else:
                    R("swap!", rho_ref, c, t); rho_lit = L.swap_dm(rho_lit, c, t)
            assert np.max(np.abs(rho_ref - rho_lit)) < 1e-14, (n, step, kind)
        assert np.max(np.abs(R("mp", rho_ref) - L.mp_dm(rho_lit))) < 1e-14
        for t in range(1, n + 1):
            for name, f in (("measure_z", L.measure_z_dm), ("measure_x", L.measure_x_dm), ("measure_y", L.measure_y_dm)):
                assert np.max(np.abs(np.array(R(name, rho_ref, t)) - np.array(f(rho_lit, t)))) < 1e-14, (n, t, name)


@needs_ref
def test_host_qinfo_mirrors_agree_with_the_executed_reference(pkg, ref):
    """quromorphic.jl_b200/qinfo.py (the host mirrors, and the yardstick of the batched device measures) against the reference's own
    QInfo.jl functions on random density matrices: full-rank, and pure-state reductions (rank-deficient)"""
    QI = pkg.qinfo
    MR, it = ref
    R = lambda name, *a: MR.ref_call(it, name, *a)
    rng = np.random.default_rng(77)

    def random_rho(d, rank=None):
        a = rng.normal(size=(d, rank or d)) + 1j * rng.normal(size=(d, rank or d))
        m = a @ a.conj().T
        return np.ascontiguousarray(m / np.trace(m).real)

    for d in (2, 4, 8):
        rho, sigma = random_rho(d), random_rho(d)
        for name in ("von_neumann_entropy", "min_entropy", "max_entropy", "l1_norm_coherence"):
            assert abs(getattr(QI, name)(rho) - R(name, rho)) < 1e-10, (d, name)
        for alpha in (0.5, 2.0, 3.0):
            assert abs(QI.renyi_entropy(rho, alpha) - R("renyi_entropy", rho, alpha)) < 1e-10, (d, alpha)
        assert abs(QI.fidelity(rho, sigma) - R("fidelity", rho, sigma)) < 1e-9
        assert abs(QI.trace_distance(rho, sigma) - R("trace_distance", rho, sigma)) < 1e-10
        assert abs(QI.relative_entropy(rho, sigma) - R("relative_entropy", rho, sigma)) < 1e-9
        low = random_rho(d, rank=1)
        assert abs(QI.von_neumann_entropy(low) - R("von_neumann_entropy", low)) < 1e-8
        assert abs(QI.max_entropy(low) - R("max_entropy", low)) < 1e-12
    for dA, dB in ((2, 2), (2, 4), (4, 2)):
        rho = random_rho(dA * dB)
        for name in ("quantum_mutual_information", "conditional_quantum_entropy", "coherent_information"):
            assert abs(getattr(QI, name)(rho, dA, dB) - R(name, rho, dA, dB)) < 1e-9, (dA, dB, name)
        for sub in (1, 2):
            assert np.max(np.abs(QI.partial_trace_matrix(rho, dA, dB, sub) - R("partial_trace", rho, dA, dB, sub))) < 1e-14


@needs_ref
def test_lsm_driver_and_ridge_readout_agree_with_the_executed_reference(pkg):
    """src/Neuromorphic/LSM.jl executed (struct, step!, collect_states, train_readout!, predict with INJECTED weights — its own
    constructor draws them from Julia's MersenneTwister stream, which nothing here can reproduce) against the host mirrors
    reservoir.LSMNet / ridge_readout that drive config C5 (LSM.jl:7-19, :39-72)"""
    it = J.Interp()
    it.load_file(os.path.join(REF, "src", "Neuromorphic", "LSM.jl"))
    it.using("LSM")
    g = it.globals.vars
    rng = np.random.default_rng(5)
    N, D, O, T = 24, 3, 2, 50
    Win = rng.uniform(-1, 1, (N, D))
    Wres = rng.uniform(-1, 1, (N, N)) * (rng.random((N, N)) < 0.2)
    Wres *= 0.9 / max(abs(np.linalg.eigvals(Wres)))
    inputs, targets = rng.uniform(0, 1, (D, T)), rng.normal(size=(O, T))
    lsm = it.call(g["LSMNet"], [D, N, O, 0.2, 0.9, 0.3, Win.copy(), Wres.copy(), np.zeros((O, N)), np.zeros(O), np.zeros(N)])
    states = it.call(g["collect_states"], [lsm, inputs])
    assert abs(it.call(g["get_spectral_radius"], [lsm]) - 0.9) < 1e-12
    R = pkg.reservoir
    mirror = R.LSMNet(Win, Wres, leak_rate=0.3, output_dim=O)
    st2 = mirror.collect_states(inputs)
    assert states.shape == st2.shape == (N, T) and np.max(np.abs(states - st2)) < 1e-14
    it.call(g["train_readout!"], [lsm, inputs, targets], {"reg": 1e-6})
    pred = it.call(g["predict"], [lsm, inputs])
    Wout, bout = R.ridge_readout(st2, targets, 1e-6)
    scale = np.max(np.abs(Wout))
    assert np.max(np.abs(Wout - lsm.fields["Wout"])) < 1e-8 * scale and np.max(np.abs(bout - lsm.fields["bout"])) < 1e-8 * scale
    assert np.max(np.abs(Wout @ st2 + bout[:, None] - pred)) < 1e-9 * max(1.0, np.max(np.abs(pred)))
    it.call(g["reset!"], [lsm])
    assert not lsm.fields["state"].any()


@needs_ref
def test_c_oracle_agrees_with_the_executed_reference_on_fresh_random_sequences(ref):
    """beyond the committed fixtures: 40 new seeded call sequences (2-7 qubits, every gate kind, any ordered pair) through the
    reference's own functions and through the C oracle — amplitudes, mp, measure_z/x/y of every qubit at 1e-12"""
    from oracle import oracle as O
    MR, it = ref
    R = lambda name, *a: MR.ref_call(it, name, *a)
    worst = 0.0
    for seed in range(40):
        n = 2 + seed % 6
        m = (seed * 7) % (1 << n)
        ops = MR.MG.make_ops(n, 30 + seed, 9000 + seed, None)
        s_ref = R("statevector", n, m)
        s_orc = O.statevector(n, m)
        for op in ops:
            MR.apply_op(it, s_ref, op)
            name = op[0]
            if isinstance(op[-1], str): getattr(O, name)(s_orc, *op[1:-1], float.fromhex(op[-1]))
            else: getattr(O, name)(s_orc, *op[1:])
        worst = max(worst, float(np.max(np.abs(s_ref - s_orc))), float(np.max(np.abs(R("mp", s_ref) - O.mp(s_orc)))))
        for t in range(1, n + 1):
            for name in ("measure_z", "measure_x", "measure_y"):
                worst = max(worst, float(np.max(np.abs(np.array(R(name, s_ref, t)) - np.array(getattr(O, name)(s_orc, t))))))
    assert worst < 1e-12, worst
