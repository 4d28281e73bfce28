"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol include/qm_b200.h
declares, refuses loudly without a GPU (no CPU fallback), and the host-side planner's reordering is
algebraically exact (verified by replaying its execution order on the CPU oracle)."""
import ctypes as C
import os
import re

import numpy as np

import emu  # noqa: E402  (tests/emu.py: the host emulator library)
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "qm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.lib()
    syms = header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(lib, s), s
        assert s in pkg._lib.SIGNATURES, "ctypes signature missing for " + s
    assert lib.qm_abi_version() == 4


def test_gate_struct_layout(pkg):
    assert pkg._lib.GATE_DTYPE.itemsize == 80 and O.GATE_DTYPE == pkg._lib.GATE_DTYPE


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.QMError) as e:
        pkg.B200State(4, 0)
    assert e.value.code == pkg._lib.QM_ERR_NO_DEVICE


def test_matrix_memory_order(pkg):
    U = np.array([[1 + 2j, 3 + 4j], [5 + 6j, 7 + 8j]])
    assert list(pkg._lib.mat8(U)) == [1, 2, 5, 6, 3, 4, 7, 8]     # U11 U21 U12 U22, Julia column-major


def test_host_gate_matrices_match_reference_expressions(pkg):
    Q = pkg.qsim
    from oracle import literal as L
    for th in (0.0, 0.3, np.pi / 2, np.pi, -2.2, 7.0):
        assert np.array_equal(Q.rx_gate(th), L.rx_gate(th))
        assert np.array_equal(Q.ry_gate(th), L.ry_gate(th))
        assert np.array_equal(Q.rz_gate(th), L.rz_gate(th))
    assert np.array_equal(Q.H, L.H) and np.array_equal(Q.Y, L.Y)
    assert Q.rx_gate(0.3) is Q.rx_gate(0.3)                      # memoised like RX_CACHE (QSim.jl:116-118)
    assert Q.id_(0).shape == (1, 1) and Q.id_(2).shape == (4, 4)  # test_QSim.jl:327-340


def test_effective_control_target_matches_oracle(pkg):
    for c in range(1, 8):
        for t in range(1, 8):
            if c != t:
                assert pkg.qsim.effective_control_target(c, t) == O.effective_ct(c, t)


def random_circuit(pkg, n, ngates, seed):
    rng = np.random.default_rng(seed)
    gl = pkg.circuits.GateList(n)
    Q = pkg.qsim
    for _ in range(ngates):
        k = rng.integers(0, 9)
        t = int(rng.integers(1, n + 1))
        c = int(rng.integers(1, n + 1))
        th = float(rng.uniform(-np.pi, np.pi))
        if k == 0: gl.h(t)
        elif k == 1: gl.rx(t, th)
        elif k == 2: gl.ry(t, th)
        elif k == 3: gl.rz(t, th)
        elif k == 4: gl.u2(t, Q.X)
        elif k == 5: gl.u2(t, Q.Z)
        elif c != t:
            if k == 6: gl.cnot(c, t)
            elif k == 7: gl.crz(c, t, th)
            else: gl.controlled(c, t, Q.ry_gate(th))
    return gl.array()


@pytest.mark.parametrize("n,tile,low,cost", [(6, 12, 5, 0), (10, 6, 2, 0), (12, 8, 3, 0), (14, 12, 5, 40.0),
                                             (13, 7, 0, 0), (9, 4, 1, 8.0)])
def test_planner_order_is_exact_on_cpu(pkg, n, tile, low, cost):
    gates = random_circuit(pkg, n, 300, 1000 + n)
    order, pass_of = pkg._lib.plan_gates(n, gates, 0, tile, low, cost)
    assert len(order) <= len(gates) and len(order) > 0
    assert np.all(np.diff(pass_of) >= 0)
    s = O.random_state(n, 5)
    ref = O.apply_circuit(s.copy(), gates)
    got = O.apply_circuit(s.copy(), order)
    assert np.max(np.abs(got - ref)) < 1e-13


def test_planner_c3_fuses_many_gates_per_pass(pkg):
    gates = pkg.circuits.c3_circuit(n=30, depth=20)
    order, pass_of = pkg._lib.plan_gates(30, gates, 0, 12, 5, 0.0)
    npass = pass_of[-1] + 1
    assert len(gates) == 20 * 30 + 10 * 15 + 10 * 14
    assert npass < len(gates) / 20


def test_planner_leaves_global_targets_for_remap(pkg):
    n, g = 10, 2
    gl = pkg.circuits.GateList(n, mode="corrected")
    gl.rx(3, 0.4); gl.rz(10, 0.3); gl.cnot(10, 2); gl.rx(9, 0.2); gl.h(1)
    out = np.zeros(4096, dtype=np.int32); ln = C.c_int64()
    a = gl.array()
    pkg._lib.check(pkg._lib.lib().qm_plan_describe_ex(n, g, 6, 3, 0.0, 0, a.ctypes.data, len(a),
                                                      out.ctypes.data_as(C.POINTER(C.c_int32)), len(out), C.byref(ln)))
    v = out[:ln.value]
    # the diagonal rz on global qubit 10 and the cnot controlled by it run locally; only rx(9) is left
    assert v[-2] == 1 and v[-1] == 1


@pytest.mark.parametrize("n,tile,low,cost,seed", [(13, 11, 5, 0, 1), (14, 12, 5, 24.0, 2), (15, 12, 3, 64.0, 3), (15, 11, 3, 0, 4),
                                                  (14, 12, 8, 32.0, 5), (15, 11, 4, 40.0, 6), (16, 12, 6, 96.0, 7),
                                                  (16, 12, 5, 86.0, 8), (16, 11, 5, 86.0, 9), (15, 12, 5, 120.0, 10),
                                                  (16, 11, 5, 250.0, 11), (14, 11, 3, 60.0, 12), (16, 12, 7, 180.0, 13)])
def test_encoded_pass_programs_emulated_on_cpu(pkg, emu_mode, n, tile, low, cost, seed):
    """planner + encoder of the TMA kernel, checked without a GPU: the encoded pass programs (the bytes the kernel
    is launched with) are walked over a host vector by the scalar emulation in csrc/emulate.hpp (test
    infrastructure) and must reproduce the oracle; a control predicate that is not warp-uniform, an unknown opcode
    or a pass the kernel could not take is an error"""
    gates = random_circuit(pkg, n, 300, 4000 + 17 * seed)
    psi0 = O.random_state(n, seed)
    got, npass = emu.plan_emulate(n, gates, psi0, tile_bits=tile, low_bits=low, max_cost=cost)
    ref = O.apply_circuit(psi0.copy(), gates)
    assert npass > 0
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12


def test_encoded_c3_programs_emulated_on_cpu(pkg, emu_mode):
    n = 16
    gates = pkg.circuits.c3_circuit(n=n, depth=8)
    for tile, cost in ((12, 0.0), (11, 24.0), (12, 48.0), (12, 86.0), (11, 86.0), (12, 130.0), (11, 250.0)):
        got, npass = emu.plan_emulate(n, gates, O.statevector(n, 0), tile_bits=tile, max_cost=cost)
        ref = O.apply_circuit(O.statevector(n, 0), gates)
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12


def test_gate_list_files_round_trip_bit_exactly(pkg, tmp_path):
    """the on-disk circuit forms (binary and hex-float text) give back the identical 80-byte records"""
    C = pkg.circuits
    gates = C.c3_circuit(n=12, depth=3)
    for text in (False, True):
        path = str(tmp_path / ("c.qmg" if not text else "c.qmg.txt"))
        C.save_gates(path, 12, gates, text=text)
        n, back = C.load_gates(path)
        assert n == 12 and back.dtype == gates.dtype and back.tobytes() == gates.tobytes()
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.qmg"
        bad.write_bytes(b"QMGATES1" + np.array([12, 5], dtype="<u4").tobytes() + b"x" * 7)
        C.load_gates(str(bad))


def test_host_qinfo_measures_match_reference_known_answers(pkg):
    """the QInfo measures that stay on the host (small matrices), against the assertions of test/test_QInfo.jl"""
    QI = pkg.qinfo
    ap = lambda v: pytest.approx(v, abs=1e-10)
    bell = np.array([1, 0, 0, 1]) / np.sqrt(2)
    rho_bell = np.outer(bell, bell).astype(np.complex128)
    assert QI.quantum_mutual_information(rho_bell, 2, 2) == ap(2.0)                       # TI:61-64
    assert QI.quantum_mutual_information(np.diag([0.5, 0.5, 0, 0]).astype(complex), 2, 2) == ap(0.0)   # TI:67-71
    assert QI.conditional_quantum_entropy(rho_bell, 2, 2) == ap(-1.0)                     # TI:76-79
    assert QI.conditional_quantum_entropy(np.diag([1.0, 0, 0, 0]).astype(complex), 2, 2) == ap(0.0)    # TI:82-86
    assert QI.coherent_information(rho_bell, 2, 2) == ap(1.0)                             # TI:91-94
    rho, sig = np.diag([0.7, 0.3]).astype(complex), np.diag([0.5, 0.5]).astype(complex)
    D = QI.relative_entropy(rho, sig)                                                     # TI:99-110
    assert D >= 0 and QI.relative_entropy(rho, rho) == ap(0.0)
    assert D == ap(0.7 * np.log2(0.7 / 0.5) + 0.3 * np.log2(0.3 / 0.5))
    r1, r2 = np.diag([1.0, 0]).astype(complex), np.diag([0, 1.0]).astype(complex)
    assert QI.holevo_information([r1, r2], [0.5, 0.5]) == ap(1.0)                         # TI:115-121
    assert QI.holevo_information([r1, r1], [0.5, 0.5]) == ap(0.0)                         # TI:124-125
    assert QI.quantum_relative_renyi_entropy(rho, sig, 2.0) >= 0                          # TI:149-150
    assert QI.quantum_relative_renyi_entropy(rho, rho, 2.0) == ap(0.0)                    # TI:153
    with pytest.raises(AssertionError):
        QI.quantum_relative_renyi_entropy(rho, sig, 1.0)                                  # TI:156
    th = np.pi / 4                                                                        # TI:161-167
    assert QI.quantum_fisher_information([np.cos(th / 2), np.sin(th / 2)], [-np.sin(th / 2) / 2, np.cos(th / 2) / 2]) == ap(1.0)
    assert QI.quantum_fisher_information([1.0, 0.0], [0.0, 0.0]) == ap(0.0)               # TI:170-172
    assert QI.l1_norm_coherence(np.full((2, 2), 0.5, dtype=complex)) == ap(1.0)           # TI:177-178
    assert QI.l1_norm_coherence(rho) == ap(0.0)                                           # TI:181-182
    assert QI.l1_norm_coherence(np.array([[0.6, 0.2], [0.2, 0.4]], dtype=complex)) == ap(0.4)   # TI:185-186
    assert QI.l1_norm_coherence(np.array([[0.5, 0.3j], [-0.3j, 0.5]])) == ap(0.6)         # TI:326-329
    sw = np.diag([0.3, 0.7]).astype(complex)
    assert QI.fidelity(rho, rho) == ap(1.0)                                               # TI:191-192
    assert QI.fidelity(rho, sw) == pytest.approx(0.84, abs=0.01)                          # TI:195-198
    assert QI.fidelity(r1, r2) == ap(0.0) and QI.fidelity(r1, r1) == ap(1.0)              # TI:201-206
    assert QI.trace_distance(rho, rho) == ap(0.0) and QI.trace_distance(rho, sw) == ap(0.4)     # TI:211-218
    assert QI.trace_distance(r1, r2) == ap(1.0)                                           # TI:221-223
    # GHZ on 3 qubits, 2 x 4 split (TI:298-308): the reduced state of the first factor is mixed
    ghz = np.zeros(8); ghz[0] = ghz[7] = 1 / np.sqrt(2)
    rA = QI.partial_trace_matrix(np.outer(ghz, ghz), 2, 4, 2)
    assert rA.shape == (2, 2) and QI.von_neumann_entropy(rA) > 0


def test_default_budget_packs_one_full_group_per_pass_on_c3(pkg):
    """planner quality guard (host only): at the engine's default budget the C3 workload at 30 qubits must come out
    as passes of ONE register-block group holding >= 12 gates on average (a second group never hides
    behind HBM time), every pass must be encodable for the TMA kernel, and tiles must be rebuilt around their targets
    (>= 128 contiguous amplitudes per row)"""
    L = pkg._lib
    gates = pkg.circuits.c3_circuit(n=30, depth=20)
    plan = L.plan_describe(30, gates, max_cost=80.0)
    ng = np.array([p["ngates"] for p in plan]); G = np.array([len(p["groups"]) for p in plan])
    assert ng.sum() >= 0.95 * len(gates) and ng.mean() >= 12.0
    assert G.mean() <= 1.1 and G.max() <= 2
    assert min(p["L"] for p in plan) >= 5 and np.mean([p["L"] for p in plan]) >= 7
    for p in plan:
        for rb, _ in p["groups"]:
            assert len(rb) <= 4
    h = L.plan_ophist(30, gates, max_cost=80.0)          # raises if a pass cannot be encoded
    assert h["passes"] == len(plan)
    deep = L.plan_describe(30, gates, max_cost=0.0)
    assert len(deep) < len(plan) / 3


def test_short_circuit_is_planned_at_the_budget_that_finishes_first(pkg, emu_mode):
    """A reservoir step (C4: 2-3 gates per qubit, then a readout) cannot fill one register block per pass; a circuit
    shorter than the look-ahead window is planned at the multiple of the budget the pass model says finishes first
    (planner.hpp budget_for_pending): the 33-qubit step must come out in <= 5 passes instead of 10-11 and the batched
    16-qubit step (C2) in <= 3 instead of 7, a deep circuit (C3) must keep its single-block passes, and the encoded
    programs of such steps must reproduce the oracle (host emulation, consecutive steps on one register)"""
    L = pkg._lib
    for s in range(3):
        plan = L.plan_describe(33, pkg.circuits.c4_step(33, s), max_cost=75.0)
        assert sum(p["ngates"] for p in plan) >= 75
        assert len(plan) <= 5 and max(len(p["groups"]) for p in plan) >= 2
    assert len(L.plan_describe(16, pkg.circuits.c2_step_gates(16), max_cost=75.0)) <= 3
    n = 15
    psi = O.random_state(n, 77)
    ref = psi.copy()
    for s in range(3):
        gates = pkg.circuits.c4_step(n, s)
        psi, npass = emu.plan_emulate(n, gates, psi, max_cost=75.0)
        ref = O.apply_circuit(ref, gates)
        assert npass <= 2
        assert np.max(np.abs(psi - ref)) / np.max(np.abs(ref)) < 1e-12


@pytest.mark.parametrize("n,depth,cost", [(14, 52, 80.0), (13, 60, 60.0), (15, 40, 110.0), (14, 56, 75.0)])
def test_deep_circuit_programs_emulated_on_cpu(pkg, emu_mode, n, depth, cost):
    """circuits longer than the planner's look-ahead (768 gates) keep the budget as it is — the C3 headline path:
    single-block passes filled to the budget.  Encoded programs walked by the host emulator against the oracle."""
    gates = pkg.circuits.c3_circuit(n=n, depth=depth, seed=3001 + n)
    assert len(gates) > 768
    plan = pkg._lib.plan_describe(n, gates, max_cost=cost)
    assert np.mean([len(p["groups"]) for p in plan]) <= 1.15              # deep mode, decided once for the whole circuit
    psi0 = O.random_state(n, n)
    got, npass = emu.plan_emulate(n, gates, psi0, max_cost=cost)
    ref = O.apply_circuit(psi0.copy(), gates)
    if emu_mode == "records":
        assert npass == len(plan)
    else:                                 # the compiled kernels' cost model (planner.hpp cost_model 1): fewer, fuller passes
        assert npass <= len(plan)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12


@pytest.mark.parametrize("n,seed,perm_only", [(12, 1, True), (13, 2, True), (14, 3, False), (12, 4, False), (15, 5, False),
                                              (13, 6, True), (16, 7, False)])
def test_permutation_tails_folded_into_the_stores_are_exact(pkg, emu_mode, n, seed, perm_only):
    """x!/cnot!/swap!-heavy circuits: groups end in permutation tails that encode_v4 folds into the END record's store
    offsets instead of executing them.  Pure permutation circuits must stay bit for bit (np.array_equal), mixed ones
    within 1e-12; short chunks applied one after the other so that many groups END in such a tail."""
    rng = np.random.default_rng(9000 + seed)
    Q = pkg.qsim
    psi = O.random_state(n, seed)
    ref = psi.copy()
    for chunk in range(6):
        gl = pkg.circuits.GateList(n, "reference" if seed % 2 else "corrected")
        for _ in range(int(rng.integers(8, 60))):
            k = int(rng.integers(0, 10 if not perm_only else 6))
            t, c = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
            if k <= 1: gl.x(t)
            elif k <= 4 and c != t: gl.cnot(c, t)
            elif k == 5 and c != t: gl.swap(c, t)
            elif k == 6: gl.ry(t, float(rng.uniform(-3, 3)))
            elif k == 7: gl.rz(t, float(rng.uniform(-3, 3)))
            elif k == 8 and c != t: gl.crz(c, t, float(rng.uniform(-3, 3)))
            elif k == 9: gl.h(t)
        gates = gl.array()
        if len(gates) == 0:
            continue
        psi, npass = emu.plan_emulate(n, gates, psi, max_cost=75.0)
        ref = O.apply_circuit(ref, gates)
        if perm_only:
            assert np.array_equal(psi, ref), (chunk, npass)
        else:
            assert np.max(np.abs(psi - ref)) / np.max(np.abs(ref)) < 1e-12


def test_compiled_pass_kernels_generate_and_assemble_for_sm_100a(pkg):
    """the compiled pass kernels (QM_OPT_JIT) without a GPU: every pass of a planned circuit becomes PTX text (the JIT
    template embedded in the library + the straight-line register-block groups of csrc/jit_codegen.hpp) and, where
    libnvJitLink is installed, assembles into an sm_100a cubin.  (What the text COMPUTES is checked by the
    `compiled_ptx` mode of the emulator tests above.)"""
    L = pkg._lib
    for n, cost, seed in ((14, 80.0, 1), (16, 0.0, 2), (13, 24.0, 3)):
        gates = random_circuit(pkg, n, 200, 600 + seed)
        d = L.jit_check(n, gates, max_cost=cost, compile=False, want_ptx=True)
        assert d["passes"] == d["with_ptx"] > 0 and d["distinct"] > 0
        ptx = d["ptx"]
        assert "// QMJIT_BEGIN" in ptx and "// QMJIT_END" in ptx and ".target sm_100a" in ptx
        body = ptx[ptx.index("// QMJIT_BEGIN"):ptx.index("// QMJIT_END")]
        assert body.count("ld.shared.v2.f64") % 16 == 0 and body.count("st.shared.v2.f64") == body.count("ld.shared.v2.f64")
        assert "brx.idx" not in body                       # no dispatch: that is the point
        try:
            d2 = L.jit_check(n, gates, max_cost=cost, compile=True)
        except L.QMError as e:
            if "libnvJitLink not found" in str(e):
                pytest.skip("libnvJitLink is not installed here")
            raise
        assert d2["assembled"] == d2["passes"] and 10000 < d2["max_cubin_bytes"] < 400000


def test_fold_flavour_of_the_compiled_kernels_assembles_and_leaves_the_plain_one_alone(pkg):
    """QM_OPT_ZFOLD without a GPU: the ZFOLD instantiation of the pass kernel (all-qubit <Z> sums taken inside the last pass) is a
    second PTX template with the same generated groups spliced in; it assembles for sm_100a, and the plain kernels carry none of it."""
    import re
    L = pkg._lib
    gates = pkg.circuits.c4_step(20, 2)
    plain = L.jit_check(20, gates, max_cost=80.0, compile=False, want_ptx=True)
    fold = L.jit_check(20, gates, max_cost=80.0, compile=False, want_ptx=True, zfold=True)
    assert plain["passes"] == fold["passes"] == fold["with_ptx"] > 0 and plain["structure_hash"] == fold["structure_hash"]
    entry = lambda t: re.search(r"\.entry\s+(\w+)", t).group(1)
    assert entry(plain["ptx"]) != entry(fold["ptx"]) and "k_fused_tma" in entry(fold["ptx"])
    body = lambda t: t[t.index("// QMJIT_BEGIN"):t.index("// QMJIT_END")]
    # the template's operand register names and the entry's own (mangled) name differ between the two
    strip = lambda t: re.sub(r"_ZN2qm11k_fused_tma\w+?_param_", "K_param_", re.sub(r"%\w+", "%", t))
    assert strip(body(plain["ptx"])) == strip(body(fold["ptx"]))          # same generated group code
    for marker in ("shfl.sync.bfly", "bar.sync \t8", "setp.gt.f64"):
        assert marker not in plain["ptx"].replace("bar.sync 8", "bar.sync \t8")
    assert fold["ptx"].count("shfl.sync.bfly") >= 20 and "setp.gt.f64" in fold["ptx"]
    try:
        d = L.jit_check(20, gates, max_cost=80.0, compile=True, zfold=True)
    except L.QMError as e:
        if "libnvJitLink not found" in str(e):
            pytest.skip("libnvJitLink is not installed here")
        raise
    assert d["assembled"] == d["passes"] and 10000 < d["max_cubin_bytes"] < 400000


def test_compiled_kernel_disk_cache(pkg, tmp_path):
    """qm_jit_cache_dir: assembled kernels are kept as files keyed by the complete PTX text; a second walk over the same circuit
    reads them instead of assembling again, a damaged file is ignored and replaced, and without a directory nothing is written."""
    import os
    import time
    L = pkg._lib
    gates = random_circuit(pkg, 13, 60, 4242)
    try:
        L.jit_check(13, gates, max_cost=24.0, compile=True)
    except L.QMError as e:
        if "libnvJitLink not found" in str(e):
            pytest.skip("libnvJitLink is not installed here")
        raise
    d = tmp_path / "jitcache"
    try:
        L.jit_cache_dir(d)
        s0 = L.jit_cache_stats()
        t0 = time.perf_counter()
        a = L.jit_check(13, gates, max_cost=24.0, compile=True)
        t_cold = time.perf_counter() - t0
        s1 = L.jit_cache_stats()
        files = sorted(os.listdir(d))
        # one file per distinct PTX text (coefficient offsets can differ between passes of one opcode sequence: >= distinct)
        assert a["distinct"] <= len(files) <= a["passes"] and all(f.startswith("qmjit-") and f.endswith(".cubin") for f in files)
        assert s1["disk_stores"] - s0["disk_stores"] == len(files)
        assert s1["disk_hits"] - s0["disk_hits"] == a["passes"] - len(files)
        t0 = time.perf_counter()
        b = L.jit_check(13, gates, max_cost=24.0, compile=True)
        t_warm = time.perf_counter() - t0
        s2 = L.jit_cache_stats()
        assert b == a
        assert s2["disk_hits"] - s1["disk_hits"] == a["passes"] and s2["disk_stores"] == s1["disk_stores"]
        assert t_warm < t_cold
        # a truncated file and a file with a flipped byte do not verify: assembled again and written again
        f0, f1 = d / files[0], d / files[-1]
        raw = f0.read_bytes(); f0.write_bytes(raw[:len(raw) // 2])
        if f1 != f0:
            raw1 = bytearray(f1.read_bytes()); raw1[-5] ^= 0x40; f1.write_bytes(bytes(raw1))
        c = L.jit_check(13, gates, max_cost=24.0, compile=True)
        s3 = L.jit_cache_stats()
        assert c == a and s3["disk_stores"] - s2["disk_stores"] == (1 if f1 == f0 else 2)
        assert f0.read_bytes() == raw
        assert not [f for f in os.listdir(d) if ".tmp." in f]
    finally:
        L.jit_cache_dir(None)
    s4 = L.jit_cache_stats()
    L.jit_check(13, gates, max_cost=24.0, compile=True)
    assert L.jit_cache_stats() == s4


def test_compiled_kernels_are_keyed_by_structure_not_by_angles(pkg):
    """canonical signs (encode_v4 canon): the opcode sequence of a pass program — the key of a compiled kernel — must not
    depend on the angles.  C4 reservoir steps have a structure period of 6 (gate kind (q + l) mod 3, brickwork l mod 2) and
    fresh random angles every step, incl. merged general unitaries whose lifting forms take every sign variant."""
    L = pkg._lib
    for n in (16, 21):
        h = [L.jit_check(n, pkg.circuits.c4_step(n, s), max_cost=80.0, compile=False)["structure_hash"] for s in range(18)]
        for s in range(12):
            assert h[s] == h[s + 6], (n, s)
        assert len(set(h[:6])) >= 4                      # ... and it does tell different structures apart
    # same circuit shape, random angles in [-2 pi, 2 pi]: rx / ry / rz / crz / crx with every sign of cos and sin
    rng = np.random.default_rng(11)
    hs = set()
    for rep in range(6):
        gl = pkg.circuits.GateList(14, "corrected")
        for q in range(1, 15):
            gl.rx(q, float(rng.uniform(-6.3, 6.3))); gl.rz(q, float(rng.uniform(-6.3, 6.3)))
        for q in range(1, 14):
            gl.crz(q, q + 1, float(rng.uniform(-6.3, 6.3)))
        for q in range(1, 15):
            gl.ry(q, float(rng.uniform(-6.3, 6.3)))
        for q in range(2, 14, 2):
            gl.crx(q + 1, q, float(rng.uniform(-6.3, 6.3)))
        hs.add(L.jit_check(14, gl.array(), max_cost=80.0, compile=False)["structure_hash"])
    assert len(hs) == 1
