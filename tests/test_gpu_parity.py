"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libqm_b200.so), against the
CPU oracle on the same seeded inputs.  Tolerance for ComplexF64 amplitudes and readouts: 1e-12
relative to the largest magnitude (BASELINE.json north_star); permutation gates and the seeded
sampler are compared bit for bit."""
import types

import numpy as np
import pytest

import ref_known_answers as KA
from oracle import literal as LIT
from oracle import oracle as O
from test_abi_cpu import random_circuit

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def engine_backend(pkg, opts=None):
    opts = opts or {}
    Q, QI = pkg.qsim, pkg.qinfo
    B = types.SimpleNamespace()
    B.Error = pkg.QSimError

    def sv(n, m):
        s = Q.statevector(n, m)
        for k, v in opts.items():
            s.set_option(k, v)
        return s
    B.statevector = sv
    B.nb, B.mp = Q.nb, Q.mp
    B.to_numpy = lambda s: s.to_numpy()
    for name in ("h", "x", "y", "z", "rx", "ry", "rz", "cnot", "crx", "cry", "crz", "swap",
                 "measure_z", "measure_x", "measure_y"):
        setattr(B, name, getattr(Q, name))

    def reduced(psi, keep_low, k):
        s = pkg.B200State.from_numpy(np.asarray(psi, dtype=np.complex128))
        n = s.n
        dB = (1 << k) if keep_low else (1 << (n - k))
        return QI.partial_trace(s, (1 << n) // dB, dB, 1 if keep_low else 2)
    B.reduced = reduced
    B.entropy = QI.von_neumann_entropy
    return B


@pytest.mark.parametrize("check", KA.ALL_QSIM, ids=lambda f: f.__name__)
@pytest.mark.parametrize("eager", [0, 1])
def test_reference_known_answers_through_the_engine(pkg, check, eager):
    check(engine_backend(pkg, {pkg._lib.OPT_EAGER: eager}))


def test_reference_partial_trace_known_answers_through_the_engine(pkg):
    KA.check_partial_trace(engine_backend(pkg))


def test_prstate_format(pkg, capsys):
    s = pkg.qsim.statevector(2, 0)
    pkg.qsim.h(s, 1)
    pkg.qsim.prstate(s)                                   # test_QSim.jl:319-321 (no throw) + QSim.jl:371 format
    out = capsys.readouterr().out.strip().splitlines()
    assert out[0].startswith("|00⟩: ") and out[1].startswith("|01⟩: ")


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 9, 12, 13, 14, 17, 20])
@pytest.mark.parametrize("mode", ["fused", "nofuse", "eager"])
def test_random_circuit_matches_oracle(pkg, n, mode):
    gates = random_circuit(pkg, n, 120 if n > 14 else 250, 77 + n)
    psi0 = O.random_state(n, 11 + n)
    st = pkg.B200State.from_numpy(psi0)
    if mode == "eager":
        st.set_option(pkg._lib.OPT_EAGER, 1)
    if mode == "nofuse":
        st.set_option(pkg._lib.OPT_FUSE, 0)
    st.apply_circuit(gates)
    got = st.to_numpy()
    ref = O.apply_circuit(psi0.copy(), gates)
    assert rel_err(got, ref) < RTOL
    c = st.counters()
    assert c["launches"] > 0 and c["gates"] == len(gates)


@pytest.mark.parametrize("pipe", [0, 1])
@pytest.mark.parametrize("tile,low", [(4, 0), (6, 2), (8, 3), (10, 5), (11, 3), (11, 4), (11, 5), (12, 3), (12, 5),
                                      (12, 6), (12, 8), (13, 5)])
def test_tile_shapes(pkg, tile, low, pipe):
    n = 15
    gates = random_circuit(pkg, n, 300, 5 * tile + low)
    psi0 = O.random_state(n, tile)
    st = pkg.B200State.from_numpy(psi0)
    st.set_option(pkg._lib.OPT_PIPELINE, pipe)
    st.set_option(pkg._lib.OPT_TILE_BITS, tile)
    st.set_option(pkg._lib.OPT_LOW_BITS, low)
    st.apply_circuit(gates)
    assert rel_err(st.to_numpy(), O.apply_circuit(psi0.copy(), gates)) < RTOL


def test_every_control_target_pair_reference_semantics(pkg):
    """every ordered (c, t) on 6 qubits through the QSim-shaped API: equals the literal dense-kron
    restatement, i.e. the reference's behaviour for non-adjacent pairs included (QSim.jl:181-197)"""
    n, Q = 6, pkg.qsim
    V = LIT.ry_gate(0.77) @ LIT.rz_gate(-0.4)
    for c in range(1, n + 1):
        for t in range(1, n + 1):
            if c == t:
                continue
            psi0 = O.random_state(n, 100 * c + t)
            s = pkg.B200State.from_numpy(psi0)
            Q.controlled(s, c, t, V)
            assert rel_err(s.to_numpy(), LIT.controlled(psi0.copy(), c, t, V)) < RTOL, (c, t)
            s2 = pkg.B200State.from_numpy(psi0, mode="corrected")
            Q.controlled(s2, c, t, V)
            assert rel_err(s2.to_numpy(), O.controlled(psi0.copy(), c, t, V, quirk=False)) < RTOL


@pytest.mark.parametrize("eager", [0, 1])
def test_permutation_gates_are_bit_exact(pkg, eager):
    n, Q = 11, pkg.qsim
    psi0 = O.random_state(n, 31337)
    s = pkg.B200State.from_numpy(psi0)
    s.set_option(pkg._lib.OPT_EAGER, eager)
    ref = psi0.copy()
    for (c, t) in ((1, 2), (5, 4), (11, 10), (3, 9), (10, 1)):
        Q.cnot(s, c, t); O.cnot(ref, c, t)
    Q.x(s, 7); O.x(ref, 7)
    Q.swap(s, 2, 9); O.swap(ref, 2, 9)
    Q.swap(s, 4, 4); O.swap(ref, 4, 4)
    assert np.array_equal(s.to_numpy(), ref)


@pytest.mark.parametrize("n", [1, 3, 8, 13, 14, 18, 22])
def test_readouts_match_oracle(pkg, n):
    Q = pkg.qsim
    psi = O.random_state(n, 900 + n)
    if n >= 3:
        psi[5] = 1e-6            # probability 1e-12: dropped by the 1e-10 threshold (QSim.jl:308)
    s = pkg.B200State.from_numpy(psi)
    z = s.expect_z_all()
    zr = O.expect_z_all(psi, 1e-10)
    assert np.max(np.abs(z - zr)) < RTOL
    z0 = s.expect_z_all(threshold=-1.0)
    assert np.max(np.abs(z0 - O.expect_z_all(psi, -1.0))) < RTOL
    assert abs(s.norm2() - O.norm2(psi)) < RTOL
    for t in sorted({1, (n + 1) // 2, n}):
        assert Q.measure_z(s, t) == pytest.approx(O.measure_z(psi, t), rel=RTOL, abs=1e-15)
        assert Q.measure_x(s, t) == pytest.approx(O.measure_x(psi, t), rel=RTOL, abs=1e-15)
        assert Q.measure_y(s, t) == pytest.approx(O.measure_y(psi, t), rel=RTOL, abs=1e-15)
    if n <= 18:
        assert np.array_equal(Q.mp(s), O.mp(psi))          # abs2 is computed with the same three operations


@pytest.mark.parametrize("n,k", [(2, 1), (5, 2), (8, 4), (12, 4), (12, 6), (16, 4), (16, 8), (20, 4)])
def test_reduced_density_matches_oracle(pkg, n, k):
    psi = O.random_state(n, 17 * n + k)
    s = pkg.B200State.from_numpy(psi)
    for keep_low in (True, False):
        got = s.reduced_density(keep_low, k)
        ref = O.reduced_density(psi, keep_low, k)
        assert np.max(np.abs(got - ref)) < RTOL
        assert abs(np.trace(got).real - 1.0) < 1e-12
        assert pkg.qinfo.von_neumann_entropy(got) == pytest.approx(LIT.von_neumann_entropy(ref), abs=1e-9)


@pytest.mark.parametrize("n", [1, 5, 12, 13, 19, 24])
def test_sampler_is_bit_exact(pkg, n):
    psi = O.random_state(n, 555 + n)
    s = pkg.B200State.from_numpy(psi)
    for seed in (0, 12345):
        assert np.array_equal(s.sample(seed, 257), O.sample(psi, seed, 257))


def test_sampler_skewed_distribution(pkg):
    n = 16
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[40000] = np.sqrt(0.75); psi[3] = 0.5j
    s = pkg.B200State.from_numpy(psi)
    got = s.sample(7, 1000)
    assert np.array_equal(got, O.sample(psi, 7, 1000))
    assert set(got.tolist()) == {3, 40000}


def test_batched_registers_match_per_state_oracle(pkg):
    n, B, T = 10, 37, 3
    C = pkg.circuits
    gates = C.c2_step_gates(n, seed=1603)
    w = [C.SplitMix64(5).uniform() for _ in range(n)]
    st = pkg.B200State(n, 0, batch=B)
    refs = [O.statevector(n, 0) for _ in range(B)]
    for t in range(T):
        enc = C.c2_encoding(B, n, t, w)
        st.apply_batched(gates, enc)
        z = st.expect_z_all()
        for b in range(B):
            g = gates.copy()
            for i in range(len(g)):
                if g[i]["kind"] == pkg._lib.GATE_U2_SLOT:
                    g[i]["m"] = enc[b, g[i]["q1"]]; g[i]["kind"] = pkg._lib.GATE_U2; g[i]["q1"] = 0
            O.apply_circuit(refs[b], g)
            assert np.max(np.abs(z[b] - O.expect_z_all(refs[b], 1e-10))) < RTOL
    for b in (0, 17, B - 1):
        assert rel_err(st.to_numpy(b), refs[b]) < RTOL


def test_c1_reservoir_prefix_matches_literal_reference_algorithm(pkg):
    """config C1 at a size the literal dense-kron restatement can run (n = 8, T = 5), including the
    deliberately non-adjacent crz!(n, 1, gamma)."""
    n, T, Q = 8, 5, pkg.qsim
    u, w, layer, gamma = pkg.circuits.c1_inputs(n, T)
    s = Q.statevector(n, 0)
    ref = LIT.statevector(n, 0)
    for t in range(T):
        for q in range(1, n + 1):
            Q.ry(s, q, np.pi * u[t] * w[q - 1]); LIT.u2(ref, q, LIT.ry_gate(np.pi * u[t] * w[q - 1]))
        for q in range(1, n + 1):
            Q.rx(s, q, layer[q - 1][0]); LIT.u2(ref, q, LIT.rx_gate(layer[q - 1][0]))
            Q.rz(s, q, layer[q - 1][1]); LIT.u2(ref, q, LIT.rz_gate(layer[q - 1][1]))
        for q in range(1, n):
            Q.cnot(s, q, q + 1); LIT.cnot(ref, q, q + 1)
        Q.crz(s, n, 1, gamma); LIT.controlled(ref, n, 1, LIT.rz_gate(gamma))
        for q in range(1, n + 1):
            assert Q.measure_z(s, q) == pytest.approx(LIT.measure_z(ref, q), rel=RTOL, abs=1e-15)
    assert rel_err(s.to_numpy(), ref) < RTOL


@pytest.mark.parametrize("tile,cost", [(12, 0), (11, 0), (12, 48), (11, 24)])
def test_c3_layers_at_24_qubits_match_oracle(pkg, tile, cost):
    n = 24
    gates = pkg.circuits.c3_circuit(n=n, depth=6)
    st = pkg.B200State(n, 0)
    st.set_option(pkg._lib.OPT_TILE_BITS, tile)
    st.set_option(pkg._lib.OPT_MAX_COST, cost)
    st.apply_circuit(gates)
    ref = O.apply_circuit(O.statevector(n, 0), gates)
    assert rel_err(st.to_numpy(), ref) < RTOL
    assert np.max(np.abs(st.expect_z_all() - O.expect_z_all(ref, 1e-10))) < RTOL


def test_large_register_properties_28_qubits(pkg):
    """size-independent properties at a size the oracle would need minutes for: unitarity (norm),
    circuit followed by its inverse returns the start state, X^2 = identity bit-exactly."""
    n, Q = 28, pkg.qsim
    gates = pkg.circuits.c3_circuit(n=n, depth=4, seed=99)
    st = pkg.B200State(n, 5)
    st.apply_circuit(gates)
    assert abs(st.norm2() - 1.0) < 1e-12
    inv = gates[::-1].copy()
    for i in range(len(inv)):
        U = inv[i]["m"].view(np.complex128).reshape(2, 2).T          # numpy [row, col]
        inv[i]["m"] = pkg._lib.mat8(U.conj().T)
    st.apply_circuit(inv)
    z = st.expect_z_all(threshold=-1.0)
    want = np.array([[0.0, 1.0] if (5 >> q) & 1 else [1.0, 0.0] for q in range(n)])
    assert np.max(np.abs(z - want)) < 1e-12
    samp = st.sample(1, 16)
    assert np.all(samp == 5)


def test_argument_errors(pkg):
    L = pkg._lib
    st = pkg.B200State(3, 0)
    m = L.mat8(np.eye(2))
    assert L.lib().qm_apply_u2(st._h, 3, L.dptr(m)) == L.QM_ERR_ARG
    assert "out of range" in L.last_error()
    assert L.lib().qm_apply_controlled(st._h, 1, 1, L.dptr(m)) == L.QM_ERR_ARG
    assert "same qubit" in L.last_error()
    with pytest.raises(pkg.QMError):
        st.reduced_density(True, 3)
    with pytest.raises(pkg.QSimError):
        pkg.B200State(3, 8)
    with pytest.raises(AssertionError):
        pkg.qinfo.partial_trace(st, 4, 4, 1)            # QInfo.jl:382 "Dimension mismatch"


@pytest.mark.timeout(700, method="thread")
@pytest.mark.parametrize("n", [14, 23])
def test_sharded_two_ranks(pkg, n):
    """2 ranks against the oracle.  On a box with >= 2 GPUs: one PROCESS per GPU (torchrun-style environment, CUDA IPC peer
    mapping, flags over NVLink, the overlapped exchange at 23 qubits).  On a 1-GPU box the same two ranks run as a
    same-process world on device 0 (test_sharded_same_process_world has the wider matrix) — never skipped."""
    import os, socket, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        import gpu_dist_worker as W
        R = W.reference(pkg, n)
        res = pkg.dist.run_local_world(n, 2, lambda r, st: W.check_sharded(pkg, st, r, 2, R), m=R["m"], timeout=120.0)
        assert all(out["ok"] for out in res), res
        return
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    here = os.path.dirname(os.path.abspath(__file__))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(here, "gpu_dist_worker.py"), str(n)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert all("GPU_DIST_OK" in o for o in outs), "\n".join(outs)


@pytest.mark.timeout(150, method="thread")
@pytest.mark.parametrize("n,world,overlap", [(14, 2, 1), (23, 2, 1), (23, 2, 0), (21, 4, 1), (16, 4, 1)])
def test_sharded_same_process_world(pkg, n, world, overlap):
    """The sharded path on ONE GPU: every rank is a host thread of this process with its own shard allocation, streams
    and peer window on device 0 (qm_create_dist_local).  The rank kernels are the product's own — fused passes with rank
    predicates, the in-place exchange kernel over "peer" pointers, the flag fences (stream waits: no kernel ever waits
    for another rank's kernel) and the mailbox collectives.  At 23 / 21 qubits the exchange runs slab by slab beside the
    passes (QM_OPT_OVERLAP), which is asserted.  Everything is compared with the oracle: amplitudes, <Z> with and without
    the threshold, X/Y/Z marginals incl. global qubits, both reduced densities, the sampler bit for bit, and a reset
    after an odd number of layout changes."""
    import gpu_dist_worker as W
    R = W.reference(pkg, n)
    L = pkg._lib
    res = pkg.dist.run_local_world(n, world, lambda r, st: W.check_sharded(pkg, st, r, world, R), m=R["m"],
                                   options={L.OPT_OVERLAP: overlap}, timeout=120.0)
    for r, out in enumerate(res):
        assert out["ok"], (r, out)
    g = world.bit_length() - 1
    if overlap and n >= 23:
        assert all(out["overlapped_remaps"] > 0 for out in res), res
    if not overlap:
        assert all(out["overlapped_remaps"] == 0 for out in res), res


@pytest.mark.timeout(150, method="thread")
@pytest.mark.parametrize("xun,slab_bits,depth", [(4, 1, 1), (8, 3, 3), (8, 2, 4), (4, 3, 2)])
def test_sharded_exchange_kernel_variants(pkg, xun, slab_bits, depth):
    """every variant of the overlapped exchange (pairs in flight per thread 4 / 8, 2 / 4 / 8 slabs, 1-4 passes in front of the exchange) on a same-process world, against the oracle"""
    import gpu_dist_worker as W
    n, world = 23, 2
    R = W.reference(pkg, n, seed_base=4300)
    L = pkg._lib
    res = pkg.dist.run_local_world(n, world, lambda r, st: W.check_sharded(pkg, st, r, world, R), m=R["m"],
                                   options={L.OPT_XUN: xun, L.OPT_SLAB_BITS: slab_bits, L.OPT_OVERLAP_DEPTH: depth, L.OPT_XSM: 24},
                                   timeout=120.0)
    for r, out in enumerate(res):
        assert out["ok"], (r, out)
    assert all(out["overlapped_remaps"] > 0 for out in res), res


def test_c2_batched_full_size_4096_registers_of_16_qubits(pkg):
    """config C2 at BASELINE.json's size: 4096 independent 16-qubit registers (4 GiB) evolved in lockstep with
    per-state encodings.  A sample of registers is checked against the per-state oracle, all of them through
    the norm (unitarity)."""
    n, B, T = 16, 4096, 2
    C = pkg.circuits
    gates = C.c2_step_gates(n, seed=1603)
    assert len(gates) == 63
    w = [C.SplitMix64(7).uniform() for _ in range(n)]
    st = pkg.B200State(n, 0, batch=B)
    sample = [0, 1, 777, 2048, 4095]
    refs = {b: O.statevector(n, 0) for b in sample}
    for t in range(T):
        enc = C.c2_encoding(B, n, t, w)
        st.apply_batched(gates, enc)
        z = st.expect_z_all()
        assert z.shape == (B, n, 2)
        assert np.max(np.abs(z.sum(axis=2) - 1.0)) < 1e-5          # kept mass: the 1e-10 threshold (QSim.jl:308) drops ~4e-7 at 16 qubits
        for b in sample:
            g = gates.copy()
            for i in range(len(g)):
                if g[i]["kind"] == pkg._lib.GATE_U2_SLOT:
                    g[i]["m"] = enc[b, g[i]["q1"]]; g[i]["kind"] = pkg._lib.GATE_U2; g[i]["q1"] = 0
            O.apply_circuit(refs[b], g)
            assert np.max(np.abs(z[b] - O.expect_z_all(refs[b], 1e-10))) < RTOL
    for b in sample:
        assert rel_err(st.to_numpy(b), refs[b]) < RTOL
    c = st.counters()
    assert c["passes"] >= 2 * T


def test_batched_general_per_state_unitaries(pkg):
    """per-state slots holding arbitrary unitaries (phase * rotation * phase path) and a non-unitary one (v1 path)"""
    n, B = 12, 9
    rng = np.random.default_rng(5)
    gl = pkg.circuits.GateList(n)
    gl.slot(3, 0); gl.cnot(3, 4); gl.slot(12, 1); gl.h(1); gl.slot(1, 0)
    gates = gl.array()
    for unitary in (True, False):
        mats = np.empty((B, 2, 8))
        Us = []
        for b in range(B):
            row = []
            for s in range(2):
                M = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
                if unitary:
                    M, _ = np.linalg.qr(M)
                row.append(M); mats[b, s] = pkg._lib.mat8(M)
            Us.append(row)
        st = pkg.B200State(n, 5, batch=B)
        st.apply_batched(gates, mats)
        for b in (0, 4, B - 1):
            ref = O.statevector(n, 5)
            O.u2(ref, 3, Us[b][0]); O.cnot(ref, 3, 4); O.u2(ref, 12, Us[b][1]); O.h(ref, 1); O.u2(ref, 1, Us[b][0])
            assert rel_err(st.to_numpy(b), ref) < 1e-11


def test_c5_lsm_driven_reservoir_with_qinfo_readouts(pkg):
    """config C5 at a size the oracle follows step by step (12 qubits, T = 6): LSM states drive the encoding,
    readouts are all-qubit <Z> plus the von Neumann entropies of the low-4 / high-4 reduced density matrices."""
    R = pkg.reservoir
    n, T, k = 12, 6, 4
    rng = np.random.default_rng(2001)
    lsm = R.LSMNet(rng.uniform(-1, 1, size=(50, 3)), 0.05 * rng.normal(size=(50, 50)), leak_rate=0.3)
    drive = lsm.collect_states(rng.uniform(0, 1, size=(3, T)))[:n]          # LSM.jl:49-57
    qr = R.QuantumReservoir(n, entropy_k=k)
    feats = qr.collect_states(drive)
    assert feats.shape == (n + 2, T)
    ref = O.statevector(n, 0)
    for t in range(T):
        O.apply_circuit(ref, qr.step_gates(drive[:, t]))
        z = O.expect_z_all(ref, 1e-10)
        assert np.max(np.abs(feats[:n, t] - (z[:, 0] - z[:, 1]))) < RTOL
        SB = LIT.von_neumann_entropy(O.reduced_density(ref, True, k))
        SA = LIT.von_neumann_entropy(O.reduced_density(ref, False, k))
        assert feats[n, t] == pytest.approx(SB, abs=1e-9) and feats[n + 1, t] == pytest.approx(SA, abs=1e-9)
    # ridge readout on the quantum features (LSM.jl:59-72 analogue) reproduces a linear target of them
    target = (0.3 * feats[0] - 0.2 * feats[5])[None, :]
    Wout, bout = R.ridge_readout(feats, target, reg=1e-10)
    assert np.max(np.abs(Wout @ feats + bout[:, None] - target)) < 1e-6


def test_c5_at_20_qubits_properties(pkg):
    n, k = 20, 4
    qr = pkg.reservoir.QuantumReservoir(n, entropy_k=k)
    rng = np.random.default_rng(3)
    for t in range(3):
        f = qr.step(rng.uniform(0, 1, size=n))
        assert np.all(np.abs(f[:n]) <= 1 + 1e-12) and 0 <= f[n] <= k + 1e-9 and 0 <= f[n + 1] <= k + 1e-9
    rho = pkg.qinfo.partial_trace(qr.state, 1 << (n - k), 1 << k, 1)
    assert abs(np.trace(rho).real - 1) < 1e-12 and np.max(np.abs(rho - rho.conj().T)) < 1e-15


def test_plan_cache_reuses_structure_not_values(pkg):
    """run_circuit caches plans by circuit STRUCTURE: the same layer with new angles must reuse the plan and still be
    exact, and an angle that changes a gate's class (ry(0) = identity, rx(pi) = -iX, rz only: diagonal) must not be
    served a stale plan.  13 qubits (TMA kernel), eight steps on one register against the oracle."""
    n, C = 13, pkg.circuits
    st = pkg.B200State(n, 0)
    ref = O.statevector(n, 0)
    g = C.SplitMix64(4242)
    specials = {3: 0.0, 5: np.pi, 6: 2 * np.pi}
    for step in range(8):
        gl = C.GateList(n, "reference")
        for q in range(1, n + 1):
            gl.ry(q, specials.get(step, np.pi * g.uniform()) if q % 4 == 1 else np.pi * g.uniform())
        for q in range(1, n + 1):
            gl.rx(q, specials.get(step, g.angle()) if q % 5 == 2 else g.angle())
            gl.rz(q, g.angle())
        for q in range(1, n):
            gl.cnot(q, q + 1)
        gl.crz(n, 1, g.angle())
        gates = gl.array()
        st.apply_circuit(gates)
        O.apply_circuit(ref, gates)
        assert rel_err(st.to_numpy(), ref) < RTOL, step
        assert np.max(np.abs(st.expect_z_all() - O.expect_z_all(ref, 1e-10))) < RTOL
    st.close()


def test_deep_circuit_default_budget_at_20_qubits(pkg):
    """the C3 headline path at a size the oracle follows: a circuit longer than the planner's look-ahead (1,200 gates),
    default budget, single-block HBM-bound passes on the TMA kernel"""
    n = 20
    gates = pkg.circuits.c3_circuit(n=n, depth=40, seed=3001)
    assert len(gates) > 768
    st = pkg.B200State(n, 0)
    st.apply_circuit(gates)
    c = st.counters()
    assert len(gates) / c["passes"] < 20                 # budget-limited passes (interpreter kernel at this size), not deep fusion
    ref = O.apply_circuit(O.statevector(n, 0), gates)
    assert rel_err(st.to_numpy(), ref) < RTOL
    assert np.max(np.abs(st.expect_z_all() - O.expect_z_all(ref, 1e-10))) < RTOL
    st.close()


@pytest.mark.parametrize("n,seed", [(12, 1), (14, 2), (17, 3)])
def test_permutation_circuits_bit_exact_on_the_tma_kernel(pkg, n, seed):
    """x!/cnot!/swap! through the fused TMA kernel (>= 12 qubits), where tails of such gates on register-block bits
    are folded into the store offsets instead of being executed: pure moves, so bit for bit; then a mixed circuit
    (rotations between the permutations) within 1e-12."""
    rng = np.random.default_rng(7700 + seed)
    psi0 = O.random_state(n, 100 + seed)
    st = pkg.B200State.from_numpy(psi0) if hasattr(pkg.B200State, "from_numpy") else None
    if st is None:
        st = pkg.B200State(n, 0); st.upload(psi0)
    ref = psi0.copy()
    for chunk in range(5):
        gl = pkg.circuits.GateList(n, "reference")
        for _ in range(int(rng.integers(10, 70))):
            k = int(rng.integers(0, 6))
            t, c = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
            if k <= 1: gl.x(t)
            elif k <= 4 and c != t: gl.cnot(c, t)
            elif c != t: gl.swap(c, t)
        gates = gl.array()
        st.apply_circuit(gates)
        ref = O.apply_circuit(ref, gates)
        assert np.array_equal(st.to_numpy(), ref), chunk
    gl = pkg.circuits.GateList(n, "reference")
    for q in range(1, n + 1):
        gl.ry(q, float(rng.uniform(-3, 3)))
    for q in range(1, n):
        gl.cnot(q, q + 1)
    gl.x(1); gl.cnot(n, 1)
    gates = gl.array()
    st.apply_circuit(gates)
    ref = O.apply_circuit(ref, gates)
    assert rel_err(st.to_numpy(), ref) < RTOL
    st.close()


@pytest.mark.parametrize("n", [1, 2, 5, 11, 12, 13, 16, 21, 22])
def test_all_qubit_x_and_y_marginals_match_oracle(pkg, n):
    """qm_expect_all_with: measure_x / measure_y (QSim.jl:320-330) of EVERY qubit from ceil((n-12)/9)+1 reads of the
    state, against the oracle's copy-rotate-scan per qubit; the threshold applies to the rotated probabilities."""
    psi0 = O.random_state(n, 900 + n)
    st = pkg.B200State.from_numpy(psi0)
    st.apply_circuit(random_circuit(pkg, n, 40, 31 + n))
    ref = O.apply_circuit(psi0.copy(), random_circuit(pkg, n, 40, 31 + n))
    before = st.counters()["launches"]
    for basis, fo in (("x", O.measure_x), ("y", O.measure_y), ("z", O.measure_z)):
        got = st.expect_all(basis)
        want = np.array([fo(ref, t) for t in range(1, n + 1)])
        assert got.shape == (n, 2)
        assert np.max(np.abs(got - want)) < RTOL, basis
        # no threshold at all (the literal sums)
        got0 = st.expect_all(basis, threshold=-1.0)
        want0 = np.array([fo(ref, t, -1.0) for t in range(1, n + 1)])
        assert np.max(np.abs(got0 - want0)) < RTOL, basis
    # X and Y: one kernel pair per window of qubits, not one per qubit
    windows = 1 if n <= 12 else 1 + -(-(n - 12) // 9)
    assert st.counters()["launches"] - before == 2 * (2 * 2 * windows) + 2 * 2
    # the module-level mirror
    assert np.array_equal(pkg.qsim.measure_all(st, "x"), st.expect_all("x"))


def test_all_qubit_x_marginals_apply_the_threshold_to_rotated_probabilities(pkg):
    """amplitudes chosen so that |a|^2 is above 1e-10 but the H-rotated probability of one branch is below it"""
    n = 13
    v = np.zeros(1 << n, dtype=np.complex128)
    v[0] = 1.0
    eps = 3e-6                                   # (eps / sqrt 2)^2 = 4.5e-12 < 1e-10 < eps^2 ... both dropped / kept differently
    v[5] = eps; v[5 | (1 << 12)] = -eps * (1 - 1e-3)
    v[77] = 2e-5; v[77 ^ 1] = 2e-5               # |a|^2 = 4e-10 kept in Z; X-rotated: 8e-10 and exactly 0
    st = pkg.B200State.from_numpy(v)
    for basis, fo in (("x", O.measure_x), ("y", O.measure_y)):
        got = st.expect_all(basis)
        want = np.array([fo(v, t) for t in range(1, n + 1)])
        assert np.max(np.abs(got - want)) < 1e-14, basis
        one = np.array([getattr(pkg.qsim, "measure_" + basis)(st, t) for t in range(1, n + 1)])
        assert np.max(np.abs(got - one)) < 1e-14


def test_all_qubit_x_marginals_batched(pkg):
    B, n = 5, 14
    st = pkg.B200State(n, 0, batch=B)
    refs = []
    for b in range(B):
        psi = O.random_state(n, 60 + b)
        st.upload(psi, b)
        refs.append(psi)
    got = st.expect_all("y")
    assert got.shape == (B, n, 2)
    for b in range(B):
        want = np.array([O.measure_y(refs[b], t) for t in range(1, n + 1)])
        assert np.max(np.abs(got[b] - want)) < RTOL


def test_split_readout_matches_the_synchronous_one(pkg):
    """qm_expect_z_all_begin / _end: the readout of step t is taken from the stream position of _begin even though the gates
    of step t + 1 are enqueued before _end; two may be pending, a third _begin is refused"""
    n, B = 12, 16
    C = pkg.circuits
    gates = C.c2_step_gates(n, seed=1603)
    w = [C.SplitMix64(5).uniform() for _ in range(n)]
    encs = [C.c2_encoding(B, n, t, w) for t in range(4)]
    ref_state = pkg.B200State(n, 0, batch=B)
    want = []
    for t in range(4):
        ref_state.apply_batched(gates, encs[t])
        want.append(ref_state.expect_z_all())
    st = pkg.B200State(n, 0, batch=B)
    got = []
    for t in range(4):
        st.apply_batched(gates, encs[t])
        st.expect_z_all_begin()
        if t > 0:
            got.append(st.expect_z_all_end())
    got.append(st.expect_z_all_end())
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    st.expect_z_all_begin(); st.expect_z_all_begin()
    with pytest.raises(pkg._lib.QMError):
        st.expect_z_all_begin()
    st.expect_z_all_end(); st.expect_z_all_end()
    with pytest.raises(pkg._lib.QMError):
        st.expect_z_all_end()
