"""GPU parity tests of the COMPILED pass kernels (QM_OPT_JIT, csrc/jit.cu + jit_codegen.hpp): a fused pass whose opcode
sequence repeats is compiled into straight-line code and launched instead of the interpreter.  Same bar as the interpreter:
1e-12 against the CPU oracle, permutation circuits bit for bit; every test asserts that compiled kernels really ran
(counters()["compiled_passes"]) and, where both exist, that the two paths agree to the last bit (same operation order)."""
import numpy as np
import pytest

from oracle import oracle as O
from test_abi_cpu import random_circuit

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("n,cost,seed", [(12, 0, 1), (13, 80, 2), (14, 24, 3), (17, 80, 4), (20, 120, 5), (16, 250, 6), (22, 80, 7)])
def test_compiled_passes_match_oracle_and_interpreter(pkg, n, cost, seed):
    L = pkg._lib
    gates = random_circuit(pkg, n, 300, 9100 + seed)
    psi0 = O.random_state(n, 40 + seed)
    ref = O.apply_circuit(psi0.copy(), gates)
    out = {}
    for jit in (0, 1):
        st = pkg.B200State.from_numpy(psi0)
        st.set_option(L.OPT_JIT, jit)
        st.set_option(L.OPT_MAX_COST, cost)
        st.apply_circuit(gates)
        out[jit] = st.to_numpy()
        c = st.counters()
        assert (c["compiled_passes"] > 0) == (jit == 1), c
        if jit:
            assert c["compiled_passes"] == c["passes"], c
        st.close()
    assert rel_err(out[1], ref) < RTOL and rel_err(out[0], ref) < RTOL
    # the two kernels agree far inside the parity bound (not to the last bit: the planner prices gates differently for the two,
    # and the canonical-sign form replaces the complex-multiply variant of a thread-constant phase by a lift)
    assert rel_err(out[0], out[1]) < 1e-13
    s = L.jit_stats()
    assert s["compiled"] > 0 and s["failed"] == 0, s


def test_default_threshold_compiles_on_the_second_run_of_a_structure(pkg):
    """QM_OPT_JIT default (2): the first circuit of a structure runs through the interpreter, the second one — same
    structure, NEW angles — through compiled kernels (all its passes compiled together from the cached plan).  The default
    only applies to registers of 2^26 amplitudes and more (here 1024 registers of 16 qubits); a small register never compiles."""
    L = pkg._lib
    n = 16
    small = pkg.B200State(n, 0)
    st = pkg.B200State(n, 0, batch=1024)
    ref = O.statevector(n, 0)
    rng = np.random.default_rng(3)
    seen = []
    for rep in range(3):
        gl = pkg.circuits.GateList(n)
        for q in range(1, n + 1):
            gl.ry(q, float(rng.uniform(-6.3, 6.3)))         # merged with the rx and rz into one general unitary per qubit:
            gl.rx(q, float(rng.uniform(-6.3, 6.3)))         # every sign variant of the lifting forms turns up, and the
            gl.rz(q, float(rng.uniform(-6.3, 6.3)))         # canonical-sign encoding keeps the opcode sequence the same
        for q in range(1, n):
            gl.cnot(q, q + 1)
        for q in range(1, n, 2):
            gl.crz(q, q + 1, float(rng.uniform(-6.3, 6.3)))
        g = gl.array()
        st.apply_circuit(g)
        small.apply_circuit(g)
        ref = O.apply_circuit(ref, g)
        seen.append(st.counters()["compiled_passes"])
    assert seen[0] == 0 and seen[1] > 0 and seen[2] > seen[1], seen
    for b in (0, 513, 1023):
        assert rel_err(st.to_numpy(b), ref) < RTOL
    assert rel_err(small.to_numpy(), ref) < RTOL and small.counters()["compiled_passes"] == 0


@pytest.mark.parametrize("n,seed", [(12, 1), (14, 2), (17, 3)])
def test_compiled_permutation_circuits_are_bit_exact(pkg, n, seed):
    """x!/cnot!/swap! inside the register block are register renames in the compiled code, tails are folded into the stores"""
    L = pkg._lib
    rng = np.random.default_rng(700 + seed)
    psi = O.random_state(n, seed)
    st = pkg.B200State.from_numpy(psi)
    st.set_option(L.OPT_JIT, 1)
    ref = psi.copy()
    for chunk in range(5):
        gl = pkg.circuits.GateList(n)
        for _ in range(int(rng.integers(10, 70))):
            k = int(rng.integers(0, 6))
            t, c = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
            if k <= 1: gl.x(t)
            elif k <= 4 and c != t: gl.cnot(c, t)
            elif c != t: gl.swap(c, t)
        g = gl.array()
        if len(g) == 0:
            continue
        st.apply_circuit(g)
        ref = O.apply_circuit(ref, g)
    assert np.array_equal(st.to_numpy(), ref)
    assert st.counters()["compiled_passes"] > 0


def test_compiled_c3_layers_at_24_qubits(pkg):
    L = pkg._lib
    n = 24
    gates = pkg.circuits.c3_circuit(n=n, depth=12)
    st = pkg.B200State(n, 0)
    st.set_option(L.OPT_JIT, 1)
    st.apply_circuit(gates)
    ref = O.apply_circuit(O.statevector(n, 0), gates)
    assert rel_err(st.to_numpy(), ref) < RTOL
    c = st.counters()
    assert c["compiled_passes"] == c["passes"] > 0, c
    z = st.expect_z_all()
    assert np.max(np.abs(z - O.expect_z_all(ref, 1e-10))) < RTOL


def test_compiled_batched_registers_with_per_state_slots(pkg):
    """C2-shaped batched step (SLOT records: the coefficients of a gate come from the per-state table): three steps, the
    all through compiled kernels"""
    n, B, T = 12, 37, 3
    C = pkg.circuits
    gates = C.c2_step_gates(n, seed=1603)
    w = [C.SplitMix64(5).uniform() for _ in range(n)]
    st = pkg.B200State(n, 0, batch=B)
    st.set_option(pkg._lib.OPT_JIT, 1)
    refs = [O.statevector(n, 0) for _ in range(B)]
    for t in range(T):
        enc = C.c2_encoding(B, n, t, w)
        st.apply_batched(gates, enc)
        for b in range(B):
            g = gates.copy()
            for i in range(len(g)):
                if g[i]["kind"] == pkg._lib.GATE_U2_SLOT:
                    g[i]["m"] = enc[b, g[i]["q1"]]; g[i]["kind"] = pkg._lib.GATE_U2; g[i]["q1"] = 0
            O.apply_circuit(refs[b], g)
    for b in (0, 17, B - 1):
        assert rel_err(st.to_numpy(b), refs[b]) < RTOL
    assert st.counters()["compiled_passes"] > 0


@pytest.mark.timeout(150, method="thread")
@pytest.mark.parametrize("n,world", [(23, 2), (21, 4)])
def test_compiled_passes_on_a_sharded_register(pkg, n, world):
    """sharded registers (same-process world on one GPU): rank predicates and rank-dependent phases are data of the launch,
    slab launches beside the exchange use the same compiled kernel"""
    import gpu_dist_worker as W
    R = W.reference(pkg, n)
    L = pkg._lib
    res = pkg.dist.run_local_world(n, world, lambda r, st: W.check_sharded(pkg, st, r, world, R), m=R["m"],
                                   options={L.OPT_JIT: 1}, timeout=120.0)
    for r, out in enumerate(res):
        assert out["ok"], (r, out)
    assert L.jit_stats()["failed"] == 0


def test_driver_ptx_compiler_when_nvjitlink_is_not_used(pkg):
    """without libnvJitLink the generated text goes to the driver's own PTX compiler (cuModuleLoadDataEx): same results.
    The choice is made once per process, so this runs in a fresh interpreter with QM_JIT_BACKEND=driver."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import __graft_entry__ as ge\n"
        "from oracle import oracle as O\n"
        "from test_abi_cpu import random_circuit\n"
        "pkg = ge.load_package(); L = pkg._lib\n"
        "n = 15; gates = random_circuit(pkg, n, 200, 4242); psi0 = O.random_state(n, 9)\n"
        "st = pkg.B200State.from_numpy(psi0); st.set_option(L.OPT_JIT, 1); st.apply_circuit(gates)\n"
        "err = float(np.max(np.abs(st.to_numpy() - O.apply_circuit(psi0.copy(), gates))))\n"
        "c = st.counters(); s = L.jit_stats()\n"
        "assert err < 1e-12 and c['compiled_passes'] == c['passes'] > 0 and s['backend'] == 'driver' and s['failed'] == 0, (err, c, s)\n"
        "print('DRIVER_JIT_OK', err, s)\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QM_JIT_BACKEND="driver")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "DRIVER_JIT_OK" in out.stdout, out.stdout + out.stderr


def test_precompile_circuit_compiles_ahead_and_leaves_the_register_alone(pkg):
    """qm_precompile_circuit: the kernels of a circuit's passes are compiled before its first application (all together, from the
    plan it leaves in the plan cache), the register — amplitudes and pending global phase — is not touched, and the FIRST
    application already runs compiled.  On a register below the default size gate it is a no-op."""
    L = pkg._lib
    n = 16
    gl = pkg.circuits.GateList(n)
    rng = np.random.default_rng(77)
    for q in range(1, n + 1):
        gl.rz(q, float(rng.uniform(-6.3, 6.3))); gl.ry(q, float(rng.uniform(-6.3, 6.3)))
    for q in range(1, n - 2):
        gl.crx(q, q + 2, float(rng.uniform(-6.3, 6.3)))
    for q in range(n, 1, -1):
        gl.cnot(q, q - 1)
    for q in range(1, n + 1):
        gl.rz(q, float(rng.uniform(-6.3, 6.3))); gl.rx(q, float(rng.uniform(-6.3, 6.3)))
    gates = gl.array()
    small = pkg.B200State(n, 5)
    assert small.precompile_circuit(gates) == 0
    small.apply_circuit(gates)
    assert small.counters()["compiled_passes"] == 0
    st = pkg.B200State(n, 5, batch=1024)                    # 2^26 amplitudes: the default (QM_OPT_JIT 2) applies
    st.counters_reset()
    before = L.jit_stats()
    k = st.precompile_circuit(gates)
    after = L.jit_stats()
    assert k > 0 and after["compiled"] - before["compiled"] == k and after["failed"] == 0, (k, before, after)
    c = st.counters()
    assert c["passes"] == 0 and c["launches"] == 0 and c["gates"] == 0, c
    e5 = np.zeros(1 << n, dtype=np.complex128); e5[5] = 1.0
    for b in (0, 1023):
        assert np.array_equal(st.to_numpy(b), e5)           # untouched, and no global phase owed
    assert st.precompile_circuit(gates) == 0                # everything is there already
    st.apply_circuit(gates)
    c = st.counters()
    assert c["passes"] > 0 and c["compiled_passes"] == c["passes"], c
    ref = O.apply_circuit(O.statevector(n, 5), gates)
    for b in (0, 700, 1023):
        assert rel_err(st.to_numpy(b), ref) < RTOL
    assert rel_err(small.to_numpy(), ref) < RTOL
    st.close(); small.close()


def test_disk_cache_of_compiled_kernels_across_processes(pkg, tmp_path):
    """QM_JIT_CACHE_DIR: the second PROCESS that runs a circuit of the same structure loads its kernels from the directory
    (nothing assembled, same results to the last bit)."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import __graft_entry__ as ge\n"
        "from oracle import oracle as O\n"
        "from test_abi_cpu import random_circuit\n"
        "pkg = ge.load_package(); L = pkg._lib\n"
        "n = 15; gates = random_circuit(pkg, n, 200, 777); psi0 = O.random_state(n, 9)\n"
        "st = pkg.B200State.from_numpy(psi0); st.set_option(L.OPT_JIT, 1); st.apply_circuit(gates)\n"
        "got = st.to_numpy(); err = float(np.max(np.abs(got - O.apply_circuit(psi0.copy(), gates))))\n"
        "c = st.counters(); s = L.jit_stats(); d = L.jit_cache_stats()\n"
        "assert err < 1e-12 and c['compiled_passes'] == c['passes'] > 0 and s['failed'] == 0, (err, c, s)\n"
        "import hashlib\n"
        "print('DISK_CACHE', s['backend'], s['compiled'], d['disk_hits'], d['disk_stores'], hashlib.sha1(got.tobytes()).hexdigest())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, QM_JIT_CACHE_DIR=str(tmp_path / "kernels"))
    env.pop("QM_JIT_BACKEND", None)
    runs = []
    for rep in range(2):
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "DISK_CACHE" in out.stdout, out.stdout + out.stderr
        runs.append(out.stdout[out.stdout.index("DISK_CACHE"):].split())
    (_, be0, comp0, hits0, stores0, sha0), (_, be1, comp1, hits1, stores1, sha1) = runs
    if be0 != "nvJitLink":
        pytest.skip("libnvJitLink is not installed here: the driver compiler takes PTX text, there is no cubin to keep")
    assert int(stores0) > 0 and int(hits1) == int(comp1) == int(comp0) and int(stores1) == 0, runs
    assert sha0 == sha1
    assert len(os.listdir(tmp_path / "kernels")) == int(stores0)
