"""The Julia shim EXECUTED (CPU only): quromorphic.jl_b200/julia/B200QSim.jl cannot run for real here — there is no Julia runtime and,
in this suite, no GPU — but its host logic can: the Julia-subset interpreter (oracle/jl_subset.py) loads the reference's own
QSim.jl / QInfo.jl, then the shim on top of them (its `import Quromorphic.QSim: u2!, ...` adds METHODS to the reference's functions,
as in Julia), with `ccall` handed to a recorder instead of the library.  The same sequence of reference-API calls is then made
through the tested Python mirror (qsim.py / qinfo.py) with ITS library object replaced by a recorder, and the two call logs must
be identical: same C symbols in the same order, same 0-based bits after the reference's control/target map, same thresholds and
batch indices, and bit-identical 2x2 matrices (rx_gate etc. are the reference's own functions on the Julia side, the mirror's on
the Python side).  What the GPU suite establishes for the mirror therefore carries over to the shim call by call."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
from oracle import jl_subset as J  # noqa: E402

REF = os.environ.get("QM_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "Quantum", "QSim.jl")),
                                reason="the reference sources are not on this machine")

MATRIX_ARGS = {"qm_apply_u2": (2,), "qm_apply_controlled": (3,), "qm_measure_with": (2,), "qm_expect_all_with": (1,), "qm_dm_measure": (2,)}
HANDLE = 0xBEEF


def make_calls(n, count, seed):
    """a seeded sequence of reference-API calls (1-based qubits): (name, args)"""
    rng = np.random.default_rng(seed)
    calls = []
    for _ in range(count):
        k = int(rng.integers(0, 18))
        t = int(rng.integers(1, n + 1))
        c = int(rng.integers(1, n))
        c = c + 1 if c >= t else c
        th = float(rng.uniform(-6.3, 6.3))
        if k < 4: calls.append((("h", "x", "y", "z")[k], (t,)))
        elif k < 7: calls.append((("rx", "ry", "rz")[k - 4], (t, th)))
        elif k == 7: calls.append(("cnot", (c, t)))
        elif k < 11: calls.append((("crx", "cry", "crz")[k - 8], (c, t, th)))
        elif k == 11: calls.append(("swap", (c, t)))
        elif k == 12: calls.append((("measure_z", "measure_x", "measure_y")[int(rng.integers(0, 3))], (t,)))
        elif k == 13: calls.append(("mp", ()))
        elif k == 15: calls.append(("expect_z_all", ()))
        elif k == 16: calls.append(("expect_all", (("z", "x", "y")[int(rng.integers(0, 3))],)))
        elif k == 17: calls.append(("sample", (int(rng.integers(1, 1 << 40)), int(rng.integers(1, 200)))))
        else: calls.append(("partial_trace", (1 << (n - 2), 4, 1) if rng.random() < 0.5 else (4, 1 << (n - 2), 2)))
    return calls


def julia_log(n, m, calls, density):
    import make_ref_exec_golden as MR
    it = MR.load_reference()
    it.modules["Quromorphic"] = it.globals
    log = []

    def handler(name, args):
        rec = [name]
        for i, a in enumerate(args):
            if isinstance(a, J.JRef):
                if name == "qm_create": a.value = HANDLE
                rec.append("ref")
            elif isinstance(a, np.ndarray):
                if i in MATRIX_ARGS.get(name, ()):
                    rec.append(tuple(np.ascontiguousarray(np.asarray(a, dtype=np.complex128).ravel(order="F")).view(np.float64).tolist()))
                else:
                    rec.append("out")
            elif i in MATRIX_ARGS.get(name, ()) and a == 0: rec.append(None)          # Ptr{ComplexF64}(C_NULL)
            elif a == HANDLE and i < 2: rec.append("H")
            else: rec.append(a)
        log.append(tuple(rec))
        return 0

    it.ccall_handler = handler
    it.current_dir = os.path.join(ROOT, "quromorphic.jl_b200", "julia")
    it.load_file(os.path.join(it.current_dir, "B200QSim.jl"))
    it.using("B200QSim")
    lines = ["s = B200Density(%d, %d)" % (n, m) if density else "s = B200State(%d, %d)" % (n, m)]
    for name, args in calls:
        a = ", ".join((":" + x) if isinstance(x, str) else repr(x) for x in args)          # a str is a Symbol (:x)
        bang = "!" if name in ("h", "x", "y", "z", "rx", "ry", "rz", "cnot", "crx", "cry", "crz", "swap") else ""
        lines.append("%s%s(s%s)" % (name, bang, ", " + a if a else ""))
    it.run("\n".join(lines))
    return log


class FakeLib:
    """stands in for the ctypes library object of quromorphic.jl_b200/_lib.py: records every call, returns QM_OK"""

    def __init__(self, log): self._log = log

    def __getattr__(self, name):
        def call(*args):
            rec = [name]
            for i, a in enumerate(args):
                if isinstance(a, C.c_void_p): rec.append("H")
                elif i in MATRIX_ARGS.get(name, ()):
                    rec.append(None if a is None else tuple(float(a[j]) for j in range(8)))
                elif isinstance(a, (int, np.integer)): rec.append(int(a))
                elif isinstance(a, float): rec.append(a)
                elif a is None: rec.append(None)
                elif type(a).__name__ == "CArgObject": rec.append("ref")
                else: rec.append("out")
            self._log.append(tuple(rec))
            return 0
        return call


def python_log(pkg, n, m, calls, density):
    L, Q, QI = pkg._lib, pkg.qsim, pkg.qinfo
    log = []
    saved = L._lib
    L._lib = FakeLib(log)
    try:
        s = Q.density_matrix(n, m) if density else Q.statevector(n, m)
        for name, args in calls:
            if name == "partial_trace": QI.partial_trace(s, *args)
            elif name in ("expect_z_all", "expect_all", "sample"): getattr(s, name)(*args)
            else: getattr(Q, name)(s, *args)
        # the recorder must not be asked to destroy anything after it has been taken away
        for obj in ([s.vec] if density else [s]):
            obj._h = None
    finally:
        L._lib = saved
    return log


def normalise(log, density):
    out = []
    for rec in log:
        if rec[0] == "qm_destroy": continue
        out.append(tuple("H" if (i in (1, 2) and x == "H") else x for i, x in enumerate(rec)))
    return out


@pytest.mark.parametrize("n,m,seed", [(5, 3, 1), (7, 0, 2), (4, 9, 3)])
def test_shim_and_python_mirror_make_the_same_library_calls(pkg, n, m, seed):
    calls = make_calls(n, 80, seed)
    jl = normalise(julia_log(n, m, calls, False), False)
    py = normalise(python_log(pkg, n, m, calls, False), False)
    assert len(jl) == len(py) and len(jl) > 80
    for a, b in zip(jl, py):
        assert a == b, (a, b)


@pytest.mark.parametrize("q,m,seed", [(3, 2, 4), (4, 0, 5)])
def test_shim_and_python_mirror_agree_on_density_matrices(pkg, q, m, seed):
    """Matrix{ComplexF64} methods: rho on q qubits is a 2q-qubit register holding vec(rho); op on the row bits, conj(op) on the column bits"""
    calls = [c for c in make_calls(q, 70, seed) if c[0] not in ("partial_trace", "expect_z_all", "expect_all", "sample")]
    jl = normalise(julia_log(q, m, calls, True), True)
    py = normalise(python_log(pkg, q, m, calls, True), True)
    assert len(jl) == len(py) and len(jl) > 60
    for a, b in zip(jl, py):
        assert a == b, (a, b)


def test_shim_checks_arguments_like_the_reference(pkg):
    """range first, then c == t (QSim.jl:177-178), as ErrorExceptions; the shim also rejects t < 1"""
    import make_ref_exec_golden as MR
    it = MR.load_reference()
    it.modules["Quromorphic"] = it.globals
    it.ccall_handler = lambda name, args: ([setattr(a, "value", HANDLE) for a in args if isinstance(a, J.JRef)], 0)[1]
    it.current_dir = os.path.join(ROOT, "quromorphic.jl_b200", "julia")
    it.load_file(os.path.join(it.current_dir, "B200QSim.jl"))
    it.using("B200QSim")
    it.run("s = B200State(3, 0)")
    for src, msg in (("h!(s, 4)", "out of range"), ("h!(s, 0)", "out of range"), ("cnot!(s, 2, 2)", "same qubit"), ("cnot!(s, 4, 4)", "out of range"),
                     ("crz!(s, 1, 5, 0.1)", "out of range"), ("swap!(s, 0, 1)", "out of range"), ("measure_z(s, 9)", "out of range"),
                     ("B200State(2, 4)", "BoundsError")):
        with pytest.raises(J.JuliaError, match=msg):
            it.run(src)
    assert it.run("swap!(s, 2, 2) === s") is True


def test_shim_sharded_register_calls_match_the_python_mirror(pkg):
    """one rank of a sharded register: id, creation (argument ORDER of qm_create_dist), a few collective calls, the download of
    this rank's slice (2^n / world amplitudes, not 2^n), the optional peer-memory export"""
    import make_ref_exec_golden as MR
    n, world, rank, m = 9, 4, 2, 5
    it = MR.load_reference()
    it.modules["Quromorphic"] = it.globals
    jl = []

    def handler(name, args):
        rec = [name]
        for i, a in enumerate(args):
            if isinstance(a, J.JRef):
                a.value = HANDLE; rec.append("ref")
            elif isinstance(a, np.ndarray):
                if i in MATRIX_ARGS.get(name, ()): rec.append("mat")
                else: rec.append(("buf", int(a.size)))
            elif a == HANDLE and i == 0: rec.append("H")
            else: rec.append(a)
        jl.append(tuple(rec))
        return 0

    it.ccall_handler = handler
    it.current_dir = os.path.join(ROOT, "quromorphic.jl_b200", "julia")
    it.load_file(os.path.join(it.current_dir, "B200QSim.jl"))
    it.using("B200QSim")
    it.run("""
    id = dist_unique_id()
    s = B200State(%d, %d, %d, id; m = %d, device = 1)
    h!(s, 9); cnot!(s, 9, 1); crz!(s, 2, 8, 0.3)
    z = expect_z_all(s)
    p = measure_x(s, 9)
    v = Vector{ComplexF64}(s)
    pr = mp(s)
    e = dist_ipc_export(s)
    w = s.world
    """ % (n, rank, world, m))
    assert it.globals.vars["w"] == world and len(it.globals.vars["v"]) == (1 << n) // world == len(it.globals.vars["pr"])

    L, Q = pkg._lib, pkg.qsim
    py = []

    class Rec:
        def __getattr__(self, name):
            def call(*args):
                rec = [name]
                for i, a in enumerate(args):
                    if isinstance(a, C.c_void_p): rec.append("H")
                    elif i in MATRIX_ARGS.get(name, ()): rec.append("mat" if a is not None else None)
                    elif isinstance(a, (int, np.integer)): rec.append(int(a))
                    elif isinstance(a, float): rec.append(a)
                    elif type(a).__name__ == "CArgObject": rec.append("ref")
                    elif isinstance(a, C.Array): rec.append(("buf", len(a)))
                    else: rec.append(("buf", None))
                py.append(tuple(rec))
                return 0
            return call

    saved = L._lib
    L._lib = Rec()
    try:
        idb = Q.B200State.unique_id()
        s = Q.B200State.create_dist(n, rank, world, idb, m=m, device=1)
        Q.h(s, 9); Q.cnot(s, 9, 1); Q.crz(s, 2, 8, 0.3)
        s.expect_z_all()
        Q.measure_x(s, 9)
        v = s.to_numpy()
        Q.mp(s)
        s._h = None
    finally:
        L._lib = saved
    assert len(v) == (1 << n) // world
    # same symbols, same order, same scalars (buffers: only their presence; the Julia side also exported its IPC handle)
    strip = lambda log: [tuple(x if not (isinstance(x, tuple) and x[0] == "buf") else "buf" for x in rec) for rec in log if rec[0] != "qm_destroy"]
    a, b = strip(jl), strip(py)
    assert a[-1][0] == "qm_dist_ipc_export" and a[:-1] == b, (a, b)
    assert ("qm_create_dist", n, m, 1, rank, world, "buf", "ref") in a
    assert ("qm_download", "H", "buf", (1 << n) // world, 0) in a


def test_shim_reservoir_loop_calls_match_the_python_mirror(pkg):
    """reset!, norm2, the split readout, batched QInfo, the one-call step and the per-register batched application: symbols, order and
    scalar arguments as the Python mirror's (buffers: presence only)"""
    import make_ref_exec_golden as MR
    n, batch = 6, 3
    it = MR.load_reference()
    it.modules["Quromorphic"] = it.globals
    jl = []

    def handler(name, args):
        rec = [name]
        for i, a in enumerate(args):
            if isinstance(a, J.JRef):
                if name == "qm_create": a.value = HANDLE
                rec.append("ref")
            elif isinstance(a, (np.ndarray, list)): rec.append("buf")
            elif a == HANDLE and i == 0: rec.append("H")
            else: rec.append(a)
        jl.append(tuple(rec))
        return 0

    it.ccall_handler = handler
    it.current_dir = os.path.join(ROOT, "quromorphic.jl_b200", "julia")
    it.load_file(os.path.join(it.current_dir, "B200QSim.jl"))
    it.using("B200QSim")
    it.globals.vars["gates"] = []                      # a Vector{QMGate}: only its length crosses the recorder
    it.globals.vars["mats"] = np.zeros((8, 2, batch))
    it.run("""
    s = B200State(%d, 0; batch = %d)
    reset!(s, 5)
    apply_batched!(s, gates, mats)
    expect_z_all_begin!(s)
    apply_circuit!(s, gates)
    z = expect_z_all_end!(s)
    z2 = apply_circuit_expect_z!(s, gates; threshold = -1.0)
    nrm = norm2(s; batch_idx = 2)
    rb = reduced_density_batch(s, true, 2)
    e0 = entropy_batch(s, false, 2)
    e1 = entropy_batch(s, true, 3; kind = :renyi, alpha = 0.5)
    k = precompile_circuit!(s, gates)
    sync!(s)
    shapes = (size(z), size(z2), size(rb), length(e0))
    """ % (n, batch))
    assert it.globals.vars["shapes"] == ((2, n, batch), (2, n, batch), (4, 4, batch), batch)

    L = pkg._lib
    py = []

    class Rec:
        def __getattr__(self, name):
            def call(*args):
                rec = [name]
                gate_array = name in ("qm_apply_batched", "qm_apply_circuit", "qm_apply_circuit_expect_z", "qm_precompile_circuit")
                for i, a in enumerate(args):
                    if isinstance(a, C.c_void_p): rec.append("H")
                    elif gate_array and i == 1: rec.append("buf")              # gates.ctypes.data: an address
                    elif isinstance(a, bool): rec.append(int(a))
                    elif isinstance(a, (int, np.integer)): rec.append(int(a))
                    elif isinstance(a, float): rec.append(a)
                    elif type(a).__name__ == "CArgObject": rec.append("ref")
                    else: rec.append("buf")
                py.append(tuple(rec))
                return 0
            return call

    saved = L._lib
    L._lib = Rec()
    try:
        s = pkg.B200State(n, 0, batch=batch)
        gates = np.zeros(0, dtype=L.GATE_DTYPE)
        s.reset(5)
        s.apply_batched(gates, np.zeros((batch, 2, 8)))
        s.expect_z_all_begin()
        s.apply_circuit(gates)
        s.expect_z_all_end()
        s.apply_circuit_expect_z(gates, -1.0)
        s.norm2(2)
        s.reduced_density_batch(True, 2)
        s.entropy_batch(False, 2)
        s.entropy_batch(True, 3, kind="renyi", alpha=0.5)
        s.precompile_circuit(gates)
        s.sync()
        s._h = None
    finally:
        L._lib = saved
    a = [r for r in jl if r[0] != "qm_destroy"]
    b = [r for r in py if r[0] != "qm_destroy"]
    assert a == b, [(x, y) for x, y in zip(a, b) if x != y]
    assert ("qm_entropy_batch", "H", 1, 3, 1, 0.5, "buf") in a and ("qm_reset", "H", 5) in a
