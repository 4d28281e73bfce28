"""Static checks of the Julia shim quromorphic.jl_b200/julia/B200QSim.jl (CPU only; no Julia runtime exists here, so the shim
cannot be RUN — but it can be PARSED with the Julia-subset parser of oracle/jl_subset.py and checked against the C ABI and
against the reference it extends):
  * the file parses;
  * every `ccall((:name, LIB), Ret, (ArgTypes...), args...)` names a symbol include/qm_b200.h declares, passes as many arguments
    as it lists types, lists as many types as the C function has parameters, and each Julia type matches the C parameter's type
    (the same table the Python mirror's ctypes signatures are built from);
  * every function the shim adds methods to by `import Quromorphic.QSim: ...` / `QInfo: ...` exists in the reference's own source
    (executed by the interpreter), and each such method takes a number of positional arguments the reference's function also
    accepts — what makes `s = B200State(n, m)` a drop-in for `s = statevector(n, m)` in code written against QSim."""
import ctypes as C
import os
import re
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
from oracle import jl_subset as J  # noqa: E402

SHIM = os.path.join(ROOT, "quromorphic.jl_b200", "julia", "B200QSim.jl")
REF = os.environ.get("QM_REFERENCE_ROOT", "/root/reference")
HAVE_REF = os.path.exists(os.path.join(REF, "src", "Quantum", "QSim.jl"))


def walk(node):
    """every tuple node of the AST (the JType objects of annotations are leaves)"""
    if isinstance(node, tuple):
        yield node
        for x in node:
            yield from walk(x)
    elif isinstance(node, list):
        for x in node:
            yield from walk(x)
    elif isinstance(node, dict):
        for x in node.values():
            yield from walk(x)


@pytest.fixture(scope="module")
def shim_ast():
    with open(SHIM, encoding="utf-8") as f:
        return J.parse(f.read())


def type_name(node):
    """Julia type expression in an argument-type tuple -> 'Int32', 'Ptr', ..."""
    if node[0] == "id": return node[1]
    if node[0] == "curly": return type_name(node[1])
    raise AssertionError("unexpected type expression %r" % (node,))


def c_kind(ct):
    if ct in (C.c_int32,): return "Int32"
    if ct in (C.c_int64,): return "Int64"
    if ct in (C.c_uint64,): return "UInt64"
    if ct in (C.c_uint32,): return "UInt32"
    if ct in (C.c_double,): return "Float64"
    if ct in (C.c_void_p, C.c_char_p) or (isinstance(ct, type) and issubclass(ct, C._Pointer)): return "Ptr"
    raise AssertionError("unmapped ctypes type %r" % (ct,))


def check_ccalls(pkg, ast):
    """every ccall of the tree against the ctypes signature table; returns the set of symbols called"""
    sigs = pkg._lib.SIGNATURES
    calls = [n for n in walk(ast) if n[0] == "call" and n[1] == ("id", "ccall")]
    seen = set()
    for c in calls:
        args = c[2]
        target, ret, argtypes, actual = args[0], args[1], args[2], args[3:]
        assert target[0] == "tuple" and target[1][0][0] == "str" and target[1][1] == ("id", "LIB"), target
        name = target[1][0][1]
        seen.add(name)
        assert name in sigs, "ccall of %s: not declared in include/qm_b200.h" % name
        restype, cargs = sigs[name]
        assert type_name(ret) == c_kind(restype), (name, ret)
        types = argtypes[1] if argtypes[0] == "tuple" else [argtypes[1]]
        assert len(types) == len(cargs), "%s: %d Julia argument types for %d C parameters" % (name, len(types), len(cargs))
        assert len(actual) == len(types), "%s: %d arguments for %d types" % (name, len(actual), len(types))
        for i, (jt, ct) in enumerate(zip(types, cargs)):
            jn = type_name(jt)
            jn = {"Ref": "Ptr", "Cstring": "Ptr", "Cint": "Int32", "Cdouble": "Float64", "Csize_t": "UInt64"}.get(jn, jn)
            assert jn == c_kind(ct), "%s: argument %d is %s in the shim, %s in the C ABI" % (name, i + 1, jn, c_kind(ct))
    return seen, len(calls)


def test_shim_parses_and_every_ccall_matches_the_c_abi(pkg, shim_ast):
    seen, ncalls = check_ccalls(pkg, shim_ast)
    assert ncalls >= 25
    # the path of the reference's API is all there
    for must in ("qm_create", "qm_destroy", "qm_upload", "qm_download", "qm_apply_u2", "qm_apply_controlled", "qm_apply_circuit",
                 "qm_apply_circuit_expect_z", "qm_precompile_circuit", "qm_probs", "qm_measure_with", "qm_expect_z_all", "qm_reduced_density",
                 "qm_sample", "qm_set_option", "qm_last_error"):
        assert must in seen, must


@pytest.mark.skipif(not HAVE_REF, reason="the reference sources are not on this machine")
def test_shim_methods_extend_real_reference_functions_with_matching_arity(shim_ast):
    import make_ref_exec_golden as MR
    it = MR.load_reference()
    src = open(SHIM, encoding="utf-8").read()
    imported = {}
    for mod, names in re.findall(r"^import Quromorphic\.(\w+):((?:[^\n]*,\s*\n)*[^\n]*)", src, flags=re.M):
        for nm in re.split(r"[,\s]+", re.sub(r"#.*", "", names)):
            if nm: imported[nm] = mod
    assert {"u2!", "controlled!", "swap!", "mp", "measure_z", "measure_x", "measure_y", "nb", "partial_trace"} <= set(imported)
    for nm, mod in imported.items():
        assert nm in it.modules[mod].vars, "%s is imported from the reference's %s but the reference does not define it" % (nm, mod)
    module = shim_ast[1][0]
    assert module[0] == "module" and module[1] == "B200QSim"
    checked = 0
    for st in module[2][1]:
        if st[0] != "function" or st[1] not in imported: continue
        ref_fn = it.modules[imported[st[1]]].vars[st[1]]
        ref_arities = {len([p for p in m.params if p[2] is None]) for m in ref_fn.methods}
        mine = len([p for p in st[2] if p[2] is None])
        # QInfo functions take the matrix rho; the shim's methods take the state instead of rho: same positional count
        assert mine in ref_arities, "%s: the shim's method takes %d positional arguments, the reference's take %s" % (st[1], mine, sorted(ref_arities))
        # ... and must not have the positional SIGNATURE of one of the reference's own methods: it would replace that method
        # (keyword parameters do not take part in dispatch) — how `density_matrix(q::Int, m::Int; ...)` was caught
        my_sig = [repr(p[1]) for p in st[2] if len(p) < 4]
        for m in ref_fn.methods:
            assert [repr(p[1]) for p in m.params if len(p) < 4] != my_sig, "%s%s would replace the reference's own method" % (st[1], tuple(my_sig))
        checked += 1
    assert checked >= 25


def test_integration_md_excerpts_parse_and_agree_with_the_abi(pkg):
    """the Julia excerpts INTEGRATION.md shows a maintainer are held to the same checks as the shim itself"""
    with open(os.path.join(ROOT, "INTEGRATION.md"), encoding="utf-8") as f:
        blocks = re.findall(r"```julia\n(.*?)```", f.read(), flags=re.S)
    assert len(blocks) >= 2
    total = 0
    for blk in blocks:
        seen, n = check_ccalls(pkg, J.parse(blk))
        total += n
    assert total >= 3
