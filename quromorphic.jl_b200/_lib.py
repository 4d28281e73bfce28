"""ctypes binding of libqm_b200.so — one-to-one with include/qm_b200.h.

The library is built in-tree by quromorphic.jl_b200/csrc/Makefile (see __graft_entry__.build()).
Importing this module on a box without a GPU works (so the CPU tests can check the exported
symbols and the host-only planner); creating a state there raises QMError(QM_ERR_NO_DEVICE):
there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# QM_B200_LIB: an alternative build of the same ABI (A/B timing of kernel versions in tools/sweep.py)
SO_PATH = os.environ.get("QM_B200_LIB") or os.path.join(HERE, "libqm_b200.so")

QM_OK, QM_ERR_ARG, QM_ERR_CUDA, QM_ERR_NOMEM, QM_ERR_NO_DEVICE, QM_ERR_STATE, QM_ERR_COMM = range(7)
GATE_U2, GATE_CU2, GATE_SWAP, GATE_U2_SLOT = 0, 1, 2, 3
BASIS_Z, BASIS_X, BASIS_Y = 0, 1, 2
OPT_EAGER, OPT_TILE_BITS, OPT_LOW_BITS, OPT_FUSE, OPT_TRACE, OPT_PIPELINE, OPT_MAX_COST, OPT_TIME_PASSES, OPT_SHORT_PCT = range(9)
OPT_OVERLAP, OPT_XSM, OPT_XUN, OPT_SLAB_BITS, OPT_OVERLAP_DEPTH, OPT_JIT, OPT_DEFER_PASS, OPT_ZFOLD = 9, 10, 11, 12, 13, 14, 15, 16

GATE_DTYPE = np.dtype([("kind", np.int32), ("q0", np.int32), ("q1", np.int32), ("flags", np.int32),
                       ("m", np.float64, (8,))])
assert GATE_DTYPE.itemsize == 80


class Counters(C.Structure):
    _fields_ = [("gates", C.c_uint64), ("passes", C.c_uint64), ("launches", C.c_uint64),
                ("bytes_hbm", C.c_uint64), ("bytes_link", C.c_uint64), ("bytes_h2d", C.c_uint64),
                ("bytes_d2h", C.c_uint64), ("ms_device", C.c_double),
                ("ms_plan", C.c_double), ("ms_remap", C.c_double), ("overlapped_remaps", C.c_uint64), ("overlapped_passes", C.c_uint64),
                ("compiled_passes", C.c_uint64), ("deferred_passes", C.c_uint64), ("folded_readouts", C.c_uint64)]


class QMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("qm_b200 error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes): every symbol include/qm_b200.h declares
SIGNATURES = {
    "qm_create": (C.c_int32, [C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "qm_destroy": (C.c_int32, [_vp]),
    "qm_reset": (C.c_int32, [_vp, C.c_uint64]),
    "qm_set_option": (C.c_int32, [_vp, C.c_int32, C.c_int64]),
    "qm_num_qubits": (C.c_int32, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "qm_upload": (C.c_int32, [_vp, _dp, C.c_uint64, C.c_int32]),
    "qm_download": (C.c_int32, [_vp, _dp, C.c_uint64, C.c_int32]),
    "qm_apply_u2": (C.c_int32, [_vp, C.c_int32, _dp]),
    "qm_apply_controlled": (C.c_int32, [_vp, C.c_int32, C.c_int32, _dp]),
    "qm_apply_swap": (C.c_int32, [_vp, C.c_int32, C.c_int32]),
    "qm_apply_circuit": (C.c_int32, [_vp, _vp, C.c_int32]),
    "qm_apply_circuit_expect_z": (C.c_int32, [_vp, _vp, C.c_int32, C.c_double, _dp]),
    "qm_precompile_circuit": (C.c_int32, [_vp, _vp, C.c_int32, C.POINTER(C.c_int32)]),
    "qm_apply_batched": (C.c_int32, [_vp, _vp, C.c_int32, _dp, C.c_int32]),
    "qm_probs": (C.c_int32, [_vp, _dp, C.c_int32]),
    "qm_measure": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_double, C.c_int32, _dp]),
    "qm_measure_with": (C.c_int32, [_vp, C.c_int32, _dp, C.c_double, C.c_int32, _dp]),
    "qm_expect_z_all": (C.c_int32, [_vp, C.c_double, _dp]),
    "qm_expect_z_all_begin": (C.c_int32, [_vp, C.c_double]),
    "qm_expect_z_all_end": (C.c_int32, [_vp, _dp]),
    "qm_expect_all": (C.c_int32, [_vp, C.c_int32, C.c_double, _dp]),
    "qm_expect_all_with": (C.c_int32, [_vp, _dp, C.c_double, _dp]),
    "qm_reduced_density": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, _dp]),
    "qm_reduced_density_batch": (C.c_int32, [_vp, C.c_int32, C.c_int32, _dp]),
    "qm_entropy_batch": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_double, _dp]),
    "qm_herm_measure_batch": (C.c_int32, [C.c_int32, C.c_int32, C.c_double, _dp, _dp, C.c_int32, C.c_int32, _dp]),
    "qm_sample": (C.c_int32, [_vp, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
    "qm_norm2": (C.c_int32, [_vp, C.c_int32, _dp]),
    "qm_dm_from_state": (C.c_int32, [_vp, _vp]),
    "qm_dm_diag": (C.c_int32, [_vp, _dp]),
    "qm_dm_measure": (C.c_int32, [_vp, C.c_int32, _dp, C.c_double, _dp]),
    "qm_dist_unique_id": (C.c_int32, [C.POINTER(C.c_uint8)]),
    "qm_create_dist": (C.c_int32, [C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_uint8), C.POINTER(_vp)]),
    "qm_create_dist_local": (C.c_int32, [C.c_int32, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.POINTER(_vp)]),
    "qm_dist_local_connect": (C.c_int32, [_vp]),
    "qm_dist_ipc_export": (C.c_int32, [_vp, C.POINTER(C.c_uint8)]),
    "qm_dist_ipc_import": (C.c_int32, [_vp, C.POINTER(C.c_uint8)]),
    "qm_plan_describe_ex": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, _vp,
                                        C.c_int32, C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int64)]),
    "qm_plan_gates": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, _vp, C.c_int32, _vp,
                                  C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_int32)]),
    "qm_pass_times": (C.c_int32, [_vp, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32)]),
    "qm_trace_read": (C.c_int32, [_vp, C.POINTER(C.c_uint64), C.c_int64]),
    "qm_plan_ophist": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_double, _vp, C.c_int32, C.POINTER(C.c_int64)]),
    "qm_sync": (C.c_int32, [_vp]),
    "qm_last_error": (C.c_int32, [C.c_char_p, C.c_int32]),
    "qm_counters": (C.c_int32, [_vp, C.POINTER(Counters)]),
    "qm_counters_reset": (C.c_int32, [_vp]),
    "qm_jit_stats": (C.c_int32, [C.POINTER(C.c_uint64), _dp]),
    "qm_jit_cache_dir": (C.c_int32, [C.c_char_p]),
    "qm_jit_cache_stats": (C.c_int32, [C.POINTER(C.c_uint64)]),
    "qm_jit_check": (C.c_int32, [C.c_int32, C.c_int32, C.c_double, _vp, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_char_p, C.c_int64]),
    "qm_abi_version": (C.c_int32, []),
    "qm_device_ptr": (C.c_int32, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
}

_lib = None


def lib():
    """Load the engine.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a).  There is no CPU fallback." % SO_PATH)
        l = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(l, name) and "QM_B200_LIB" in os.environ:
                continue
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    lib().qm_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc):
    if rc != QM_OK:
        raise QMError(rc, last_error())


def dptr(a):
    assert a.dtype in (np.float64, np.complex128) and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def mat8(U):
    """2x2 complex (numpy [row, col]) -> Julia Matrix{ComplexF64} memory: U11 U21 U12 U22, (re, im)."""
    U = np.asarray(U, dtype=np.complex128)
    assert U.shape == (2, 2)
    return np.ascontiguousarray(U.T).view(np.float64).reshape(8).copy()


def plan_gates(n, gates, global_bits=0, tile_bits=12, low_bits=5, max_cost=0.0):
    """host-only: the planner's execution order as (gates, pass_of)"""
    cap = 4 * len(gates) + 16
    out = np.zeros(cap, dtype=GATE_DTYPE)
    po = np.zeros(cap, dtype=np.int32)
    ln = C.c_int32()
    check(lib().qm_plan_gates(n, global_bits, tile_bits, low_bits, max_cost, gates.ctypes.data, len(gates),
                              out.ctypes.data, po.ctypes.data_as(C.POINTER(C.c_int32)), cap, C.byref(ln)))
    return out[:ln.value].copy(), po[:ln.value].copy()


def plan_describe(n, gates, global_bits=0, tile_bits=12, low_bits=5, max_cost=0.0, pipeline=1):
    """host-only: the planner's passes as a list of dicts {ngates, m, L, hb, groups: [(rb, nsub)]}"""
    cap = 64 * len(gates) + 1024
    out = np.zeros(cap, dtype=np.int32)
    ln = C.c_int64()
    check(lib().qm_plan_describe_ex(n, global_bits, tile_bits, low_bits, max_cost, pipeline, gates.ctypes.data,
                                    len(gates), out.ctypes.data_as(C.POINTER(C.c_int32)), cap, C.byref(ln)))
    v, i, steps = out[:ln.value].tolist(), 0, []
    while i < len(v):
        if v[i] == 1:
            steps.append({"remap_left": v[i + 1]})
            i += 2
            continue
        ng, m, L, nh = v[i + 1:i + 5]
        hb = v[i + 5:i + 5 + nh]
        i += 5 + nh
        G = v[i]
        i += 1
        groups = []
        for _ in range(G):
            nrb = v[i]
            rb = v[i + 1:i + 1 + nrb]
            groups.append((rb, v[i + 1 + nrb]))
            i += 2 + nrb
        steps.append({"ngates": ng, "m": m, "L": L, "hb": hb, "groups": groups})
    return steps


def jit_check(n, gates, low_bits=5, max_cost=0.0, compile=True, want_ptx=False, zfold=False):
    """host-only: {passes, with_ptx, assembled, distinct, max_cubin_bytes} (+ the PTX text of the first pass) for the compiled
    pass kernels of this circuit"""
    out = (C.c_int64 * 6)()
    buf = C.create_string_buffer(4 << 20) if want_ptx else None
    check(lib().qm_jit_check(n, low_bits, max_cost, gates.ctypes.data, len(gates), (1 if compile else 0) | (2 if zfold else 0), out, buf, (4 << 20) if want_ptx else 0))
    d = dict(zip(("passes", "with_ptx", "assembled", "distinct", "max_cubin_bytes", "structure_hash"), list(out)))
    if want_ptx:
        d["ptx"] = buf.value.decode()
    return d


def jit_stats():
    out = (C.c_uint64 * 4)()
    ms = C.c_double()
    check(lib().qm_jit_stats(out, C.byref(ms)))
    return {"compiled": out[0], "failed": out[1], "cached": out[2], "backend": {0: "none", 1: "nvJitLink", 2: "driver"}[int(out[3])],
            "compile_ms": ms.value}


def jit_cache_dir(path):
    """keep compiled pass kernels as files under `path` (None = no disk cache); process-wide, see include/qm_b200.h"""
    check(lib().qm_jit_cache_dir(None if path is None else str(path).encode()))


def jit_cache_stats():
    out = (C.c_uint64 * 2)()
    check(lib().qm_jit_cache_stats(out))
    return {"disk_hits": out[0], "disk_stores": out[1]}


def plan_ophist(n, gates, tile_bits=12, low_bits=5, max_cost=0.0):
    """host-only: records of the encoded pass programs by family"""
    h = (C.c_int64 * 9)()
    check(lib().qm_plan_ophist(n, tile_bits, low_bits, max_cost, gates.ctypes.data, len(gates), h))
    return dict(zip(("ROT", "ROTX", "PHASE", "PERMX", "THR", "SLOT", "CTRL", "groups", "passes"), list(h)))
