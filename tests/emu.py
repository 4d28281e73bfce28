"""TEST INFRASTRUCTURE: ctypes binding of tests/native/libqm_emu.so — the host emulator of the fused-pass programs
(the engine's planner + encoder, then a scalar walk over the encoded bytes).  A library of its own: the shipped
libqm_b200.so neither contains nor calls it."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "native", "libqm_emu.so")
_dp = C.POINTER(C.c_double)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", os.path.join(HERE, "native")])
        l = C.CDLL(SO)
        l.qm_plan_emulate.restype = C.c_int32
        l.qm_plan_emulate.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32, _dp, C.POINTER(C.c_int32)]
        l.qm_plan_emulate_sharded.restype = C.c_int32
        l.qm_plan_emulate_sharded.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_int32, _dp,
                                              C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        l.qm_emu_set_defer.restype = C.c_int32
        l.qm_emu_set_defer.argtypes = [C.c_int32]
        l.qm_emu_last_deferred.restype = C.c_int32
        l.qm_emu_set_jit.restype = C.c_int32
        l.qm_emu_set_jit.argtypes = [C.c_int32]
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        buf = C.create_string_buffer(512)
        lib().qm_emu_last_error(buf, 512)
        raise RuntimeError("qm_emu error %d: %s" % (rc, buf.value.decode("utf-8", "replace")))


def plan_emulate(n, gates, psi, tile_bits=12, low_bits=5, max_cost=0.0):
    """the planned + encoded pass programs walked over psi (complex128, length 2^n): (state, passes)"""
    out = np.ascontiguousarray(psi, dtype=np.complex128).copy()
    npass = C.c_int32()
    _check(lib().qm_plan_emulate(n, tile_bits, low_bits, max_cost, gates.ctypes.data, len(gates),
                                 out.ctypes.data_as(_dp), C.byref(npass)))
    return out, npass.value


def plan_emulate_sharded(n, global_bits, gates, psi, tile_bits=12, low_bits=5, max_cost=0.0):
    """plan_emulate for a register sharded over 2^global_bits ranks, all emulated in this process on the whole
    state vector: (state, passes, remaps)"""
    out = np.ascontiguousarray(psi, dtype=np.complex128).copy()
    npass, nrem = C.c_int32(), C.c_int32()
    _check(lib().qm_plan_emulate_sharded(n, global_bits, tile_bits, low_bits, max_cost, gates.ctypes.data, len(gates),
                                         out.ctypes.data_as(_dp), C.byref(npass), C.byref(nrem)))
    return out, npass.value, nrem.value


def set_jit(on):
    """0: walk the encoded records (the interpreter kernel's view); 1: execute the straight-line PTX text generated for the
    compiled pass kernels (csrc/jit_codegen.hpp) with the PTX-subset interpreter of tests/native/ptx_interp.hpp"""
    _check(lib().qm_emu_set_jit(1 if on else 0))


def set_defer(on):
    """sharded emulation: allow (default) or forbid postponing the last pass before an exchange (planner.hpp try_defer_last_pass)"""
    _check(lib().qm_emu_set_defer(1 if on else 0))


def last_deferred():
    """passes postponed across an exchange by the last plan_emulate_sharded call"""
    return int(lib().qm_emu_last_deferred())


def plan_count_sharded(n, global_bits, gates, low_bits=5, max_cost=80.0, cost_model=1):
    """planning only, any size: (passes, exchanges, postponed passes, [gates per pass, -1 = exchange])"""
    l = lib()
    l.qm_emu_plan_count_sharded.restype = C.c_int32
    l.qm_emu_plan_count_sharded.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_int32,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]
    a, b, d = C.c_int32(), C.c_int32(), C.c_int32()
    seq = (C.c_int32 * 64)()
    _check(l.qm_emu_plan_count_sharded(n, global_bits, low_bits, max_cost, cost_model, gates.ctypes.data, len(gates),
                                       C.byref(a), C.byref(b), C.byref(d), seq, 64))
    return a.value, b.value, d.value, list(seq[:a.value + b.value])
