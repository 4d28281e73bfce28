"""oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT.

ctypes front-end of oracle/libqsim_oracle.so (built from oracle/qsim_oracle.c by oracle/Makefile),
shaped like the reference's QSim API: 1-based qubits, in-place on a numpy complex128 vector.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import literal as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqsim_oracle.so")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "qsim_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(_SO)
        l.orc_num_threads.restype = C.c_int
        l.orc_set_threads.argtypes = [C.c_int]
        l.orc_statevector.argtypes = [_dp, C.c_int, C.c_uint64]
        l.orc_u2.argtypes = [_dp, C.c_int, C.c_int, _dp]
        l.orc_u2.restype = C.c_int
        l.orc_controlled.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, C.c_int]
        l.orc_controlled.restype = C.c_int
        l.orc_swap.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int]
        l.orc_swap.restype = C.c_int
        l.orc_mp.argtypes = [_dp, C.c_int, _dp]
        l.orc_measure_z_seq.argtypes = [_dp, C.c_int, C.c_int, C.c_double, _dp]
        l.orc_measure_z_seq.restype = C.c_int
        l.orc_measure_z.argtypes = [_dp, C.c_int, C.c_int, C.c_double, _dp]
        l.orc_measure_z.restype = C.c_int
        l.orc_expect_z_all.argtypes = [_dp, C.c_int, C.c_double, _dp]
        l.orc_measure_rot.argtypes = [_dp, C.c_int, C.c_int, _dp, C.c_double, C.c_int, _dp]
        l.orc_measure_rot.restype = C.c_int
        l.orc_reduced_density.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp]
        l.orc_sample.argtypes = [_dp, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
        l.orc_apply_circuit.argtypes = [_dp, C.c_int, C.c_void_p, C.c_int]
        l.orc_apply_circuit.restype = C.c_int
        l.orc_norm2.argtypes = [_dp, C.c_int]
        l.orc_norm2.restype = C.c_double
        l.orc_random_state.argtypes = [_dp, C.c_int, C.c_uint64]
        l.orc_splitmix64.argtypes = [C.c_uint64]
        l.orc_splitmix64.restype = C.c_uint64
        l.orc_effective_ct.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib = l
    return _lib


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def _mat(U):
    """2x2 complex -> Julia Matrix{ComplexF64} memory (column-major interleaved)."""
    U = np.asarray(U, dtype=np.complex128)
    return np.ascontiguousarray(U.T).view(np.float64).reshape(8).copy()


class OracleError(RuntimeError):
    """mirrors the reference's ErrorException"""


def _check(rc):
    if rc == 1 or rc == 2:
        raise OracleError("Qubit index out of range")
    if rc == 3:
        raise OracleError("Control and target cannot be the same qubit")
    if rc:
        raise OracleError("oracle error %d" % rc)


def num_threads():
    return lib().orc_num_threads()


def set_threads(t):
    lib().orc_set_threads(int(t))


def nb(s):
    return int(round(np.log2(len(s))))


def statevector(n, m):
    s = np.empty(1 << n, dtype=np.complex128)
    lib().orc_statevector(_p(s), n, m)
    return s


def random_state(n, seed):
    s = np.empty(1 << n, dtype=np.complex128)
    lib().orc_random_state(_p(s), n, seed)
    return s


def u2(s, t, U):
    m = _mat(U)
    _check(lib().orc_u2(_p(s), nb(s), t, _p(m)))
    return s


def h(s, t): return u2(s, t, L.H)
def x(s, t): return u2(s, t, L.X)
def y(s, t): return u2(s, t, L.Y)
def z(s, t): return u2(s, t, L.Z)
def rx(s, t, th): return u2(s, t, L.rx_gate(th))
def ry(s, t, th): return u2(s, t, L.ry_gate(th))
def rz(s, t, th): return u2(s, t, L.rz_gate(th))


def controlled(s, c, t, V, quirk=True):
    m = _mat(V)
    _check(lib().orc_controlled(_p(s), nb(s), c, t, _p(m), 1 if quirk else 0))
    return s


def cnot(s, c, t, quirk=True): return controlled(s, c, t, L.X, quirk)
def crx(s, c, t, th, quirk=True): return controlled(s, c, t, L.rx_gate(th), quirk)
def cry(s, c, t, th, quirk=True): return controlled(s, c, t, L.ry_gate(th), quirk)
def crz(s, c, t, th, quirk=True): return controlled(s, c, t, L.rz_gate(th), quirk)


def swap(s, q1, q2, quirk=True):
    _check(lib().orc_swap(_p(s), nb(s), q1, q2, 1 if quirk else 0))
    return s


def effective_ct(c, t):
    ce, te = C.c_int(), C.c_int()
    lib().orc_effective_ct(c, t, C.byref(ce), C.byref(te))
    return ce.value, te.value


def mp(s):
    out = np.empty(len(s), dtype=np.float64)
    lib().orc_mp(_p(s), nb(s), _p(out))
    return out


def measure_z(s, t, threshold=1e-10, seq=False):
    out = np.empty(2)
    f = lib().orc_measure_z_seq if seq else lib().orc_measure_z
    _check(f(_p(s), nb(s), t, threshold, _p(out)))
    return (float(out[0]), float(out[1]))


def measure_x(s, t, threshold=1e-10, seq=False):
    out = np.empty(2)
    m = _mat(L.H)
    _check(lib().orc_measure_rot(_p(s), nb(s), t, _p(m), threshold, int(seq), _p(out)))
    return (float(out[0]), float(out[1]))


def measure_y(s, t, threshold=1e-10, seq=False):
    out = np.empty(2)
    m = _mat(L.rx_gate(np.pi / 2))
    _check(lib().orc_measure_rot(_p(s), nb(s), t, _p(m), threshold, int(seq), _p(out)))
    return (float(out[0]), float(out[1]))


def expect_z_all(s, threshold=1e-10):
    n = nb(s)
    out = np.empty(2 * n)
    lib().orc_expect_z_all(_p(s), n, threshold, _p(out))
    return out.reshape(n, 2)


def reduced_density(s, keep_low, k):
    d = 1 << k
    out = np.empty(d * d, dtype=np.complex128)
    lib().orc_reduced_density(_p(s), nb(s), 1 if keep_low else 0, k, _p(out))
    return out.reshape(d, d).T.copy()      # column-major -> numpy [i, j]


def sample(s, seed, nshots):
    out = np.empty(nshots, dtype=np.uint64)
    lib().orc_sample(_p(s), nb(s), seed, nshots, out.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out


def norm2(s):
    return lib().orc_norm2(_p(s), nb(s))


GATE_DTYPE = np.dtype([("kind", np.int32), ("q0", np.int32), ("q1", np.int32), ("flags", np.int32),
                       ("m", np.float64, (8,))])


def apply_circuit(s, gates):
    """gates: structured array of GATE_DTYPE (0-based textbook bits — the engine's wire format)."""
    assert gates.dtype == GATE_DTYPE and gates.flags["C_CONTIGUOUS"]
    rc = lib().orc_apply_circuit(_p(s), nb(s), gates.ctypes.data, len(gates))
    _check(rc)
    return s
