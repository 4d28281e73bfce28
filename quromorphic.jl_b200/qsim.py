"""qsim.py — host-side mirror of the reference's QSim API over the B200 engine.

Same function names as /root/reference/src/Quantum/QSim.jl with the `!` dropped (Python has no
bang identifiers): statevector, nb, id_, u2, h, x, y, z, rx, ry, rz, controlled, cnot, crx, cry,
crz, swap, mp, measure_z, measure_x, measure_y, prstate, and the exported constants X, Y, Z, I2.
Qubit arguments are the reference's 1-based indices; the state is a device-resident `B200State`
(the third "method family" next to Vector{ComplexF64} and Matrix{ComplexF64}).
The Julia shim quromorphic.jl_b200/julia/B200QSim.jl is the same code in the reference's own
language; this module is what can be executed and tested in an image without Julia.

Everything numerical happens in libqm_b200.so (CUDA, sm_100a).  This module only
  * builds the 2x2 gate matrices on the host with the reference's own expressions
    (QSim.jl:10-17, :114-148) so the engine receives the same bits Julia would pass,
  * converts t -> t-1 and applies the reference's index map for non-adjacent controlled gates
    (QSim.jl:181-197; `mode="reference"`, default) or passes (c, t) through (`mode="corrected"`),
  * performs the reference's argument checks in the reference's order and raises QSimError
    (standing in for Julia's ErrorException).
"""
import ctypes as C

import numpy as np

from . import _lib as L

# ---- constants, QSim.jl:8-17 -------------------------------------------------------------------
c1 = np.complex128(1)
im_F64 = np.complex128(1j)
I2 = np.array([[1, 0], [0, 1]], dtype=np.complex128)
H = np.complex128(1 / np.sqrt(2)) * np.array([[1, 1], [1, -1]], dtype=np.complex128)
X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -im_F64], [im_F64, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)

# rotation caches, QSim.jl:25-27 (kept host-side; the engine takes matrices by value)
RX_CACHE, RY_CACHE, RZ_CACHE = {}, {}, {}


class QSimError(RuntimeError):
    """Julia's ErrorException at this boundary (test_QSim.jl:351-372 pins only the type)."""


def id_(n):
    """id(n), QSim.jl:29-31"""
    return np.array([[c1]]) if n == 0 else np.eye(1 << n, dtype=np.complex128)


def _rot(cache, theta, sigma):
    th = float(theta)
    g = cache.get(th)
    if g is None:
        c = np.complex128(np.cos(th / 2))
        s = np.complex128(np.sin(th / 2))
        g = c * I2 - im_F64 * s * sigma
        cache[th] = g
    return g


def rx_gate(theta): return _rot(RX_CACHE, theta, X)     # QSim.jl:114-124
def ry_gate(theta): return _rot(RY_CACHE, theta, Y)     # QSim.jl:126-136
def rz_gate(theta): return _rot(RZ_CACHE, theta, Z)     # QSim.jl:138-148


def effective_control_target(c, t):
    """What controlled!(s, c, t, V) really acts on (QSim.jl:181-197): with b = max(c, t) the 4x4
    block sits on qubits (b-1, b); c < t puts the control on b-1, c > t on b."""
    b = max(c, t)
    return (b - 1, b) if c < t else (b, b - 1)


class B200State:
    """Device-resident state vector(s): the drop-in for Vector{ComplexF64} (QSim.jl:33-37).

    B200State(n, m)        == statevector(n, m)
    B200State.from_numpy(v) uploads a host vector; .to_numpy() downloads it.
    batch > 1 holds `batch` independent n-qubit registers evolved in lockstep.
    mode: "reference" reproduces the reference's non-adjacent controlled-gate behaviour,
          "corrected" applies the textbook gate on (c, t).
    """

    def __init__(self, n, m=0, batch=1, device=0, mode="reference", threshold=1e-10, _handle=None):
        assert mode in ("reference", "corrected")
        self.n, self.batch, self.device, self.mode, self.threshold = int(n), int(batch), device, mode, threshold
        if _handle is not None:
            self._h = _handle
            return
        h = L._vp()
        if m < 0 or m >= (1 << n):
            raise QSimError("BoundsError: basis index %d out of range for %d qubits" % (m, n))
        L.check(L.lib().qm_create(self.n, int(m), self.batch, int(device), C.byref(h)))
        self._h = h

    @classmethod
    def create_dist(cls, n, rank, world, nccl_id, m=0, device=0, mode="reference", threshold=1e-10):
        """this rank's shard of a register sharded over `world` GPUs (one process per GPU); nccl_id is the
        128-byte id from unique_id(), broadcast by the host"""
        h = L._vp()
        idb = (C.c_uint8 * 128).from_buffer_copy(bytes(nccl_id))
        L.check(L.lib().qm_create_dist(int(n), int(m), int(device), int(rank), int(world), idb, C.byref(h)))
        st = cls(n, m, 1, device, mode, threshold, _handle=h)
        st.world, st.rank = world, rank
        return st

    @classmethod
    def create_dist_local(cls, n, rank, world, key, m=0, device=0, mode="reference", threshold=1e-10):
        """this rank's shard when ALL ranks live in this process (one host thread per rank; one device or several):
        every rank calls this with the same `key`, then — after all have returned — .connect_local()."""
        h = L._vp()
        L.check(L.lib().qm_create_dist_local(int(n), int(m), int(device), int(rank), int(world), int(key), C.byref(h)))
        st = cls(n, m, 1, device, mode, threshold, _handle=h)
        st.world, st.rank = world, rank
        return st

    def connect_local(self):
        L.check(L.lib().qm_dist_local_connect(self._h))
        self.p2p = True

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        L.check(L.lib().qm_dist_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_numpy(cls, v, **kw):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        n = int(round(np.log2(len(v))))
        if (1 << n) != len(v):
            raise QSimError("state length must be a power of two")
        st = cls(n, 0, **kw)
        st.upload(v)
        return st

    def upload(self, v, batch_idx=0):
        v = np.ascontiguousarray(v, dtype=np.complex128)
        L.check(L.lib().qm_upload(self._h, L.dptr(v), len(v), batch_idx))

    def to_numpy(self, batch_idx=0):
        out = np.empty(self.local_len, dtype=np.complex128)
        L.check(L.lib().qm_download(self._h, L.dptr(out), len(out), batch_idx))
        return out

    @property
    def local_len(self):
        w = getattr(self, "world", 1)
        return (1 << self.n) // w

    def __len__(self):
        return 1 << self.n

    def set_option(self, key, value):
        L.check(L.lib().qm_set_option(self._h, key, int(value)))

    def reset(self, m=0):
        L.check(L.lib().qm_reset(self._h, int(m)))

    def sync(self):
        L.check(L.lib().qm_sync(self._h))

    def counters(self):
        c = L.Counters()
        L.check(L.lib().qm_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in L.Counters._fields_}

    def pass_times(self, cap=65536):
        """diagnostics (OPT_TIME_PASSES): array [npasses, 8] of {ms, groups, gates, planner cost, records: full, half, xor, thr}"""
        out = np.zeros((cap, 8), dtype=np.float32)
        n = C.c_int32()
        L.check(L.lib().qm_pass_times(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), cap, C.byref(n)))
        return out[:min(n.value, cap)].copy()

    def counters_reset(self):
        L.check(L.lib().qm_counters_reset(self._h))

    def precompile_circuit(self, gates):
        """compile the pass kernels `gates` will need (QM_OPT_JIT) without touching the register; returns how many were compiled"""
        assert gates.dtype == L.GATE_DTYPE and gates.flags["C_CONTIGUOUS"]
        n = C.c_int32()
        L.check(L.lib().qm_precompile_circuit(self._h, gates.ctypes.data, len(gates), C.byref(n)))
        return n.value

    def apply_circuit_expect_z(self, gates, threshold=None):
        """apply_circuit(gates) then expect_z_all(threshold) in one call — a reservoir step.  When the last fused pass runs as
        a compiled kernel the <Z> sums are taken inside it (QM_OPT_ZFOLD; counters()["folded_readouts"])."""
        assert gates.dtype == L.GATE_DTYPE and gates.flags["C_CONTIGUOUS"]
        out = np.empty((self.batch, self.n, 2), dtype=np.float64)
        thr = self.threshold if threshold is None else threshold
        L.check(L.lib().qm_apply_circuit_expect_z(self._h, gates.ctypes.data, len(gates), thr, L.dptr(out)))
        return out if self.batch > 1 else out[0]

    def apply_circuit(self, gates):
        """gates: numpy structured array of _lib.GATE_DTYPE, 0-based textbook bits (wire format)."""
        assert gates.dtype == L.GATE_DTYPE and gates.flags["C_CONTIGUOUS"]
        L.check(L.lib().qm_apply_circuit(self._h, gates.ctypes.data, len(gates)))
        return self

    def apply_batched(self, gates, per_state_mats):
        """per_state_mats: float64 array [batch, nslots, 8]"""
        assert gates.dtype == L.GATE_DTYPE and gates.flags["C_CONTIGUOUS"]
        p = np.ascontiguousarray(per_state_mats, dtype=np.float64)
        assert p.ndim == 3 and p.shape[0] == self.batch and p.shape[2] == 8
        L.check(L.lib().qm_apply_batched(self._h, gates.ctypes.data, len(gates), L.dptr(p), p.shape[1]))
        return self

    def expect_z_all(self, threshold=None):
        """[batch, n, 2] array of (p0, p1) — all-qubit measure_z in one pass."""
        out = np.empty((self.batch, self.n, 2), dtype=np.float64)
        thr = self.threshold if threshold is None else threshold
        L.check(L.lib().qm_expect_z_all(self._h, thr, L.dptr(out)))
        return out if self.batch > 1 else out[0]

    def expect_z_all_begin(self, threshold=None):
        """first half of a split readout: enqueue the all-qubit <Z> of the state as it is now and return at once (gate calls
        may follow immediately; at most two readouts pending)"""
        thr = self.threshold if threshold is None else threshold
        L.check(L.lib().qm_expect_z_all_begin(self._h, thr))

    def expect_z_all_end(self):
        """second half: the oldest pending readout, [batch, n, 2]"""
        out = np.empty((self.batch, self.n, 2), dtype=np.float64)
        L.check(L.lib().qm_expect_z_all_end(self._h, L.dptr(out)))
        return out if self.batch > 1 else out[0]

    def expect_all(self, basis, threshold=None):
        """[batch, n, 2] array of (p0, p1): measure_z / measure_x / measure_y of EVERY qubit (QSim.jl:297-330).
        basis: "z", "x", "y" or a 2x2 rotation matrix R (the marginals of R applied to each qubit in turn).  The
        X/Y matrices are the host's own H and rx_gate(pi/2), as in measure_x / measure_y below."""
        out = np.empty((self.batch, self.n, 2), dtype=np.float64)
        thr = self.threshold if threshold is None else threshold
        if isinstance(basis, str):
            b = basis.lower()
            R = None if b == "z" else (H if b == "x" else rx_gate(np.pi / 2))
        else:
            R = np.asarray(basis, dtype=np.complex128)
        Rp = L.dptr(L.mat8(R)) if R is not None else None
        L.check(L.lib().qm_expect_all_with(self._h, Rp, thr, L.dptr(out)))
        return out if self.batch > 1 else out[0]

    def reduced_density(self, keep_low, k, batch_idx=0):
        d = 1 << k
        out = np.empty(d * d, dtype=np.complex128)
        L.check(L.lib().qm_reduced_density(self._h, 1 if keep_low else 0, k, batch_idx, L.dptr(out)))
        return out.reshape(d, d).T.copy()          # column-major -> numpy [i, j]

    def reduced_density_batch(self, keep_low, k):
        """reduced density matrix of EVERY register of the batch: array [batch, 2^k, 2^k] (numpy [i, j])"""
        d = 1 << k
        out = np.empty((self.batch, d * d), dtype=np.complex128)
        L.check(L.lib().qm_reduced_density_batch(self._h, 1 if keep_low else 0, k, L.dptr(out)))
        return out.reshape(self.batch, d, d).transpose(0, 2, 1).copy()

    def entropy_batch(self, keep_low, k, kind="von_neumann", alpha=2.0):
        """entropy of the reduced matrix of every register of the batch, computed on the device (one Jacobi warp per
        register): kind in von_neumann, renyi, min, max — QInfo.jl:47-50, :199-204, :353-357, :332-337"""
        kinds = {"von_neumann": 0, "renyi": 1, "min": 2, "max": 3}
        out = np.empty(self.batch, dtype=np.float64)
        L.check(L.lib().qm_entropy_batch(self._h, 1 if keep_low else 0, k, kinds[kind], float(alpha), L.dptr(out)))
        return out

    def sample(self, seed, nshots, batch_idx=0):
        out = np.empty(nshots, dtype=np.uint64)
        L.check(L.lib().qm_sample(self._h, int(seed), nshots, batch_idx, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def norm2(self, batch_idx=0):
        v = C.c_double()
        L.check(L.lib().qm_norm2(self._h, batch_idx, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None):
            L.lib().qm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class B200Density:
    """Device-resident density matrix: the drop-in for the Matrix{ComplexF64} methods of QSim.jl
    (u2! :67-79, controlled! :209-232, swap! :277-290, mp :295, measure_z/x/y :332-365, prstate :375-383).

    rho on q qubits lives in the engine as the 2q-qubit register vec(rho) in Julia's column-major Matrix
    memory (amplitude r + c 2^q = rho[r, c]); `rho <- op rho op'` is `op` on the row bits and conj(op) on
    the column bits, fused by the planner like any other gates.  q <= 16 on one B200 (4^16 x 16 B = 64 GiB).
    """

    def __init__(self, q, m=0, device=0, mode="reference", threshold=1e-10):
        if m < 0 or m >= (1 << q):
            raise QSimError("BoundsError: basis index %d out of range for %d qubits" % (m, q))
        self.q, self.mode, self.threshold = int(q), mode, threshold
        self.vec = B200State(2 * q, m * ((1 << q) + 1), device=device, mode="corrected", threshold=threshold)

    @classmethod
    def from_statevector(cls, s):
        """statevector_to_density_matrix(s) = s * s', QSim.jl:39-41, formed on the device"""
        rho = cls(s.n, 0, device=s.device, mode=s.mode, threshold=s.threshold)
        L.check(L.lib().qm_dm_from_state(s._h, rho.vec._h))
        return rho

    @classmethod
    def from_numpy(cls, M, **kw):
        M = np.asarray(M, dtype=np.complex128)
        q = int(round(np.log2(M.shape[0])))
        if M.shape != (1 << q, 1 << q):
            raise QSimError("density matrix must be 2^q x 2^q")
        rho = cls(q, 0, **kw)
        rho.vec.upload(np.ascontiguousarray(M.T).reshape(-1))       # column-major
        return rho

    def to_numpy(self):
        d = 1 << self.q
        return self.vec.to_numpy().reshape(d, d).T.copy()

    def set_option(self, key, value):
        self.vec.set_option(key, value)

    def counters(self):
        return self.vec.counters()

    def close(self):
        self.vec.close()


def density_matrix(n, m, **kw):
    """density_matrix(n, m) = |m><m| on n qubits (QSim.jl:43-46)"""
    return B200Density(n, m, **kw)


def statevector_to_density_matrix(s):
    """statevector_to_density_matrix(s), QSim.jl:39-41"""
    return B200Density.from_statevector(s)


# ---- the reference's functions --------------------------------------------------------------------

def statevector(n, m, **kw):
    """statevector(n, m), QSim.jl:33-37"""
    return B200State(n, m, **kw)


def nb(s):
    """nb(s), QSim.jl:48 (Matrix: round(Int, log2(size(rho, 1))))"""
    return s.q if isinstance(s, B200Density) else s.n


def _range1(s, t):
    # reference: `t <= q || error(...)` (QSim.jl:54); it does not test t >= 1 and fails later with a
    # different exception — the shim rejects it here with the same ErrorException.
    if not (1 <= t <= s.n):
        raise QSimError("Qubit index out of range")


def u2(s, t, U):
    """u2!(s, t, U), QSim.jl:52-63;  u2!(rho, t, U) = op rho op', QSim.jl:67-79"""
    if isinstance(s, B200Density):
        if not (1 <= t <= s.q):
            raise QSimError("Qubit index out of range")
        U = np.asarray(U, dtype=np.complex128)
        m, mc = L.mat8(U), L.mat8(np.conj(U))
        L.check(L.lib().qm_apply_u2(s.vec._h, t - 1, L.dptr(m)))                 # op on the row index
        L.check(L.lib().qm_apply_u2(s.vec._h, s.q + t - 1, L.dptr(mc)))          # conj(op) on the column index
        return s
    _range1(s, t)
    m = L.mat8(U)
    L.check(L.lib().qm_apply_u2(s._h, t - 1, L.dptr(m)))
    return s


def h(s, t): return u2(s, t, H)                      # QSim.jl:81-83
def x(s, t): return u2(s, t, X)                      # QSim.jl:89-91
def y(s, t): return u2(s, t, Y)                      # QSim.jl:97-99
def z(s, t): return u2(s, t, Z)                      # QSim.jl:105-107
def rx(s, t, theta): return u2(s, t, rx_gate(theta))  # QSim.jl:150-152
def ry(s, t, theta): return u2(s, t, ry_gate(theta))  # QSim.jl:158-160
def rz(s, t, theta): return u2(s, t, rz_gate(theta))  # QSim.jl:166-168


def controlled(s, c, t, V):
    """controlled!(s, c, t, V), QSim.jl:175-203 (Matrix: :209-232, the same 4x4 block placement).  Checks in
    the reference's order: range, then c == t."""
    n = nb(s)
    if not (c <= n and t <= n) or c < 1 or t < 1:
        raise QSimError("Qubit index out of range")
    if c == t:
        raise QSimError("Control and target cannot be the same qubit")
    ce, te = effective_control_target(c, t) if s.mode == "reference" else (c, t)
    m = L.mat8(V)
    if isinstance(s, B200Density):
        mc = L.mat8(np.conj(np.asarray(V, dtype=np.complex128)))
        L.check(L.lib().qm_apply_controlled(s.vec._h, ce - 1, te - 1, L.dptr(m)))
        L.check(L.lib().qm_apply_controlled(s.vec._h, s.q + ce - 1, s.q + te - 1, L.dptr(mc)))
        return s
    L.check(L.lib().qm_apply_controlled(s._h, ce - 1, te - 1, L.dptr(m)))
    return s


def cnot(s, c, t): return controlled(s, c, t, X)                     # QSim.jl:205-207
def crx(s, c, t, theta): return controlled(s, c, t, rx_gate(theta))  # QSim.jl:238-240
def cry(s, c, t, theta): return controlled(s, c, t, ry_gate(theta))  # QSim.jl:246-248
def crz(s, c, t, theta): return controlled(s, c, t, rz_gate(theta))  # QSim.jl:254-256


def swap(s, q1, q2):
    """swap!(s, q1, q2), QSim.jl:262-275: range check, no-op when equal, else cnot!(q1,q2); cnot!(q2,q1);
    cnot!(q1,q2) — each through controlled!, so the reference mode inherits its index map (Matrix: :277-290)."""
    n = nb(s)
    if not (q1 <= n and q2 <= n) or q1 < 1 or q2 < 1:
        raise QSimError("Qubit index out of range")
    if q1 == q2:
        return s
    cnot(s, q1, q2)
    cnot(s, q2, q1)
    cnot(s, q1, q2)
    return s


def mp(s, batch_idx=0):
    """mp(s) = abs2.(s), QSim.jl:293;  mp(rho) = real.(diag(rho)), QSim.jl:295"""
    if isinstance(s, B200Density):
        out = np.empty(1 << s.q, dtype=np.float64)
        L.check(L.lib().qm_dm_diag(s.vec._h, L.dptr(out)))
        return out
    out = np.empty(s.local_len, dtype=np.float64)      # a sharded register returns this rank's contiguous slice
    L.check(L.lib().qm_probs(s._h, L.dptr(out), batch_idx))
    return out


def _measure(s, t, R, batch_idx):
    if not (1 <= t <= nb(s)):
        raise QSimError("Qubit index out of range")
    out = np.empty(2, dtype=np.float64)
    Rp = L.dptr(L.mat8(R)) if R is not None else None
    if isinstance(s, B200Density):       # QSim.jl:332-365: diagonal of R rho R' by bit t, entries <= 1e-10 dropped
        L.check(L.lib().qm_dm_measure(s.vec._h, t - 1, Rp, s.threshold, L.dptr(out)))
        return (float(out[0]), float(out[1]))
    L.check(L.lib().qm_measure_with(s._h, t - 1, Rp, s.threshold, batch_idx, L.dptr(out)))
    return (float(out[0]), float(out[1]))


def measure_z(s, t, batch_idx=0):
    """measure_z(s, t), QSim.jl:297-318: (p0, p1), summing only probabilities > 1e-10."""
    return _measure(s, t, None, batch_idx)


def measure_x(s, t, batch_idx=0):
    """measure_x(s, t) = measure_z(h!(copy(s), t), t), QSim.jl:320-324 (no copy is made: the kernel
    rotates each amplitude pair in registers)."""
    return _measure(s, t, H, batch_idx)


def measure_y(s, t, batch_idx=0):
    """measure_y(s, t) = measure_z(rx!(copy(s), t, pi/2), t), QSim.jl:326-330"""
    return _measure(s, t, rx_gate(np.pi / 2), batch_idx)


def measure_all(s, basis="z"):
    """measure_z / measure_x / measure_y for every qubit t = 1..n in one call: array [n, 2] (or [batch, n, 2])"""
    return s.expect_all(basis)


def prstate(s, batch_idx=0, file=None):
    """prstate(s), QSim.jl:367-373: '|bits⟩: amplitude' for abs(a) > 1e-10, MSB-first bit string;
    prstate(rho), QSim.jl:375-383: '|bits⟩ : probability' from mp(rho)."""
    if isinstance(s, B200Density):
        p = mp(s)
        for i in range(len(p)):
            if abs(p[i]) > 1e-10:
                print("|" + format(i, "b").rjust(s.q, "0") + "⟩ : " + repr(float(p[i])), file=file)
        return None
    v = s.to_numpy(batch_idx)
    for i in range(len(v)):
        a = v[i]
        if abs(a) > 1e-10:
            print("|" + format(i, "b").rjust(s.n, "0") + "⟩: " + repr(complex(a)), file=file)
