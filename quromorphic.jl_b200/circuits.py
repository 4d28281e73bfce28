"""circuits.py — the synthetic reservoir workloads of BASELINE.json / SURVEY.md §8(d), as gate lists in
the engine's wire format (0-based textbook bits, after the reference's index map for non-adjacent
controlled gates has been applied — i.e. exactly what the Julia shim would send).

Random numbers come from SplitMix64 -> u = (x >> 11) * 2^-53 so that every consumer (this harness,
the C oracle, a Julia driver) can regenerate identical circuits; angles are theta = 2*pi*u.
"""
import numpy as np

from . import _lib as L
from . import qsim as Q

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & MASK

    def next_u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def uniform(self):
        return (self.next_u64() >> 11) * (2.0 ** -53)

    def angle(self):
        return 2.0 * np.pi * self.uniform()


class GateList:
    """append-only builder speaking the reference's 1-based API, emitting wire-format gates"""

    def __init__(self, n, mode="reference"):
        self.n, self.mode, self.rows = n, mode, []

    def u2(self, t, U):
        assert 1 <= t <= self.n
        self.rows.append((L.GATE_U2, t - 1, 0, 0, L.mat8(U)))

    def slot(self, t, slot):
        self.rows.append((L.GATE_U2_SLOT, t - 1, slot, 0, np.zeros(8)))

    def controlled(self, c, t, V):
        assert 1 <= c <= self.n and 1 <= t <= self.n and c != t
        ce, te = Q.effective_control_target(c, t) if self.mode == "reference" else (c, t)
        self.rows.append((L.GATE_CU2, ce - 1, te - 1, 0, L.mat8(V)))

    def rx(self, t, th): self.u2(t, Q.rx_gate(th))
    def ry(self, t, th): self.u2(t, Q.ry_gate(th))
    def rz(self, t, th): self.u2(t, Q.rz_gate(th))
    def h(self, t): self.u2(t, Q.H)
    def x(self, t): self.u2(t, Q.X)
    def y(self, t): self.u2(t, Q.Y)
    def z(self, t): self.u2(t, Q.Z)
    def cnot(self, c, t): self.controlled(c, t, Q.X)
    def crx(self, c, t, th): self.controlled(c, t, Q.rx_gate(th))
    def cry(self, c, t, th): self.controlled(c, t, Q.ry_gate(th))
    def crz(self, c, t, th): self.controlled(c, t, Q.rz_gate(th))

    def swap(self, q1, q2):
        """swap!, QSim.jl:262-275: no-op on equal qubits, else three cnot! (each under the reference's index map)"""
        assert 1 <= q1 <= self.n and 1 <= q2 <= self.n
        if q1 != q2:
            self.cnot(q1, q2); self.cnot(q2, q1); self.cnot(q1, q2)

    def array(self):
        a = np.zeros(len(self.rows), dtype=L.GATE_DTYPE)
        for i, r in enumerate(self.rows):
            a[i] = r
        return a

    def __len__(self):
        return len(self.rows)


def reservoir_layer_params(n, seed):
    """fixed reservoir layer of configs C1/C2/C5: rx(alpha_q); rz(beta_q) on every qubit"""
    g = SplitMix64(seed)
    return [(g.angle(), g.angle()) for _ in range(n)]


def c1_step(gl, n, u_t, w, layer, gamma, non_adjacent=True):
    """one time step of config C1: encoding ry, reservoir layer, CNOT chain, and one
    deliberately non-adjacent crz!(n, 1, gamma) that pins the reference's index map."""
    for q in range(1, n + 1):
        gl.ry(q, np.pi * u_t * w[q - 1])
    for q in range(1, n + 1):
        gl.rx(q, layer[q - 1][0])
        gl.rz(q, layer[q - 1][1])
    for q in range(1, n):
        gl.cnot(q, q + 1)
    if non_adjacent and n >= 3:
        gl.crz(n, 1, gamma)


def c1_inputs(n=12, T=100):
    gu, gw, gg = SplitMix64(1201), SplitMix64(1202), SplitMix64(1203)
    u = [gu.uniform() for _ in range(T)]
    w = [gw.uniform() for _ in range(n)]
    layer = [(gg.angle(), gg.angle()) for _ in range(n)]
    gamma = gg.angle()
    return u, w, layer, gamma


def c3_circuit(n=30, depth=200, seed=3001, mode="reference"):
    """config C3: per layer l = 1..depth one of rx/ry/rz per qubit (type (q + l) mod 3, fresh angle),
    then a brickwork of cnot! (l odd, pairs (1,2),(3,4),...) or crz! (l even, pairs (2,3),(4,5),...)."""
    g = SplitMix64(seed)
    gl = GateList(n, mode)
    for l in range(1, depth + 1):
        for q in range(1, n + 1):
            th = g.angle()
            k = (q + l) % 3
            (gl.rx if k == 0 else gl.ry if k == 1 else gl.rz)(q, th)
        if l % 2 == 1:
            for q in range(1, n, 2):
                gl.cnot(q, q + 1)
        else:
            for q in range(2, n, 2):
                gl.crz(q, q + 1, g.angle())
    return gl.array()


def c4_step(n, step, seed=3601, mode="reference"):
    """config C4: one reservoir step = encoding layer (ry on every qubit) + one C3-style layer"""
    g = SplitMix64(seed + 7919 * step)
    gl = GateList(n, mode)
    for q in range(1, n + 1):
        gl.ry(q, np.pi * g.uniform())
    l = step + 1
    for q in range(1, n + 1):
        th = g.angle()
        k = (q + l) % 3
        (gl.rx if k == 0 else gl.ry if k == 1 else gl.rz)(q, th)
    if l % 2 == 1:
        for q in range(1, n, 2):
            gl.cnot(q, q + 1)
    else:
        for q in range(2, n, 2):
            gl.crz(q, q + 1, g.angle())
    return gl.array()


def c2_step_gates(n=16, seed=1603):
    """config C2 shared gate list for one time step: slot-encoded ry per qubit (slot q-1), then the
    shared reservoir layer and the adjacent CNOT chain.  63 gates for n = 16."""
    layer = reservoir_layer_params(n, seed)
    gl = GateList(n)
    for q in range(1, n + 1):
        gl.slot(q, q - 1)
    for q in range(1, n + 1):
        gl.rx(q, layer[q - 1][0])
        gl.rz(q, layer[q - 1][1])
    for q in range(1, n):
        gl.cnot(q, q + 1)
    return gl.array()


def c2_encoding(batch, n, t, w, seed0=1601):
    """per-state encoding matrices of step t: ry(pi * u_{b,t} * w_q) as [batch, n, 8] doubles"""
    out = np.empty((batch, n, 8), dtype=np.float64)
    for b in range(batch):
        g = SplitMix64((seed0 + b) * 1000003 + t)
        u = g.uniform()
        for q in range(n):
            out[b, q] = L.mat8(Q.ry_gate(np.pi * u * w[q]))
    return out


# ---- gate-list files: one circuit, identical bits for every consumer ----------------
# Binary form: 16-byte header  b"QMGATES1" | u32 n_qubits | u32 n_gates  (little endian), then n_gates records of the
# 80-byte wire format (include/qm_b200.h qm_gate: i32 kind, q0, q1, flags; f64 m[8]).
# Text form: "# qmgates 1 <n_qubits> <n_gates>" then one line per gate: kind q0 q1 and the eight matrix doubles as
# C99 hex floats (float.hex()), so that no consumer ever parses a decimal.
GATEFILE_MAGIC = b"QMGATES1"


def save_gates(path, n_qubits, gates, text=False):
    """write a gate list (wire format: 0-based textbook bits) to `path`"""
    assert gates.dtype == L.GATE_DTYPE
    if text:
        with open(path, "w") as f:
            f.write("# qmgates 1 %d %d\n" % (n_qubits, len(gates)))
            for g in gates:
                f.write("%d %d %d %s\n" % (g["kind"], g["q0"], g["q1"], " ".join(float(x).hex() for x in g["m"])))
        return
    with open(path, "wb") as f:
        f.write(GATEFILE_MAGIC)
        f.write(np.array([n_qubits, len(gates)], dtype="<u4").tobytes())
        f.write(np.ascontiguousarray(gates).tobytes())


def load_gates(path):
    """-> (n_qubits, gates); accepts both forms"""
    with open(path, "rb") as f:
        head = f.read(8)
        if head == GATEFILE_MAGIC:
            n, ng = (int(x) for x in np.frombuffer(f.read(8), dtype="<u4"))
            raw = f.read()
            if len(raw) != ng * L.GATE_DTYPE.itemsize:
                raise ValueError("%s: expected %d gate records, file holds %d bytes" % (path, ng, len(raw)))
            return n, np.frombuffer(raw, dtype=L.GATE_DTYPE).copy()
    with open(path, "r") as f:
        first = f.readline().split()
        if first[:3] != ["#", "qmgates", "1"]:
            raise ValueError("%s: not a qmgates file" % path)
        n, ng = int(first[3]), int(first[4])
        gates = np.zeros(ng, dtype=L.GATE_DTYPE)
        for i in range(ng):
            t = f.readline().split()
            if len(t) != 11:
                raise ValueError("%s: gate %d: expected 11 fields" % (path, i))
            gates[i] = (int(t[0]), int(t[1]), int(t[2]), 0, [float.fromhex(x) for x in t[3:]])
        return n, gates
