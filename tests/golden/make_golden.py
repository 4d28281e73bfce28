#!/usr/bin/env python
"""tests/golden/make_golden.py — writes tests/golden/qsim_golden.json.

The reference (Julia) cannot run in this image, so these vectors do not come from the reference binary: they come from
oracle/literal.py, the numpy transcription of the reference's LITERAL algorithm (QSim.jl:52-63 `kron(id(r), kron(U, id(l)))`
then a dense product; QSim.jl:175-203 incl. the non-adjacent control/target map; QSim.jl:297-330 measurements with the
1e-10 threshold; QInfo.jl:381-405 partial_trace of s*s').  They are frozen here so that the C oracle, the host emulator
and the CUDA path are all held to the same committed numbers (a drift of the restatement itself shows up as a diff of
this file).  Floats are stored as C99 hex strings: exact bits.

Each case: a seeded sequence of reference-API calls (1-based qubits) on statevector(n, m), then the final amplitudes,
mp, measure_z/x/y of every qubit, both reduced density matrices for k = 1..min(3, n-1).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import literal as L  # noqa: E402

ONE_Q = ["h", "x", "y", "z", "rx", "ry", "rz"]
TWO_Q = ["cnot", "crx", "cry", "crz", "swap"]


class SplitMix64:
    def __init__(self, seed): self.s = seed & 0xFFFFFFFFFFFFFFFF
    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    def uniform(self): return (self.next() >> 11) * 2.0 ** -53
    def below(self, k): return int(self.uniform() * k)


def make_ops(n, nops, seed, kinds=None):
    g = SplitMix64(seed)
    ops = []
    for _ in range(nops):
        two = n >= 2 and g.uniform() < 0.45
        pool = [k for k in (TWO_Q if two else ONE_Q) if kinds is None or k in kinds] or [k for k in ONE_Q + TWO_Q if k in kinds]
        name = pool[g.below(len(pool))]
        if name in ONE_Q:
            op = [name, 1 + g.below(n)]
        else:
            a = 1 + g.below(n); b = 1 + g.below(n - 1)
            if b >= a: b += 1                      # any ordered pair, adjacent or not (the reference's map differs for them)
            op = [name, a, b]
        if name in ("rx", "ry", "rz", "crx", "cry", "crz"):
            op.append(float.hex(2.0 * np.pi * g.uniform()))
        ops.append(op)
    return ops


def apply_op(s, op):
    name = op[0]
    th = float.fromhex(op[-1]) if isinstance(op[-1], str) else None
    if name == "h": L.u2(s, op[1], L.H)
    elif name == "x": L.u2(s, op[1], L.X)
    elif name == "y": L.u2(s, op[1], L.Y)
    elif name == "z": L.u2(s, op[1], L.Z)
    elif name == "rx": L.u2(s, op[1], L.rx_gate(th))
    elif name == "ry": L.u2(s, op[1], L.ry_gate(th))
    elif name == "rz": L.u2(s, op[1], L.rz_gate(th))
    elif name == "cnot": L.cnot(s, op[1], op[2])
    elif name == "crx": L.controlled(s, op[1], op[2], L.rx_gate(th))
    elif name == "cry": L.controlled(s, op[1], op[2], L.ry_gate(th))
    elif name == "crz": L.controlled(s, op[1], op[2], L.rz_gate(th))
    elif name == "swap": L.swap(s, op[1], op[2])
    else: raise ValueError(name)


def hx(a):
    return [float.hex(float(v)) for v in np.asarray(a, dtype=np.float64).ravel()]


def hxc(a):
    a = np.asarray(a, dtype=np.complex128).ravel()
    return hx(np.stack([a.real, a.imag], axis=1))


CASES = [  # (name, n, basis m, number of calls, seed, restrict to kinds)
    ("n3_mixed", 3, 0, 40, 101, None), ("n4_mixed_from_basis_5", 4, 5, 60, 102, None), ("n5_mixed", 5, 0, 80, 103, None),
    ("n6_mixed", 6, 0, 90, 104, None), ("n7_mixed", 7, 3, 100, 105, None), ("n8_mixed", 8, 0, 110, 106, None),
    ("n6_permutations_only", 6, 37, 70, 107, ("x", "cnot", "swap")),           # bit-exact case
    ("n5_controlled_every_pair", 5, 0, 0, 108, None),                          # filled below: every ordered (c, t)
    ("n2_bell", 2, 0, 0, 109, None),
    # 12 qubits: the smallest register the fused TMA kernel takes (dense 4096 x 4096 operators: ~0.3 s per call here)
    ("n12_mixed_tma_kernel_size", 12, 0, 36, 110, None),
    # config C1 of BASELINE.json (the reference's own CPU-runnable case): 12-qubit reservoir, three time steps of 48 gates each
    # (ry encoding, rx/rz layer, CNOT chain, the non-adjacent crz!(12, 1, gamma)) from |0..0>
    ("c1_reservoir_12q_3_steps", 12, 0, 0, 111, None),
]


def c1_ops(n=12, steps=3):
    """the gate calls of config C1 (quromorphic.jl_b200/circuits.py c1_step / c1_inputs) as golden-file ops"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("qm_circuits_for_golden", os.path.join(ROOT, "quromorphic.jl_b200", "circuits.py"))
    src = open(spec.origin).read()
    # circuits.py imports the package's ctypes layer for the wire format; only its pure-Python generators are needed here
    ns = {"__name__": "qm_circuits_for_golden"}
    exec(compile(src.replace("from . import _lib as L", "L = None").replace("from . import qsim as Q", "Q = None"), spec.origin, "exec"), ns)

    class Rec:
        def __init__(self): self.ops = []
        def ry(self, q, th): self.ops.append(["ry", q, float.hex(float(th))])
        def rx(self, q, th): self.ops.append(["rx", q, float.hex(float(th))])
        def rz(self, q, th): self.ops.append(["rz", q, float.hex(float(th))])
        def cnot(self, c, t): self.ops.append(["cnot", c, t])
        def crz(self, c, t, th): self.ops.append(["crz", c, t, float.hex(float(th))])
    u, w, layer, gamma = ns["c1_inputs"](n, steps)
    rec = Rec()
    for t in range(steps):
        ns["c1_step"](rec, n, u[t], w, layer, gamma)
    return rec.ops


def main():
    out = {"generator": "tests/golden/make_golden.py (oracle/literal.py: literal kron transcription of QSim.jl / QInfo.jl)",
           "cases": []}
    for name, n, m, nops, seed, kinds in CASES:
        if name == "n5_controlled_every_pair":
            g = SplitMix64(seed)
            ops = [["h", q] for q in range(1, n + 1)] + [["ry", q, float.hex(2 * np.pi * g.uniform())] for q in range(1, n + 1)]
            for c in range(1, n + 1):
                for t in range(1, n + 1):
                    if c != t:
                        ops.append(["crx", c, t, float.hex(2 * np.pi * g.uniform())])
                        ops.append(["cnot", t, c])
        elif name == "n2_bell":
            ops = [["h", 1], ["cnot", 1, 2]]
        elif name.startswith("c1_reservoir"):
            ops = c1_ops(n, 3)
        else:
            ops = make_ops(n, nops, seed, kinds)
        s = L.statevector(n, m)
        for op in ops:
            apply_op(s, op)
        rho = L.statevector_to_density_matrix(s) if n <= 8 else None
        case = {"name": name, "n": n, "m": m, "ops": ops, "amps": hxc(s), "mp": hx(L.mp(s)),
                "measure_z": [hx(L.measure_z(s, t)) for t in range(1, n + 1)],
                "measure_x": [hx(L.measure_x(s, t)) for t in range(1, n + 1)],
                "measure_y": [hx(L.measure_y(s, t)) for t in range(1, n + 1)],
                "reduced": []}
        for k in range(1, (min(3, n - 1) if n <= 8 else 0) + 1):
            dB = 1 << k; dA = (1 << n) // dB
            rB = L.partial_trace(rho, dA, dB, 1)          # keep the low k qubits (trace out A)
            rA = L.partial_trace(rho, dB, dA, 2)          # keep the high k qubits
            case["reduced"].append({"k": k, "rho_B_low": hxc(rB), "rho_A_high": hxc(rA),
                                    "S_B": float.hex(float(L.von_neumann_entropy(rB))), "S_A": float.hex(float(L.von_neumann_entropy(rA)))})
        out["cases"].append(case)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "qsim_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes,", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
