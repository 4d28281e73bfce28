"""Backends for tests/ref_known_answers.py: literal numpy restatement and the C oracle."""
import types

import numpy as np

from oracle import literal as L
from oracle import oracle as O


def literal_backend():
    B = types.SimpleNamespace()
    B.Error = RuntimeError
    B.statevector, B.nb, B.mp = L.statevector, L.nb, L.mp
    B.to_numpy = lambda s: s
    B.h = lambda s, t: L.u2(s, t, L.H)
    B.x = lambda s, t: L.u2(s, t, L.X)
    B.y = lambda s, t: L.u2(s, t, L.Y)
    B.z = lambda s, t: L.u2(s, t, L.Z)
    B.rx = lambda s, t, th: L.u2(s, t, L.rx_gate(th))
    B.ry = lambda s, t, th: L.u2(s, t, L.ry_gate(th))
    B.rz = lambda s, t, th: L.u2(s, t, L.rz_gate(th))
    B.cnot = L.cnot
    B.crx = lambda s, c, t, th: L.controlled(s, c, t, L.rx_gate(th))
    B.cry = lambda s, c, t, th: L.controlled(s, c, t, L.ry_gate(th))
    B.crz = lambda s, c, t, th: L.controlled(s, c, t, L.rz_gate(th))
    B.swap = L.swap
    B.measure_z, B.measure_x, B.measure_y = L.measure_z, L.measure_x, L.measure_y

    def reduced(psi, keep_low, k):
        n = L.nb(psi)
        dB = (1 << k) if keep_low else (1 << (n - k))
        dA = (1 << n) // dB
        return L.partial_trace(L.statevector_to_density_matrix(psi), dA, dB, 1 if keep_low else 2)
    B.reduced = reduced
    B.entropy = L.von_neumann_entropy
    return B


def oracle_backend(seq=True):
    B = types.SimpleNamespace()
    B.Error = O.OracleError
    B.statevector, B.nb, B.mp = O.statevector, O.nb, O.mp
    B.to_numpy = lambda s: s
    for name in ("h", "x", "y", "z", "rx", "ry", "rz", "cnot", "crx", "cry", "crz", "swap"):
        setattr(B, name, getattr(O, name))
    B.measure_z = lambda s, t: O.measure_z(s, t, 1e-10, seq)
    B.measure_x = lambda s, t: O.measure_x(s, t, 1e-10, seq)
    B.measure_y = lambda s, t: O.measure_y(s, t, 1e-10, seq)
    B.reduced = lambda psi, keep_low, k: O.reduced_density(np.ascontiguousarray(psi), keep_low, k)
    B.entropy = L.von_neumann_entropy
    return B
