#!/usr/bin/env python
"""tests/golden/make_ref_exec_golden.py — writes tests/golden/ref_exec_golden.json.

These vectors come from EXECUTING THE REFERENCE'S OWN SOURCE FILES, /root/reference/src/Quantum/QSim.jl and QInfo.jl, read where
they lie at generation time (nothing of them is copied into this repository), with oracle/jl_subset.py — a generic interpreter for
the subset of Julia those files are written in (no Julia runtime exists in this image).  The interpreter knows nothing about gates:
every number below was computed by the reference's `u2!`, `controlled!`, `swap!`, `measure_z/x/y`, `mp`,
`statevector_to_density_matrix`, `partial_trace`, `von_neumann_entropy` as written by its authors (dense `kron` operators,
the non-adjacent control/target map, the 1e-10 threshold), with numpy doing the dense algebra underneath.

Same cases, same call sequences and same schema as qsim_golden.json (which was made from oracle/literal.py, a hand
transcription of the same algorithm), so the two files can be compared entry by entry: tests/test_ref_exec.py does, and holds
the C oracle to this file as well.  Needs /root/reference; run from the repo root:
    python tests/golden/make_ref_exec_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import jl_subset as J  # noqa: E402
import make_golden as MG  # noqa: E402  (the case list and the seeded call sequences)

REF = os.environ.get("QM_REFERENCE_ROOT", "/root/reference")


def load_reference():
    it = J.Interp()
    it.load_file(os.path.join(REF, "src", "Quantum", "QSim.jl"))
    it.load_file(os.path.join(REF, "src", "Quantum", "QInfo.jl"))
    it.using("QSim")
    it.using("QInfo")
    return it


def ref_call(it, name, *args):
    return it.call(it.globals.lookup(name), list(args))


def apply_op(it, s, op):
    name = op[0]
    th = float.fromhex(op[-1]) if isinstance(op[-1], str) else None
    if name in ("h", "x", "y", "z"): ref_call(it, name + "!", s, op[1])
    elif name in ("rx", "ry", "rz"): ref_call(it, name + "!", s, op[1], th)
    elif name in ("cnot", "swap"): ref_call(it, name + "!", s, op[1], op[2])
    elif name in ("crx", "cry", "crz"): ref_call(it, name + "!", s, op[1], op[2], th)
    else: raise ValueError(name)


def build_ops(name, n, nops, seed, kinds):
    if name == "n5_controlled_every_pair":
        g = MG.SplitMix64(seed)
        ops = [["h", q] for q in range(1, n + 1)] + [["ry", q, float.hex(2 * np.pi * g.uniform())] for q in range(1, n + 1)]
        for c in range(1, n + 1):
            for t in range(1, n + 1):
                if c != t:
                    ops.append(["crx", c, t, float.hex(2 * np.pi * g.uniform())])
                    ops.append(["cnot", t, c])
        return ops
    if name == "n2_bell":
        return [["h", 1], ["cnot", 1, 2]]
    if name.startswith("c1_reservoir"):
        return MG.c1_ops(n, 3)
    return MG.make_ops(n, nops, seed, kinds)


def main():
    it = load_reference()
    out = {"generator": "tests/golden/make_ref_exec_golden.py: the reference's own src/Quantum/QSim.jl and QInfo.jl executed by "
                        "oracle/jl_subset.py (a Julia-subset interpreter over numpy)", "cases": []}
    for name, n, m, nops, seed, kinds in MG.CASES:
        ops = build_ops(name, n, nops, seed, kinds)
        s = ref_call(it, "statevector", n, m)
        for op in ops:
            apply_op(it, s, op)
        case = {"name": name, "n": n, "m": m, "ops": ops, "amps": MG.hxc(s), "mp": MG.hx(ref_call(it, "mp", s)),
                "measure_z": [MG.hx(ref_call(it, "measure_z", s, t)) for t in range(1, n + 1)],
                "measure_x": [MG.hx(ref_call(it, "measure_x", s, t)) for t in range(1, n + 1)],
                "measure_y": [MG.hx(ref_call(it, "measure_y", s, t)) for t in range(1, n + 1)],
                "reduced": []}
        rho = ref_call(it, "statevector_to_density_matrix", s) if n <= 8 else None
        for k in range(1, (min(3, n - 1) if n <= 8 else 0) + 1):
            dB = 1 << k; dA = (1 << n) // dB
            rB = ref_call(it, "partial_trace", rho, dA, dB, 1)          # keep the low k qubits (trace out A)
            rA = ref_call(it, "partial_trace", rho, dB, dA, 2)          # keep the high k qubits
            case["reduced"].append({"k": k, "rho_B_low": MG.hxc(rB), "rho_A_high": MG.hxc(rA),
                                    "S_B": float.hex(float(ref_call(it, "von_neumann_entropy", rB))),
                                    "S_A": float.hex(float(ref_call(it, "von_neumann_entropy", rA)))})
        out["cases"].append(case)
        print("case", name, "done", flush=True)
    path = os.path.join(HERE, "ref_exec_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes,", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
