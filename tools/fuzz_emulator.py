"""tools/fuzz_emulator.py — CPU fuzzing of the planner + encoder + code generator (no GPU).

Random circuits with a wider gate mix than the unit tests (general unitaries, every named rotation, controlled general
unitaries, swaps, long stretches on few qubits, repeated gates, identity-like angles, angles at multiples of pi/2 where the
lifting forms change class) are planned and encoded by the engine's own planner/encoder and walked by the host emulator
(tests/native: record walk = the interpreter kernel's view, generated PTX text = the compiled kernels' view), unsharded and
sharded over 2/4/8 emulated ranks, at random planner options.  Every result is compared with the CPU oracle at 1e-12.

    python tools/fuzz_emulator.py --minutes 20 --seed 1 [--out gpurun_out/fuzz.log]

A failure prints the reproducing parameters and saves the gate list (circuits.save_gates) next to the log."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

SPECIAL = [0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi, 2 * np.pi, -2 * np.pi, 3 * np.pi / 2, 4 * np.pi, 1e-9, -1e-9, np.pi - 1e-9,
           np.pi / 4, 1e-15, np.pi + 1e-15]


def angle(rng):
    if rng.random() < 0.25:
        return float(SPECIAL[rng.integers(0, len(SPECIAL))])
    return float(rng.uniform(-2 * np.pi, 2 * np.pi))


def random_unitary(rng):
    a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def fuzz_circuit(pkg, n, ngates, rng):
    gl = pkg.circuits.GateList(n)
    Q = pkg.qsim
    style = rng.integers(0, 4)       # 0 uniform, 1 local stretches, 2 few hot qubits, 3 layered
    hot = [int(x) for x in rng.choice(np.arange(1, n + 1), size=min(n, int(rng.integers(2, 6))), replace=False)]
    t_prev = 1
    while len(gl) < ngates:
        if style == 1 and rng.random() < 0.8:
            t = int(np.clip(t_prev + rng.integers(-2, 3), 1, n))
        elif style == 2 and rng.random() < 0.7:
            t = hot[rng.integers(0, len(hot))]
        else:
            t = int(rng.integers(1, n + 1))
        t_prev = t
        c = int(rng.integers(1, n + 1))
        if style == 1 and rng.random() < 0.7:
            c = int(np.clip(t + rng.choice([-3, -2, -1, 1, 2, 3]), 1, n))
        k = int(rng.integers(0, 16))
        th = angle(rng)
        if k == 0: gl.h(t)
        elif k == 1: gl.rx(t, th)
        elif k == 2: gl.ry(t, th)
        elif k == 3: gl.rz(t, th)
        elif k == 4: gl.x(t)
        elif k == 5: gl.z(t)
        elif k == 6: gl.y(t)
        elif k == 7: gl.u2(t, random_unitary(rng))
        elif k == 8: gl.u2(t, np.diag([np.exp(1j * angle(rng)), np.exp(1j * angle(rng))]))
        elif c != t:
            if k == 9: gl.cnot(c, t)
            elif k == 10: gl.crz(c, t, th)
            elif k == 11: gl.crx(c, t, th)
            elif k == 12: gl.cry(c, t, th)
            elif k == 13: gl.controlled(c, t, random_unitary(rng))
            elif k == 14: gl.swap(c, t)
            else: gl.controlled(c, t, Q.Z)
    return gl.array()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-qubits", type=int, default=17)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    pkg = ge.load_package()
    import emu
    from oracle import oracle as O
    rng = np.random.default_rng(args.seed)
    log = open(args.out, "a") if args.out else None

    def say(s):
        print(s, flush=True)
        if log:
            log.write(s + "\n"); log.flush()

    t_end = time.time() + 60.0 * args.minutes
    it, fails = 0, 0
    worst = 0.0
    while time.time() < t_end:
        it += 1
        sharded = rng.random() < 0.4
        g = int(rng.integers(1, 4)) if sharded else 0
        n = int(rng.integers(12 + g, args.max_qubits + 1))
        ngates = int(rng.choice([5, 20, 60, 150, 300, 500]))
        jit = int(rng.integers(0, 2))
        low = int(rng.choice([3, 4, 5]))
        cost = float(rng.choice([0.0, 9.0, 24.0, 40.0, 80.0, 120.0, 160.0, 250.0]))
        defer = int(rng.integers(0, 2))
        seed = int(rng.integers(0, 2 ** 31))
        crng = np.random.default_rng(seed)
        try:
            gl = fuzz_circuit(pkg, n, ngates, crng)
        except AttributeError as e:
            say("circuit builder lacks a gate: %r" % (e,))
            raise
        psi = O.random_state(n, seed % 1000)
        ref = O.apply_circuit(psi.copy(), gl)
        desc = "it=%d n=%d g=%d gates=%d jit=%d low=%d cost=%g defer=%d seed=%d" % (it, n, g, len(gl), jit, low, cost, defer, seed)
        try:
            emu.set_jit(jit)
            if sharded:
                emu.set_defer(defer)
                got, npass, nrem = emu.plan_emulate_sharded(n, g, gl, psi, 12, low, cost)
            else:
                got, npass = emu.plan_emulate(n, gl, psi, 12, low, cost)
                nrem = 0
            err = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
        except Exception as e:      # noqa: BLE001
            err = float("inf")
            say("EXC  %s : %r" % (desc, e))
        worst = max(worst, err if np.isfinite(err) else 0.0)
        if not err < 1e-12:
            fails += 1
            path = (args.out or "fuzz") + ".fail%d.gates" % fails
            pkg.circuits.save_gates(path, n, gl)
            say("FAIL %s err=%.3e  -> %s" % (desc, err, path))
        elif it % 25 == 0:
            say("ok   %s passes=%d remaps=%d err=%.1e  (worst so far %.1e, fails %d)" % (desc, npass, nrem, err, worst, fails))
    emu.set_jit(0); emu.set_defer(1)
    say("DONE iterations=%d fails=%d worst_err=%.2e seed=%d" % (it, fails, worst, args.seed))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
