"""tools/bench_first_sight.py — what the FIRST application of a circuit costs (config C3 by default), three ways, each in a
fresh process (the kernel cache is process-wide):

  default      nothing prepared: step 1 runs through the interpreter kernel, step 2 compiles (QM_OPT_JIT 2), step 3 is steady state
  precompile   qm_precompile_circuit first (timed), then step 1 already runs compiled
  disk         the same with QM_JIT_CACHE_DIR pointing at the directory the `precompile` run filled: nothing is assembled

    python tools/bench_first_sight.py [--qubits 30] [--depth 200] [--out gpurun_out/first_sight.json]

Prints one JSON object; every mode checks <Z> of qubit 1 against the `default` mode's steady-state step (same circuit)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(mode, n, depth):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.load_package()
    L = pkg._lib
    gates = pkg.circuits.c3_circuit(n=n, depth=depth, seed=3001)
    st = pkg.B200State(n, 0)
    st.sync()
    out = {"mode": mode, "n_qubits": n, "gates": int(len(gates))}
    if mode != "default":
        t0 = time.perf_counter()
        out["kernels_compiled"] = st.precompile_circuit(gates)
        out["precompile_s"] = time.perf_counter() - t0
    steps = []
    for i in range(3 if mode == "default" else 2):
        st.reset(0)
        st.counters_reset()
        t0 = time.perf_counter()
        st.apply_circuit(gates)
        z = st.expect_z_all()
        dt = time.perf_counter() - t0
        c = st.counters()
        steps.append({"wall_s": dt, "passes": int(c["passes"]), "compiled_passes": int(c["compiled_passes"]),
                      "gates_per_s": len(gates) / dt * 2.0 ** (n - 30)})
    out["steps"] = steps
    out["z0"] = [float(z[0][0]), float(z[0][1])]
    out["jit"] = L.jit_stats()
    out["disk"] = L.jit_cache_stats()
    st.close()
    print("FIRST_SIGHT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=200)
    ap.add_argument("--out", default="")
    ap.add_argument("--child", default="")
    args = ap.parse_args()
    if args.child:
        return child(args.child, args.qubits, args.depth)
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for mode in ("default", "precompile", "disk"):
            env = dict(os.environ)
            env.pop("QM_JIT_CACHE_DIR", None)
            if mode != "default":
                env["QM_JIT_CACHE_DIR"] = os.path.join(d, "kernels")
            t0 = time.perf_counter()
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", mode, "--qubits", str(args.qubits),
                                "--depth", str(args.depth)], env=env, capture_output=True, text=True, timeout=900)
            line = [x for x in p.stdout.splitlines() if x.startswith("FIRST_SIGHT ")]
            if p.returncode != 0 or not line:
                res[mode] = {"error": (p.stdout + p.stderr)[-800:]}
                continue
            res[mode] = json.loads(line[0][len("FIRST_SIGHT "):])
            res[mode]["process_wall_s"] = time.perf_counter() - t0
    ok = all("z0" in res.get(m, {}) for m in res)
    if ok:
        ref = res["default"]["z0"]
        res["readouts_agree"] = all(abs(res[m]["z0"][0] - ref[0]) < 1e-12 and abs(res[m]["z0"][1] - ref[1]) < 1e-12 for m in ("precompile", "disk"))
    s = json.dumps(res)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")
    return 0 if ok and res.get("readouts_agree") else 1


if __name__ == "__main__":
    sys.exit(main())
