#!/bin/bash
# tools/gpu_round2_final.sh A|B PREFIX — the two GPU calls of the last session of round 2 (6.5 GPU-minutes left):
#   A: the folded readout (tests + A/B timing of the 33-qubit reservoir step)
#   B: whole GPU suite, smoke, the bench line, first-sight timings
MODE=$1; P=$2
mkdir -p gpurun_out
if [ "$MODE" = "A" ]; then
  timeout 170 python -m pytest tests/test_gpu_zfold.py -x -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/${P}_zfold_pytest.log; tail -6 gpurun_out/${P}_zfold_pytest.log
  timeout 150 python tools/bench_zfold.py --out gpurun_out/${P}_zfold_c4_33q.json > gpurun_out/${P}_zfold_c4_33q.log 2>&1; echo "zfold bench rc=$?"; tail -c 1500 gpurun_out/${P}_zfold_c4_33q.log
else
  timeout 400 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/${P}_pytest.log; tail -4 gpurun_out/${P}_pytest.log
  timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -3
  timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${P}_bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "roofline", "clocks", "e2e", "c4_base", "parity", "c2_batched"):
        print(k, json.dumps(d.get(k))[:400])
except Exception as e:
    print("no bench line:", e)
PY
  tail -3 gpurun_out/${P}_bench.err
  timeout 200 python tools/bench_first_sight.py --out gpurun_out/${P}_first_sight.json > gpurun_out/${P}_first_sight.log 2>&1; echo "first sight rc=$?"; tail -c 1200 gpurun_out/${P}_first_sight.log
  # ncu --set full of one reservoir step with the folded readout and one without (30 qubits: passes of ~6 ms), after the plain runs above
  timeout 150 ncu --set full --clock-control none --kernel-name-base mangled -k regex:k_fused_tma -s 56 -c 16 -f -o gpurun_out/${P}_zfold \
      python tools/bench_zfold.py --qubits 30 --steps 2 --rounds 1 > gpurun_out/${P}_zfold_ncu.log 2>&1; echo "ncu rc=$?"
  timeout 60 python tools/ncu_summary.py gpurun_out/${P}_zfold.ncu-rep > gpurun_out/${P}_ncu_k_fused_tma_zfold_30q.txt 2>&1; head -c 1500 gpurun_out/${P}_ncu_k_fused_tma_zfold_30q.txt
fi
