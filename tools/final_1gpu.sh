#!/bin/bash
# tools/final_1gpu.sh PREFIX — the one-GPU evidence for the current binary: GPU test suite, smoke, the bench line, the
# reference arm, the ncu launch list + one --set full capture of the default fused pass, and the readout kernels' rooflines.
P=$1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${P}_pytest.log; tail -3 gpurun_out/${P}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${P}_bench_reference_arm.json 2> gpurun_out/${P}_bench_reference_arm.err; echo "ref rc=$?"
timeout 600 bash tools/capture_ncu.sh ${P} 2>&1 | tail -8
timeout 300 python tools/bench_readouts.py --qubits 28 > gpurun_out/${P}_readouts.jsonl 2> gpurun_out/${P}_readouts.err; echo "readouts rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${P}_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "roofline", "clocks", "e2e", "deep_fusion", "c4_base", "cpu_baseline", "parity", "c2_batched", "c1", "c5"):
    print(k, json.dumps(d.get(k))[:600])
PY
cat gpurun_out/${P}_readouts.jsonl | cut -c1-400; tail -3 gpurun_out/${P}_bench.err gpurun_out/${P}_readouts.err
head -30 gpurun_out/${P}_ncu_k_fused_tma_default_budget80.txt
