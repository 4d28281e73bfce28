#!/bin/bash
# tools/run_dist_sweep2.sh N OUT_PREFIX "name:args" ... — the sharded bench at N GPUs, one torchrun per named setting
N=$1; P=$2; shift; shift; PORT=29600
for spec in "$@"; do
  name=${spec%%:*}; args=${spec#*:}
  timeout 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 3 --warmup 3 $args > gpurun_out/${P}_${name}.json 2> gpurun_out/${P}_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${P}_${name}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "passes_per_step", "overlapped_remaps_per_step", "weak_scaling_eff", "norm2_ok")},
          "link", {k: d["link"].get(k) for k in ("achieved", "frac", "ms_remap_per_step")}, "hbm", round(d["roofline"]["frac"], 3),
          "parity", d.get("parity", {}).get("ok"), d.get("parity", {}).get("max_abs_err"), d.get("parity", {}).get("overlapped_remaps"))
except Exception as e:
    print("no line:", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/${P}_${name}.err | tail -n 3 | cut -c1-300
  PORT=$((PORT+1))
done
