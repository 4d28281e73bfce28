#!/bin/bash
# tools/run_dist_sweep.sh N OUT_PREFIX — the sharded bench at N GPUs over the exchange settings (one torchrun per setting, each under
# its own timeout so that a hang cannot hold the box): exchange between passes, then overlapped with 16 / 24 / 32 SMs and 4 / 8 slabs.
N=$1; P=$2; PORT=29500
run() {  # name, extra args
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/${P}_${name}.json 2> gpurun_out/${P}_${name}.err
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
  tail -n 3 gpurun_out/${P}_${name}.err | cut -c1-300
  PORT=$((PORT+1))
}
run nooverlap --overlap 0 --no-alt
run ov_x16_s2 --overlap 1 --xsm 16 --slab-bits 2
run ov_x24_s2 --overlap 1 --xsm 24 --slab-bits 2 --no-alt
run ov_x16_s3 --overlap 1 --xsm 16 --slab-bits 3 --no-alt
run ov_x32_s3 --overlap 1 --xsm 32 --slab-bits 3 --no-alt
run ov_x16_s2_un4 --overlap 1 --xsm 16 --slab-bits 2 --xun 4 --no-alt
