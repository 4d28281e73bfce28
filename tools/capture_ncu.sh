#!/bin/bash
# tools/capture_ncu.sh PREFIX — the ncu evidence for the CURRENT binary (run on the GPU box, one GPU): the plain command first
# (must exit 0), then the launch list, then one --set full capture of three default fused passes, summarised in place.
P=$1
CMD="python bench.py --steps 1 --warmup 3 --depth 40 --no-cpu --no-alt"
$CMD > gpurun_out/${P}_plain_depth40.json 2> gpurun_out/${P}_plain_depth40.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${P}_launches_bench_depth40.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_tma -s 150 -c 3 -f -o gpurun_out/${P}_fused $CMD > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${P}_fused.ncu-rep > gpurun_out/${P}_ncu_k_fused_tma_default_budget80.txt 2>&1
ncu -i gpurun_out/${P}_fused.ncu-rep --page source --csv > gpurun_out/${P}_fused_source.csv 2>/dev/null
(python tools/ncu_regions_jit.py gpurun_out/${P}_fused_source.csv || python tools/ncu_regions.py gpurun_out/${P}_fused_source.csv) >> gpurun_out/${P}_ncu_k_fused_tma_default_budget80.txt 2>&1
rm -f gpurun_out/${P}_fused_source.csv
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${P}_launches_bench_depth40.csv")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
SC = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= vi: continue
    try: v = float(r[vi].replace(",", "")) * SC.get(r[ui].strip(), 1.0)
    except ValueError: continue
    k = r[ki].split("(")[0]; t[k][0] += 1; t[k][1] += v
tot = sum(v[1] for v in t.values())
for k, (c, v) in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print("%-60s launches %5d  total %10.3f ms  share %6.2f %%  mean %8.3f ms" % (k[:60], c, v / 1e6, 100 * v / tot, v / c / 1e6))
PY
