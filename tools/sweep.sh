#!/bin/bash
# BASELINE configs[1] (10K x 100K, K=6) and configs[4] (K sweep at 100K x 100K) on one B200.
out=gpurun_out/sweep_r1.jsonl; : > $out
python bench.py --individuals 10000 --snps 100000 --k 6 --steps 3 --warmup 3 --no-cpu-baseline >> $out 2>/dev/null
for k in 2 4 6 8 10 12 16 20; do
  python bench.py --individuals 100000 --snps 100000 --k $k --steps 3 --warmup 3 --no-cpu-baseline >> $out 2>/dev/null
done
python - <<'PY'
import json
for l in open("gpurun_out/sweep_r1.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"], "| us/SVI-iter %.1f | genotypes/s %.3e | roofline frac %.3f" % (d["us_per_svi_iteration"], d["value"], d["roofline"]["frac"]))
PY
