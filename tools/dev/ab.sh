#!/bin/bash
# A/B of library variants on one box: TSGPU_LIB=<variant> bench.py on a shortened workload.
# usage: tools/dev/ab.sh name1 name2 ...   (terastructure_b200/lib/libtsgpu_<name>.so)
out=gpurun_out/ab.jsonl; : > $out
for rep in 1 2; do
for v in "$@"; do
  TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_$v.so python bench.py --snps 50000 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v rep$rep us/SVI-iter %.2f e2e %.3e' % (d['us_per_svi_iteration'], d['e2e']['value']))" | tee -a $out
done
done
