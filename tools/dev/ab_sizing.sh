#!/bin/bash
# Shard-sizing A/B on one box (see profiles/r1_summary.md): us per SVI iteration for different shard
# sizes / individuals-per-thread choices.
run() { python bench.py --snps 50000 --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('us/SVI-iter %.2f' % d['us_per_svi_iteration'])"; }
for rep in 1 2; do
  echo default-125K; run --individuals 125000
  echo i3_288-125K; TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_i3_288.so run --individuals 125000
done
echo default-100K; run
echo default-80K; run --individuals 80000
echo 80K-ipt3; TSGPU_IPT=3 run --individuals 80000
echo default-60K; run --individuals 60000
echo 60K-ipt1; TSGPU_IPT=1 run --individuals 60000
