#!/bin/bash
# round 2, GPU call 1b (one B200): the round-1 library against the current tree on the SAME box
mkdir -p gpurun_out
O=gpurun_out/r2c1b
for rep in 1 2; do
  ( cd tools/dev/_r1tree && timeout 300 python bench.py --snps 50000 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('r1tree rep$rep us/SVI-iter %.2f' % d['us_per_svi_iteration'])" ) | tee -a ${O}_ab.jsonl
done
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh base nofence pre prenf sg all
( cd tools/dev/_r1tree && timeout 120 python tools/dev/trace_persist.py 100000 ) > ${O}_trace_r1.txt 2>&1
for v in "" _nofence _prenf; do
  echo "== trace lib$v" >> ${O}_trace.txt
  TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu$v.so timeout 120 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
done
grep -E "per SNP|mean/round|gamma phase" ${O}_trace_r1.txt; grep -E "==|per SNP|mean/round|gamma phase" ${O}_trace.txt
