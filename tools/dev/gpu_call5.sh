#!/bin/bash
# round 2, GPU call 5 (one B200): full suite after the tier-loop / fence-window changes, A/B, large shards, traces
mkdir -p gpurun_out
O=gpurun_out/r2c5
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh base nofence
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$*: us/SVI-iter %.2f genotypes/s %.3e frac %.3f parity %s kernel %s' % (d['us_per_svi_iteration'], d['value'], d['roofline']['frac'], d['parity_check']['ok'], d['roofline']['kernel'][:60]))" | tee -a ${O}_sweep.txt; }
run --individuals 125000 --snps 50000
run --individuals 200000 --snps 20000
run --individuals 400000 --snps 20000
run --individuals 1000000 --snps 20000
TSGPU_IPT=0 run --k 20 --individuals 100000 --snps 100000
run --k 20 --individuals 100000 --snps 100000
for v in "" _nofence; do
  echo "== trace lib$v" >> ${O}_trace.txt
  TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu$v.so timeout 120 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
done
grep -E "==|per SNP|mean/round|gamma phase|round 0|round 1:|round 9" ${O}_trace.txt
timeout 900 python tools/dev/big_bed_test.py --out ${O}_big_bed.json 2>&1 | tail -30
