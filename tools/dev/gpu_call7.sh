#!/bin/bash
# round 2, GPU call 7 (one B200): full suite (staged cross-checks through the test-only library), bench, sweep
mkdir -p gpurun_out
O=gpurun_out/r2c7
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh base
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$*: us/SVI-iter %.2f genotypes/s %.3e frac %.3f parity %s kernel %s' % (d['us_per_svi_iteration'], d['value'], d['roofline']['frac'], d['parity_check']['ok'], d['roofline']['kernel'][:60]))" | tee -a ${O}_sweep.txt; }
run --individuals 125000 --snps 50000
run --individuals 200000 --snps 20000
run --individuals 400000 --snps 20000
run --individuals 1000000 --snps 20000
for k in 2 4 6 8 10 12 16 20; do run --k $k --individuals 100000 --snps 100000; done
echo "== trace 400000" >> ${O}_trace.txt; timeout 200 python tools/dev/trace_persist.py 400000 >> ${O}_trace.txt 2>&1
echo "== trace 100000" >> ${O}_trace.txt; timeout 200 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
grep -E "==|per SNP|mean/round|gamma phase" ${O}_trace.txt
