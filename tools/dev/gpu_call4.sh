#!/bin/bash
# round 2, GPU call 4 (one B200): full GPU suite (tiered kernel, row ring), A/B of the fence placement,
# large shards and the K sweep's upper end, ncu capture of the tiered kernel at 1M individuals
mkdir -p gpurun_out
O=gpurun_out/r2c4
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh base nofence
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$*: us/SVI-iter %.2f genotypes/s %.3e frac %.3f parity %s kernel %s' % (d['us_per_svi_iteration'], d['value'], d['roofline']['frac'], d['parity_check']['ok'], d['roofline']['kernel'][:60]))" | tee -a ${O}_sweep.txt; }
run --individuals 200000 --snps 20000
run --individuals 400000 --snps 20000
run --individuals 1000000 --snps 20000
TSGPU_TIER_J=0 run --individuals 400000 --snps 20000
for k in 12 13 16 20 24; do run --k $k --individuals 100000 --snps 100000; done
run --k 6 --individuals 10000 --snps 100000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persist -s 3 -c 1 -f -o ${O}_prof_n1000000 \
   python bench.py --individuals 1000000 --steps 1 --warmup 3 --batch 20 --snps 20000 --no-extras --no-cpu-baseline > ${O}_ncu_full_1M.log 2>&1
tail -2 ${O}_ncu_full_1M.log
