#!/bin/bash
# round 2, GPU call 17 (one B200): full GPU suite on the rebuilt product (table-driven control path, single-/multi-GPU
# instantiations, no trace stamps, pruned exchange modes), A/B of the pinned fixed-point constants (t8 against t7)
mkdir -p gpurun_out
O=gpurun_out/r2c17
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
cp terastructure_b200/lib/libtsgpu.so terastructure_b200/lib/libtsgpu_new.so
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh t7 t8 new
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- t7 t8
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$*: us/SVI-iter %.2f genotypes/s %.3e frac %.3f parity %s kernel %s' % (d['us_per_svi_iteration'], d['value'], d['roofline']['frac'], d['parity_check']['ok'], d['roofline']['kernel'][:60]))" | tee -a ${O}_sweep.txt; }
run --individuals 400000 --snps 20000
run --individuals 1000000 --snps 20000
for k in 6 16 20; do run --k $k --individuals 100000 --snps 100000; done
