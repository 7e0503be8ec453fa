#!/bin/bash
# round 2, GPU call 31 (one B200): weights through the exponent field (t30), + gamma rows prefetched into L1 (t31), unequal
# halves (TSGPU_ASYM) on t31; the unequal-halves oracle tests
mkdir -p gpurun_out
O=gpurun_out/r2c31
one() {  # $1 = label, $2 = lib suffix, $3 = TSGPU_ASYM or "-", rest = bench args
  lbl=$1; lib=$PWD/terastructure_b200/lib/libtsgpu$2.so; asym=$3; shift 3
  if [ "$asym" == "-" ]; then unset TSGPU_ASYM; else export TSGPU_ASYM=$asym; fi
  TSGPU_LIB=$lib timeout 300 python bench.py --snps 50000 --steps 5 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$lbl $*: us/SVI-iter %.2f e2e %.3e parity %s %s' % (d['us_per_svi_iteration'], d['e2e']['value'], d['parity_check']['ok'], d['roofline']['kernel'][:28]))" | tee -a ${O}_ab.txt
  unset TSGPU_ASYM
}
for rep in 1 2; do
one product "" -
one t30 _t30 -
one t31 _t31 -
one t31_asym24 _t31 2,4
one t31_asym34 _t31 3,4
done
one product "" - --individuals 125000
one t31 _t31 - --individuals 125000
one t31_asym34 _t31 3,4 --individuals 125000
TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_t31.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "unequal_halves and not 20" 2>&1 | tail -3
