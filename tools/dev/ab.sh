#!/bin/bash
# A/B of library variants on one box: TSGPU_LIB=<variant> bench.py on a shortened workload.
# usage: tools/dev/ab.sh [bench args --] name1 name2 ...   (terastructure_b200/lib/libtsgpu<name>.so; "" = the product)
out=${AB_OUT:-gpurun_out/ab.jsonl}
extra=()
while [ $# -gt 0 ] && [ "$1" != "--" ] && [[ "$1" == --* ]]; do extra+=("$1" "$2"); shift 2; done
[ "$1" == "--" ] && shift
for rep in 1 2; do
for v in "$@"; do
  [ "$v" == "base" ] && lib=$PWD/terastructure_b200/lib/libtsgpu.so || lib=$PWD/terastructure_b200/lib/libtsgpu_$v.so
  TSGPU_LIB=$lib timeout 300 python bench.py --snps 50000 --steps 5 --warmup 3 --no-cpu-baseline --no-extras "${extra[@]}" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$v ${extra[*]} rep$rep us/SVI-iter %.2f e2e %.3e parity %s' % (d['us_per_svi_iteration'], d['e2e']['value'], d['parity_check']['ok']))" | tee -a $out
done
done
