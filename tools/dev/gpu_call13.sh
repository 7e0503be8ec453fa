#!/bin/bash
# round 2, GPU call 13 (one B200): integer convergence decision (t2) and staggered polls in flight (t3-t5: depth 2-4)
mkdir -p gpurun_out
O=gpurun_out/r2c13
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh r2base t2 t3 t4 t5
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- t2 t3 t4
for v in t2 t3 t4; do
echo "== trace 100000 $v" >> ${O}_trace.txt; TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_$v.so timeout 200 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
done
grep -E "==|per SNP|mean/round|gamma phase" ${O}_trace.txt
