#!/bin/bash
# round 2, GPU call 14 (one B200): control block with the table-driven b and the convergence sum interleaved (t6)
mkdir -p gpurun_out
O=gpurun_out/r2c14
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh r2base t6
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- r2base t6
for v in t6; do
echo "== trace 100000 $v" >> ${O}_trace.txt; TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_$v.so timeout 200 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
done
grep -E "==|per SNP|mean/round|gamma phase|round 3" ${O}_trace.txt
