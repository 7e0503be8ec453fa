#!/bin/bash
# round 2, GPU call 15 (one B200): single-GPU instantiation without exchange code and without trace stamps (t7)
mkdir -p gpurun_out
O=gpurun_out/r2c15
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh r2base t7
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- r2base t7
echo "== trace 100000 trace-lib" >> ${O}_trace.txt; TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_trace.so timeout 200 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
grep -E "==|per SNP|mean/round|gamma phase|round 3" ${O}_trace.txt
