#!/bin/bash
# round 2, GPU call 12 (one B200): table-driven control path (ts_ftab.cuh) -- full GPU suite, A/B against the
# library of commit 83a7ee3 (libtsgpu_r2base.so), phase trace
mkdir -p gpurun_out
O=gpurun_out/r2c12
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
cp terastructure_b200/lib/libtsgpu.so terastructure_b200/lib/libtsgpu_new.so
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh r2base new
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- r2base new
echo "== trace 100000 new" >> ${O}_trace.txt; timeout 200 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
echo "== trace 100000 r2base" >> ${O}_trace.txt; TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_r2base.so timeout 200 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
grep -E "==|per SNP|mean/round|gamma phase" ${O}_trace.txt
