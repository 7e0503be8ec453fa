#!/bin/bash
# round 2, GPU call 1 (one B200): the whole GPU suite incl. the new oracle parity tests, the same
# parity tests on the all-options build, A/B of the kernel variants, phase traces, barrier micro-benchmarks
mkdir -p gpurun_out
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > ${O}_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_tests.log 2>&1
tail -5 ${O}_tests.log
( TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_all.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "forced or bench_geometry or config1 or single_round or staged or trajectory" ) > ${O}_tests_all.log 2>&1
tail -3 ${O}_tests_all.log
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh base sg n1 pre all
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- base all
for v in "" _all; do
  for n in 100000 125000; do
    echo "== trace lib$v n=$n" >> ${O}_trace.txt
    TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu$v.so timeout 120 python tools/dev/trace_persist.py $n >> ${O}_trace.txt 2>&1
  done
done
timeout 120 tools/dev/ubench_xchg 1 > ${O}_xchg1.txt 2>&1
timeout 120 tools/dev/ubench5 > ${O}_ubench5.txt 2>&1
cat ${O}_trace.txt | grep -E "==|per SNP|mean/round|gamma phase"
cat ${O}_xchg1.txt ${O}_ubench5.txt
