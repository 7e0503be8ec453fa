#!/bin/bash
# round 2, GPU call 6 (one B200): tier-loop tuning A/B at 400K and 1M individuals, phase trace of the tiered kernel
mkdir -p gpurun_out
O=gpurun_out/r2c6
AB_OUT=${O}_ab400.jsonl tools/dev/ab.sh --individuals 400000 --snps 20000 -- base u1 u3 i1u3 i1u4
AB_OUT=${O}_ab1m.jsonl tools/dev/ab.sh --individuals 1000000 --snps 20000 -- base u3 i1u4
AB_OUT=${O}_ab100.jsonl tools/dev/ab.sh base
for n in 400000 200000; do
  echo "== trace n=$n" >> ${O}_trace.txt
  timeout 200 python tools/dev/trace_persist.py $n >> ${O}_trace.txt 2>&1
done
grep -E "==|per SNP|mean/round|gamma phase|round 0|round 1:|round 9" ${O}_trace.txt
