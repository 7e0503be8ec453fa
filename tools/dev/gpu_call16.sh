#!/bin/bash
# round 2, GPU call 16 (one B200): ncu --set full of the single-GPU instantiation (t7) at 100 000 individuals
mkdir -p gpurun_out
O=gpurun_out/r2c16
export TSGPU_LIB=$PWD/terastructure_b200/lib/libtsgpu_t7.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persist -s 3 -c 1 -f -o ${O}_prof_n100000 \
   python bench.py --individuals 100000 --steps 1 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_full.log 2>&1
tail -2 ${O}_ncu_full.log
