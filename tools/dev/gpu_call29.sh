#!/bin/bash
# round 2, GPU calls 29a/29b (one B200): ONE ncu --set full capture per call (a capture with source is 37 MB; 64 MiB travel back)
# usage: gpu_call29.sh <individuals> [launch-list]
mkdir -p gpurun_out
O=gpurun_out/r2c29
n=$1
if [ "$2" == "launches" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv \
   python bench.py --steps 2 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_launch.log 2>&1
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persist -s 3 -c 1 -f -o ${O}_prof_n$n \
   python bench.py --individuals $n --steps 1 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_full_$n.log 2>&1
tail -1 ${O}_ncu_full_$n.log
du -sh gpurun_out
