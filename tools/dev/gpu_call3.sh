#!/bin/bash
# round 2, GPU call 3 (one B200): full GPU suite on the current tree (row ring, one-Newton E-step),
# the default bench line, ncu launch list and --set full captures of the two benchmarked instantiations
mkdir -p gpurun_out
O=gpurun_out/r2c3
( time timeout 900 python -m pytest tests -m gpu -x -q ) > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh base nofence
timeout 600 python bench.py --steps 10 --warmup 3 > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; tail -c 1500 ${O}_bench_1gpu.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv \
   python bench.py --steps 2 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_launch.log 2>&1
for n in 100000 125000; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persist -s 3 -c 1 -f -o ${O}_prof_n$n \
     python bench.py --individuals $n --steps 1 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_full_$n.log 2>&1
  tail -2 ${O}_ncu_full_$n.log
done
ls -la gpurun_out/ | grep r2c3
