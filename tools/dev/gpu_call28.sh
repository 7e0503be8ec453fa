#!/bin/bash
# round 2, GPU call 28 (one B200): call 27 again without the third --set full capture (its outputs exceeded the 64 MiB that
# travel back): smoke, full GPU suite, the driver's bench command, ncu launch list, two --set full captures
mkdir -p gpurun_out
O=gpurun_out/r2c28
( time python -c "import __graft_entry__ as g; g.smoke()" ) > ${O}_smoke.log 2>&1; grep -E "smoke ok|Error|error" ${O}_smoke.log | tail -3
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_tests.log 2>&1; grep -E "passed|failed" ${O}_tests.log | tail -2
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; tail -c 3500 ${O}_bench_1gpu.json; tail -3 ${O}_bench_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv \
   python bench.py --steps 2 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_launch.log 2>&1
for n in 100000 125000; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persist -s 3 -c 1 -f -o ${O}_prof_n$n \
     python bench.py --individuals $n --steps 1 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_full_$n.log 2>&1
  tail -1 ${O}_ncu_full_$n.log
done
ls -la gpurun_out | tail -12
