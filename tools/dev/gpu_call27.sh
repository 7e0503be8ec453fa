#!/bin/bash
# round 2, GPU call 27 (one B200): verification of the final kernels -- smoke, full GPU suite, the driver's bench commands
# (ours and the reference arm), ncu launch list, --set full captures of the benchmarked instantiations, sweeps, trace
mkdir -p gpurun_out
O=gpurun_out/r2c27
( time python -c "import __graft_entry__ as g; g.smoke()" ) > ${O}_smoke.log 2>&1; tail -3 ${O}_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_tests.log 2>&1; tail -4 ${O}_tests.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; tail -c 3000 ${O}_bench_1gpu.json; tail -3 ${O}_bench_1gpu.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > ${O}_bench_ref.json 2> ${O}_bench_ref.err; tail -c 1500 ${O}_bench_ref.json; tail -3 ${O}_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv \
   python bench.py --steps 2 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_launch.log 2>&1
for n in 100000 125000; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persist -s 3 -c 1 -f -o ${O}_prof_n$n \
     python bench.py --individuals $n --steps 1 --warmup 3 --batch 50 --snps 50000 --no-extras --no-cpu-baseline > ${O}_ncu_full_$n.log 2>&1
  tail -1 ${O}_ncu_full_$n.log
done
timeout 600 ncu --set full --clock-control none -k regex:k_persist -s 3 -c 1 -f -o ${O}_prof_n1000000 \
   python bench.py --individuals 1000000 --steps 1 --warmup 3 --batch 20 --snps 20000 --no-extras --no-cpu-baseline > ${O}_ncu_full_1000000.log 2>&1
tail -1 ${O}_ncu_full_1000000.log
run() { timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras "$@" 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$*: us/SVI-iter %.2f genotypes/s %.3e frac %.3f parity %s kernel %s' % (d['us_per_svi_iteration'], d['value'], d['roofline']['frac'], d['parity_check']['ok'], d['roofline']['kernel'][:60]))" | tee -a ${O}_sweep.txt; }
run --individuals 125000 --snps 50000
run --individuals 200000 --snps 20000
run --individuals 400000 --snps 20000
run --individuals 1000000 --snps 20000
for k in 2 4 6 8 10 12 16 20; do run --k $k --individuals 100000 --snps 100000; done
run --k 6 --individuals 10000 --snps 100000
echo "== trace 100000" >> ${O}_trace.txt; timeout 200 python tools/dev/trace_persist.py 100000 >> ${O}_trace.txt 2>&1
echo "== trace 400000" >> ${O}_trace.txt; timeout 200 python tools/dev/trace_persist.py 400000 >> ${O}_trace.txt 2>&1
grep -E "==|per SNP|mean/round|gamma phase" ${O}_trace.txt
