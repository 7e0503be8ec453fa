#!/bin/bash
# round 2, GPU call 35 (one B200): bench.py after its last edit (kernel label), short workload
mkdir -p gpurun_out
timeout 200 python bench.py --snps 50000 --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2> gpurun_out/r2c35_bench.err | tail -1 > gpurun_out/r2c35_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2c35_bench.json')); print(d['us_per_svi_iteration'], d['value'], d['roofline']['kernel'], d['roofline']['traffic'], d['roofline']['fp64_pipe_pct'], d['parity_check']['ok'])"; tail -3 gpurun_out/r2c35_bench.err
