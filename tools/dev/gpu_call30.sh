#!/bin/bash
# round 2, GPU call 30 (eight B200s): the default line of bench.py --gpus 8 (1M x 1M, convergence run) with the final kernels
mkdir -p gpurun_out
O=gpurun_out/r2c30
TSGPU_TIMEOUT_S=30 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus 8 --steps 10 --warmup 3 2> ${O}_bench_full.err | tail -1 > ${O}_bench_full.json
python - <<P
import json
try:
    d = json.load(open("${O}_bench_full.json"))
    print("full: us/SVI-iter %.2f value %.3e e2e %.3e parity %s frac %.3f" % (d["us_per_svi_iteration"], d["value"], d["e2e"]["value"], d["parity_check"]["ok"], d["roofline"]["frac"]))
    for k in ("nccl_allreduce_2k_f64_us", "same_shard_1gpu_us", "exchange_us_per_round", "efficiency_vs_same_shard_alone", "convergence"):
        if k in d: print("   ", k, d[k])
except Exception as ex:
    print("full: FAILED", ex); print(open("${O}_bench_full.err").read()[-1500:])
P
