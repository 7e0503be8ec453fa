#!/bin/bash
# round 2, GPU call 8 (eight B200s): exchange schemes in isolation at 8 ranks, bench at 8 GPUs for the
# peer-store and the NVLS multicast-store exchange, then the full default line (1M x 1M, convergence run)
mkdir -p gpurun_out
O=gpurun_out/r2c10
echo skip ubench
run() {  # $1 = tag, $2 = TSGPU_XCHG, rest = bench args
  tag=$1; m=$2; shift 2
  TSGPU_XCHG=$m TSGPU_TIMEOUT_S=30 timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
     bench.py --gpus 8 "$@" 2> ${O}_bench_$tag.err | tail -1 > ${O}_bench_$tag.json
  python - <<P
import json
try:
    d = json.load(open("${O}_bench_$tag.json"))
    print("$tag: us/SVI-iter %.2f value %.3e e2e %.3e parity %s exchange %s" % (d["us_per_svi_iteration"], d["value"], d["e2e"]["value"], d["parity_check"]["ok"], d["config"]["exchange"].get("exchange")))
    for k in ("nccl_allreduce_2k_f64_us", "same_shard_1gpu_us", "exchange_us_per_round", "efficiency_vs_same_shard_alone", "convergence"):
        if k in d: print("   ", k, d[k])
except Exception as ex:
    print("$tag: FAILED", ex); print(open("${O}_bench_$tag.err").read()[-1500:])
P
}
run gacc_short gacc --snps 50000 --steps 5 --warmup 3 --no-extras
run mcacc_short mcacc --snps 50000 --steps 5 --warmup 3 --no-extras
run full auto --steps 10 --warmup 3
TSGPU_TIMEOUT_S=30 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/dev/trace_mp.py > ${O}_trace.log 2>&1
cat gpurun_out/trace_mp_8_default.txt
