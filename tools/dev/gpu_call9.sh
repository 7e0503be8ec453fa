#!/bin/bash
# round 2, GPU call 9 (two B200s): the replicated-accumulator exchange (GACC / MCACC) against the slot
# exchange, the multi-device tests, the loader and CLI tests that call 7 did not reach
mkdir -p gpurun_out
O=gpurun_out/r2c9
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or two_shards or cli_012 or fanout or cli_synthetic or staged or tiered_kernel_large" ) > ${O}_tests.log 2>&1; tail -5 ${O}_tests.log
for m in mcacc slots; do
  ( TSGPU_XCHG=$m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or cli_synthetic" ) > ${O}_tests_$m.log 2>&1; echo "$m: $(tail -1 ${O}_tests_$m.log)"
done
run() {  # $1 = TSGPU_XCHG, rest = bench args
  m=$1; shift
  TSGPU_XCHG=$m TSGPU_TIMEOUT_S=20 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --snps 50000 --steps 5 --warmup 3 "$@" 2> ${O}_bench_$m.err | tail -1 > ${O}_bench_$m.json
  python - <<P
import json
try:
    d = json.load(open("${O}_bench_$m.json"))
    print("$m: us/SVI-iter %.2f value %.3e parity %s exchange %s" % (d["us_per_svi_iteration"], d["value"], d["parity_check"]["ok"], d["config"]["exchange"].get("exchange")))
    for k in ("nccl_allreduce_2k_f64_us", "same_shard_1gpu_us", "exchange_us_per_round", "efficiency_vs_same_shard_alone"):
        if k in d: print("   ", k, d[k])
except Exception as ex:
    print("$m: FAILED", ex); print(open("${O}_bench_$m.err").read()[-1500:])
P
}
run gacc --converge-seconds 0
run mcacc --no-extras
run slots --no-extras
run gacc --no-extras
for m in gacc mcacc; do
  TSGPU_XCHG=$m TSGPU_TIMEOUT_S=20 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dev/trace_mp.py > ${O}_trace_$m.log 2>&1
  cat gpurun_out/trace_mp_2_$m.txt
done
