#!/bin/bash
# round 2, GPU call 2 (two B200s): exchange schemes in isolation, NCCL latency, the real two-device
# tests, bench at 2 GPUs for every exchange mode, phase traces
mkdir -p gpurun_out
O=gpurun_out/r2c2
nvidia-smi topo -m > ${O}_topo.txt 2>&1
timeout 180 tools/dev/ubench_xchg 2 > ${O}_xchg2.txt 2>&1; cat ${O}_xchg2.txt
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or two_shards or cli_012 or cli_side" ) > ${O}_tests.log 2>&1; tail -5 ${O}_tests.log
run() {  # $1 = TSGPU_XCHG, rest = bench args
  m=$1; shift
  TSGPU_XCHG=$m TSGPU_TIMEOUT_S=20 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus 2 --snps 50000 --steps 5 --warmup 3 "$@" 2> ${O}_bench_$m.err | tail -1 > ${O}_bench_$m.json
  python - <<P
import json
try:
    d = json.load(open("${O}_bench_$m.json"))
    print("$m: us/SVI-iter %.2f value %.3e parity %s exchange %s" % (d["us_per_svi_iteration"], d["value"], d["parity_check"], d["config"]["exchange"]))
    for k in ("nccl_allreduce_2k_f64_us", "same_shard_1gpu_us", "exchange_us_per_round", "efficiency_vs_same_shard_alone"):
        if k in d: print("   ", k, d[k])
except Exception as ex:
    print("$m: FAILED", ex); print(open("${O}_bench_$m.err").read()[-1500:])
P
}
run mcred --converge-seconds 0
run mcslot --no-extras
run slots --no-extras
run ipc --no-extras
run mcred --no-extras
for m in mcred slots; do
  TSGPU_XCHG=$m TSGPU_TIMEOUT_S=20 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dev/trace_mp.py > ${O}_trace_$m.log 2>&1
  cat gpurun_out/trace_mp_2_$m.txt
done
