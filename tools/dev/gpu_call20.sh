#!/bin/bash
# round 2, GPU call 20 (two B200s): "the last arrival forwards" (atomic with return value instead of a polling CTA 0)
# against the previous product (libtsgpu_prev.so: CTA 0 polls the local words, then forwards)
mkdir -p gpurun_out
O=gpurun_out/r2c20
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or two_shards or cli_012 or cli_synthetic" ) > ${O}_tests.log 2>&1; tail -3 ${O}_tests.log
( TSGPU_XCHG=gacc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or cli_synthetic" ) > ${O}_tests_gacc.log 2>&1; echo "gacc: $(tail -1 ${O}_tests_gacc.log)"
run() {  # $1 = tag, $2 = TSGPU_XCHG, $3 = library, rest = bench args
  tag=$1; m=$2; lib=$3; shift 3
  TSGPU_LIB=$PWD/terastructure_b200/lib/$lib TSGPU_XCHG=$m TSGPU_TIMEOUT_S=30 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
     bench.py --gpus 2 "$@" 2> ${O}_bench_$tag.err | tail -1 > ${O}_bench_$tag.json
  python - <<P
import json
try:
    d = json.load(open("${O}_bench_$tag.json"))
    print("$tag: us/SVI-iter %.2f value %.3e e2e %.3e parity %s exchange %s" % (d["us_per_svi_iteration"], d["value"], d["e2e"]["value"], d["parity_check"]["ok"], d["config"]["exchange"].get("exchange")))
    for k in ("nccl_allreduce_2k_f64_us", "same_shard_1gpu_us", "exchange_us_per_round", "efficiency_vs_same_shard_alone"):
        if k in d: print("   ", k, d[k])
except Exception as ex:
    print("$tag: FAILED", ex); print(open("${O}_bench_$tag.err").read()[-1500:])
P
}
for rep in 1 2; do
run prev_mc_$rep auto libtsgpu_prev.so --snps 50000 --steps 5 --warmup 3 --no-extras
run new_mc_$rep auto libtsgpu.so --snps 50000 --steps 5 --warmup 3 --no-extras
done
run prev_gacc gacc libtsgpu_prev.so --snps 50000 --steps 5 --warmup 3 --no-extras
run new_gacc gacc libtsgpu.so --snps 50000 --steps 5 --warmup 3 --no-extras
run new_full auto libtsgpu.so --snps 50000 --steps 5 --warmup 3 --converge-seconds 0
TSGPU_TIMEOUT_S=30 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/dev/trace_mp.py > ${O}_trace.log 2>&1
cat gpurun_out/trace_mp_2_default.txt
