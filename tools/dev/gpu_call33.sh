#!/bin/bash
# round 2, GPU call 33 (two B200s): copies of the local words on several GPUs (TS_MG_SUB = 4: product; 2: t36; 1: t35 = one
# word per statistic as before), multi-device tests on the product
mkdir -p gpurun_out
O=gpurun_out/r2c33
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or two_shards or cli_012 or cli_synthetic" ) > ${O}_tests.log 2>&1; tail -2 ${O}_tests.log
( TSGPU_XCHG=gacc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or cli_synthetic" ) > ${O}_tests_gacc.log 2>&1; echo "gacc: $(tail -1 ${O}_tests_gacc.log)"
run() {  # $1 = tag, $2 = TSGPU_XCHG, $3 = library, rest = bench args
  tag=$1; m=$2; lib=$3; shift 3
  TSGPU_LIB=$PWD/terastructure_b200/lib/$lib TSGPU_XCHG=$m TSGPU_TIMEOUT_S=30 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
     bench.py --gpus 2 "$@" 2> ${O}_bench_$tag.err | tail -1 > ${O}_bench_$tag.json
  python - <<P
import json
try:
    d = json.load(open("${O}_bench_$tag.json"))
    print("$tag: us/SVI-iter %.2f value %.3e e2e %.3e parity %s" % (d["us_per_svi_iteration"], d["value"], d["e2e"]["value"], d["parity_check"]["ok"]))
except Exception as ex:
    print("$tag: FAILED", ex); print(open("${O}_bench_$tag.err").read()[-1500:])
P
}
run sub1_mc auto libtsgpu_t35.so --snps 50000 --steps 5 --warmup 3 --no-extras
run sub2_mc auto libtsgpu_t36.so --snps 50000 --steps 5 --warmup 3 --no-extras
run sub4_mc auto libtsgpu.so --snps 50000 --steps 5 --warmup 3 --no-extras
run sub1_mc_b auto libtsgpu_t35.so --snps 50000 --steps 5 --warmup 3 --no-extras
run sub4_mc_b auto libtsgpu.so --snps 50000 --steps 5 --warmup 3 --no-extras
run sub4_gacc gacc libtsgpu.so --snps 50000 --steps 5 --warmup 3 --no-extras
