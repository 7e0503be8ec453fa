#!/bin/bash
# round 2, GPU call 24 (two B200s): delayed first poll -- on GPU 0 alone 200 (t21) / 300 (t17) / 400 (t22) cycles;
# on both, the ranks' accumulator first polled 800 (t21) / 1600 (t22) cycles after the arrival's return
mkdir -p gpurun_out
O=gpurun_out/r2c24
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh t14 t21 t17 t22
run() {  # $1 = tag, $2 = TSGPU_XCHG, $3 = library, rest = bench args
  tag=$1; m=$2; lib=$3; shift 3
  TSGPU_LIB=$PWD/terastructure_b200/lib/$lib TSGPU_XCHG=$m TSGPU_TIMEOUT_S=30 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
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
for rep in 1 2; do
run t14_mc_$rep auto libtsgpu_t14.so --snps 50000 --steps 5 --warmup 3 --no-extras
run t21_mc_$rep auto libtsgpu_t21.so --snps 50000 --steps 5 --warmup 3 --no-extras
run t22_mc_$rep auto libtsgpu_t22.so --snps 50000 --steps 5 --warmup 3 --no-extras
done
