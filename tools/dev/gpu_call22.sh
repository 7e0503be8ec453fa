#!/bin/bash
# round 2, GPU call 22 (one B200): two polls of a word pair in flight, half a round trip apart (t15: 225 cycles, t16: 350)
mkdir -p gpurun_out
O=gpurun_out/r2c22
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh t14 t15 t16
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- t14 t15
