#!/bin/bash
# round 2, GPU call 23 (one B200): fewer polls -- first poll delayed by 300 (t17) / 600 (t18) cycles, 300 + a gap of 150
# between polls (t19), a gap of 200 alone (t20)
mkdir -p gpurun_out
O=gpurun_out/r2c23
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh t14 t17 t18 t19 t20
