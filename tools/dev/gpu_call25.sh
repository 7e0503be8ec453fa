#!/bin/bash
# round 2, GPU call 25 (one B200): first poll after 400 cycles (t22), 500 (t26); 400 + 4 cycles per missing arrival between
# polls (t23), 300 + 8 (t24), 0 + 6 (t25)
mkdir -p gpurun_out
O=gpurun_out/r2c25
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh t14 t22 t26 t23 t24 t25
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- t14 t22 t23
