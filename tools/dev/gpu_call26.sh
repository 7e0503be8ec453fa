#!/bin/bash
# round 2, GPU call 26 (one B200): with the delayed first poll, low words a slice apart from the high words (t28; t29 with
# a delay of 500) against the pairs (t27 = product)
mkdir -p gpurun_out
O=gpurun_out/r2c26
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh t27 t28 t29
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- t27 t28
