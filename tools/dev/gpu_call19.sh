#!/bin/bash
# round 2, GPU call 19 (one B200): genotype bytes loaded one SNP ahead (t11), + the two words of a statistic adjacent
# and polled with one 16-byte load (t12), against the product of commit "Round loop tail ..." (t10)
mkdir -p gpurun_out
O=gpurun_out/r2c19
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh t10 t11 t12
AB_OUT=${O}_ab125.jsonl tools/dev/ab.sh --individuals 125000 -- t10 t11 t12
