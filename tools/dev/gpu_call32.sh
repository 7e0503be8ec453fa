#!/bin/bash
# round 2, GPU call 32 (one B200): the E-step's weights through the exponent field alone (t32) against the same source
# without (t33) and the product library
mkdir -p gpurun_out
O=gpurun_out/r2c32
cp terastructure_b200/lib/libtsgpu.so terastructure_b200/lib/libtsgpu_new.so
AB_OUT=${O}_ab.jsonl tools/dev/ab.sh new t33 t32
