#!/bin/bash
# round 2, GPU call 36 (one B200): the full GPU suite on the final build
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q ) > gpurun_out/r2c36_tests.log 2>&1; grep -E "passed|failed|error" gpurun_out/r2c36_tests.log | tail -3
