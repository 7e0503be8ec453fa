#!/bin/bash
# round 2, GPU call 34 (one B200): the new device test of the control path's table and analytic fallback; smoke
mkdir -p gpurun_out
O=gpurun_out/r2c34
( timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "control_path_table_and_fallback or first_snp_step" ) > ${O}_tests.log 2>&1; tail -15 ${O}_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
