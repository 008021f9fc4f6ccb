#!/usr/bin/env bash
# round 2: vectorised entry loads in the Schur list kernel (modes 11 / 12): parity, timing
set -u
mkdir -p gpurun_out
(CCM_SCHUR=11 timeout 300 python -m pytest tests/test_gpu_ba.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/vec_parity.log
cat gpurun_out/vec_parity.log
(timeout 400 python tools/schur_probe2.py cfg5 "untiled prefetch, unsorted" "vectorised" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/vec_cfg5.log
cut -c 1-330 gpurun_out/vec_cfg5.log
(timeout 200 python tools/schur_probe2.py cfg4 "untiled prefetch, unsorted" "vectorised" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/vec_cfg4.log
cut -c 1-330 gpurun_out/vec_cfg4.log
