#!/usr/bin/env bash
# round 2: the row-synchronous Schur schedule (CCM_SCHUR=10): parity suites, then timing against the default
set -u
mkdir -p gpurun_out
(CCM_SCHUR=10 timeout 300 python -m pytest tests/test_gpu_ba.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/rowsync_parity.log
cat gpurun_out/rowsync_parity.log
(timeout 400 python tools/schur_probe2.py cfg5 "untiled prefetch, unsorted" "row-synchronous" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/rowsync_cfg5.log
cat gpurun_out/rowsync_cfg5.log
(timeout 200 python tools/schur_probe2.py cfg4 "untiled prefetch, unsorted" "row-synchronous" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/rowsync_cfg4.log
cat gpurun_out/rowsync_cfg4.log
