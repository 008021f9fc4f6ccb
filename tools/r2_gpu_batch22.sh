#!/usr/bin/env bash
# round 2: more wavefront-saving variants of the Schur list kernel (modes 13 / 14)
set -u
mkdir -p gpurun_out
(CCM_SCHUR=14 timeout 300 python -m pytest tests/test_gpu_ba.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/vec2_parity.log
cat gpurun_out/vec2_parity.log
(CCM_SCHUR=13 timeout 300 python -m pytest tests/test_gpu_ba.py -m gpu -x -q 2>&1 | tail -3) >> gpurun_out/vec2_parity.log
tail -2 gpurun_out/vec2_parity.log
(timeout 400 python tools/schur_probe2.py cfg5 "vectorised entries u8" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/vec2_cfg5.log
cut -c 1-330 gpurun_out/vec2_cfg5.log
(timeout 200 python tools/schur_probe2.py cfg4 "vectorised entries u8" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/vec2_cfg4.log
cut -c 1-330 gpurun_out/vec2_cfg4.log
