#!/usr/bin/env bash
# round 2, GPU call 8 (1 GPU): GJB=16 coarse inversion, refresh sweep, exact column windows; parity suites first
set -u
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ba.py tests/test_golden.py tests/test_gpu_frontend.py -m gpu -x -q 2>&1 | tail -6) > gpurun_out/b8_parity.log
cat gpurun_out/b8_parity.log
(timeout 400 python tools/pcg_probe.py cfg5 "-" "CCM_PCG_REFRESH=3" "CCM_PCG_REFRESH=4" "CCM_PCG_NC=320,CCM_PCG_REFRESH=2" "CCM_PCG_NC=320,CCM_PCG_REFRESH=4" "CCM_PCG_NC=384,CCM_PCG_REFRESH=4" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b8_cfg5.log
cat gpurun_out/b8_cfg5.log
(timeout 200 python tools/pcg_probe.py cfg4 "-" "CCM_PCG_REFRESH=4" "CCM_PCG_REFRESH=8" "CCM_PCG_NC=64,CCM_PCG_REFRESH=4" "CCM_PCG_IMPL=1" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b8_cfg4.log
cat gpurun_out/b8_cfg4.log
(timeout 200 python tools/pcg_probe.py cfg3 "-" "CCM_PCG_REFRESH=4" "CCM_PCG_IMPL=1" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b8_cfg3.log
cat gpurun_out/b8_cfg3.log
