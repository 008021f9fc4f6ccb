#!/usr/bin/env bash
# round 2, GPU call 2: the second-generation PCG kernel (pcg2.cuh) on one GPU: parity suites, then cfg5 / cfg4 probes
set -u
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ba.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pcg2_parity.log
cat gpurun_out/pcg2_parity.log
(timeout 400 python tools/pcg_probe.py cfg5 "CCM_PCG_IMPL=1" "-" "CCM_PCG_NC=256" "CCM_PCG_NC=384" "CCM_PCG_NC=384,CCM_PCG_REFRESH=8" "CCM_PCG_NC=256,CCM_PCG_REFRESH=8" 2>&1 | grep -E "RESULT|Error|error" ) > gpurun_out/pcg2_cfg5.log
cat gpurun_out/pcg2_cfg5.log
(timeout 200 python tools/pcg_probe.py cfg4 "CCM_PCG_IMPL=1" "-" "CCM_PCG_NC=64" "CCM_PCG2_GRID=74" "CCM_PCG2_GRID=37" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/pcg2_cfg4.log
cat gpurun_out/pcg2_cfg4.log
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/gpu_suite2.log
cat gpurun_out/gpu_suite2.log
