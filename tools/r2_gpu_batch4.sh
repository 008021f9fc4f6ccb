#!/usr/bin/env bash
# round 2, GPU call 4 (1 GPU): tiled Schur schedule + sorted product lists, PCG refresh / coarse-size sweep, ncu of k_pcg2
set -u
mkdir -p gpurun_out
(CCM_SCHUR=9 timeout 300 python -m pytest tests/test_gpu_ba.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -6) > gpurun_out/schur_tiled_parity.log
cat gpurun_out/schur_tiled_parity.log
(timeout 400 python tools/schur_probe2.py cfg5 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/schur_tiled_cfg5.log
cat gpurun_out/schur_tiled_cfg5.log
(timeout 400 python tools/pcg_probe.py cfg5 "-" "CCM_PCG_REFRESH=2" "CCM_PCG_REFRESH=1" "CCM_PCG_NC=256,CCM_PCG_REFRESH=2" "CCM_PCG_NC=256,CCM_PCG_REFRESH=1" "CCM_PCG_NC=128,CCM_PCG_REFRESH=1" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/pcg2_sweep_cfg5.log
cat gpurun_out/pcg2_sweep_cfg5.log
(timeout 200 python tools/pcg_probe.py cfg4 "CCM_PCG_IMPL=1" "-" "CCM_PCG_REFRESH=1" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/pcg2_sweep_cfg4.log
cat gpurun_out/pcg2_sweep_cfg4.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_pcg2 -s 9 -c 1 -o gpurun_out/prof_r2_k_pcg2 python tools/pcg_probe.py cfg5 "-" > gpurun_out/ncu_pcg2.log 2>&1
tail -3 gpurun_out/ncu_pcg2.log
