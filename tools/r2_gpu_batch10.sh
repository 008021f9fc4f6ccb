#!/usr/bin/env bash
# round 2, GPU call 10 (1 GPU): the landmark-synchronous Schur panel kernel: parity suites, then timing on cfg5 / cfg4 / cfg3
set -u
mkdir -p gpurun_out
(CCM_SCHUR_PANEL=1 timeout 300 python -m pytest tests/test_gpu_ba.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/panel_parity.log
cat gpurun_out/panel_parity.log
for cfg in cfg5 cfg4 cfg3; do
(timeout 300 python tools/pcg_probe.py $cfg "-" "CCM_SCHUR_PANEL=1" "CCM_SCHUR_PANEL=1,CCM_SCHUR_PANEL_FACTOR=100" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/panel_$cfg.log
cat gpurun_out/panel_$cfg.log
done
