#!/usr/bin/env bash
# round 2: the one-hop grid barrier (last arriver releases a separate word) in k_pcg and k_pcg2: parity, then timing on all sizes
set -u
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_ba.py tests/test_golden.py tests/test_gpu_frontend.py tests/test_gpu_single.py -m gpu -x -q 2>&1 | tail -6) > gpurun_out/b20_parity.log
cat gpurun_out/b20_parity.log
(timeout 300 python tools/pcg_probe.py cfg5 "-" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b20_cfg5.log
cut -c 1-700 gpurun_out/b20_cfg5.log
(timeout 200 python tools/pcg_probe.py cfg4 "-" "CCM_PCG_IMPL=2" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b20_cfg4.log
cut -c 1-700 gpurun_out/b20_cfg4.log
(timeout 200 python tools/pcg_probe.py cfg3 "-" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b20_cfg3.log
cut -c 1-700 gpurun_out/b20_cfg3.log
(timeout 200 python tools/pcg_probe.py cfg2 "-" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b20_cfg2.log
cut -c 1-700 gpurun_out/b20_cfg2.log
