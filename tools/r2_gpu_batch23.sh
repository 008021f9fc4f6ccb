#!/usr/bin/env bash
# round 2: Schur default = vectorised entry loads (+ batched diagonal blocks): whole GPU suite, then timing
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/b23_suite.log
cat gpurun_out/b23_suite.log
(timeout 300 python tools/pcg_probe.py cfg5 "-" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b23_cfg5.log
cut -c 1-600 gpurun_out/b23_cfg5.log
(timeout 200 python tools/pcg_probe.py cfg4 "-" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b23_cfg4.log
cut -c 1-600 gpurun_out/b23_cfg4.log
(timeout 200 python tools/schur_probe2.py cfg5 "vectorised entries u8" "untiled prefetch, unsorted" 2>&1 | grep -E "RESULT|Error|error") > gpurun_out/b23_schur.log
cut -c 1-300 gpurun_out/b23_schur.log
