#!/bin/bash
# round 2, batch 25: grouped Schur lists (CCM_SCHUR=16/17) and the shared-memory entry broadcast (15): parity suite + cfg5 timings
mkdir -p gpurun_out
CCM_SCHUR=17 timeout 600 python -m pytest tests/test_gpu_ba.py -m gpu -x -q > gpurun_out/quad17_tests.log 2>&1; tail -3 gpurun_out/quad17_tests.log
CCM_SCHUR=16 timeout 600 python -m pytest tests/test_gpu_ba.py -m gpu -x -q > gpurun_out/quad16_tests.log 2>&1; tail -3 gpurun_out/quad16_tests.log
CCM_SETUP_PROF=1 timeout 900 python tools/schur_probe2.py cfg5 "baseline" "entries through shared" "grouped" > gpurun_out/quad_cfg5.log 2>&1
grep -E "RESULT|grouped" gpurun_out/quad_cfg5.log
timeout 600 python tools/schur_probe2.py cfg4 "baseline" "grouped" > gpurun_out/quad_cfg4.log 2>&1
grep RESULT gpurun_out/quad_cfg4.log
