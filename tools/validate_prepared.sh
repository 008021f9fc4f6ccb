#!/usr/bin/env bash
# First GPU call of the next round: validate the switches that were prepared without a device (DESIGN.md §7 "Prepared, not yet run").
# Usage (1 GPU):   gpurun --timeout 420 -- 'bash tools/validate_prepared.sh single'
#       (2 GPUs):  gpurun --gpus 2 --timeout 300 -- 'bash tools/validate_prepared.sh dist'
# Everything is wrapped in `timeout`; logs go to gpurun_out/.
set -u
mkdir -p gpurun_out
mode=${1:-single}
if [ "$mode" = single ]; then
  # 1. piecewise-linear coarse prolongation: parity first (BA suites incl. golden), then iterations / ms on cfg5 for three coarse sizes
  (CCM_PCG_PROLONG=1 timeout 120 python -m pytest tests/test_gpu_ba.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/prolong_parity.log
  for nc in 0 192 384; do
    if [ $nc = 0 ]; then env_nc=""; label="pc384 (default)"; else env_nc="CCM_PCG_PROLONG=1 CCM_PCG_NC=$nc"; label="pl$nc"; fi
    (echo "== $label"; env $env_nc timeout 90 python tools/schur_variants.py cfg5 2>&1 | tail -1) >> gpurun_out/prolong_cfg5.log
  done
  # 1b. the distributed PCG kernel against its own window (one rank): everything but the NVLink hop, on one GPU
  (CCM_PCG_DIST=1 timeout 120 python -m pytest tests/test_gpu_ba.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/dist_selfwindow_parity.log
  (CCM_PCG_DIST=1 CCM_PCG_PROLONG=1 timeout 120 python -m pytest tests/test_gpu_ba.py -m gpu -x -q 2>&1 | tail -5) >> gpurun_out/dist_selfwindow_parity.log
  (echo "== dist self-window"; CCM_PCG_DIST=1 timeout 90 python tools/schur_variants.py cfg5 2>&1 | tail -1) >> gpurun_out/prolong_cfg5.log
  # 1c. device-side window search of Fuse / SearchBySim3 (k_window_best)
  (CCM_MATCH_WINDOW=1 timeout 120 python -m pytest tests/test_gpu_widen.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/match_window.log
  # 1d. the shims over the real device entry points next to the reference's ORBmatcher.cpp / Optimizer.cpp (opt-in file)
  (CCM_TEST_UNVALIDATED=1 timeout 120 python -m pytest tests/test_gpu_zz_dropin.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/dropin_gpu.log
  # 1e. the map update after a global BA (k_map_update_points), bit for bit against the oracle
  (CCM_TEST_UNVALIDATED=1 timeout 120 python -m pytest tests/test_gpu_zz_map_update.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/map_update_gpu.log
  # 2. the whole GPU suite with the current defaults (CTA-128 Schur kernel, new golden / SearchForInitialization tests)
  (timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/gpu_suite.log
  cat gpurun_out/prolong_parity.log gpurun_out/match_window.log gpurun_out/dropin_gpu.log gpurun_out/map_update_gpu.log gpurun_out/dist_selfwindow_parity.log gpurun_out/prolong_cfg5.log gpurun_out/gpu_suite.log
else
  # 3. row-distributed PCG over peer memory: parity of every rank against the oracle, replicated vs distributed, then one bench line each
  run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 "$@"; }
  (run tools/multirank_check.py 2>&1 | grep -E "OK|FAIL|Error|error" | tail -12) > gpurun_out/dist_off.log
  (CCM_PCG_DIST=1 run tools/multirank_check.py 2>&1 | grep -E "OK|FAIL|Error|error" | tail -12) > gpurun_out/dist_on.log
  (CCM_PCG_DIST=1 CCM_PCG_PROLONG=1 run tools/multirank_check.py 2>&1 | grep -E "OK|FAIL|Error|error" | tail -12) > gpurun_out/dist_on_pl.log
  (run bench.py --gpus 2 --steps 2 --warmup 3 2>&1 | tail -1) > gpurun_out/bench_n2_replicated.json
  (CCM_PCG_DIST=1 run bench.py --gpus 2 --steps 2 --warmup 3 2>&1 | tail -1) > gpurun_out/bench_n2_dist.json
  cat gpurun_out/dist_off.log gpurun_out/dist_on.log gpurun_out/dist_on_pl.log gpurun_out/bench_n2_replicated.json gpurun_out/bench_n2_dist.json
fi
