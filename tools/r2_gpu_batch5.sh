#!/usr/bin/env bash
# round 2, multi-GPU call: phase timing of the row-distributed PCG at N ranks, kf-store GPU test
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
(timeout 200 python -m pytest tests/test_kf_store.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/kf_store_gpu.log
cat gpurun_out/kf_store_gpu.log
(run tools/pcg_probe.py cfg5 "-" "CCM_PCG_NC=256,CCM_PCG_REFRESH=2" "CCM_PCG_IMPL=1" 2>&1 | grep -E "RESULT|Error|error|Traceback") > gpurun_out/pcg2_dist_probe_n$N.log
cat gpurun_out/pcg2_dist_probe_n$N.log
(run tools/pcg_probe.py cfg4 "-" "CCM_PCG_IMPL=1" 2>&1 | grep -E "RESULT|Error|error|Traceback") > gpurun_out/pcg2_dist_probe_cfg4_n$N.log
cat gpurun_out/pcg2_dist_probe_cfg4_n$N.log
