#!/usr/bin/env bash
# round 2, N-GPU call: low-latency packet protocol of the distributed PCG: parity of every rank, then phase cycles
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
(run tools/multirank_check.py 2>&1 | grep -E "OK|FAIL|Error|error|Traceback" | tail -40) > gpurun_out/ll_check_n$N.log
cat gpurun_out/ll_check_n$N.log
(run tools/pcg_probe.py cfg5 "-" "CCM_PCG_IMPL=1" 2>&1 | grep -E "RESULT|Error|error|Traceback") > gpurun_out/ll_probe_n$N.log
cat gpurun_out/ll_probe_n$N.log
(run tools/pcg_probe.py cfg4 "-" "CCM_PCG_IMPL=1" 2>&1 | grep -E "RESULT|Error|error|Traceback") > gpurun_out/ll_probe_cfg4_n$N.log
cat gpurun_out/ll_probe_cfg4_n$N.log
