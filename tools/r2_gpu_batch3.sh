#!/usr/bin/env bash
# round 2, multi-GPU call: row-distributed PCG (pcg2.cuh over IPC windows) at N ranks: parity of every rank against the oracle, then bench lines
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
(run tools/multirank_check.py 2>&1 | grep -E "OK|FAIL|Error|error|Traceback" | tail -40) > gpurun_out/dist2_check_n$N.log
cat gpurun_out/dist2_check_n$N.log
(CCM_PCG_IMPL=1 run tools/multirank_check.py 2>&1 | grep -E "OK|FAIL|Error|error|Traceback" | tail -40) > gpurun_out/dist2_check_n${N}_replicated.log
cat gpurun_out/dist2_check_n${N}_replicated.log
(run bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/bench_n${N}_dist.err | tail -1) > gpurun_out/bench_n${N}_dist.json
tail -5 gpurun_out/bench_n${N}_dist.err; cat gpurun_out/bench_n${N}_dist.json | cut -c 1-1500
(CCM_PCG_IMPL=1 run bench.py --gpus $N --steps 3 --warmup 3 --no-parity 2>gpurun_out/bench_n${N}_repl.err | tail -1) > gpurun_out/bench_n${N}_repl.json
cat gpurun_out/bench_n${N}_repl.json | cut -c 1-600
