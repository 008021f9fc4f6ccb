#!/usr/bin/env bash
# round 2, N-GPU call: bench line (with parity block) and PCG phase cycles at N ranks, low-latency packet protocol
set -u
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
(run bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 3 2>gpurun_out/bench_n${N}_ll.err | tail -1) > gpurun_out/bench_n${N}_ll.json
grep "step:" gpurun_out/bench_n${N}_ll.err | cut -c 1-120; cut -c 1-300 gpurun_out/bench_n${N}_ll.json
(run tools/pcg_probe.py cfg5 "-" 2>&1 | grep -E "RESULT|Error|error|Traceback") > gpurun_out/ll_probe_n$N.log
cut -c 1-900 gpurun_out/ll_probe_n$N.log
