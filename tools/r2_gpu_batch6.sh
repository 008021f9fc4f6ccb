#!/usr/bin/env bash
# round 2, N-GPU call: where the 1 -> N scaling stands: bench line (with its parity block) and PCG phase cycles at N ranks
set -u
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
(timeout 200 python -m pytest tests/test_kf_store.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/kf_store_gpu.log
cat gpurun_out/kf_store_gpu.log
(run tools/pcg_probe.py cfg5 "-" 2>&1 | grep -E "RESULT|Error|error|Traceback") > gpurun_out/pcg2_dist_probe_n$N.log
cat gpurun_out/pcg2_dist_probe_n$N.log
(run bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/bench_n${N}_dist.err | tail -1) > gpurun_out/bench_n${N}_dist.json
tail -3 gpurun_out/bench_n${N}_dist.err; cat gpurun_out/bench_n${N}_dist.json | cut -c 1-900
