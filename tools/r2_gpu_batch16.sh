#!/usr/bin/env bash
# round 2, N-GPU call: (N = 2: every rank of the streamed / distributed solve against the oracle first) then the bench line at N ranks
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
if [ "$N" = 2 ]; then
(CCM_PCG_IMPL=2 run tools/multirank_check.py 2>&1 | grep -E "OK|FAIL|Error|error|Traceback" | tail -40) > gpurun_out/final_check_n$N.log
cat gpurun_out/final_check_n$N.log
fi
(run bench.py --gpus $N --steps 5 --warmup 3 --e2e-steps 3 2>gpurun_out/bench_n${N}_final.err | tail -1) > gpurun_out/bench_n${N}_final.json
grep "step:" gpurun_out/bench_n${N}_final.err | cut -c 1-120; cut -c 1-260 gpurun_out/bench_n${N}_final.json
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n${N}_final.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "parity", d["parity"])
print({k: round(v["avg_ms"] * v["launches"] / d["steps"], 2) for k, v in d["kernels"].items()})
PY
