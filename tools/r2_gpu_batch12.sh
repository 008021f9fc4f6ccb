#!/usr/bin/env bash
# round 2, GPU call 12 (1 GPU): the default bench line end to end (parity block, cfg4, front end), PGO traces, CPU-suite-adjacent GPU tests
set -u
mkdir -p gpurun_out
(timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_n1.err | tail -1) > gpurun_out/bench_n1.json
tail -12 gpurun_out/bench_n1.err; cut -c 1-600 gpurun_out/bench_n1.json
(timeout 120 python tools/pgo_trace.py 2>&1 | tail -60) > gpurun_out/pgo_trace.log
cat gpurun_out/pgo_trace.log
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/gpu_suite3.log
cat gpurun_out/gpu_suite3.log
