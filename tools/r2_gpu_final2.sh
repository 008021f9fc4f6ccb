#!/usr/bin/env bash
# round 2, closing single-GPU sweep after the last library change: whole GPU suite, smoke(), the default bench line, the ncu launch list of the same command
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/gpu_suite_final.log
cat gpurun_out/gpu_suite_final.log
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3) > gpurun_out/smoke_final.log
cat gpurun_out/smoke_final.log
(timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_n1_final.err | tail -1) > gpurun_out/bench_n1_final.json
grep "step:" gpurun_out/bench_n1_final.err | cut -c 1-140; cut -c 1-400 gpurun_out/bench_n1_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-parity --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c 1-200; wc -l gpurun_out/launches_r2.csv
