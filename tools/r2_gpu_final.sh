#!/usr/bin/env bash
# round 2, final single-GPU validation sweep: whole GPU suite, smoke(), one ncu --set full capture of the shipped Schur kernel, the default bench line
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/gpu_suite_final.log
cat gpurun_out/gpu_suite_final.log
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3) > gpurun_out/smoke_final.log
cat gpurun_out/smoke_final.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_schur_mma -s 2 -c 1 -o gpurun_out/prof_r2_k_schur_mma python tools/schur_variants.py cfg5 --capture > gpurun_out/ncu_schur.log 2>&1
tail -2 gpurun_out/ncu_schur.log | cut -c 1-200
(timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_n1_final.err | tail -1) > gpurun_out/bench_n1_final.json
grep "step:" gpurun_out/bench_n1_final.err | cut -c 1-140; cut -c 1-400 gpurun_out/bench_n1_final.json
