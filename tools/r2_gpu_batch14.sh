#!/usr/bin/env bash
# the full default bench line twice (is the 292 ms step of the first full run reproducible?), then the launch list under ncu
set -u
mkdir -p gpurun_out
for i in 1 2; do
(timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_n1_run$i.err | tail -1) > gpurun_out/bench_n1_run$i.json
grep "step:" gpurun_out/bench_n1_run$i.err | cut -c 1-120; cut -c 1-330 gpurun_out/bench_n1_run$i.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-parity --no-extras --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c 1-200; wc -l gpurun_out/launches_r2.csv
