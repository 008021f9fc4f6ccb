#!/usr/bin/env bash
# why does the timed region of bench.py take 292 ms per step when the kernels sum to 136 ms?  sampler on / off, profile on / off
set -u
mkdir -p gpurun_out
(timeout 600 python bench.py --steps 4 --warmup 3 --no-parity --no-extras --no-cpu-baseline --e2e-steps 2 2>&1 | grep -E "\[bench\] step|value" | cut -c 1-200) > gpurun_out/b13_sampler_on.log
cat gpurun_out/b13_sampler_on.log
(CCM_BENCH_NO_SAMPLER=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-parity --no-extras --no-cpu-baseline --e2e-steps 2 2>&1 | grep -E "\[bench\] step|value" | cut -c 1-200) > gpurun_out/b13_sampler_off.log
cat gpurun_out/b13_sampler_off.log
