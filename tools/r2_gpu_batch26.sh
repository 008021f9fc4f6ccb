#!/usr/bin/env bash
# round 2, batch 26: end-to-end steps with the measurement upload overlapped with the structure build (default) and on one stream
set -u
mkdir -p gpurun_out
for ov in 1 0 1 0; do
echo "CCM_MEAS_OVERLAP=$ov" >> gpurun_out/e2e_overlap.log
CCM_MEAS_OVERLAP=$ov timeout 600 python bench.py --steps 2 --warmup 3 --no-parity --no-extras --no-cpu-baseline --e2e-steps 9 2>&1 | grep -E "e2e step|^\{" | cut -c1-330 >> gpurun_out/e2e_overlap.log
done
CCM_SETUP_PROF=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-parity --no-extras --no-cpu-baseline --e2e-steps 3 2>&1 | grep -E "ccm_ba_create|e2e step" | tail -40 >> gpurun_out/e2e_overlap.log
cat gpurun_out/e2e_overlap.log
