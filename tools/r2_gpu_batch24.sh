#!/bin/bash
# round 2, batch 24: layouts of the rows of Z under the default Schur kernel (CCM_Z_LAYOUT), ms per launch + parity of one Global BA
mkdir -p gpurun_out
timeout 900 python tools/schur_probe2.py cfg5 "layout" > gpurun_out/zlayout_cfg5.log 2>&1
grep RESULT gpurun_out/zlayout_cfg5.log
