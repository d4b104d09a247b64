#!/bin/bash
mkdir -p gpurun_out/r2e
( time CUOPT_B200_PLACEMENT_PROBE=1 timeout 300 python scripts/exp_kernel_variants.py c4 "" ) > gpurun_out/r2e/probe_c4.txt 2>&1
cat gpurun_out/r2e/probe_c4.txt
