#!/bin/bash
mkdir -p gpurun_out/r2b
( time timeout 900 python scripts/exp_kernel_variants.py c4 "" "CUOPT_B200_L2_WARM=1" "CUOPT_B200_L2_WARM=2" \
   "CUOPT_B200_L2_WARM=1,CUOPT_B200_GATHER_BLOCK_BYTES=0" "CUOPT_B200_GATHER_BLOCK_BYTES=0" \
   "CUOPT_B200_L2_WARM=1,CUOPT_B200_GATHER_LDG=1" "CUOPT_B200_L2_WARM=1,CUOPT_B200_GATHER_BLOCK_BYTES=0,CUOPT_B200_GATHER_LDG=1" \
   "CUOPT_B200_L2_WARM=1,CUOPT_B200_GATHER_BLOCK_BYTES=20000000" ) > gpurun_out/r2b/variants_c4.txt 2>&1
( time timeout 300 python scripts/exp_kernel_variants.py c2 "" "CUOPT_B200_L2_WARM=1" "CUOPT_B200_GATHER_LDG=1" ) > gpurun_out/r2b/variants_c2.txt 2>&1
timeout 120 python -m pytest tests/test_zero_movement.py -m gpu -q > gpurun_out/r2b/pytest_zero.txt 2>&1
cat gpurun_out/r2b/variants_c4.txt gpurun_out/r2b/variants_c2.txt; tail -5 gpurun_out/r2b/pytest_zero.txt
