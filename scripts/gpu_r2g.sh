#!/bin/bash
mkdir -p gpurun_out/r2g
( time timeout 600 scripts/_bin/spmv_lab 10000000 1000000 ) > gpurun_out/r2g/spmv_lab.txt 2>&1
grep -E "==|PRODUCTION|base|bicsr hint \+prefetch  " gpurun_out/r2g/spmv_lab.txt
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2g/pytest_gpu.txt 2>&1
tail -30 gpurun_out/r2g/pytest_gpu.txt
( time timeout 600 python scripts/exp_kernel_variants.py c4 "" "CUOPT_B200_GATHER_BLOCK_BYTES=0" "CUOPT_B200_GATHER_BLOCK_BYTES=27000000" ) > gpurun_out/r2g/variants_c4.txt 2>&1
( time timeout 300 python scripts/exp_kernel_variants.py c2 "" ) > gpurun_out/r2g/variants_c2.txt 2>&1
cat gpurun_out/r2g/variants_c4.txt gpurun_out/r2g/variants_c2.txt
