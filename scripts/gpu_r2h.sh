#!/bin/bash
mkdir -p gpurun_out/r2h
( time timeout 600 python scripts/exp_kernel_variants.py c4 "CUOPT_B200_SPMV_NPRE=1" "CUOPT_B200_SPMV_NPRE=2" "" ) > gpurun_out/r2h/variants_c4.txt 2>&1
( time timeout 300 python scripts/exp_kernel_variants.py c2 "CUOPT_B200_SPMV_NPRE=1" "CUOPT_B200_SPMV_NPRE=2" ) > gpurun_out/r2h/variants_c2.txt 2>&1
cat gpurun_out/r2h/variants_c4.txt gpurun_out/r2h/variants_c2.txt
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2h/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2h/pytest_gpu.txt
