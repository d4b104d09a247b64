#!/bin/bash
# round 2, GPU call A: SpMV lab, cuSPARSE comparator, first GPU run of the Methodical1 / save_best_primal tests
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a/smi.txt 2>&1
nproc > gpurun_out/r2a/nproc.txt
( time timeout 600 scripts/_bin/spmv_lab 10000000 1000000 ) > gpurun_out/r2a/spmv_lab.txt 2>&1
( time timeout 300 scripts/_bin/cusparse_pdhg 1000000 10000000 ) > gpurun_out/r2a/cusparse_pdhg.txt 2>&1
( time CUOPT_B200_RUN_EXPERIMENTAL=1 CUOPT_B200_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_methodical1_experimental.py tests/test_save_best_primal.py -m gpu -q ) > gpurun_out/r2a/pytest_unvalidated.txt 2>&1
tail -30 gpurun_out/r2a/spmv_lab.txt
tail -8 gpurun_out/r2a/cusparse_pdhg.txt
tail -30 gpurun_out/r2a/pytest_unvalidated.txt
