#!/bin/bash
mkdir -p gpurun_out/r2k
( time CUOPT_B200_TRACE=1 timeout 400 python scripts/exp_e2e_trace.py c4 200 ) > gpurun_out/r2k/e2e_trace_c4.txt 2>&1
tail -45 gpurun_out/r2k/e2e_trace_c4.txt
( time timeout 900 python -m pytest tests -m gpu -q -k "not headline_lp_converges" ) > gpurun_out/r2k/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2k/pytest_gpu.txt
( time timeout 700 python scripts/exp_time_to_gap.py c4 1e-6 2000000 500 ) > gpurun_out/r2k/ttg_c4_1e-6.txt 2>&1
tail -3 gpurun_out/r2k/ttg_c4_1e-6.txt
