#!/bin/bash
mkdir -p gpurun_out/r2j
( time CUOPT_B200_TRACE=1 timeout 400 python scripts/exp_e2e_trace.py c4 200 ) > gpurun_out/r2j/e2e_trace_c4.txt 2>&1
cat gpurun_out/r2j/e2e_trace_c4.txt | tail -70
( time timeout 900 python scripts/exp_time_to_gap.py c4 1e-7 2000000 600 ) > gpurun_out/r2j/ttg_c4_1e-7.txt 2>&1
tail -5 gpurun_out/r2j/ttg_c4_1e-7.txt
( time timeout 300 python scripts/exp_time_to_gap.py c2 1e-7 2000000 200 ) > gpurun_out/r2j/ttg_c2_1e-7.txt 2>&1
tail -3 gpurun_out/r2j/ttg_c2_1e-7.txt
