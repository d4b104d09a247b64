#!/bin/bash
# 1 GPU: where the end-to-end time of a C-ABI solve goes (configs[3]); the whole GPU test suite
O=gpurun_out/r2q; mkdir -p $O
( time CUOPT_B200_TRACE=1 timeout 400 python scripts/exp_e2e_trace.py c4 2000 ) > $O/e2e_trace_c4.txt 2>&1
tail -45 $O/e2e_trace_c4.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.txt 2>&1
tail -25 $O/pytest_gpu.txt
