#!/bin/bash
mkdir -p gpurun_out/r2i
( time timeout 400 python scripts/exp_time_to_gap.py c2 1e-6 2000000 300 ) > gpurun_out/r2i/ttg_c2.txt 2>&1
tail -4 gpurun_out/r2i/ttg_c2.txt
( time timeout 200 python -m pytest tests/test_methodical1.py -m gpu -q ) > gpurun_out/r2i/pytest_m1.txt 2>&1
tail -12 gpurun_out/r2i/pytest_m1.txt
