#!/bin/bash
mkdir -p gpurun_out/r2f
( time timeout 600 scripts/_bin/spmv_lab 10000000 1000000 ) > gpurun_out/r2f/spmv_lab.txt 2>&1
( time timeout 300 scripts/_bin/cusparse_pdhg 1000000 10000000 ) > gpurun_out/r2f/cusparse_pdhg.txt 2>&1
cat gpurun_out/r2f/spmv_lab.txt; tail -5 gpurun_out/r2f/cusparse_pdhg.txt
