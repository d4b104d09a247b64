#!/bin/bash
mkdir -p gpurun_out/r2c
( time timeout 600 scripts/_bin/spmv_lab 10000000 1000000 ) > gpurun_out/r2c/spmv_lab.txt 2>&1
cat gpurun_out/r2c/spmv_lab.txt
