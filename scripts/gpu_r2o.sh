#!/bin/bash
# 2 GPUs: NVLink store microbenchmark + the whole sharded test file (no -x)
O=gpurun_out/r2o; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
( time timeout 120 scripts/_bin/nvlink_store_bench 5000000 ) > $O/nvlink_store_bench.txt 2>&1
cat $O/nvlink_store_bench.txt
( time timeout 1500 python -m pytest tests/test_gpu_dist.py -q ) > $O/pytest_dist.txt 2>&1
tail -40 $O/pytest_dist.txt
