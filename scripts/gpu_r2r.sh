#!/bin/bash
# 8 GPUs: element-wise check of the packed gather transport on 8 ranks, configs[3] at N = 8 (bench line, slot + setup trace)
O=gpurun_out/r2r; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_dist.py -q -k "many_gpu_iterates and 8" ) > $O/pytest_dist_8.txt 2>&1
tail -6 $O/pytest_dist_8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29537 bench.py --gpus 8 --steps 2 --warmup 3 --gap-iteration-limit 0 > $O/bench_c4_n8_gather.json 2> $O/bench_c4_n8_gather.err
tail -c 700 $O/bench_c4_n8_gather.json; tail -3 $O/bench_c4_n8_gather.err
CUOPT_B200_TRACE=1 CUOPT_B200_DIST_TRACE=1 timeout 400 $TR --master-port 29547 bench.py --gpus 8 --steps 1 --warmup 3 --iters 200 --gap-iteration-limit 0 > $O/trace_c4_n8_gather.json 2> $O/trace_c4_n8_gather.err
grep "dist trace" $O/trace_c4_n8_gather.err | tail -8
grep "cuopt-b200 trace" $O/trace_c4_n8_gather.err | tail -28
( time timeout 120 scripts/_bin/nvlink_store_bench 10000000 ) > $O/nvlink_store_bench_8.txt 2>&1
tail -8 $O/nvlink_store_bench_8.txt
