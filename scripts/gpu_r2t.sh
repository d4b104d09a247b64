#!/bin/bash
# 2 GPUs: send kernel with 8 independent chains per thread on 64 CTAs, 32 reserved SpMV slots
O=gpurun_out/r2t; mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_dist.py -q -x -k "iterates_track and gather and not many_gpu" ) > $O/pytest_dist_gather.txt 2>&1
tail -5 $O/pytest_dist_gather.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --gap-iteration-limit 0 > $O/bench_c4_n2_gather.json 2> $O/bench_c4_n2_gather.err
tail -c 300 $O/bench_c4_n2_gather.json; tail -3 $O/bench_c4_n2_gather.err
CUOPT_B200_DIST_TRACE=1 timeout 300 $TR --master-port 29527 bench.py --gpus 2 --steps 1 --warmup 3 --iters 200 --gap-iteration-limit 0 > $O/trace_c4_n2_gather.json 2> $O/trace_c4_n2_gather.err
grep "dist trace" $O/trace_c4_n2_gather.err | tail -2
