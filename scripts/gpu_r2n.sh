#!/bin/bash
# 2 GPUs: the gather transport (default) against single GPU / the other transports; bench configs[3] at N = 2; slot trace
O=gpurun_out/r2n; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_dist.py -x -q ) > $O/pytest_dist.txt 2>&1
tail -8 $O/pytest_dist.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in gather p2p; do
  CUOPT_B200_DIST_MODE=$mode timeout 400 $TR --master-port 2951$((RANDOM%10)) bench.py --gpus 2 --steps 2 --warmup 3 --gap-iteration-limit 0 \
     > $O/bench_c4_n2_$mode.json 2> $O/bench_c4_n2_$mode.err
  tail -c 1500 $O/bench_c4_n2_$mode.json; tail -3 $O/bench_c4_n2_$mode.err
done
for mode in gather p2p; do
  CUOPT_B200_DIST_TRACE=1 CUOPT_B200_DIST_MODE=$mode timeout 300 $TR --master-port 2952$((RANDOM%10)) bench.py --gpus 2 --steps 1 --warmup 3 --iters 400 --gap-iteration-limit 0 \
     > $O/trace_c4_n2_$mode.json 2> $O/trace_c4_n2_$mode.err
  grep "dist trace" $O/trace_c4_n2_$mode.err
done
