#!/bin/bash
# 1 GPU, final tree: smoke(), then the bench line with every leg (shorter caps on the time-to-gap and simplex legs)
O=gpurun_out/r2w; mkdir -p $O
( time timeout 120 python __graft_entry__.py smoke ) > $O/smoke.txt 2>&1; tail -3 $O/smoke.txt
( time timeout 330 python bench.py --gap-time-limit 60 --simplex-cap 15 ) > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err
tail -c 1500 $O/bench_c4_n1.json; tail -4 $O/bench_c4_n1.err
