#!/bin/bash
# round 2, re-entry: state of the tree on a fresh B200 — GPU tests, default bench line, ncu evidence for profiles/r2
O=gpurun_out/r2m; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
( time timeout 500 python bench.py ) > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err
tail -c 3000 $O/bench_c4_n1.json
( time timeout 200 python bench.py --workload c2 --steps 3 --simplex-cap 0 --gap-iteration-limit 0 ) > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err
KRE='regex:k_primal_step|k_dual_step|k_transpose_step|k_block_pass'
timeout 400 ncu --set full --clock-control none --import-source on -k "$KRE" -s 100 -c 5 -f -o $O/ncu_full_c4 \
   python scripts/profile_target.py --workload c4 --warmup 45 --reps 2 > $O/ncu_full_c4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "$KRE" -s 90 -c 3 -f -o $O/ncu_full_c2 \
   python scripts/profile_target.py --workload c2 --warmup 45 --reps 2 > $O/ncu_full_c2.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launch_list_c4_bench.csv \
   python bench.py --steps 1 --warmup 3 --iters 200 --no-cpu-baseline --gap-iteration-limit 0 --simplex-cap 0 --comparator off --profile-reps 3 > $O/bench_under_ncu.log 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu.txt 2>&1
tail -15 $O/pytest_gpu.txt
ls -la $O
