#!/bin/bash
# 2 GPUs, final defaults (gather transport, fused packed stores, device block cache): the gather tests, bench at N = 2
O=gpurun_out/r2v; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_dist.py -q -k "gather and not many_gpu" ) > $O/pytest_dist_gather.txt 2>&1
tail -6 $O/pytest_dist_gather.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 --gap-iteration-limit 0 > $O/bench_c4_n2_gather.json 2> $O/bench_c4_n2_gather.err
python - <<'PY'
import json
for l in open("gpurun_out/r2v/bench_c4_n2_gather.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["detail"]["solver_seconds_per_step"], d["detail"]["setup_seconds_per_step"])
PY
tail -3 $O/bench_c4_n2_gather.err
