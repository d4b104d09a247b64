#!/bin/bash
# 1 GPU: device block cache + final defaults — whole GPU suite, e2e trace, short bench line
O=gpurun_out/r2u; mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/pytest_gpu.txt 2>&1
tail -6 $O/pytest_gpu.txt
( time CUOPT_B200_TRACE=1 timeout 300 python scripts/exp_e2e_trace.py c4 2000 ) > $O/e2e_trace_c4.txt 2>&1
grep rep $O/e2e_trace_c4.txt | cut -c1-330; grep "so far" $O/e2e_trace_c4.txt | tail -1
( time timeout 300 python bench.py --steps 3 --gap-iteration-limit 0 --simplex-cap 0 --comparator off --no-cpu-baseline ) > $O/bench_c4_n1_short.json 2> $O/bench_c4_n1_short.err
python - <<'PY'
import json
for l in open("gpurun_out/r2u/bench_c4_n1_short.json"):
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["detail"]["solver_seconds_per_step"], d["detail"]["setup_seconds_per_step"])
PY
