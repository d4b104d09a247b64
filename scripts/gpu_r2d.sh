#!/bin/bash
mkdir -p gpurun_out/r2d
( time timeout 600 scripts/_bin/spmv_lab 10000000 0 ) > gpurun_out/r2d/spmv_lab.txt 2>&1
# launch list of the production kernels in situ on the SAME box (cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 300 --csv --log-file gpurun_out/r2d/launches_c4.csv \
   python scripts/exp_kernel_variants.py c4 "" > gpurun_out/r2d/variants_under_ncu.txt 2>&1
( time timeout 300 python scripts/exp_kernel_variants.py c4 "" "CUOPT_B200_L2_HINTS=0" ) > gpurun_out/r2d/variants_c4.txt 2>&1
grep -E "base|production|bicsr ldg \+prefetch  |==" gpurun_out/r2d/spmv_lab.txt
cat gpurun_out/r2d/variants_c4.txt
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2d/launches_c4.csv')) if len(r)>5 and r[0].isdigit()]
acc=collections.defaultdict(list)
for r in rows:
    name=r[4]; val=float(r[-1].replace(',',''))
    acc[name.split('(')[0]].append(val)
for k,v in sorted(acc.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k[:70]:70s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.1f} us")
PY
