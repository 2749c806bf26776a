#!/bin/bash
for f in test_gpu_parity test_gpu_round2; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider > gpurun_out/r2f_$f.log 2>&1; echo "$f rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2f_$f.log | tail -25
done
bash tools/r2_run_e.sh 2>&1 | grep -v "^\"" 
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2e_post_launches.csv', errors='replace')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); h=rows[hi]
iN,iV,iG=h.index("Kernel Name"),h.index("Metric Value"),h.index("Grid Size")
for r in rows[hi+1:hi+6]:
    print(r[iN].split("(")[0], r[iG], float(r[iV].replace(",",""))/1e3, "us")
PY
python bench.py --config stress --no-cpu-baseline > gpurun_out/r2f_stress.json 2> gpurun_out/r2f_stress.err; cut -c1-250 gpurun_out/r2f_stress.json; tail -3 gpurun_out/r2f_stress.err
