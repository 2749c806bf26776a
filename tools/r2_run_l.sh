#!/bin/bash
# per-kernel times of the stress configuration (10 647 candidates per image)
python bench.py --config stress --steps 2 --warmup 3 > gpurun_out/r2l_stress_plain.json 2> gpurun_out/r2l_stress.err || tail -3 gpurun_out/r2l_stress.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2l_stress_launches.csv \
    python bench.py --config stress --steps 2 --warmup 3 > gpurun_out/r2l_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2l_stress_launches.csv")) if len(r) > 5 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:40], r[-1], r[-2] if len(r) > 2 else "")
PY
ncu --set full --clock-control none --import-source on -k regex:"nms_mask" -s 3 -c 1 -o gpurun_out/r2l_mask -f python bench.py --config stress --steps 2 --warmup 3 > gpurun_out/r2l_ncu2.log 2>&1
ncu -i gpurun_out/r2l_mask.ncu-rep --page raw --csv > gpurun_out/r2l_mask_raw.csv 2>/dev/null
ncu -i gpurun_out/r2l_mask.ncu-rep --page source --csv > gpurun_out/r2l_mask_source.csv 2>/dev/null; rm -f gpurun_out/r2l_mask.ncu-rep
ls -la gpurun_out/r2l_*
