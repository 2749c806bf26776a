#!/bin/bash
# Final single-GPU evidence of the build (run under gpurun; results land in gpurun_out/ev7_*).
python -m pytest tests -x -q -m gpu > gpurun_out/ev7_pytest.log 2>&1; tail -2 gpurun_out/ev7_pytest.log
python tools/gpu_check.py --tile-n 256 > gpurun_out/ev7_check.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/ev7_bench.json 2> gpurun_out/ev7_bench.err
python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/ev7_bench_300.json 2>> gpurun_out/ev7_bench.err
python bench.py --size 608 --batch 40 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ev7_bench_608_b40.json 2>> gpurun_out/ev7_bench.err
python bench.py --size 608 --batch 160 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ev7_bench_608_b160.json 2>> gpurun_out/ev7_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ev7_bench_ref.json 2>> gpurun_out/ev7_bench.err
bash tools/trace_forward.sh > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev7_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ev7_ncu1.log 2>&1
cat > /tmp/fwd2.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from face_vijnana_yolov3_b200 import arch, synth
from face_vijnana_yolov3_b200.engine import Engine
import torch
eng = Engine(416, 416, nb_class=1, max_batch=40)
eng.load_weights(synth.darknet_stream(arch.yolo3_table(1), 0, synth.INIT_KERAS_DEFAULT))
xd = torch.from_numpy(synth.images(40, 416, 416, 1)).cuda()
for _ in range(2):
    eng.forward(xd, want_outputs=False)
PY
FVY_GRAPH=0 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"conv_igemm|conv_chain|stem_" -c 200 --csv --log-file gpurun_out/ev7_traffic.csv python /tmp/fwd2.py > gpurun_out/ev7_ncu4.log 2>&1
FVY_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:stem_strip -s 1 -c 1 -o gpurun_out/ev7_stem -f python /tmp/fwd2.py > gpurun_out/ev7_ncu5.log 2>&1
ncu -i gpurun_out/ev7_stem.ncu-rep --page raw --csv > gpurun_out/ev7_stem_raw.csv 2>/dev/null
rm -f gpurun_out/ev7_stem.ncu-rep
cut -c1-700 gpurun_out/ev7_bench.json; echo; for f in 300 608_b40 608_b160 ref; do cut -c1-160 gpurun_out/ev7_bench_$f.json; echo; done; tail -2 gpurun_out/ev7_check.log; cat gpurun_out/ev7_bench.err | tail -5
