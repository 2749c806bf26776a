#!/bin/bash
# Regenerates the measured evidence of the current build (run under gpurun; results land in gpurun_out/ev6_*).
set -x
python tools/gpu_check.py --tile-n 256 > gpurun_out/ev6_check.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/ev6_bench.json 2> gpurun_out/ev6_bench.err
python bench.py --steps 300 --warmup 5 > gpurun_out/ev6_bench_300.json 2>> gpurun_out/ev6_bench.err
FVY_DBG=1 python tools/run_layer.py --layers 1,3,4,6,8,9,10,11,27,28,45,58 --iters 10 > gpurun_out/ev6_timeline.log 2>&1
bash tools/trace_forward.sh > /dev/null 2>&1
# (1) launch list of a bench step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev6_launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ev6_ncu1.log 2>&1
# (2) full capture of representative layers (isolated launches; the forward that fills the buffers runs un-chained so the skip count is fixed)
FVY_CHAIN=0 python tools/run_layer.py --layers 28,11,27,3,45,10,1,6 --iters 1 > gpurun_out/ev6_plain2.log 2>&1 || exit 1
FVY_CHAIN=0 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 74 -c 16 -o gpurun_out/ev6_full -f \
    python tools/run_layer.py --layers 28,11,27,3,45,10,1,6 --iters 1 > gpurun_out/ev6_ncu2.log 2>&1
ncu -i gpurun_out/ev6_full.ncu-rep --page raw --csv > gpurun_out/ev6_full_raw.csv 2>/dev/null
# (3) the chain kernel inside a forward, and DRAM traffic of every conv launch of one forward
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
FVY_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:conv_chain -s 3 -c 3 -o gpurun_out/ev6_chain -f python /tmp/fwd2.py > gpurun_out/ev6_ncu3.log 2>&1
ncu -i gpurun_out/ev6_chain.ncu-rep --page raw --csv > gpurun_out/ev6_chain_raw.csv 2>/dev/null
FVY_GRAPH=0 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"conv_igemm|conv_chain|stem_" -c 200 --csv --log-file gpurun_out/ev6_traffic.csv python /tmp/fwd2.py > gpurun_out/ev6_ncu4.log 2>&1
rm -f gpurun_out/ev6_full.ncu-rep
ls -la gpurun_out/ev6_*; cat gpurun_out/ev6_bench.json | cut -c1-600; tail -3 gpurun_out/ev6_check.log
