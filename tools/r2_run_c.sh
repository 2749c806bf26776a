#!/bin/bash
# round 2, run C: pin down the crash in fvy_destroy (letterbox test) with a host-debug build, micro cases in separate processes
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O1 -std=c++17 -fmad=false -Xcompiler -fPIC -Xcompiler -g -shared \
     -o /tmp/libfvy_dbg.so face_vijnana_yolov3_b200/csrc/fvy_api.cu > gpurun_out/r2c_build.log 2>&1 || tail -5 gpurun_out/r2c_build.log
cat > /tmp/micro.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
from face_vijnana_yolov3_b200 import _lib as L
from face_vijnana_yolov3_b200.engine import Engine
case = sys.argv[1]
eng = Engine(416, 416, head=L.HEAD_FD6, max_batch=3)
rng = np.random.default_rng(8)
if case in ("lb", "lb_err", "lb_big"):
    im = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    eng.letterbox(im, 0, 416, 312, 52, 0)
    eng.staged_to_host(1)
if case == "lb_big":
    im = rng.integers(0, 256, (200, 1500, 3), dtype=np.uint8)
    eng.letterbox(im, 1, 416, 55, 180, 0)
if case == "lb_err":
    try:
        eng.letterbox(im, 3, 416, 312, 52, 0)
    except ValueError as e:
        print("expected:", e)
print(case, "closing", flush=True)
eng.close()
print(case, "closed OK", flush=True)
PY
for c in plain lb lb_big lb_err; do python /tmp/micro.py $c > gpurun_out/r2c_micro_$c.log 2>&1; echo "micro $c rc=$?"; tail -2 gpurun_out/r2c_micro_$c.log; done
FVY_LIB_PATH=/tmp/libfvy_dbg.so timeout 300 cuda-gdb -batch -ex "set pagination off" -ex run -ex bt --args python -m pytest tests/test_gpu_parity.py -q -m gpu -k "letterbox_gpu" -p no:cacheprovider > gpurun_out/r2c_gdb.log 2>&1
grep -n "SIGSEGV" -A12 gpurun_out/r2c_gdb.log | head -30
timeout 600 compute-sanitizer --tool memcheck --report-api-errors all python /tmp/micro.py lb > gpurun_out/r2c_memcheck.log 2>&1; grep -v "^=========     " gpurun_out/r2c_memcheck.log | tail -20
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "dropin" -p no:cacheprovider > gpurun_out/r2c_dropin.log 2>&1; grep -E "^E |Error|passed|failed" gpurun_out/r2c_dropin.log | head -20
