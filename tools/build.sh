#!/bin/bash
# Builds libfvy.so (sm_100a only) and the C oracle in-tree.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
     -Xcompiler -fPIC -shared ${FVY_NVCC_EXTRA} \
     -o face_vijnana_yolov3_b200/libfvy.so face_vijnana_yolov3_b200/csrc/fvy_api.cu
make -s -C oracle -B liboracle_postproc.so
echo "built face_vijnana_yolov3_b200/libfvy.so oracle/liboracle_postproc.so"
