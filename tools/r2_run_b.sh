#!/bin/bash
# round 2, run B: GPU suite per file (a crash in one file does not hide the others), gdb backtrace of the letterbox test, bench
for f in test_gpu_parity test_gpu_round2 test_train; do
  timeout 1200 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider > gpurun_out/r2b_$f.log 2>&1; echo "$f rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2b_$f.log | tail -25
done
timeout 300 cuda-gdb -batch -ex "set pagination off" -ex run -ex bt -ex "info sharedlibrary libfvy" --args python -m pytest tests/test_gpu_parity.py -q -m gpu -k "letterbox_gpu" -p no:cacheprovider > gpurun_out/r2b_gdb.log 2>&1; tail -40 gpurun_out/r2b_gdb.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; cut -c1-300 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
