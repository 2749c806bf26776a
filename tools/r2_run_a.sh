#!/bin/bash
# round 2, run A: whole GPU suite, baseline bench, FVY_CHAIN_SCHED A/B (results in gpurun_out/r2a_*)
python -m pytest tests -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; tail -15 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; cut -c1-400 gpurun_out/r2a_bench.json
FVY_CHAIN_SCHED=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_sched.json 2>> gpurun_out/r2a_bench.err; cut -c1-300 gpurun_out/r2a_bench_sched.json
FVY_CHAIN_SCHED=1 python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2a_pytest_sched.log 2>&1; tail -3 gpurun_out/r2a_pytest_sched.log
tail -5 gpurun_out/r2a_bench.err
