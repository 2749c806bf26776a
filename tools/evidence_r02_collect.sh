#!/bin/bash
# Copies the summaries of tools/evidence_r02.sh / tools/r2_run_multi.sh from gpurun_out/ (scratch) into profiles/ (tracked).
cd "$(dirname "$0")/.."
T=${1:-r2e}
for f in bench bench_608x320 bench_stress bench_train bench_ref; do [ -s gpurun_out/${T}_$f.json ] && cp gpurun_out/${T}_$f.json profiles/r02_$f.json; done
[ -s gpurun_out/${T}_pytest.log ] && cp gpurun_out/${T}_pytest.log profiles/r02_pytest_gpu.log
[ -s gpurun_out/${T}_check.log ] && cp gpurun_out/${T}_check.log profiles/r02_layer_times.log
[ -s gpurun_out/${T}_launches.csv ] && python tools/summarize_ncu.py launches gpurun_out/${T}_launches.csv profiles/r02_launches.csv
[ -s gpurun_out/${T}_conv_traffic.json ] && cp gpurun_out/${T}_conv_traffic.json profiles/conv_traffic.json && cp gpurun_out/${T}_conv_traffic_per_launch.csv profiles/r02_conv_traffic_per_launch.csv
[ -s gpurun_out/${T}_full_raw.csv ] && python tools/summarize_ncu.py full gpurun_out/${T}_full_raw.csv profiles/r02_ncu_full_kernel.csv
[ -s gpurun_out/${T}_chain_raw.csv ] && python tools/summarize_ncu.py full gpurun_out/${T}_chain_raw.csv profiles/r02_ncu_full_chain.csv
[ -s gpurun_out/${T}_post_raw.csv ] && python tools/summarize_ncu.py full gpurun_out/${T}_post_raw.csv profiles/r02_ncu_full_post.csv
for f in gpurun_out/r2m_*.json; do [ -s "$f" ] && cp "$f" profiles/r02_$(basename $f | sed 's/^r2m_//'); done
[ -s gpurun_out/r2m_sharded.log ] && cp gpurun_out/r2m_sharded.log profiles/r02_sharded_identity_multi_gpu.log
ls profiles | grep r02
