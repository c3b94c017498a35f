#!/bin/bash
# Session-3 first contact: parity, smoke, bench line (with cpu baseline), per-phase traces of the TC pair kernel, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== smoke"; timeout -s KILL 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "=== trace"; MDIL_TC_TRACE=1 timeout -s KILL 300 python tools/trace_tc.py 2>&1 | tee gpurun_out/trace_tc.log | tail -40
echo "=== bench"; timeout -s KILL 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-600; tail -5 gpurun_out/bench.err
echo "=== ncu launch list"; timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -3 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv
python tools/launch_summary.py gpurun_out/launches.csv 40 | tee gpurun_out/launch_summary.txt
