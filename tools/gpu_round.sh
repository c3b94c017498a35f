#!/bin/bash
# Full GPU pass: parity tests, smoke, bench line, ncu launch list of one step.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15
echo "=== smoke"; timeout -s KILL 300 python __graft_entry__.py --smoke 2>&1 | grep -v "^hi" | tail -5
echo "=== bench"; timeout -s KILL 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "=== ncu launch list"; timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -3 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv
