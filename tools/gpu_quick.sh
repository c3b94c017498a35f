#!/bin/bash
# quick GPU pass: all gpu tests + smoke + bench line (no baselines) [+ ncu launch list with "ncu"]
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
echo "=== pytest -m gpu"; timeout -s KILL 600 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=line 2>&1 | grep -v "^hi" | tail -25
echo "=== parity"; cat gpurun_out/parity.jsonl 2>/dev/null | cut -c1-330
echo "=== smoke"; timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | grep -v "^hi" | tail -2
echo "=== bench"; timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_kind_ms_per_step'])"; tail -3 gpurun_out/bench.err | cut -c1-300
if [ "$1" = "ncu" ]; then
echo "=== ncu launch list"; timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
python tools/launch_summary.py gpurun_out/launches.csv 30 | tee gpurun_out/launch_summary.txt
fi
