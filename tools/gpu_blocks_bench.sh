#!/bin/bash
# fast iteration pass: nb1d block parity + smoke + short bench (per-kind pair-kernel times) [+ ncu launch list with "ncu"]
mkdir -p gpurun_out
echo "=== blocks"; timeout -s KILL 300 python -m pytest tests/test_gpu_blocks.py -m gpu -q -p no:cacheprovider --tb=line -k "nb1d" 2>&1 | grep -v "^hi" | tail -8
echo "=== smoke"; timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | grep -v "^hi" | tail -1
echo "=== bench"; timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); n={'C=16':2,'C=64':7,'C=128':8}
print({k[10:]: round(v*1000/n[k[10:].split('>')[0]],1) for k,v in d['roofline']['per_kind_ms_per_step'].items()})"; tail -3 gpurun_out/bench.err | cut -c1-300
if [ "$1" = "ncu" ]; then
echo "=== ncu launch list"; timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
python tools/launch_summary.py gpurun_out/launches.csv 30 | tee gpurun_out/launch_summary.txt
fi
