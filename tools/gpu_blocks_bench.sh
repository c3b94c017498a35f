#!/bin/bash
# fast iteration pass: nb1d block parity + smoke + short bench (per-kind pair-kernel times)
echo "=== blocks"; timeout -s KILL 300 python -m pytest tests/test_gpu_blocks.py -m gpu -q -p no:cacheprovider --tb=line -k "nb1d" 2>&1 | grep -v "^hi" | tail -8
echo "=== smoke"; timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | grep -v "^hi" | tail -1
echo "=== bench"; timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k[10:]: round(v*1000/ (5 if 'C=64' in k and 'fwd' in k else 1),1) for k,v in d['roofline']['per_kind_ms_per_step'].items()})"; tail -3 gpurun_out/bench.err | cut -c1-300
