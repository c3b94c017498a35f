#!/bin/bash
# round-2 call A: sharpened parity tests at the benchmark shapes, smoke, bench with the PyTorch-eager (cuDNN) baseline,
# per-role wait counters of the pipelined tensor-core kernels
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40
echo "=== parity numbers"; cat gpurun_out/parity.jsonl
echo "=== smoke"; timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | grep -v "^hi" | tail -3
echo "=== bench step1"; timeout -s KILL 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench_step1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_baseline'), d.get('cpu_baseline'))"; tail -3 gpurun_out/bench.err | cut -c1-300
echo "=== trace"; MDIL_TC_TRACE=1 timeout -s KILL 120 python tools/trace_tc.py 2>&1 | grep -v "^hi" | tail -40
