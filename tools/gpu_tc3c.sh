#!/bin/bash
mkdir -p gpurun_out
echo "=== net"; timeout -s KILL 300 python -m pytest tests/test_gpu_net.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
for d in 0 3; do echo "== DBG=$d"; MDIL_TC3_DBG=$d MDIL_TC_TRACE=1 timeout -s KILL 120 python tools/trace_tc.py 2>&1 | grep -E "pair_tc3" | sed -n '1p;3p;5p;7p' | cut -c1-330; done
echo "=== bench"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_tc3.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['per_kind_ms_per_step'])"; tail -3 gpurun_out/bench.err | cut -c1-300
