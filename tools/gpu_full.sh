#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout -s KILL 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6
echo "=== bench"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_tc.json | cut -c1-400; tail -5 gpurun_out/bench.err
