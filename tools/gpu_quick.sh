#!/bin/bash
mkdir -p gpurun_out
echo "=== nb1d blocks"; timeout -s KILL 300 python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "nb1d" -p no:cacheprovider -x 2>&1 | tail -15
echo "=== net"; timeout -s KILL 300 python -m pytest tests/test_gpu_net.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "=== bench"; timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_tc.json; tail -5 gpurun_out/bench.err
