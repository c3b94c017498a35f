#!/bin/bash
# ncu --set full capture of the fused pair kernel (C=128 and C=64 variants) inside one training step.
mkdir -p gpurun_out
timeout -s KILL 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:pair_kernel -c 12 -o gpurun_out/pair_full python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log; ls -la gpurun_out/*.ncu-rep
echo "=== step2 bench"; timeout -s KILL 900 python bench.py --workload step2 --steps 5 --warmup 3 2>gpurun_out/bench2.err | tee gpurun_out/bench_step2.json; tail -3 gpurun_out/bench2.err
