#!/bin/bash
# ncu --set full of fused pair kernel launches inside one bench step: $1 = kernel-name regex (base name), $2 = skip, $3 = count, $4 = output name
# launch order of pair_h3_kernel in a step-1 iteration: 0-9 C=64 fwd (encoder), 10-25 C=128 fwd, 26-29 C=64 fwd (decoder),
# 30-33 C=64 bwd (decoder), 34-49 C=128 bwd, 50-59 C=64 bwd (encoder)
mkdir -p gpurun_out
timeout -s KILL 500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k "regex:$1" -s $2 -c $3 -o gpurun_out/$4 -f python bench.py --ncu-step --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_$4.log 2>&1
tail -2 gpurun_out/ncu_$4.log
python tools/ncu_summary.py gpurun_out/$4.ncu-rep gpurun_out/$4.md | tail -16
