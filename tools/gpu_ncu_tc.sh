#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:pair_tc_kernel -c 8 -o gpurun_out/pair_tc_full python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_pair_tc.log 2>&1
tail -2 gpurun_out/ncu_pair_tc.log
