#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:pair_tc_kernel -c 4 -o gpurun_out/pair_tc_v2 python tools/trace_tc.py > gpurun_out/ncu_pair_tc.log 2>&1
tail -2 gpurun_out/ncu_pair_tc.log
