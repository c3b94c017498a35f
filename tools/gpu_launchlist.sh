#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv
