#!/bin/bash
# Round-end GPU pass: parity, smoke, bench (with the CPU baseline), step-2 bench, ncu launch list, ncu --set full of the
# dominant kernels (summaries only: gpurun_out is capped at 64 MiB).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout -s KILL 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"; timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | grep -v "^hi" | tail -2
echo "=== bench step1"; timeout -s KILL 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench_step1.json | cut -c1-300; tail -2 gpurun_out/bench.err
echo "=== bench step2"; timeout -s KILL 300 python bench.py --workload step2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench2.err | tee gpurun_out/bench_step2.json | cut -c1-200; tail -2 gpurun_out/bench2.err
echo "=== bench reference arm"; timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_reference.json | cut -c1-300
echo "=== ncu launch list"; timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
python tools/launch_summary.py gpurun_out/launches.csv 40 > gpurun_out/launch_summary.txt; head -12 gpurun_out/launch_summary.txt
cap() { name=$1; regex=$2; skip=$3; cnt=$4; src=$5
  timeout -s KILL 400 ncu --profile-from-start off --set full --clock-control none $src \
    -k regex:$regex -s $skip -c $cnt -o gpurun_out/$name -f python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.md > /dev/null 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null; }
cap tc3_c64_fwd pair_tc3_kernel 0 2 "--import-source on"
cap tc3_c128_fwd pair_tc3_kernel 10 2 ""
cap tc3_c128_bwd pair_tc3_kernel 34 2 ""
cap tc3_c64_bwd pair_tc3_kernel 58 2 "--import-source on"
cap wgrad_tc wgrad_tc_kernel 0 4 ""
cap small "conv_mma|wgrad_mma|pair_kernel" 0 12 ""
rm -f gpurun_out/tc3_c128_fwd.ncu-rep gpurun_out/tc3_c128_bwd.ncu-rep gpurun_out/wgrad_tc.ncu-rep gpurun_out/small.ncu-rep
du -sh gpurun_out
