#!/bin/bash
# Round-2 evidence pass (one GPU): parity, smoke, bench lines (step-1 with CPU and PyTorch-eager baselines, step-2, step-3,
# multi-task full-res, reference arm), ncu launch list of one step, ncu --set full of every fused-pair kind and of the
# weight-gradient kernel (summaries + raw csv; gpurun_out is capped at 64 MiB), per-role wait counters.
mkdir -p gpurun_out; rm -f gpurun_out/parity.jsonl gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "=== smoke"; timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | grep -v "^hi" | tail -2 | tee gpurun_out/smoke.txt
echo "=== bench step1"; timeout -s KILL 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench_step1.json | cut -c1-300; tail -2 gpurun_out/bench.err
echo "=== bench step2"; timeout -s KILL 300 python bench.py --workload step2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench2.err | tee gpurun_out/bench_step2.json | cut -c1-200
echo "=== bench step3"; timeout -s KILL 300 python bench.py --workload step3 --batch 3 --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench2.err | tee gpurun_out/bench_step3.json | cut -c1-200
echo "=== bench multitask full-res"; timeout -s KILL 300 python bench.py --workload multitask --full-res --batch 4 --steps 6 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench2.err | tee gpurun_out/bench_multitask_1024x2048.json | cut -c1-200
echo "=== bench reference arm"; timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/bench_reference.json | cut -c1-300
echo "=== ncu launch list"; timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log
python tools/launch_summary.py gpurun_out/launches.csv 40 > gpurun_out/launch_summary.txt; head -14 gpurun_out/launch_summary.txt
cap() { name=$1; regex=$2; skip=$3; cnt=$4; keep=$5
  timeout -s KILL 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k "regex:$regex" -s $skip -c $cnt -o gpurun_out/$name -f python bench.py --ncu-step --no-cpu-baseline --no-gpu-baseline > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.md > /dev/null 2>&1
  if [ "$keep" != "keep" ]; then rm -f gpurun_out/$name.ncu-rep; fi; }
# launch order of pair_h3_kernel in a step-1 iteration: 0-9 C=64 fwd (encoder), 10-25 C=128 fwd, 26-29 C=64 fwd (decoder),
# 30-33 C=16 fwd (packed-4 view), 34-37 C=16 bwd, 38-41 C=64 bwd (decoder), 42-57 C=128 bwd, 58-67 C=64 bwd (encoder)
cap h3_c64_fwd pair_h3_kernel 0 2 keep
cap h3_c128_fwd pair_h3_kernel 10 2 drop
cap h3_c16_fwd pair_h3_kernel 30 2 drop
cap h3_c16_bwd pair_h3_kernel 34 2 drop
cap h3_c128_bwd pair_h3_kernel 42 2 drop
cap h3_c64_bwd pair_h3_kernel 58 2 keep
cap wgrad_tc wgrad_tc_kernel 0 8 drop
cap conv_tc "^conv_tc_kernel" 0 9 drop
cap small "bn_act_fused|stats_kernel" 0 12 drop
cap head "outconv|ce2d_phase" 0 4 drop
cap bn_bwd "bn_bwd_apply_fused|stats_kernel|bn_act_fused" 30 8 drop
echo "=== trace"; MDIL_TC_TRACE=1 timeout -s KILL 120 python tools/trace_tc.py 2>&1 | grep -v "^hi" | grep -A3 "pair_h3\|wgrad_tc" > gpurun_out/trace_counters.txt; head -8 gpurun_out/trace_counters.txt
du -sh gpurun_out
