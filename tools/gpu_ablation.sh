#!/bin/bash
# same-box ablation of the round's switches: step-1 bench (10 steps) with ONE switch thrown at a time
mkdir -p gpurun_out; : > gpurun_out/ablation.txt
one() { name=$1; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-34s %7.1f crops/s  %6.2f ms/step  e2e %7.1f  launches/step %d' % ('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']//d['steps']))" | tee -a gpurun_out/ablation.txt; }
one "default" MDIL_NOOP=1
one "MDIL_PREPACK=0" MDIL_PREPACK=0
one "MDIL_S16=0" MDIL_S16=0
one "MDIL_HEAD_FUSED=0" MDIL_HEAD_FUSED=0
one "MDIL_WGRAD_GATHER=0" MDIL_WGRAD_GATHER=0
one "MDIL_CONV_TC=0" MDIL_CONV_TC=0
one "MDIL_P4=0" MDIL_P4=0
one "all five off + MDIL_P4=0" MDIL_PREPACK=0 MDIL_S16=0 MDIL_HEAD_FUSED=0 MDIL_WGRAD_GATHER=0 MDIL_CONV_TC=0 MDIL_P4=0
one "MDIL_PAIR_IMPL=tc3 (round 1 kernel)" MDIL_PAIR_IMPL=tc3
one "default (again)" MDIL_NOOP=2
