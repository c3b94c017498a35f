#!/bin/bash
# same-box A/B of an environment switch: bash tools/gpu_ab.sh VAR A B  (two alternating bench runs each)
VAR=$1; A=$2; B=$3
mkdir -p gpurun_out
for rep in 1 2; do
for v in $A $B; do
  echo "=== $VAR=$v"
  env $VAR=$v timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); n={'C=16':2,'C=64':7,'C=128':8}
print({k[10:]: round(v*1000/n[k[10:].split('>')[0]],1) for k,v in d['roofline']['per_kind_ms_per_step'].items()})"
done; done
