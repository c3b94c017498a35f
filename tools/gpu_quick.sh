#!/bin/bash
# quick GPU pass: all gpu tests + bench line (no cpu baseline) + launch summary
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout -s KILL 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
echo "=== bench"; timeout -s KILL 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_kind_ms_per_step'])"; tail -3 gpurun_out/bench.err | cut -c1-300
if [ "$1" = "ncu" ]; then
echo "=== ncu launch list"; timeout -s KILL 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; tail -2 gpurun_out/ncu_launches.log
python tools/launch_summary.py gpurun_out/launches.csv 30 | tee gpurun_out/launch_summary.txt
fi
