#!/bin/bash
# pipelined pair kernel: block parity per cluster size, traces, then net tests + bench
mkdir -p gpurun_out
for cl in 1 2 4; do
  echo "=== nb1d blocks CL=$cl"; MDIL_TC_CLUSTER=$cl timeout -s KILL 240 python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "nb1d" -p no:cacheprovider -x 2>&1 | tail -6
done
for cl in 1 2 4; do
  echo "=== trace CL=$cl"; MDIL_TC_CLUSTER=$cl MDIL_TC_TRACE=1 timeout -s KILL 120 python tools/trace_tc.py 2>&1 | grep -E "pair_tc3|Error|error" | head -12
done
echo "=== net"; timeout -s KILL 300 python -m pytest tests/test_gpu_net.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"; timeout -s KILL 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for cl in 1 2 4; do
echo "=== bench CL=$cl"; MDIL_TC_CLUSTER=$cl timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_tc3_cl$cl.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['per_kind_ms_per_step'])"; tail -3 gpurun_out/bench.err
done
