#!/bin/bash
# 8-GPU box: BASELINE configs at their stated GPU counts + the weak-scaling points of configs[1] (run with gpurun --gpus 8)
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((n-1))) timeout -k 10 -s TERM 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n \
    --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" 2>gpurun_out/$name.err | grep "^{" > gpurun_out/$name.json
  python -c "
import json,sys
d=json.load(open('gpurun_out/$name.json')); print('$name', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d.get('cuda_graph'), d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 gpurun_out/$name.err | cut -c1-300; }
echo '=== multi-GPU tests'; timeout -k 10 -s TERM 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_multi.txt
run bench_step1_2gpu 2 --steps 20 --warmup 5
run bench_step1_4gpu 4 --steps 20 --warmup 5
run bench_step1_8gpu 8 --steps 20 --warmup 5
run bench_step2_2gpu 2 --workload step2 --steps 10 --warmup 3
run bench_step3_8gpu 8 --workload step3 --batch 3 --steps 10 --warmup 3
run bench_multitask_1024x2048_8gpu 8 --workload multitask --full-res --batch 4 --steps 6 --warmup 3
