#!/bin/bash
# First-contact GPU run: each group in its own process (a sticky CUDA error must not cascade) with its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout -s KILL 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run head    python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "output_conv or cpu_tensor" -p no:cacheprovider
run down    python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "downsampler" -p no:cacheprovider
run up      python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "upsampler" -p no:cacheprovider
run nb1d    python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "nb1d" -p no:cacheprovider
run net     python -m pytest tests/test_gpu_net.py -q -m gpu -p no:cacheprovider
run smoke   python __graft_entry__.py --smoke
