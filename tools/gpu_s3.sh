echo "=== pytest -m gpu"; timeout -s KILL 200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -6
echo "=== bench step3"; timeout -s KILL 100 python bench.py --workload step3 --batch 3 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench3.err | tee gpurun_out/bench_step3.json | cut -c1-260; tail -2 gpurun_out/bench3.err
echo "=== bench step1"; timeout -s KILL 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | cut -c1-200; tail -2 gpurun_out/bench.err
