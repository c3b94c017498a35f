for cl in 1 2 4; do echo "== CL=$cl"; MDIL_TC_CLUSTER=$cl timeout -s KILL 90 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['per_kind_ms_per_step']; print(round(d['value'],1), round(d['ms_per_step'],2), {a[10:]:round(b,2) for a,b in k.items() if '16' not in a})"; done
