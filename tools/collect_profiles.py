"""Copy the round's evidence from gpurun_out/ (scratch) into profiles/ (tracked) and rebuild profiles/pair_traffic.json
(measured dram__bytes_read.sum + dram__bytes_write.sum per launch of every fused-pair kind, from the ncu --set full
summaries) -- bench.py reports it as roofline.traffic.   usage: python tools/collect_profiles.py r2"""
import json, os, re, shutil, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, P = os.path.join(R, "gpurun_out"), os.path.join(R, "profiles")
copies = {"bench_step1.json": "bench_step1.json", "bench_step2.json": "bench_step2.json", "bench_step3.json": "bench_step3.json",
          "bench_multitask_1024x2048.json": "bench_multitask_1024x2048.json", "bench_reference.json": "bench_reference_arm.json",
          "launches.csv": "step1_launches.csv", "launch_summary.txt": "step1_launch_summary.txt", "gpu.txt": "gpu.txt",
          "parity.jsonl": "parity_pretrained_train_step.jsonl", "trace_counters.txt": "trace_counters.txt",
          "pytest_gpu.txt": "pytest_gpu.txt", "smoke.txt": "smoke.txt", "pytest_gpu_multi.txt": "pytest_gpu_multi.txt"}
for n in ("step1_2gpu", "step1_4gpu", "step1_8gpu", "step2_2gpu", "step3_8gpu", "multitask_1024x2048_8gpu"):
    copies[f"bench_{n}.json"] = f"bench_{n}.json"
for name in ("h3_c64_fwd", "h3_c64_bwd", "h3_c128_fwd", "h3_c128_bwd", "h3_c16_fwd", "h3_c16_bwd", "wgrad_tc", "conv_tc", "small", "head", "bn_bwd"):
    copies[name + ".md"] = "ncu_full_" + name + ".md"
for src, dst in copies.items():
    s = os.path.join(G, src)
    if os.path.exists(s):
        shutil.copyfile(s, os.path.join(P, f"{tag}_{dst}"))
def traffic(md):
    t = open(md).read()
    rd = re.search(r"dram_rd\s+([\d.]+) (\w+)", t); wr = re.search(r"dram_wr\s+([\d.]+) (\w+)", t)
    unit = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    return float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)]
out = {"_source": f"ncu --set full --clock-control none captures of round {tag} (profiles/{tag}_ncu_full_h3_*.md): first two launches of "
                  "each kind inside one step-1 iteration (forward: pair 1 and pair 2; backward: pair 2 and pair 1), averaged",
       "_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"}
for c, key in ((16, "c16"), (64, "c64"), (128, "c128")):
    for d, phases in (("fwd", ("fwd pair1", "fwd pair2")), ("bwd", ("bwd pair2", "bwd pair1"))):
        md = os.path.join(G, f"h3_{key}_{d}.md")
        if os.path.exists(md):
            for ph in phases:
                out[f"nb1d_pair<C={c}> {ph}"] = traffic(md)
json.dump(out, open(os.path.join(P, "pair_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
