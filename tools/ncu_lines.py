"""Stall samples per CUDA source line from `ncu -i rep --page source --csv --print-source cuda,sass` (needs --import-source on
and -lineinfo).  usage: python tools/ncu_lines.py cuda_sass.csv [file-substring] [top_n]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cur_file, hdr, ix = None, None, None
per = collections.defaultdict(lambda: [0, collections.Counter(), "", 0])
first_kernel = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]; continue
    if r[0] == "Function Name":
        if first_kernel is None: first_kernel = r[1]
        cur_fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ix = {}
        for i, n in enumerate(hdr):
            ix.setdefault(n, i)
        stall = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None or len(r) < len(hdr) or cur_fn != first_kernel:
        continue
    if want and want not in (cur_file or ""):
        continue
    if not r[0].isdigit():
        continue
    line = int(r[0])
    s = r[ix["# Samples"]]
    if not s.isdigit():
        continue
    key = (cur_file.split("/")[-1], line)
    e = per[key]
    e[0] += int(s)
    e[2] = r[1].strip()[:90]
    ie = r[ix["Instructions Executed"]]
    e[3] += int(ie) if ie.isdigit() else 0
    for n in stall:
        v = r[ix[n]]
        if v.isdigit() and int(v): e[1][n[6:]] += int(v)
tot = sum(e[0] for e in per.values())
print("kernel:", first_kernel, " total samples:", tot)
for key, e in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{e[0]:6d} {100*e[0]/max(tot,1):5.1f}%  inst {e[3]:9d}  {key[0]}:{key[1]:<5d} {e[2]:90s} {dict(e[1].most_common(3))}")
