"""Per-source-line stall samples and shared-memory wavefronts from an ncu report (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [file-substring] [kernel-instance]"""
import csv, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ".cu"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = None; hdr = None; out = []; seen_files = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; seen_files[cur] = seen_files.get(cur, 0) + 1; continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; hdr_list = r; continue
    if r[0] == "Kernel Name": continue
    if hdr is None or cur is None or want not in cur or seen_files[cur] > 1: continue
    if r[0] != "" and len(r) >= len(hdr_list):
        out.append(r)
H = hdr
def I(x):
    try: return int(x)
    except ValueError: return 0
stalls = [n for n in hdr_list if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(I(r[H["# Samples"]]) for r in out)
print(f"total samples {tot}")
print("--- top lines by samples")
for r in sorted(out, key=lambda r: -I(r[H["# Samples"]]))[:40]:
    st = sorted([(I(r[H[s]]), s[6:]) for s in stalls], reverse=True)[:3]
    print(r[H["# Samples"]].rjust(6), ("L" + r[0]).rjust(5), r[1].strip()[:90].ljust(90), [x for x in st if x[0] > 0])
print("--- shared-memory wavefronts (actual / ideal) by line")
for r in sorted(out, key=lambda r: -I(r[H["L1 Wavefronts Shared"]]))[:20]:
    print(r[H["L1 Wavefronts Shared"]].rjust(9), r[H["L1 Wavefronts Shared Ideal"]].rjust(9), ("L" + r[0]).rjust(5), r[1].strip()[:100])
