"""Per-role stall breakdown of a warp-specialised kernel from `ncu --page source --csv` output (needs --import-source on).
usage: python tools/ncu_stalls.py source.csv [top_n]   -- prints barrier-wait loops (by mbarrier offset) and the hottest instructions."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_idx[0]]
ix = {n: i for i, n in enumerate(h)}
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
data = [r for r in rows[hdr_idx[0] + 1:end] if len(r) > ix["# Samples"]]
S, SRC = ix["# Samples"], ix["Source"]
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
def smp(r): return int(r[S]) if r[S].isdigit() else 0
tot = sum(smp(r) for r in data)
print("total samples", tot, "instructions", len(data))
c = collections.Counter()
for r in data:
    for n in stall_cols:
        if r[ix[n]].isdigit(): c[n[6:]] += int(r[ix[n]])
print("stall reasons:", c.most_common(10))
print("--- mbarrier wait loops (samples in the TRYWAIT + following 3 instructions)")
for i, r in enumerate(data):
    if "TRYWAIT" in r[SRC]:
        s = sum(smp(data[j]) for j in range(i, min(i + 4, len(data))))
        if s >= 5: print(f"  [{i}] {s:6d}  {r[SRC].strip()[:80]}")
print("--- hottest instructions")
for i in sorted(range(len(data)), key=lambda i: -smp(data[i]))[:topn]:
    r = data[i]
    st = {n[6:]: int(r[ix[n]]) for n in stall_cols if r[ix[n]].isdigit() and int(r[ix[n]]) > 0}
    print(f"  [{i}] {smp(r):6d}  {r[SRC].strip()[:70]:70s} {st}")
