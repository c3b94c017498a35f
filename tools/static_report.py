"""Static evidence for the shipped library (no GPU needed): per-kernel registers / spills / shared memory from
`cuobjdump --dump-resource-usage`, and counts of the SASS mnemonics that prove tensor-core (tcgen05 = UTC*), TMA / bulk
copy (UBLKCP / UTMA*), mbarrier (SYNCS) and warp-level MMA (HMMA/ tf32 MMA) use.

    python tools/static_report.py > profiles/r1_static_resources.md
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB = os.path.join(REPO, "mdil_ss_b200", "libmdil_b200.so")
PATTERNS = collections.OrderedDict([
    ("UTCMMA/UTC* (tcgen05.mma, commit, ld/st, alloc)", re.compile(r"\b(UTC[A-Z0-9]+|LDTM|STTM)\b")),
    ("UBLKCP / UTMA* (cp.async.bulk, TMA)", re.compile(r"\b(UBLKCP|UTMALDG|UTMASTG|UBLKRED)\b")),
    ("SYNCS (mbarrier)", re.compile(r"\bSYNCS\b")),
    ("HMMA / MMA (mma.sync tf32)", re.compile(r"\b(HMMA|IMMA|DMMA)\b")),
    ("RED / ATOM (global reductions)", re.compile(r"\b(RED|ATOM|ATOMG|REDG)\b")),
])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("void ", "")
    depth, out = 0, []
    for ch in name:                       # cut the argument list: the first "(" outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out).strip()


def ptxas_spills():
    """{mangled name: (spill stores, spill loads)} from `nvcc -Xptxas -v` over the library's sources (objects go to a
    temporary directory; the shipped build is untouched)."""
    import concurrent.futures
    import tempfile
    sys.path.insert(0, REPO)
    from mdil_ss_b200 import build
    csrc = os.path.join(REPO, "mdil_ss_b200", "csrc")
    tmp = tempfile.mkdtemp(prefix="mdil_ptxas_")

    def one(src):
        cmd = ["nvcc"] + list(build.NVCC_FLAGS) + ["-Xptxas", "-v", "-c", os.path.join(csrc, os.path.basename(src)), "-o", os.path.join(tmp, os.path.basename(src) + ".o")]
        return subprocess.run(cmd, capture_output=True, text=True).stderr

    out = {}
    with concurrent.futures.ThreadPoolExecutor(8) as ex:
        for text in ex.map(one, build.SOURCES):
            cur = None
            for line in text.splitlines():
                m = re.search(r"Compiling entry function '(\S+)'", line)
                if m:
                    cur = m.group(1)
                m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
                if m and cur:
                    out[cur] = (int(m.group(1)), int(m.group(2)))
    return out


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    rows = []
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            rows.append((fn,) + tuple(int(x) for x in m.groups()))
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(lambda: collections.Counter())
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        for label, pat in PATTERNS.items():
            if pat.search(line):
                counts[cur][label] += 1
    names = demangle([r[0] for r in rows])
    spills = ptxas_spills()
    print("# Static resources and SASS evidence of `mdil_ss_b200/libmdil_b200.so` (sm_100a)\n")
    print("Produced by `tools/static_report.py` from `cuobjdump --dump-resource-usage` and `cuobjdump -sass` of the shipped build.")
    print("Spill bytes come from `nvcc -Xptxas -v` over the same sources and flags; a stack frame without spills is an indexed")
    print("local array. Static smem excludes the dynamic shared memory requested at launch. The last five columns count SASS")
    print("instructions by family: tcgen05 (UTC*: mma / commit / alloc, LDTM/STTM: tcgen05.ld/st), bulk copies (UBLKCP), mbarrier")
    print("(SYNCS), warp-level `mma.sync` (HMMA family) and global reductions (RED/ATOM).\n")
    print("| kernel | regs | stack B | spill st/ld B | static smem B | tcgen05 | bulk/TMA | mbarrier | mma.sync | red/atom |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for fn, reg, stack, shared, local in sorted(rows, key=lambda r: short(names[r[0]])):
        c = counts.get(fn, {})
        cells = [str(c.get(label, 0)) for label in PATTERNS]
        st, ld = spills.get(fn, (0, 0))
        print(f"| `{short(names[fn])}` | {reg} | {stack} | {st}/{ld} | {shared} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    sys.exit(main())
