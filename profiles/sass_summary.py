"""SASS evidence for the shipped library: per-kernel counts of the mnemonics that prove the Blackwell paths
(UTCIMMA/UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,
SYNCS = mbarrier, VABSDIFF(4), IDP) plus an excerpt of the instructions around the first MMA issue of each
tensor-core kernel.  Usage: python profiles/sass_summary.py [libcucudecide.so] > profiles/rNN_sass_rmd_tc2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fast-cu-decision-hevc_b200", "libcucudecide.so")
KEYS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "VABSDIFF4", "VABSDIFF", "IDP",
        "PRMT", "FADD", "STL", "LDL"]

txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur is not None and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).strip())

demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip() or s
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a)\n")
print("| kernel | instructions | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for f, ins in funcs.items():
    ops = collections.Counter()
    for l in ins:
        t = re.sub(r"^/\*[0-9a-f]+\*/\s*", "", l).split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        op = op.split(".")[0].rstrip(";")
        ops[op] += 1
    name = re.sub(r"\(.*", "", demangle(f).replace("(anonymous namespace)::", "")).replace("cucd::", "").replace("void ", "")
    print(f"| `{name}` | {len(ins)} | " + " | ".join(str(ops.get(k, 0)) for k in KEYS) + " |")

for f, ins in funcs.items():
    idx = [i for i, l in enumerate(ins) if "UTCIMMA" in l or "UTCHMMA" in l]
    if not idx or "frame" not in f:
        continue
    print(f"\n## {demangle(f)[:120]}: around the first tcgen05.mma issue\n")
    a = max(0, idx[0] - 14)
    for l in ins[a:idx[0] + 14]:
        print("    " + l)
    ld = [i for i, l in enumerate(ins) if "LDTM" in l]
    if ld:
        print("\n   ... first TMEM load and the instructions behind it\n")
        for l in ins[ld[0] - 2:ld[0] + 12]:
            print("    " + l)
