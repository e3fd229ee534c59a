"""Per-kernel SASS instruction histogram of libmps_b200.so (cuobjdump -sass): which kernels use DMMA (FP64 tensor), UBLKCP (bulk copy
on the TMA engine), UTMALDG (tensor-map TMA), cluster barriers, local-memory spills.  usage: python scripts/sass_histogram.py > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "tnqvm_b200", "lib", "libmps_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
cur, hist, order = None, {}, []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); order.append(cur); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "ATOM", "RED", "MUFU"]
print("kernel | total | " + " | ".join(KEYS))
for k in order:
    h = hist[k]
    tot = sum(h.values())
    fam = collections.Counter()
    for op, c in h.items():
        fam[op.split(".")[0]] += c
    name = demangle(k).replace("(anonymous namespace)::", "").replace("mpsb200::", "").replace("void ", "")
    name = name.split("(")[0]
    print("%s | %d | %s" % (name, tot, " | ".join(str(fam.get(x, 0)) for x in KEYS)))
