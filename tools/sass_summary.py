"""Per-kernel SASS evidence of the built library: counts of the Blackwell-native mnemonics (B200_PROFILING.md) per kernel.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "deformablelka_b200", "libdlka_b200.so")
MN = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "LDG", "LDS", "STS", "SHFL", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if cur and m:
        op = m.group(1)
        for k in MN:
            if op == k or (k in ("LDG", "LDS", "STS", "SHFL") and op.startswith(k)) or (k == "FFMA" and op == "FFMA"):
                counts[cur][k] += 1
                break
        counts[cur]["_total"] += 1
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
print(f"# cuobjdump -sass deformablelka_b200/libdlka_b200.so (git {head}): SASS instruction counts per kernel")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops;")
print("# HMMA (legacy mma.sync) must be 0 everywhere")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in MN) + "   total")
tot = collections.Counter()
for k, c in counts.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(anonymous namespace\)::|dlka::", "", name)
    name = name.split("(")[0][:70]
    print(f"{name:70s} " + " ".join(f"{c[m]:8d}" for m in MN) + f" {c['_total']:7d}")
    tot.update(c)
print(f"{'ALL':70s} " + " ".join(f"{tot[m]:8d}" for m in MN) + f" {tot['_total']:7d}")
