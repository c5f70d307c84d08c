"""Per-kernel counts of the SASS mnemonics that show what the kernels are built from (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): bulk (TMA) copies, async stores with mbarrier completion, cluster barriers, warp match /
vote / shuffle, shared-memory atomics.      python profiles/sass_summary.py > profiles/r02/sass_summary.txt"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(REPO, "mean-field-multi-agent-reinforcement-learning_b200", "build", "libmagent.so")
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "STAS", "SYNCS", "UCGABAR", "MATCH", "VOTE", "SHFL", "REDUX", "ATOMS", "ATOMG",
         "RED", "MUFU.EX2", "BAR.SYNC", "LDG", "STG", "LDS", "STS", "CCTL", "MEMBAR", "ERRBAR"]

out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                kernels[cur][w] += 1
print("SASS of %s (cuobjdump -sass), instruction counts per kernel" % os.path.relpath(SO, REPO))
print("%-58s %6s  %s" % ("kernel", "instr", "of which"))
for name, c in kernels.items():
    print("%-58s %6d  %s" % (name[:58], c["_total"], "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])))
