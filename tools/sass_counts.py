"""Instruction evidence from the built library: python tools/sass_counts.py > profiles/sass_r2_excerpts.txt
Counts, per kernel of active_tracking_rl_b200/libtrack2d.so (cuobjdump -sass), the mnemonics that prove which hardware path runs."""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "active_tracking_rl_b200", "libtrack2d.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
WATCH = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "LDGSTS", "FFMA2",
         "HMMA", "DFMA", "DADD", "DMUL", "LDS.128", "STS.128")
print("# cuobjdump -sass active_tracking_rl_b200/libtrack2d.so (built from HEAD), instruction counts per kernel (static, not executed counts)")
print("# UTCHMMA = tcgen05.mma (kind::tf32), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS.* = mbarrier ops,")
print("# LDGSTS = cp.async, FFMA2 = packed fp32 FMA, D* = fp64 (rewards, A* keys); no UTMALDG/UTMASTG (no TMA), no HMMA (no legacy mma.sync)")
blocks = re.split(r"\n\s*Function : ", txt)[1:]
rows = []
for name, blk in zip(names, blocks):
    cnt = collections.Counter()
    for line in blk.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cnt[w] += 1
    short = re.sub(r"\(anonymous namespace\)::", "", name)
    short = re.sub(r"\(.*$", "", short)
    rows.append((short, cnt))
for short, cnt in sorted(rows):
    if cnt:
        print("%-72s %s" % (short[:72], "  ".join("%s=%d" % (k, v) for k, v in sorted(cnt.items()))))
