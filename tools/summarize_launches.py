"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, data = None, []
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
data = data[first:last]
agg = collections.defaultdict(lambda: [0, 0.0])
for d in data:
    k = d["Kernel Name"]
    k = k.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    k = k[:100]
    agg[k][0] += 1
    agg[k][1] += float(d["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("launches %d..%d of %s: %d launches, %.3f ms of kernel time (ncu: serialised, cold cache -- compare SHARES)" % (first, last, sys.argv[1], len(data), tot / 1e6))
print("%8s %10s %7s  %s" % ("count", "total_us", "share", "kernel"))
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print("%8d %10.1f %6.1f%%  %s" % (n, t / 1e3, 100 * t / tot, k))
