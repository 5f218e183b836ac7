"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  Usage: launch_summary.py file.csv [first-kernel-substring]
With a second argument only the launches from the LAST occurrence of that kernel on are counted (e.g. the second forward)."""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("ddw::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
    seq.append((name, v))
if len(sys.argv) > 2:
    idx = [i for i, (n, _) in enumerate(seq) if sys.argv[2] in n]
    seq = seq[idx[-1]:] if idx else seq
agg = collections.defaultdict(list)
for n, v in seq:
    agg[n].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{len(seq)} launches, {tot / 1e3:.2f} ms of kernel time")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    busy = [x for x in v if x > 10.0]
    print(f"{k[:58]:58s} n={len(v):5d} total={sum(v) / 1e3:9.2f} ms share={sum(v) / tot:6.1%} avg(>10us)={(sum(busy) / len(busy)) if busy else 0:8.1f} us")
