"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares of the step.
usage: python tools/launch_shares.py launches.csv > profiles/rN_launch_shares.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(SeistOp.*|\(const .*", "", r[ix["Kernel Name"]])
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ix["Metric Unit"]], 1.0)
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print("kernel,launches,total_us,share")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"\"{k}\",{cnt[k]},{v:.1f},{v / s:.3f}")
print(f"# {sum(cnt.values())} launches, {s:.1f} us serialised", file=sys.stderr)
