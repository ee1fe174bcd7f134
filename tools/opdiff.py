"""Compare two per-op device-time tables written by bench.py --op-times (largest changes first)."""
import json
import sys

a = {(r["phase"], r["name"]): r for r in json.load(open(sys.argv[1]))["rows"]}
b = {(r["phase"], r["name"]): r for r in json.load(open(sys.argv[2]))["rows"]}
flt = sys.argv[3] if len(sys.argv) > 3 else ""
rows = []
for k, ra in a.items():
    if k in b and flt in k[1]:
        rows.append((b[k]["ms"] - ra["ms"], k, ra["ms"], b[k]["ms"], b[k]["flops"] / b[k]["ms"] / 1e9 if b[k]["ms"] else 0))
rows.sort()
print("sum a %.2f  sum b %.2f" % (sum(r[2] for r in rows), sum(r[3] for r in rows)))
for d, k, x, y, tf in rows[:25] + rows[-8:]:
    print("%-50s %7.3f -> %7.3f  (%+.3f)  %.1f TF/s" % (k[1], x, y, d, tf))
