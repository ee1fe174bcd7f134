"""Top sampled SASS instructions of the first kernel section of an `ncu --page source --print-source sass --csv` export.
usage: python tools/sass_hot.py file.csv [N] [--ctx K]"""
import csv, sys
f = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 25
ctx = int(sys.argv[sys.argv.index("--ctx") + 1]) if "--ctx" in sys.argv else 0
rows = list(csv.reader(open(f)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = rows[starts[0]:(starts[1] if len(starts) > 1 else len(rows))]
print(sec[0][1][:100])
h = sec[1]; ix = {n: i for i, n in enumerate(h)}
body = [r for r in sec[2:] if len(r) == len(h)]
tot = sum(int(r[ix["# Samples"]]) for r in body) or 1
toti = sum(int(r[ix["Instructions Executed"]]) for r in body) or 1
stalls = [c for c in h if c.startswith("stall_")]
print(f"{len(body)} SASS instr, {toti} warp-instr executed, {tot} samples")
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print("stall mix: " + "  ".join(f"{s[6:]}={100*v/tot:.1f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:N]
for i in sorted(order) if ctx == 0 else order:
    r = body[i]
    top = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    for j in range(max(0, i - ctx), i):
        print(f"          {j:5d}  {body[j][ix['Source']].strip()[:80]}")
    print(f"{100*int(r[ix['# Samples']])/tot:5.1f}%  {i:5d}  {r[ix['Source']].strip()[:80]:80s} x{int(r[ix['Instructions Executed']])/max(1,int(body[0][ix['Instructions Executed']])):.1f}  {top}")
