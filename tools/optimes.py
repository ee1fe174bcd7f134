import json,sys
from collections import defaultdict
d=json.load(open(sys.argv[1]))
rows=d["rows"]
print("step_ms", round(d["step_ms"],2), "isolated total", round(d["total_ms_isolated"],2))
def cat(n):
    if n.startswith("stem"):
        if "dconv" in n: return "stem.dconv"
        return "stem.pw"
    if n.startswith("out_head"): return "head.conv"
    if ".conv:" in n or n.endswith(".conv"): return "enc.gconv"
    if "core" in n: return "att"
    if "prepare" in n or "finalize" in n: return "bn"
    if ".0.proj" in n or "aggr.proj" in n: return "enc.poolpw"
    return "enc.pw"
k=defaultdict(float); by=defaultdict(float)
for r in rows:
    kind={1:"fwd",2:"bwd_data",3:"bwd_w",4:"res_bwd"}.get(r["kind"],str(r["kind"]))
    k[(cat(r["name"]),kind)]+=r["ms"]; by[(cat(r["name"]),kind)]+=r["bytes"]
tot=sum(k.values())
for a,b in sorted(k.items(), key=lambda x:-x[1])[:18]: print(f"{a[0]:12s} {a[1]:9s} {b:7.2f} ms {100*b/tot:5.1f}%  {by[a]/b/1e6:7.0f} GB/s")
n=int(sys.argv[2]) if len(sys.argv)>2 else 25
for r in rows[:n]:
    print(f'{r["phase"]} {r["name"]:50s} {r["ms"]:7.3f} ms {r["bytes"]/r["ms"]/1e6:7.1f} GB/s {r["flops"]/r["ms"]/1e9:6.2f} TF/s')
