#!/bin/bash
# usage: tools/knob_sweep.sh TAG "ENV=VAL ENV2=VAL" ...   -> gpurun_out/sweep_TAG_<i>.json (+ op tables)
# Runs bench.py (no CPU baseline) once per environment setting; measurement knobs only.
tag=$1; shift
i=0
for envs in "$@"; do
  out=gpurun_out/sweep_${tag}_${i}
  env $envs timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --op-times ${out}_ops.json > ${out}.json 2> ${out}.err
  python - "$envs" ${out}.json ${out}_ops.json <<'PY'
import json,sys
envs,f,fo=sys.argv[1:4]
try:
    d=json.loads(open(f).read())
    fam={}
    for r in json.load(open(fo))['rows']:
        fam[r['family']]=fam.get(r['family'],0)+r['ms']
    top=sorted(fam.items(),key=lambda kv:-kv[1])[:9]
    print(f"[{envs}] ms/step={d['ms_per_step']:.2f} iso={sum(fam.values()):.1f} "+' '.join(f"{k.split('(')[0]}={v:.2f}" for k,v in top))
except Exception as e:
    print(f"[{envs}] FAILED {e}")
PY
  i=$((i+1))
done
