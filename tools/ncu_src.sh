#!/bin/bash
# usage (on the GPU box): tools/ncu_src.sh TAG "<ENV=VAL ...>" <op name> [<op name> ...]
# one `ncu --set full` capture of the named plan ops (bench configuration, 1 launch each); exports the raw page and the
# per-line CUDA + SASS source pages as CSV under gpurun_out/ (the .ncu-rep itself stays on the box: it exceeds the
# merge-back limit).
tag=$1; envs=$2; shift 2
rep=/tmp/ncu_${tag}
env $envs timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $rep \
    python tools/run_ops.py seist_m_dpk 512 1 "$@" > gpurun_out/ncu_${tag}.log 2>&1
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
n=$(($(wc -l < gpurun_out/ncu_${tag}_raw.csv) - 2))
for ((i = 0; i < n; i++)); do
  ncu -i $rep.ncu-rep --page source --print-source cuda --csv --launch-skip $i --launch-count 1 > gpurun_out/ncu_${tag}_cuda_$i.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --print-source sass --csv --launch-skip $i --launch-count 1 > gpurun_out/ncu_${tag}_sass_$i.csv 2>/dev/null
done
tail -3 gpurun_out/ncu_${tag}.log
ls -la gpurun_out/ncu_${tag}_* | head -20
