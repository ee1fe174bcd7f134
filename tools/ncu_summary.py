"""Summarise an .ncu-rep (read on the CPU box): per-kernel roofline-relevant metrics + stall mix, and
optionally the SASS opcode / stall histogram of one launch.
usage: python tools/ncu_summary.py report.ncu-rep [--sass <launch-index>]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
M = {
    "dur_us": "gpu__time_duration.sum", "grid": "launch__grid_size", "regs": "launch__registers_per_thread",
    "warps_active_%": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram_rd_MB": "dram__bytes_read.sum", "dram_wr_MB": "dram__bytes_write.sum",
    "dram_%": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "ipc": "sm__inst_executed.avg.per_cycle_elapsed", "fma_%": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_%": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l2_MB": "lts__t_bytes.sum", "inst": "smsp__inst_executed.sum",
}
STALLS = ["long_scoreboard", "wait", "short_scoreboard", "barrier", "no_instruction", "not_selected", "dispatch_stall",
          "math_pipe_throttle", "mio_throttle", "lg_throttle", "branch_resolving", "imc_miss"]
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(d, name):
    if name not in ix:
        return float("nan")
    v, u = d[ix[name]].replace(",", ""), units[ix[name]]
    try:
        x = float(v)
    except ValueError:
        return float("nan")
    scale = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
    return x * scale.get(u, 1.0)


for li, d in enumerate(data):
    name = d[ix["Kernel Name"]]
    print(f"[{li}] {name[:90]}")
    print("    " + "  ".join(f"{k}={val(d, m):.4g}" for k, m in M.items()))
    st = []
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in ix:
            st.append((float(d[ix[k]] or 0), s))
    st.sort(reverse=True)
    print("    stalls/issue: " + "  ".join(f"{s}={v:.2f}" for v, s in st[:6]))

if "--sass" in sys.argv:
    li = int(sys.argv[sys.argv.index("--sass") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # the source page may hold several kernels: sections start with a "Kernel Name" row followed by a header row
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = rows[starts[0]:(starts[1] if len(starts) > 1 else len(rows))]
    print("source page of:", sec[0][1])
    h = sec[1]
    ix2 = {n: i for i, n in enumerate(h)}
    body = [r for r in sec[2:] if len(r) == len(h)]
    ops, samp = Counter(), Counter()
    tot = sum(int(r[ix2["Instructions Executed"]]) for r in body) or 1
    tots = sum(int(r[ix2["# Samples"]]) for r in body) or 1
    for r in body:
        toks = r[ix2["Source"]].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += int(r[ix2["Instructions Executed"]])
        samp[op] += int(r[ix2["# Samples"]])
    print(f"SASS of launch {li}: {len(body)} instructions, {tot} warp-instr executed")
    for o, c in ops.most_common(12):
        print(f"    {o:10s} {100 * c / tot:5.1f}% of executed   {100 * samp[o] / tots:5.1f}% of samples")
    body.sort(key=lambda r: -int(r[ix2["# Samples"]]))
    for r in body[:12]:
        print(f"    {r[ix2['# Samples']]:>6s} smp  {r[ix2['Source']][:70]:70s} long_sb={r[ix2['stall_long_sb']]} wait={r[ix2['stall_wait']]} "
              f"short_sb={r[ix2['stall_short_sb']]} barrier={r[ix2['stall_barrier']]}")
