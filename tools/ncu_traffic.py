"""Turn an `ncu --set full` capture of tools/run_ops.py (reps = 1, launch order = op order printed by it) into
profiles/dram_traffic.json: per-op DRAM bytes (read + write) per launch.
usage: python tools/ncu_traffic.py report.ncu-rep <model> <batch> <out.json> <op name> [<op name> ...]
The op names must be given in the order run_ops.py ran them (it prints "ran <name>" in plan order)."""
import csv
import io
import json
import subprocess
import sys

rep, model, batch, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
names = sys.argv[5:]
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def val(d, name):
    return float(d[ix[name]].replace(",", "")) * scale[units[ix[name]]]


assert len(data) == len(names), (len(data), len(names))
ops = {}
for d, n in zip(data, names):
    ops[n] = {"kernel": d[ix["Kernel Name"]], "dram_bytes": val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum"),
              "dram_read_bytes": val(d, "dram__bytes_read.sum"), "dram_write_bytes": val(d, "dram__bytes_write.sum"),
              "duration_us_under_ncu": float(d[ix["gpu__time_duration.sum"]].replace(",", "")) *
              {"ns": 1e-3, "us": 1.0, "ms": 1e3}[units[ix["gpu__time_duration.sum"]]]}
json.dump({"model": model, "batch": batch, "source": rep.split("/")[-1] + " (ncu --set full --clock-control none, tools/run_ops.py)",
           "ops": ops}, open(out, "w"), indent=1)
print(json.dumps(ops, indent=1))
