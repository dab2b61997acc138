#!/usr/bin/env python
"""Turn an ncu capture of a bench.py run into the record bench.py reads for `roofline.traffic`.

    python tools/ncu_traffic.py <capture.ncu-rep> <config> <path-steps per launch> [source note]

Writes / updates profiles/ncu_traffic.json[<config>] = {kernel, dram bytes read / written per launch, per path-step,
src_hash}.  src_hash is the hash of the kernel's source files at capture time (bench.kernel_source_hash): bench.py only
reports the traffic while the kernel sources are unchanged, otherwise null.  Needs ncu (no GPU) to read the report."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, config, units = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
note = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, unit_row, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
row = data[-1]


def val(name):
    v = float(row[col[name]].replace(",", ""))
    u = unit_row[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
try:
    allrec = json.load(open(path))
    if "kernel" in allrec:  # round-1 single-record layout
        allrec = {}
except (OSError, ValueError):
    allrec = {}
allrec[str(config)] = {
    "kernel": row[col["Kernel Name"]], "dram_bytes_read": rd, "dram_bytes_write": wr,
    "dram_bytes_per_launch": rd + wr, "path_steps_per_launch": units, "dram_bytes_per_unit": (rd + wr) / units,
    "duration_ms_under_ncu": (float(row[col["gpu__time_duration.sum"]].replace(",", "")) *
                              {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit_row[col["gpu__time_duration.sum"]].lower()]
                              if "gpu__time_duration.sum" in col else None),
    "src_hash": bench.kernel_source_hash(config), "source": note,
}
json.dump(allrec, open(path, "w"), indent=1)
print(json.dumps(allrec[str(config)], indent=1))
