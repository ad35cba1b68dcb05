#!/usr/bin/env python
"""profiles/k3_dram_traffic.json from an `ncu --set full` capture of the fused render kernel inside bench.py:

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/update_k3_traffic.py raw.csv [workload=dtu] [views=8]

The file is keyed on the sha256 of the kernel's sources (bench._kernel_digest): bench.py reports `roofline.traffic` only while
the committed capture belongs to the code it is timing."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rows = list(csv.reader(open(sys.argv[1])))
workload = sys.argv[2] if len(sys.argv) > 2 else "dtu"
views = int(sys.argv[3]) if len(sys.argv) > 3 else 8
head, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(head)}
r = [x for x in body if "render_tc" in x[col["Kernel Name"]]][0]


def mbytes(k):
    v, u = float(r[col[k]].replace(",", "")), units[col[k]]
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


rd, wr = mbytes("dram__bytes_read.sum"), mbytes("dram__bytes_write.sum")
out = {"source": f"ncu --set full of the K3 launch inside bench.py ({sys.argv[1]}), {views} {workload} target views per launch",
       "kernel": r[col["Kernel Name"]], "workload": workload, "views_per_launch": views, "kernel_source_sha256": bench._kernel_digest(),
       "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr}
with open(os.path.join(ROOT, "profiles", "k3_dram_traffic.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
