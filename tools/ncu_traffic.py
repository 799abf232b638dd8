#!/usr/bin/env python
"""Record the DRAM traffic of one trace_tiles launch from an `ncu --set full` capture in profiles/dram_traffic.json
(bench.py reports it as roofline.traffic and, for the 50 M-triangle soup, divides it by the kernel time).

    python tools/ncu_traffic.py <workload> <file.ncu-rep> [<committed summary the numbers can be checked against>]"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, rep = sys.argv[1], sys.argv[2]
source = sys.argv[3] if len(sys.argv) > 3 else os.path.relpath(rep, ROOT)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def metric(key):
    i = hdr.index(key)
    return float(vals[i].replace(",", "")) * scale[units[i]]


rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
dur_i = hdr.index("gpu__time_duration.sum")
path = os.path.join(ROOT, "profiles", "dram_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[name] = {"bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
              "kernel_duration_under_ncu": "%s %s" % (vals[dur_i], units[dur_i]), "kernel": vals[hdr.index("Kernel Name")],
              "source": source}
json.dump(data, open(path, "w"), indent=1, sort_keys=True)
print(name, data[name])
