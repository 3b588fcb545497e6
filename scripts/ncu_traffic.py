#!/usr/bin/env python
"""profiles/r2_traffic.json: DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per pair of the fill kernels from this round's
ncu --set full captures; bench.py scales it to its launch size for `roofline.traffic`.
usage: ncu_traffic.py name=report.ncu-rep:pairs[:summary-file] ...  > profiles/r2_traffic.json"""
import csv, json, subprocess, sys
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {"_comment": "DRAM traffic of the fill kernel from ncu --set full captures of round 2 (dram__bytes_read.sum + dram__bytes_write.sum of the one profiled launch), per pair of that launch; bench.py scales it to its launch size"}
for a in sys.argv[1:]:
    name, rest = a.split("=")
    parts = rest.split(":")
    rep, pairs = parts[0], int(parts[1])
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    g = lambda k: float(v[h.index(k)]) * UNIT.get(u[h.index(k)], 1)
    tot = g("dram__bytes_read.sum") + g("dram__bytes_write.sum")
    out[name] = {"kernel": v[h.index("Kernel Name")].split("(")[0].replace("void ", ""), "capture": parts[2] if len(parts) > 2 else rep, "pairs": pairs,
                 "dram_bytes": tot, "bytes_per_pair": tot / pairs, "gpu_time_ms": g("gpu__time_duration.sum") * (1e-6 if u[h.index("gpu__time_duration.sum")] in ("ns", "nsecond") else 1)}
print(json.dumps(out, indent=1))
