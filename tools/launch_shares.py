#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share of the step.
    python tools/launch_shares.py gpurun_out/r1_launches.csv profiles/r1_launch_shares.txt"""
import csv
import sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
t = defaultdict(list)
for r in data:
    t[r[k].split("(")[0]].append(float(r[v]) / 1e3)
tot = sum(sum(x) for x in t.values())
with open(dst, "w") as f:
    f.write("steady-state launch list (ncu --metrics gpu__time_duration.sum --clock-control none, serialized, cold cache): "
            "share of step time per kernel\nnote: inside a step k_val runs on a side stream beside k_encode, so its "
            "serialized share here is not on the critical path\n")
    for name, x in sorted(t.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{name:50s} launches={len(x):3d} mean={sum(x) / len(x):9.2f} us  max={max(x):9.2f} us  share={100 * sum(x) / tot:5.1f}%\n")
print(open(dst).read())
