#!/usr/bin/env python
"""Print per-kernel durations from an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes])."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[i0]
d, order = {}, []
for r in rows[i0 + 1:]:
    if len(r) < len(h):
        continue
    id_ = int(r[0]); name = r[h.index("Kernel Name")]; m = r[h.index("Metric Name")]; v = float(r[h.index("Metric Value")])
    if id_ not in d:
        d[id_] = {"name": name.split('(')[0][-34:], "grid": r[h.index("Grid Size")]}
        order.append(id_)
    d[id_][m] = v
tot = 0
brief = len(sys.argv) > 2
line = []
for id_ in order:
    e = d[id_]; t = e.get("gpu__time_duration.sum", 0) / 1000; tot += t
    if brief:
        line.append(f"{t:.0f}")
    else:
        print(f"{id_:3d} {e['name']:36s} {e['grid']:>14s} {t:8.1f} us  rd {e.get('dram__bytes_read.sum', 0) / 1e6:7.1f} MB wr {e.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB")
if brief:
    print(" ".join(line))
print("total us %.1f" % tot)
