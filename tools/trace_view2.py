"""Raw per-role timeline of traced tc_conv launches (tools/trace_run.py): cycles since kernel entry."""
import sys
from collections import defaultdict
rows = [tuple(int(x) for x in l.split()) for l in open(sys.argv[1])]
ROLE = {0: "prod", 1: "mma", 2: "epi0", 3: "epi11"}
by = defaultdict(list)
for k, role, ev, tile, stage, clk in rows:
    by[k].append((role, ev, tile, stage, clk))
for k, evs in sorted(by.items()):
    entry = [e for e in evs if e[0] == 1 and e[1] == 6]
    t0 = entry[0][4] if entry else min(e[4] for e in evs if e[1] < 8)
    g0 = [e[4] for e in evs if e[1] == 8]
    g1 = [e[4] for e in evs if e[1] == 10]
    print(f"== launch {k}: entry..exit {((([e[4] for e in evs if e[1]==9] or [t0])[0]-t0)&0xffffffff)} cyc; globaltimer span {((g1[0]-g0[0])&0xffffffff) if g0 and g1 else -1} ns")
    seq = sorted(((e[4] - t0) & 0xffffffff, ROLE[e[0]], e[1], e[2], e[3]) for e in evs if e[1] not in (8, 10))
    lim = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    for d, r, ev, tile, st in seq[:lim]:
        print(f"   {d:8d}  {r:5s} ev{ev} tile {tile} stage {st}")
    if len(seq) > lim:
        print("   ...")
        for d, r, ev, tile, st in seq[-12:]:
            print(f"   {d:8d}  {r:5s} ev{ev} tile {tile} stage {st}")
