"""Summarise a tc_conv timeline written by tools/trace_run.py: per launch, per role, where the cycles of a tile go."""
import sys
from collections import defaultdict
rows = [tuple(int(x) for x in l.split()) for l in open(sys.argv[1])]
ROLE = {0: "producer", 1: "mma", 2: "epi_w0", 3: "epi_w11"}
by = defaultdict(list)
for k, role, ev, tile, stage, clk in rows:
    by[(k, role)].append((ev, tile, stage, clk))
for (k, role), evs in sorted(by.items()):
    t0 = evs[0][3]
    def d(a, b):
        return (b - a) & 0xffffffff
    tiles = sorted({e[1] for e in evs})
    span = d(evs[0][3], evs[-1][3])
    print(f"launch {k} {ROLE[role]:8s} events {len(evs)} tiles {len(tiles)} span {span} cyc ({span / max(len(tiles), 1):.0f} per tile)")
    # per-event-type gaps: time from previous event to this one, averaged over the steady state (skip first 2 tiles)
    acc = defaultdict(lambda: [0, 0])
    prev = None
    for e in evs:
        if prev is not None and e[1] >= 2:
            key = (prev[0], e[0])
            acc[key][0] += d(prev[3], e[3]); acc[key][1] += 1
        prev = e
    for (a, b), (tot, n) in sorted(acc.items()):
        print(f"      ev{a}->ev{b}: avg {tot / n:8.0f} cyc  x{n}  (total {tot})")
    if len(sys.argv) > 2 and role in (0, 1, 2):
        for e in evs[:int(sys.argv[2])]:
            print("        ", e[0], e[1], e[2], d(t0, e[3]))
