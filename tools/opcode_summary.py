"""Per-kernel SASS opcode counts of the built library: the Blackwell-specific instructions that show which kernels use the
5th-generation tensor cores (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st on TMEM, UTCBAR = tcgen05.commit),
TMA (UTMALDG = tensor copy, UBLKCP = bulk copy), mbarriers (SYNCS) and warp shuffles (SHFL).
    python tools/opcode_summary.py > profiles/<round>_opcode_summary.txt"""
import collections, os, re, subprocess, sys
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tinyvc_b200", "libtinyvc_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMAPF", "LDGSTS", "SYNCS", "SHFL", "DADD", "FFMA", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts, size = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        size[cur] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        size[cur] += 1
        op = m.group(1).split(".")[0]
        if op in OPS:
            counts[cur][op] += 1
def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = re.sub(r"\(anonymous namespace\)::", "", r)
    r = re.sub(r"\(.*", "", r)
    return r.replace("void ", "").replace("tvc::", "")
print(f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
tot = collections.Counter()
for k, c in counts.items():
    print(f"{demangle(k)[:58]:58s} {size[k]:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
    tot.update(c)
print(f"{'TOTAL':58s} {sum(size.values()):7d} " + " ".join(f"{tot[o]:8d}" for o in OPS))
