#!/usr/bin/env python
"""Per-launch headline metrics + top stall instructions from an .ncu-rep (read here, no GPU)."""
import csv, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
keys = ['launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'sm__cycles_elapsed.max', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum']
for n, r in enumerate(rows[2:]):
    print('--- launch', n, r[h.index('Kernel Name')][:60])
    for k in keys:
        if k in h:
            print('   ', k, r[h.index(k)])
    st = [k for k in h if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio')]
    d = sorted([(float(r[h.index(k)]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in st], reverse=True)[:6]
    print('    stalls', d)
if which is not None:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    # split per kernel
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(r)
    blk = blocks[which]
    h = blk[0]; body = blk[1:]
    si = h.index("# Samples"); sc = h.index("Source"); ie = h.index("Instructions Executed")
    stall_cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[si]) for r in body)
    print("total samples", tot, "sass instrs", len(body), "warp-instr executed", sum(int(r[ie]) for r in body))
    for idx, r in sorted(enumerate(body), key=lambda x: -int(x[1][si]))[:30]:
        stl = sorted([(int(r[i]), h[i]) for i in stall_cols], reverse=True)[:2]
        print(str(idx).rjust(5), r[si].rjust(6), r[ie].rjust(8), r[sc][:80].ljust(80), stl)
