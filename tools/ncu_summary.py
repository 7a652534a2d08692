#!/usr/bin/env python
"""Summarise an .ncu-rep (test infrastructure): key metrics + stall mix + hot SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_op_red.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for r in data:
    name = r[hdr.index('Kernel Name')]
    if kern and kern not in name: continue
    print('=====', name[:70])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"  {w:70s} {r[i][:18]:>18s} {units[i]}")
if kern:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # may contain several kernels; take the first block
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr): break
        data.append(r)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = {s: 0 for s in stalls}; total = 0
    for r in data:
        if not r[ix['# Samples']].isdigit(): continue
        total += int(r[ix['# Samples']])
        for s in stalls:
            if r[ix[s]].isdigit(): tot[s] += int(r[ix[s]])
    print('samples', total, ' '.join(f"{s[6:]}={100*v/total:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    top = sorted(data, key=lambda r: -int(r[ix['# Samples']]) if r[ix['# Samples']].isdigit() else 0)[:n]
    for r in top:
        st = {s[6:]: int(r[ix[s]]) for s in stalls if r[ix[s]].isdigit() and int(r[ix[s]]) > 0}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(f"{r[ix['# Samples']]:>6s} x{r[ix['Instructions Executed']]:>9s} {r[ix['Source']][:60]:60s} {st}")
