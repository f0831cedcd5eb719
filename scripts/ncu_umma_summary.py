import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; idx={h:i for i,h in enumerate(hdr)}
want=['gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_active.avg','smsp__inst_executed.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']
stall=[h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][40:100])
    for c in want:
        if c in idx: print(f"  {c}: {r[idx[c]]}")
    st=sorted(((float(r[idx[h]] or 0),h) for h in stall), reverse=True)[:6]
    for v,h in st: print(f"  stall {h.split('issue_stalled_')[1].split('_per_')[0]}: {v:.2f}")
