#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + per-opcode dynamic instruction mix): python tools/ncu_summary.py REP WARP_UNIT_ITERS"""
import csv, sys, collections, re, subprocess
rep=sys.argv[1]; units_iter=float(sys.argv[2])
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,r=rows[0],rows[1],rows[2]
keys=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__warps_eligible.avg.per_cycle_active','dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio']
for i,h in enumerate(hdr):
    if h in keys or ('issue_stalled' in h and 'per_issue_active' in h): 
        try: v=float(r[i])
        except: v=r[i]
        if 'issue_stalled' in h and isinstance(v,float) and v<0.05: continue
        print(f"{h:85s} {units[i]:12s} {r[i]}")
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))[2:]
hist=collections.Counter(); tot=0
for row in rows:
    try: n=int(row[5])
    except: continue
    ins=re.sub(r'^@!?U?P\d+\s+','',row[1].strip())
    hist[ins.split()[0].rstrip(';').split('.')[0]]+=n; tot+=n
ie=[float(r[i]) for i,h in enumerate(hdr) if h=='smsp__inst_executed.sum'][0]
scale=ie/tot
print('source total',tot,'inst_executed',ie,'scale',scale)
print('per 32 units:', ie/units_iter)
print(' '.join(f"{k}:{v*scale/units_iter:.1f}" for k,v in hist.most_common(32)))
