import csv,collections,re,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
d=dict(zip(rows[0],rows[2]))
def g(k): return d.get(k,'?')
print('duration us',g('gpu__time_duration.sum'),'cycles',g('sm__cycles_elapsed.max'),'inst',g('smsp__inst_executed.sum'))
print('ipc active',g('sm__inst_executed.avg.per_cycle_active'),'issue active %',g('smsp__issue_active.avg.pct_of_peak_sustained_active'))
for k in sorted(d):
    if 'warps_issue_stalled' in k and k.endswith('per_issue_active.ratio'):
        v=float(d[k])
        if v>0.05: print('  %.3f %s'%(v,k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
print('fma pipe %',g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),'alu',g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),'lsu',g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'))
print('dram read',g('dram__bytes_read.sum'),'write',g('dram__bytes_write.sum'),'smem conflicts',g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
cur=None;hdr=None;agg={}
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>2 and r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<10 or r[0]=='': continue
    try: agg[(cur,int(r[0]),r[1].strip())]=(int(r[hdr.index('Instructions Executed')]),int(r[hdr.index('# Samples')]),int(r[hdr.index('stall_no_inst')]))
    except Exception: pass
tot=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
print('instr',tot,'samples',ts)
for (f,ln,s),(n,sm,ni) in sorted(agg.items(),key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print('%5.1f%% samp %5.1f%% instr noinst %4d | %s:%d  %s'%(100*sm/ts,100*n/tot,ni,f,ln,s[:95]))
