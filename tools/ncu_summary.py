import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
keys=['Kernel Name','launch__grid_size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','smsp__cycles_active.avg','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','lts__t_sectors_srcunit_tex_op_read.sum']
for r in rows[2:]:
    for k in keys:
        if k in hdr: print('%-70s %-10s %s'%(k, units[hdr.index(k)], r[hdr.index(k)][:90]))
    print('---')
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=rows[1] if 'Source' in rows[1] else rows[0]
hi=rows.index(h)
body=[]
for r in rows[hi+1:]:
    if r==h: break
    if len(r)==len(h): body.append(r)
def I(x):
    try: return int(x)
    except: return 0
si=h.index('# Samples'); so=h.index('Source')
stall=[i for i,c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot=sum(I(r[si]) for r in body)
agg={h[i]:sum(I(r[i]) for r in body) for i in stall}
print('samples',tot)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8]: print('  %-26s %6d %5.1f%%'%(k,v,100*v/max(tot,1)))
for r in sorted(body,key=lambda r:-I(r[si]))[:int(sys.argv[2]) if len(sys.argv)>2 else 16]:
    st=sorted(((I(r[i]),h[i]) for i in stall),reverse=True)[:1]
    print('%6d %5.1f%%  %-60s %s'%(I(r[si]),100*I(r[si])/max(tot,1),r[so][:60],st))
