"""Summarise an ncu report: key raw metrics + stall samples per CUDA source line."""
import csv, sys, subprocess, collections, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__cycles_elapsed.avg','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.avg.per_cycle_elapsed']
for h,u,v in zip(hdr,units,vals):
    if h in want: print(f'{h} [{u}] = {v}')
cs = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(cs)))
hdr = rows[2]; si = hdr.index('# Samples')
stall_idx = {h:i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h}
agg = collections.Counter(); st = collections.defaultdict(collections.Counter); tot_st = collections.Counter()
cur = None; cur_src = {}
for r in rows[3:]:
    if len(r) < si+1: continue
    if r[0].strip().isdigit() and r[2].strip() in ('-',''):
        cur = int(r[0]); cur_src[cur] = r[1]; continue
    try: n = int(r[si])
    except: continue
    key = cur
    if 'DMMA' in r[3]: key = 'DMMA'
    agg[key] += n
    for s,i in stall_idx.items():
        try: v = int(r[i])
        except: v = 0
        st[key][s[6:]] += v; tot_st[s[6:]] += v
tot = sum(agg.values()); print('total samples', tot)
print('stalls:', ', '.join(f'{k}:{100*v/tot:.1f}%' for k,v in tot_st.most_common(9)))
for ln,n in agg.most_common(topn):
    top = ', '.join(f'{k}:{v}' for k,v in st[ln].most_common(3))
    src = cur_src.get(ln,'') if ln != 'DMMA' else 'DMMA instructions'
    print(f'{n:8d} {100*n/tot:5.1f}% L{ln}: {src.strip()[:88]}   [{top}]')
