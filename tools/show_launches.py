import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; iname=hdr.index('Kernel Name'); ival=hdr.index('Metric Value'); imet=hdr.index('Metric Name'); iid=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iid],{'name':r[iname].replace('void unnamed>::','')[:24]})[r[imet]]=r[ival]
tot=0
for k,v in d.items():
    us=float(v['gpu__time_duration.sum'].replace(',',''))/1e3; tot+=us
    print(f"{k:>3} {v['name']:24s} {us:9.1f} us  thr/inst {v['smsp__thread_inst_executed_per_inst_executed.ratio']:>6}  warp-inst {float(v['smsp__inst_executed.sum'].replace(',',''))/1e6:8.1f}M  warps% {v['sm__warps_active.avg.pct_of_peak_sustained_active']:>6} issue% {v['smsp__issue_active.avg.pct_of_peak_sustained_active']:>6}")
print('total us', tot)
