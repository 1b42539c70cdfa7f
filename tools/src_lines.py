import csv, re, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
cur_file=None; cur_fn=None; hdr=None
agg=collections.defaultdict(lambda: [0,0,0,''])
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': cur_fn=re.sub(r'rtb::|<unnamed>::|\(rtb.*','',r[1])[:40]; continue
    if r[0]=='Line No': hdr=r; iI=hdr.index('Instructions Executed'); iT=hdr.index('Thread Instructions Executed'); iN=hdr.index('# Samples'); continue
    if hdr and r[0]!='' and len(r)>iT:
        try:
            k=(cur_fn,cur_file,int(r[0])); a=agg[k]; a[0]+=int(r[iI]); a[1]+=int(r[iT]); a[2]+=int(r[iN]); a[3]=r[1].strip()[:100]
        except ValueError: pass
fns=sorted(set(k[0] for k in agg))
for fn in fns:
    items=[(k,v) for k,v in agg.items() if k[0]==fn]
    tot=sum(v[0] for k,v in items); 
    if tot==0: continue
    print('==== kernel',fn,'total warp-inst %.1fM'%(tot/1e6))
    for k,v in sorted(items,key=lambda kv:-kv[1][0])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
        print(f"{v[0]/1e6:8.1f}M {100*v[0]/tot:5.1f}% thr {v[1]/max(1,v[0]):5.1f} smp {v[2]:6d} {k[1]}:{k[2]} {v[3]}")
