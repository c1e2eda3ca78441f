import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=None
agg=collections.defaultdict(lambda:[0,0.0,[]])
for r in rows:
    if r[0]=='ID': hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
    if u=='ns': v/=1e3
    elif u=='ms': v*=1e3
    elif u=='s': v*=1e6
    k=d['Kernel Name'][:70]
    agg[k][0]+=1; agg[k][1]+=v; agg[k][2].append(v)
tot=sum(v[1] for v in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'min_us':>9s} {'max_us':>9s} share")
for k,v in sorted(agg.items(), key=lambda x:-x[1][1]): print(f"{k:70s} {v[0]:5d} {v[1]:11.1f} {v[1]/v[0]:9.1f} {min(v[2]):9.1f} {max(v[2]):9.1f} {v[1]/tot:.3f}")
