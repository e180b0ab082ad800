import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1], errors='ignore')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr=rows[hi]; idx={h:i for i,h in enumerate(hdr)}
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hi+1:]:
    if len(r)<len(hdr): continue
    if r[idx['Metric Name']]!='gpu__time_duration.sum': continue
    name=r[idx['Kernel Name']].split('(')[0][-40:]
    v=float(r[idx['Metric Value']].replace(',',''))
    u=r[idx['Metric Unit']]
    v = v/1000 if u in ('ns','nsecond') else v
    tot[name]+=v; cnt[name]+=1
for k,v in tot.most_common(20): print(f"{k:42s} n={cnt[k]:4d} total_us={v:10.1f} avg={v/cnt[k]:8.2f}")
print('sum', sum(tot.values()))
