import csv, collections, re, sys
path=sys.argv[1]
with open(path) as f:
    lines=[l for l in f if l.startswith('"')]
r=csv.reader(lines); hdr=next(r)
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0]); seq=[]
for row in r:
    name=re.sub(r'\(.*','',row[ki]); t=float(row[vi].replace(',',''))
    agg[name][0]+=1; agg[name][1]+=t; seq.append((name,t))
tot=sum(v[1] for v in agg.values())
print('total us %.1f launches %d'%(tot/1e3, len(seq)))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print('%8.1f us %5.1f%% n=%4d avg %7.1f  %s'%(v[1]/1e3, 100*v[1]/tot, v[0], v[1]/1e3/v[0], k[:100]))
