import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot_s=sum(f(r,'# Samples') for r in data); tot_i=sum(f(r,'Instructions Executed') for r in data)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={s:sum(f(r,s) for r in data) for s in stalls}
print("samples",tot_s,"instr",tot_i,{k[6:]:int(v) for k,v in sorted(agg.items(),key=lambda kv:-kv[1])[:8]})
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.01
cum=0
for n,r in enumerate(data):
    s=f(r,'# Samples'); cum+=s
    if s/tot_s>thr:
        st=sorted(stalls,key=lambda s_:-f(r,s_))[:2]
        print("%4d %5.1f%% ins%5.2f%% thr%4.0f %-64s %s"%(n,100*s/tot_s,100*f(r,'Instructions Executed')/tot_i,f(r,'Avg. Threads Executed'),r[ix['Source']][:64]," ".join("%s=%d"%(s_[6:],f(r,s_)) for s_ in st if f(r,s_)>0)))
