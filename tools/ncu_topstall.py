import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr_idx=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hdr_idx[0]]
end=hdr_idx[1]-1 if len(hdr_idx)>1 else len(rows)
data=[r for r in rows[hdr_idx[0]+1:end] if len(r)>10]
ci={n:i for i,n in enumerate(h)}
tot=sum(int(r[ci['# Samples']]) for r in data)
te=sum(int(r[ci['Instructions Executed']]) for r in data)
stall_cols=[n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg=collections.Counter()
for r in data:
    for n in stall_cols: agg[n]+=int(r[ci[n]] or 0)
print({k:round(v/tot*100,1) for k,v in agg.most_common(12)})
N=int(sys.argv[2]) if len(sys.argv)>2 else 40
top=sorted(range(len(data)), key=lambda i:-int(data[i][ci['# Samples']]))[:N]
for i in sorted(top):
    r=data[i]
    st={n:int(r[ci[n]] or 0) for n in stall_cols}
    m=sorted(st.items(), key=lambda kv:-kv[1])[:2]
    print(i, f"{int(r[ci['# Samples']])/tot*100:5.2f}% ex {int(r[ci['Instructions Executed']])/te*100:5.2f}%", r[ci['Source']].strip()[:60], m)
