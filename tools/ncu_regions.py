import csv,sys,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr_idx=[i for i,r in enumerate(rows) if r and r[0]=='Address']
h=rows[hdr_idx[0]]
end=hdr_idx[1]-1 if len(hdr_idx)>1 else len(rows)
data=[r for r in rows[hdr_idx[0]+1:end] if len(r)>10]
ci={n:i for i,n in enumerate(h)}
S=[int(r[ci['# Samples']]) for r in data]; E=[int(r[ci['Instructions Executed']]) for r in data]
src=[r[ci['Source']].strip() for r in data]
tot=sum(S); te=sum(E)
first_u16=next(i for i,s in enumerate(src) if 'LDS.U16' in s)
brx=[i for i,s in enumerate(src) if s.startswith('BRX')]
first_brx,last_brx=brx[0],brx[-1]
sts=[i for i,s in enumerate(src) if 'STS.128' in s and i>last_brx][-16:]
print('landmarks',first_u16,first_brx,last_brx,sts[0],sts[-1],len(src))
def reg(a,b,name):
    s=sum(S[a:b]); e=sum(E[a:b])
    print(f"{name:34s} instr[{a},{b}) samples {s/tot*100:5.1f}%  exec {e/te*100:5.1f}%")
g0=first_u16-14
reg(0,g0,'setup+producer+tile-level')
reg(g0,first_brx+1,'group prologue (addr+16 LDS+entry)')
reg(first_brx+1,sts[0]-40,'gate bodies')
reg(sts[0]-40,sts[-1]+1,'group epilogue (16 STS)')
reg(sts[-1]+1,len(src),'barrier + tile end')
b0,b1=first_brx+1,sts[0]-40
def op(i):
    t=src[i].split(); return t[1] if t[0].startswith('@') else t[0]
d=sum(S[i] for i in range(b0,b1) if op(i).split('.')[0] in('DFMA','DMUL'))
print('  in bodies: DFMA/DMUL samples %.1f%%, other %.1f%%'%(d/tot*100,(sum(S[b0:b1])-d)/tot*100))
c=collections.Counter()
for i in range(b0,b1): c[op(i).split('.')[0]]+=S[i]
print({k:round(v/tot*100,1) for k,v in c.most_common(12)})
