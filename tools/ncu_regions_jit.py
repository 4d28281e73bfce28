"""tools/ncu_regions_jit.py SOURCE.csv — warp-sample split of a COMPILED pass kernel (k_fused_tma<..., true> with spliced
straight-line groups) from `ncu --page source --csv`: everything before the first amplitude load of a group (setup, producer,
tile level incl. waiting for HBM), the group's 16 LDS.128, the FP64 stream between loads and stores, the 16 STS.128, and
what follows (team barrier, fence, arrive).  Prints per-mnemonic sample shares inside the FP64 stream and the stall reasons
of the kernel's hottest instructions."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hdr_idx[0]]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
data = [r for r in rows[hdr_idx[0] + 1:end] if len(r) > 10]
ci = {n: i for i, n in enumerate(h)}
S = [int(r[ci['# Samples']]) for r in data]
E = [int(r[ci['Instructions Executed']]) for r in data]
src = [r[ci['Source']].strip() for r in data]
tot, te = sum(S), sum(E)


def op(i):
    t = src[i].split()
    return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]


dfma = [i for i in range(len(src)) if op(i) in ('DFMA', 'DMUL')]
lds128 = [i for i in range(len(src)) if src[i].startswith('LDS.128') or ' LDS.128' in src[i][:14]]
sts128 = [i for i in range(len(src)) if 'STS.128' in src[i][:14]]
first_lds = min(i for i in lds128 if i < dfma[0] + 64) if lds128 else dfma[0]
last_sts = max(sts128) if sts128 else dfma[-1]
print('instructions', len(src), 'DFMA/DMUL', len(dfma), 'LDS.128', len(lds128), 'STS.128', len(sts128), 'first load', first_lds, 'last store', last_sts)


def share(idx):
    return 100.0 * sum(S[i] for i in idx) / tot, 100.0 * sum(E[i] for i in idx) / te


body = range(first_lds, last_sts + 1)
print('before the groups (setup, producer lane, waiting for the tile)   samples %5.1f %%  executed %5.1f %%' % share(range(0, first_lds)))
print('register-block groups (loads, FP64 stream, stores)               samples %5.1f %%  executed %5.1f %%' % share(body))
print('after the groups (team barrier, async-proxy fence, arrive, loop) samples %5.1f %%  executed %5.1f %%' % share(range(last_sts + 1, len(src))))
c = collections.Counter(); ce = collections.Counter()
for i in body:
    c[op(i)] += S[i]; ce[op(i)] += E[i]
print('inside the groups, by mnemonic (samples %, executed %):')
print('  ' + ', '.join('%s %.1f/%.1f' % (k, 100.0 * v / tot, 100.0 * ce[k] / te) for k, v in c.most_common(10)))
stall_cols = [n for n in h if n.startswith('stall_') and not n.endswith('(Not Issued)')]
agg = collections.Counter()
for r in data:
    for n in stall_cols:
        try:
            agg[n] += int(r[ci[n]])
        except ValueError:
            pass
st = sum(agg.values()) or 1
print('stall reasons over all samples: ' + ', '.join('%s %.1f %%' % (k[6:], 100.0 * v / st) for k, v in agg.most_common(8)))
