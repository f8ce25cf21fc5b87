"""Hot SASS lines of an exported `ncu --page source --csv --print-source sass` file: python tools/ncu_hot.py file.csv [kernel#] [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        blocks.append(cur)
        continue
    if cur is not None:
        cur['rows'].append(r)
b = blocks[which]
hdr = b['rows'][0]
data = [r for r in b['rows'][1:] if len(r) == len(hdr)]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in data)
print(b['name'][:90], 'samples', tot, 'warp-instr', sum(int(r[iex]) for r in data), 'sass', len(data))
agg = {}
for r in data:
    for j in stall:
        if r[j].isdigit():
            agg[hdr[j][6:]] = agg.get(hdr[j][6:], 0) + int(r[j])
print('stall totals:', sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:topn]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[j]) if r[j].isdigit() else 0, hdr[j][6:]) for j in stall), reverse=True)[:2]
    print(i, '%4.1f%%' % (100 * int(r[isamp]) / tot), r[iex], r[isrc][:64], st)
