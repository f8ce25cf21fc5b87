"""Print the per-launch table of an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_*` log (tools/gpu_check.sh)."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    ker = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = ker.setdefault(int(r[idx['ID']]), {'name': r[idx['Kernel Name']], 'grid': r[idx['Grid Size']]})
        d[r[idx['Metric Name']]] = (float(r[idx['Metric Value']].replace(',', '')), r[idx['Metric Unit']])
    out = []
    for k, d in ker.items():
        t = d['gpu__time_duration.sum']
        us = t[0] * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'usecond': 1, 'nsecond': 1e-3, 'msecond': 1e3}[t[1]]
        def mb(x):
            v, u = x
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u] / 1e6
        out.append((k, d['name'], d['grid'], us, mb(d['dram__bytes_read.sum']), mb(d['dram__bytes_write.sum'])))
    return out


if __name__ == "__main__":
    tab = load(sys.argv[1])
    print("total %.1f us over %d launches" % (sum(o[3] for o in tab), len(tab)))
    for k, name, grid, us, rd, wr in tab:
        short = name.replace("vqa::", "").replace("tc::", "")[:58]
        print("%3d %-58s %-14s %7.1f us  rd %7.2f MB  wr %7.2f MB" % (k, short, grid, us, rd, wr))
