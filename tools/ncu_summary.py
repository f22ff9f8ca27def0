"""Summarise an `ncu --csv --page raw` launch list: per kernel name (+grid) totals of time and DRAM bytes."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
# find header
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
def get(r, name, default=0.0):
    i = col.get(name)
    if i is None or i >= len(r): return default
    try: return float(r[i].replace(',', ''))
    except ValueError: return default
units = rows[hi + 1]
agg = collections.OrderedDict()
tot = 0.0
for r in rows[hi + 2:]:
    if len(r) < len(hdr): continue
    name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('<unnamed>::', '')
    name = name[:60]
    t = get(r, 'gpu__time_duration.sum')
    if units[col['gpu__time_duration.sum']] == 'ns': t /= 1e3
    elif units[col['gpu__time_duration.sum']] == 'ms': t *= 1e3
    rd, wr = get(r, 'dram__bytes_read.sum'), get(r, 'dram__bytes_write.sum')
    def tob(v, u):
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    rd = tob(rd, units[col['dram__bytes_read.sum']]) if 'dram__bytes_read.sum' in col else 0
    wr = tob(wr, units[col['dram__bytes_write.sum']]) if 'dram__bytes_write.sum' in col else 0
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
    tot += t
print(f'total {tot/1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / (a[1] * 1e-6) / 1e9 if a[1] > 0 else 0
    print(f'{a[1]/1e3:8.3f} ms {100*a[1]/tot:5.1f}%  n={a[0]:4d}  avg {a[1]/a[0]:7.1f} us  dram rd {a[2]/1e6:8.1f} MB wr {a[3]/1e6:8.1f} MB  {gbs:7.0f} GB/s  {name}')
