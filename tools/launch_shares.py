"""Per-kernel shares of ONE step from an `ncu --metrics gpu__time_duration.sum --csv` launch list (steps end at k_adam).

    python tools/launch_shares.py gpurun_out/x_launches.csv > profiles/x_launch_shares.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
names = [r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '') for r in rows[1:]]
vals = [float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[r[ui]] for r in rows[1:]]
ends = [i for i, n in enumerate(names) if n == 'k_adam']
a, b = (ends[-2] + 1, ends[-1] + 1) if len(ends) > 1 else (0, len(names))
agg = collections.OrderedDict()
for n, v in zip(names[a:b], vals[a:b]):
    c = agg.setdefault(n, [0, 0.0])
    c[0] += 1
    c[1] += v
tot = sum(v for _, v in agg.values())
w = csv.writer(sys.stdout)
w.writerow(['kernel', 'launches_per_step', 'us_per_step', 'share_pct'])
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    w.writerow([n[:70], c, f'{v:.1f}', f'{100 * v / tot:.2f}'])
w.writerow(['TOTAL', b - a, f'{tot:.1f}', '100'])
