"""Top source lines of one kernel in an ncu report captured with --import-source on (-lineinfo builds).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep k_edge_bwd_tc [N]

Prints the share of warp-stall samples and executed instructions per source line (inlined code is attributed to
the file/line it came from).
"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', f'regex:{kern}'],
                     capture_output=True, text=True).stdout
cur = hdr = None
agg = collections.defaultdict(lambda: [0, 0])
src = {}
first_fn = None
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1]
        continue
    if len(r) == 2 and r[0] == 'Function Name':
        if first_fn is None:
            first_fn = r[1]
        elif r[1] != first_fn:      # only the first matching launch
            pass
        continue
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        try:
            ln = int(r[0])
        except ValueError:
            continue
        key = (cur.split('/')[-1], ln)
        agg[key][0] += int(r[hdr.index('# Samples')] or 0)
        agg[key][1] += int(r[hdr.index('Instructions Executed')] or 0)
        src[key] = r[1]
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print(f'{kern}: {ts} samples, {ti} warp instructions (all captured launches of the kernel)')
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f'{k[0]:20s}{k[1]:5d} samples {100 * v[0] / ts:5.1f}%  inst {100 * v[1] / ti:5.1f}%  {src[k].strip()[:88]}')
