"""Per-source-line L1 data-pipe load of one kernel in an ncu report (--import-source on, -lineinfo): shared-memory
wavefronts, global tag requests, executed warp instructions.

    python tools/ncu_pipe.py gpurun_out/x.ncu-rep k_edge_bwd_tc [N]
"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k', f'regex:{kern}'],
                     capture_output=True, text=True).stdout
cur = hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0])
src = {}
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1]
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
        agg[key][0] += int(r[hdr.index('L1 Wavefronts Shared')] or 0)
        agg[key][1] += int(r[hdr.index('L1 Tag Requests Global')] or 0)
        agg[key][2] += int(r[hdr.index('Instructions Executed')] or 0)
        src[key] = r[1]
tot = [sum(v[i] for v in agg.values()) or 1 for i in range(3)]
print(f'{kern}: shared wavefronts {tot[0]}, global tag requests {tot[1]}, warp instructions {tot[2]}')
for col, name in ((0, 'shared wavefronts'), (1, 'global tag requests')):
    print(f'--- by {name}')
    for k, v in sorted(agg.items(), key=lambda x: -x[1][col])[:top]:
        if v[col] == 0:
            break
        print(f'{k[0]:20s}{k[1]:5d} {v[col]:12d} {100 * v[col] / tot[col]:5.1f}%  inst {v[2]:10d}  {src[k].strip()[:80]}')
