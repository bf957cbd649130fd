"""Summarise `ncu --page source --csv --print-source cuda,sass` output: per CUDA source line, the share of
executed instructions and of stall samples.  Usage: python tools/ncu_hot_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = {}
cur, H = None, None
per = defaultdict(lambda: [0, 0, '', defaultdict(int)])
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1]
        continue
    if r[0] == 'Line No':
        H = r
        ci, si = H.index('Instructions Executed'), H.index('Warp Stall Sampling (All Samples)')
        stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if H is None or len(r) < len(H):
        continue
    if r[0]:                       # a CUDA source line; the SASS rows that follow belong to it
        key = (cur.split('/')[-1], int(r[0]))
        per[key][2] = r[1].strip()[:100]
        continue
    try:
        per[key][0] += int(r[ci] or 0)
        per[key][1] += int(r[si] or 0)
        for i, h in stall_cols:
            per[key][3][h] += int(r[i] or 0)
    except (ValueError, NameError):
        pass
ti = sum(v[0] for v in per.values()) or 1
ts = sum(v[1] for v in per.values()) or 1
print('total warp instructions %d, stall samples %d' % (ti, ts))
allst = defaultdict(int)
for v in per.values():
    for h, c in v[3].items():
        allst[h] += c
print('stalls:', ', '.join('%s %.1f%%' % (h[6:], 100 * c / ts) for h, c in sorted(allst.items(), key=lambda kv: -kv[1])[:8]))
for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    st = max(v[3].items(), key=lambda kv: kv[1])[0][6:] if v[3] else ''
    print('%5.1f%% inst %5.1f%% stall(%-10s) %s:%d  %s' % (100 * v[0] / ti, 100 * v[1] / ts, st, key[0], key[1], v[2]))
