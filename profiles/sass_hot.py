#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: per kernel, contiguous SASS regions with similar
execution counts (loops), their share of executed warp instructions, SIMT efficiency and stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; ks.append(cur); continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    if cur is not None and r: cur['rows'].append(r)
k = ks[which]; h = k['hdr']
ia, it, isamp = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
R = k['rows']
tot = sum(int(r[ia]) for r in R); tots = sum(int(r[isamp]) for r in R)
print(k['name'][:80]); print('warp inst %.3f G, samples %d, sass lines %d' % (tot / 1e9, tots, len(R)))
seg = []
s = 0
for n in range(1, len(R) + 1):
    if n == len(R) or abs(int(R[n][ia]) - int(R[s][ia])) > 0.25 * max(int(R[s][ia]), 1):
        e = sum(int(r[ia]) for r in R[s:n]); t = sum(int(r[it]) for r in R[s:n]); sm = sum(int(r[isamp]) for r in R[s:n])
        seg.append((s, n, e, t, sm)); s = n
for s, n, e, t, sm in seg:
    if e > 0.004 * tot:
        print('%5d-%5d  n=%4d  exec/inst %7.2fM  share %5.1f%%  thr %4.1f  samples %5.1f%%   %s' % (s, n, n - s, int(R[s][ia]) / 1e6, 100 * e / tot, t / max(e, 1), 100 * sm / max(tots, 1), R[s][1].strip()[:40]))
