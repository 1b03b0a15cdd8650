#!/usr/bin/env python
"""Top source lines by warp-stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name': continue
    if hdr and len(r) > 8 and r[2] == '-':
        si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
        try: out.append((int(r[si]), int(r[ii]), cur, r[0], r[1].strip()[:120]))
        except ValueError: pass
tot = sum(o[0] for o in out) or 1
print('total samples', tot)
for o in sorted(out, reverse=True)[:n]:
    print(f"{o[0]:6d} {o[0]/tot:6.1%} inst={o[1]:8d} {o[2]}:{o[3]}  {o[4]}")
