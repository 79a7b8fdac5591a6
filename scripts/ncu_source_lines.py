#!/usr/bin/env python
"""Per-source-line stall samples and instruction counts of one kernel launch in an .ncu-rep (needs --import-source on, -lineinfo).
usage: python scripts/ncu_source_lines.py rep.ncu-rep [launch_index] [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
# launches are separated by repeated "File Path" headers; collect blocks per kernel
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == 'Function Name':
        cur = {'name': row[1], 'rows': [], 'hdr': None, 'file': None}; blocks.append(cur); continue
    if row[0] == 'File Path':
        pending_file = row[1]; continue
    if cur is None: continue
    if row[0] == 'Line No': cur['hdr'] = row; continue
    cur['rows'].append(row)
# group blocks that belong to the same launch: a new launch starts when the function name is a kernel (contains 'k_') after seeing sass rows
launches = []
for b in blocks:
    launches.append(b)
# every block is one (file, function) pair; the launch index is not explicit: print totals per block name and let the caller pick
agg = collections.OrderedDict()
for b in launches:
    h = b['hdr']
    if not h: continue
    i_s = h.index('# Samples'); i_x = h.index('Instructions Executed')
    for r in b['rows']:
        if r[0] == '' or not r[0].isdigit(): continue
        key = (b['name'][:40], int(r[0]), r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0])
        a[0] += int(r[i_s]) if r[i_s].isdigit() else 0; a[1] += int(r[i_x]) if r[i_x].isdigit() else 0
tot = sum(a[0] for a in agg.values()) or 1
print("total samples", tot)
for (name, line, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot:5.1f}% samp={a[0]:7d} inst={a[1]:9d}  L{line:4d} {src}")
