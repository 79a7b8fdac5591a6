#!/usr/bin/env python
"""Per-opcode histogram of executed warp instructions and stall samples from `ncu --page source --csv` (SASS view).
usage: python scripts/ncu_sass_hist.py report.ncu-rep [kernel-index]"""
import csv, io, re, subprocess, sys, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
blk = blocks[1 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
print(rows[0][1])
hdr = rows[1]
si, ii, st, wf, wfi = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('L1 Wavefronts Shared'), hdr.index('L1 Wavefronts Shared Ideal')
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = 0
for r in rows[2:]:
    if len(r) <= ii: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)', r[si])
    if not m: continue
    op = m.group(2)
    key = op.split('.')[0] if not op.startswith(('LDS', 'STS', 'LDG', 'STG', 'MUFU', 'I2F', 'F2I', 'BAR')) else op
    n = int(r[ii] or 0)
    agg[key][0] += n; agg[key][1] += int(r[st] or 0); agg[key][2] += 1
    agg[key][3] += int(r[wf] or 0); agg[key][4] += int(r[wfi] or 0)
    tot += n
stt = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot}, stall samples {stt}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{k:22s} inst={v[0]:11d} ({100*v[0]/tot:5.1f}%)  stall_samples={v[1]:6d} ({100*v[1]/max(stt,1):5.1f}%)  static={v[2]:4d}  smem_wavefronts={v[3]}/{v[4]}")
