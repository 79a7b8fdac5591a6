#!/usr/bin/env python
"""Writes profiles/r2_ncu_constants.json, the per-kernel figures bench.py quotes in its `roofline` object, from the committed
`ncu --set full` captures of this round (one launch of each kernel in the default bench workload: 1024 streams, one frame each).
usage: python scripts/ncu_constants.py <lanes.ncu-rep> <demod.ncu-rep> <tag of the committed summaries>"""
import csv, io, json, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRAMES = 1024                    # frames one captured launch processes (streams per GPU of the default workload)
VIT_STEPS_PER_FRAME = 114120     # 72 x 1542 + 4 x 774 trellis steps
FRAME_SAMPLES = 196608
ALU = ('VIADDMNMX', 'VIMNMX', 'VIMNMX3', 'VABSDIFF4', 'LOP3', 'PRMT', 'ISETP', 'SHF', 'SEL', 'VIADD', 'IADD3', 'LEA', 'FMNMX', 'FSETP', 'FSEL', 'POPC', 'FLO', 'PLOP3', 'P2R', 'R2P', 'IABS')
FMA = ('IMAD', 'FFMA', 'FMUL', 'FADD', 'HFMA2', 'IDP')


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return {h: rows[2][i] for i, h in enumerate(rows[0])}, {h: rows[1][i] for i, h in enumerate(rows[0])}


def num(v):
    return float(v.replace(',', ''))


def to_bytes(v, unit):
    return num(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def sass_counts(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blk = out.split('"Kernel Name"')[1]
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    hdr = rows[1]
    si, ii = hdr.index('Source'), hdr.index('Instructions Executed')
    c = collections.Counter()
    for r in rows[2:]:
        if len(r) <= ii:
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[si])
        if m and r[ii].isdigit():
            c[m.group(2)] += int(r[ii])
    return c


def main():
    lanes, demod, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    out = {}
    v, u = raw(lanes)
    c = sass_counts(lanes)
    tot = sum(c.values())
    alu = sum(n for k, n in c.items() if k in ALU)
    fma = sum(n for k, n in c.items() if k in FMA)
    groups32 = FRAMES * VIT_STEPS_PER_FRAME / 32.0
    dram = to_bytes(v['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + to_bytes(v['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
    out['k_viterbi_lanes'] = {
        'dram_bytes_per_frame': dram / FRAMES, 'warp_inst_per_32_steps': tot / groups32, 'alu_pipe_warp_inst_per_32_steps': alu / groups32,
        'fma_pipe_warp_inst_per_32_steps': fma / groups32, 'issue_active_pct': num(v['smsp__issue_active.avg.pct_of_peak_sustained_active']),
        'alu_pipe_pct': num(v['sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active']),
        'duration_us': num(v['gpu__time_duration.sum']) * {'us': 1, 'ms': 1e3, 'ns': 1e-3}[u['gpu__time_duration.sum']],
        'source': f'profiles/{tag}_viterbi_lanes_ncu_full.txt, profiles/{tag}_viterbi_lanes_sass_hist.txt (one launch, 1024 frames)'}
    v, u = raw(demod)
    c = sass_counts(demod)
    dram = to_bytes(v['dram__bytes_read.sum'], u['dram__bytes_read.sum']) + to_bytes(v['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
    out['k_ofdm_demod2'] = {
        'dram_bytes_per_frame': dram / FRAMES,
        'warp_inst_per_sample': sum(c.values()) * 32.0 / (FRAMES * FRAME_SAMPLES),   # thread instructions per IQ sample
        'issue_active_pct': num(v['smsp__issue_active.avg.pct_of_peak_sustained_active']),
        'duration_us': num(v['gpu__time_duration.sum']) * {'us': 1, 'ms': 1e3, 'ns': 1e-3}[u['gpu__time_duration.sum']],
        'source': f'profiles/{tag}_ofdm_demod_ncu_full.txt (one launch, 1024 frames)'}
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'r2_ncu_constants.json'), 'w'), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
