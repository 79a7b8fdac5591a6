"""Generates tests/golden/packet_fec_kat.npz with the REFERENCE's MSC_Reed_Solomon_Data_Packet_Processor and
Reed_Solomon_Decoder(8, 0x11D, 0, 1, 16, 51) (compiled into oracle/_ref/libdabref.so).
Run where /root/reference exists:  python tests/golden/make_packet_fec_golden.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import pyref  # noqa: E402

tx = importlib.import_module("sdrplusplus-dab-radio-plugin_b200.synth.dabtx")


def corrupt(rng, packets, n_errors, headers=False):
    for _ in range(n_errors):
        k = int(rng.integers(0, len(packets)))
        lo = 0 if headers else 2
        packets[k][int(rng.integers(lo, packets[k].size))] ^= int(rng.integers(1, 256))


def corrupt_rows(rng, packets, per_row):
    """exactly per_row byte errors in every row of the application data table (headers included)"""
    flat = np.concatenate(packets[:-9])
    for y in range(12):
        x = rng.choice(188, size=per_row, replace=False)
        flat[12 * x + y] ^= rng.integers(1, 256, per_row, dtype=np.uint8)
    off = 0
    for p in packets[:-9]:
        keep = p[0] & 0xC0                       # the length field decides how the stream is cut: leave it alone here
        p[:] = flat[off:off + p.size]
        p[0] = (p[0] & 0x3F) | keep
        off += p.size


def scenarios(rng):
    """-> list of buffers handed to ReadPacket one call each (the way basic_data_packet_channel.cpp:49-88 walks a frame)"""
    calls = []
    # 0: clean set; 1: a few byte errors (every row correctable); 2: burst beyond the code (some rows fail)
    calls += tx.packet_fec_set(rng)
    s = tx.packet_fec_set(rng); corrupt_rows(rng, s, 6); calls += s
    s = tx.packet_fec_set(rng); corrupt(rng, s, 60, headers=False); s[3][2:] ^= 0x5A; s[4][2:] ^= 0xA5; calls += s
    # 3: a data packet lost -> incomplete set, everything leaves uncorrected
    s = tx.packet_fec_set(rng); del s[5]; calls += s
    # 4: FEC counter sequence broken in the middle, then a stray FEC packet with a non-zero counter
    s = tx.packet_fec_set(rng); s[-4][0] ^= 0x04; calls += s
    # 5: one extra packet before a set: the oldest packets are evicted to make room, the set no longer lines up
    s = tx.packet_fec_set(rng, [24] * 94); calls += [rng.integers(0, 256, 24, dtype=np.uint8) & np.uint8(0x3F)] + s
    # 6: errors in packet headers (length field / address), corrected by the code before the packets are cut out again
    s = tx.packet_fec_set(rng); corrupt(rng, s[:-9], 30, headers=True); s[0][0] ^= 0xC0; calls += s
    # 7: FEC packets whose length field claims 96 bytes (ignored), buffers longer than the packet, short buffers
    s = tx.packet_fec_set(rng)
    for p in s[-9:]:
        p[0] |= 0xC0
    s = [np.concatenate([p, rng.integers(0, 256, 7, dtype=np.uint8)]) if i % 5 == 0 else p for i, p in enumerate(s)]
    calls += s
    calls += [np.zeros(0, dtype=np.uint8), np.array([0x40], dtype=np.uint8), np.array([0xC0, 0x10] + [1] * 40, dtype=np.uint8)]
    # 8: two clean sets back to back with all four packet lengths, one error per row in the second
    calls += tx.packet_fec_set(rng, [96] * 20 + [72] * 2 + [48] * 3 + [24] * 2)
    s = tx.packet_fec_set(rng); corrupt(rng, s[:-9], 12); calls += s
    return calls


def main():
    assert pyref.ref_available()
    rng = np.random.default_rng(2024)
    calls = scenarios(rng)
    ref = pyref.RefPacketFec()
    used, logs = [], []
    for c in calls:
        u, cb = ref.read_packet(c)
        used.append(u)
        flat = b"".join(np.array([len(p), int(ok)], dtype="<i4").tobytes() + p + b"\0" * ((-len(p)) % 4) for p, ok in cb)
        logs.append(np.frombuffer(flat, dtype=np.uint8))
    # FEC frames for the ABI-level check: [2256 application data | 192 RS data], expected = every row through the reference RS decoder
    rs = pyref.RefRS(16, 51)
    frames, fixed, counts = [], [], []
    for k in range(24):
        s = tx.packet_fec_set(rng)
        fr = np.concatenate([np.concatenate(s[:-9])] + [p[2:24 if i < 8 else 18] for i, p in enumerate(s[-9:])])
        assert fr.size == 2448
        n_err = [0, 3, 40, 96, 200, 600][k % 6]
        pos = rng.choice(2448, size=n_err, replace=False)
        fr[pos] ^= rng.integers(1, 256, n_err, dtype=np.uint8)
        out, cnt = fr.copy(), np.zeros(12, dtype=np.int32)
        for y in range(12):
            c, d, _ = rs.decode(fr[y::12])
            cnt[y] = c
            if c >= 0:
                out[y:2256:12] = d[:188]
        frames.append(fr); fixed.append(out); counts.append(cnt)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "packet_fec_kat.npz"),
                        build_info=np.array(pyref.RefLib.get().build_info()),
                        calls=np.concatenate(calls), call_len=np.array([c.size for c in calls], dtype=np.int32), used=np.array(used, dtype=np.int32),
                        logs=np.concatenate(logs), log_len=np.array([l.size for l in logs], dtype=np.int32),
                        frames=np.stack(frames), frames_fixed=np.stack(fixed), row_counts=np.stack(counts))
    print(len(calls), "calls,", sum(len(pyref.parse_packet_log(l.tobytes())) for l in logs), "callbacks;", "row counts:", np.stack(counts).tolist()[:6])


if __name__ == "__main__":
    main()
