"""Packet-mode FEC on the GPU: dabgpu_packet_fec_decode (k_packet_fec: 12 RS(204,188) rows per FEC
frame, read with stride 12 straight out of the transport order) and the MSC_Reed_Solomon_Data_Packet_Processor adapter class,
against the golden vectors recorded from the reference build.  Bar: identical bytes, row counts, callback order and flags."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN
from test_adapters_gpu import _run, driver  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def test_packet_fec_frames_match_reference_golden(gpu_ctx):
    kat = np.load(os.path.join(GOLDEN, "packet_fec_kat.npz"))
    g = gpu_ctx.DabGpu(mode=1, max_streams=1)
    fixed, counts = g.packet_fec_decode(kat["frames"])
    assert np.array_equal(counts, kat["row_counts"]), "per-row Reed-Solomon results differ from Reed_Solomon_Decoder::Decode"
    assert np.array_equal(fixed, kat["frames_fixed"]), "corrected application data table differs"
    assert np.array_equal(fixed[:, 2256:], kat["frames"][:, 2256:]), "the RS data table must be left as received"
    # batch boundaries: one frame at a time gives the same answer, an empty batch is a no-op
    one, c1 = g.packet_fec_decode(kat["frames"][5:6])
    assert np.array_equal(one[0], kat["frames_fixed"][5]) and np.array_equal(c1[0], kat["row_counts"][5])
    e, ce = g.packet_fec_decode(np.zeros((0, 2448), dtype=np.uint8))
    assert e.shape == (0, 2448) and ce.shape == (0, 12)
    g.close()


def test_packet_fec_large_batch_matches_oracle(gpu_ctx, pyref, tx):
    """512 frames (6144 codewords) with 0..10 errors per row against the C restatement's RS decoder"""
    rng = np.random.default_rng(5)
    base = tx.packet_fec_set(rng)
    frame = np.concatenate([np.concatenate(base[:-9])] + [p[2:24 if i < 8 else 18] for i, p in enumerate(base[-9:])])
    frames = np.tile(frame, (512, 1))
    for f in range(512):
        for y in range(12):
            n = int(rng.integers(0, 11))
            x = rng.choice(204, size=n, replace=False)
            frames[f, 12 * x + y] ^= rng.integers(1, 256, n, dtype=np.uint8)
    g = gpu_ctx.DabGpu(mode=1, max_streams=1)
    fixed, counts = g.packet_fec_decode(frames)
    g.close()
    rs = pyref.PortRS(16, 51)
    for f in range(0, 512, 37):
        for y in range(12):
            c, d, _ = rs.decode(frames[f, y::12])
            assert counts[f, y] == c
            exp = d[:188] if c >= 0 else frames[f, y:2256:12]
            assert np.array_equal(fixed[f, y:2256:12], exp)
    ok = counts >= 0
    assert ok.mean() > 0.7 and (~ok).any()      # rows with more than 8 errors fail, the others are repaired
    assert np.array_equal(fixed[ok.all(axis=1)][:, :2256], np.tile(frame[:2256], (int(ok.all(axis=1).sum()), 1)))


def test_packet_processor_class_matches_reference_golden(driver, pyref):  # noqa: F811
    kat = np.load(os.path.join(GOLDEN, "packet_fec_kat.npz"))
    off = np.concatenate([[0], np.cumsum(kat["call_len"])])
    loff = np.concatenate([[0], np.cumsum(kat["log_len"])])
    n = kat["call_len"].size
    payload = struct.pack("<i", n) + b"".join(struct.pack("<i", int(kat["call_len"][i])) + kat["calls"][off[i]:off[i + 1]].tobytes() for i in range(n))
    out = _run(driver, "pktfec", payload)
    pos = 0
    for i in range(n):
        used, fired = struct.unpack_from("<ii", out, pos)
        pos += 8
        got = []
        for _ in range(fired):
            ln, ok = struct.unpack_from("<ii", out, pos)
            pos += 8
            got.append((out[pos:pos + ln], bool(ok)))
            pos += ln
        exp = pyref.parse_packet_log(kat["logs"][loff[i]:loff[i + 1]].tobytes())
        assert used == int(kat["used"][i]), f"call {i}: consumed {used}"
        assert got == exp, f"call {i}: callbacks differ from the reference's"
    assert pos == len(out)
