"""GPU path against the committed golden vectors (produced by the reference build, tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def test_gpu_viterbi_golden(gpu_ctx, vit_flags):
    z = _load("viterbi_kat.npz")
    keys = sorted(k[:-5] for k in z.files if k.endswith("_soft"))
    g = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=vit_flags)
    outs, perr = g.viterbi_decode([z[k + "_soft"] for k in keys], [[tuple(int(v) for v in r) for r in z[k + "_segs"]] for k in keys])
    for i, k in enumerate(keys):
        assert np.array_equal(outs[i], z[k + "_out"]), k
        assert int(perr[i]) == int(z[k + "_err"][0]), k
    g.close()


def test_gpu_channel_chain_golden(gpu_ctx, tx, vit_flags):
    z = _load("channel_kat.npz")
    subs = [tx.Subchannel(i, int(s[0]), int(s[1]), bool(s[2]), int(s[3]), int(s[4]), bool(s[5]), bool(s[6])) for i, s in enumerate(z["subs"])]
    used = int(z["used_bits"][0])
    g = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=vit_flags)
    g.msc_configure(0, subs)
    fib_counts, fibs = z["fib_counts"], z["fibs"]
    fpos, gi, out_idx = 0, 0, 0
    out_pos = [0] * len(subs)
    ev_sizes, ev_blob = z["aac_event_sizes"], z["aac_events"]
    ev_i, ev_pos = 0, 0
    for f in range(z["fic_soft"].shape[0]):
        frame = np.zeros(230400, dtype=np.int8)
        frame[:9216] = z["fic_soft"][f]
        for c in range(4):
            frame[9216 + c * 55296: 9216 + c * 55296 + used] = z["msc_soft"][f, c]
        g.softbits_push(frame[None, :])
        g.chan_decode()
        got_fibs, ok = g.get_fic(0)
        outs = [g.get_msc(0, k) for k in range(len(subs))]
        log = g.get_dabplus_events(0, 0)
        exp_log = b""
        for c in range(4):
            n = int(fib_counts[gi]); gi += 1
            exp = [fibs[fpos + 30 * i: fpos + 30 * (i + 1)].tobytes() for i in range(n)]
            fpos += 30 * n
            assert [got_fibs[3 * c + i, :30].tobytes() for i in range(3) if ok[3 * c + i]] == exp
            for k in range(len(subs)):
                sz = int(z[f"msc_out_{k}_sizes"][out_idx])
                out, valid = outs[k]
                assert bool(valid[c]) == (sz > 0)
                if sz:
                    assert np.array_equal(out[c], z[f"msc_out_{k}"][out_pos[k]:out_pos[k] + sz])
                out_pos[k] += sz
            out_idx += 1
        # events of this frame, re-serialised the way ref_harness.cpp does (payload padded to 4)
        while ev_i < len(ev_sizes) and len(exp_log) < len(log):
            n_ev = int(ev_sizes[ev_i])
            blob = ev_blob[ev_pos:ev_pos + n_ev].tobytes()
            exp_log += blob + b"\0" * ((-(n_ev - 24)) % 4)
            ev_pos += n_ev
            ev_i += 1
        assert log == exp_log, f
    assert ev_i == len(ev_sizes)
    g.close()


def test_gpu_rs_golden(gpu_ctx):
    z = _load("rs_kat.npz")
    g = gpu_ctx.DabGpu(mode=1, max_streams=1)
    counts, fixed, _ = g.rs_decode(z["cw"])
    assert np.array_equal(counts, z["counts"]) and np.array_equal(fixed, z["out"])
    g.close()


def test_gpu_ofdm_golden(gpu_ctx):
    z = _load("ofdm_mode2_kat.npz")
    g = gpu_ctx.DabGpu(mode=2, max_streams=1)
    u8, block = z["iq_u8"], int(z["block"][0])
    got = []
    for off in range(0, u8.size // 2, block):
        g.ofdm_process(u8[None, 2 * off:2 * (off + block)], block_size=block)
        got += g.ofdm_pop_frames(0)
    assert len(got) == z["soft"].shape[0]
    for i, f in enumerate(got):
        assert f[3] == int(z["toff"][i])
        assert np.abs(f[0].astype(np.int32) - z["soft"][i].astype(np.int32)).max() <= 1
    g.close()
