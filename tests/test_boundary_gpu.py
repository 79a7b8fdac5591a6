"""Behaviour of the drop-in boundary that the reference gets from its object model (one decoder object per sub-channel, a blocking
ring between OFDM_Demod and BasicRadio): incremental sub-channel attach, frame-ring overrun handling, channel-stream snapshot.
Everything goes through the C ABI and is compared with oracle decoders started at the same CIF."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(pyref, sc):
    mk = pyref.RefMsc if pyref.ref_available() else pyref.PortMsc
    return mk(sc.start_address, sc.length, sc.is_uep, sc.uep_index, sc.eep_level, sc.eep_type_b)


def _cif(frame, c):
    return frame[9216 + c * 55296: 9216 + (c + 1) * 55296]


def test_late_subchannel_does_not_perturb_running_decoders(gpu_ctx, tx, pyref):
    """BasicRadio::UpdateAfterProcessing (basic_radio.cpp:98-131) attaches a decoder to a newly complete sub-channel and leaves the
    others alone: after dabgpu_msc_add_subchannel the bytes and DAB+ events of the sub-channels that were already running are the
    ones of a context that had been told everything from the start, and the late one equals an oracle decoder born at that frame.
    Re-declaring an unchanged table with dabgpu_msc_configure keeps the state as well; removing an entry keeps the others."""
    rng = np.random.default_rng(3)
    a = tx.Subchannel(0, 0, 48, eep_level=2)
    b = tx.Subchannel(1, 48, 54, eep_level=2, eep_type_b=True, dabplus=False)
    c_ = tx.Subchannel(2, 102, 35, is_uep=True, uep_index=4, dabplus=False)
    ens = tx.EnsembleTx(1, [a, b, c_], seed=13)
    told = gpu_ctx.DabGpu(mode=1, max_streams=1)
    incr = gpu_ctx.DabGpu(mode=1, max_streams=1)
    told.msc_configure(0, [a, b, c_])
    incr.msc_configure(0, [a])
    late_b, late_c = 5, 7
    o_b, o_c = _mk(pyref, b), _mk(pyref, c_)
    idx = {0: 0}
    compared = {0: 0, 1: 0, 2: 0}
    for f in range(12):
        frame = tx.hard_to_soft(ens.next_frame_bits(), rng, snr_db=5.0)
        if f == late_b:
            idx[1] = incr.msc_add_subchannel(0, b)
        if f == late_c:
            idx[2] = incr.msc_add_subchannel(0, c_)
            incr.msc_configure(0, [a, b, c_])      # same table again: nothing restarts
        if f == 10:
            incr.msc_remove_subchannel(0, idx[1])
            with pytest.raises(gpu_ctx.DabGpuError):
                incr.get_msc(0, idx[1])
        for g in (told, incr):
            g.softbits_push(frame[None, :])
            g.chan_decode()
        out_t, valid_t = told.get_msc(0, 0)
        out_i, valid_i = incr.get_msc(0, idx[0])
        assert np.array_equal(valid_t, valid_i), f
        for c in range(4):
            if valid_t[c]:
                assert np.array_equal(out_t[c], out_i[c]), (f, c)
                compared[0] += 1
        assert told.get_dabplus_events(0, 0) == incr.get_dabplus_events(0, idx[0]), f
        for k, oracle, born in ((1, o_b, late_b), (2, o_c, late_c)):
            if f < born or (k == 1 and f >= 10):
                continue
            out, valid = incr.get_msc(0, idx[k])
            for c in range(4):
                exp = oracle.decode_cif(_cif(frame, c))
                assert bool(valid[c]) == (exp.size > 0), (f, k, c)
                if exp.size:
                    assert np.array_equal(out[c], exp), (f, k, c)
                    compared[k] += 1
    assert idx == {0: 0, 1: 1, 2: 2}
    assert compared[0] >= 30 and compared[1] >= 4 and compared[2] >= 4, compared
    told.close()
    incr.close()


def test_frame_ring_full_and_overrun(gpu_ctx, tx, pyref):
    """The soft-bit frame ring is the reference's ThreadedRingBuffer between OFDM_Demod and BasicRadio (src/radio_block.cpp:20-44).
    dabgpu_softbits_push refuses a frame that would overwrite the history of an undecoded one.  A decoder that fell behind the
    OFDM stage skips to the newest frame, restarts its de-interleavers empty and counts the dropped frames -- it never emits
    bytes decoded from overwritten slots."""
    rng = np.random.default_rng(4)
    sc = tx.Subchannel(0, 10, 48, eep_level=2, dabplus=False)
    ens = tx.EnsembleTx(1, [sc], seed=23)
    # (1) push without decoding until the ring is full
    g = gpu_ctx.DabGpu(mode=1, max_streams=1, frame_slots=8)
    g.msc_configure(0, [sc])
    pushed = 0
    with pytest.raises(gpu_ctx.DabGpuError) as ei:
        for _ in range(10):
            g.softbits_push(tx.hard_to_soft(ens.next_frame_bits(), rng, snr_db=8.0)[None, :])
            pushed += 1
    assert ei.value.code == gpu_ctx.ERR_OVERFLOW and 2 <= pushed <= 4
    for k in range(pushed):                       # what was accepted is decodable in order
        g.chan_decode()
        assert g.chan_status(0) == (1, k)
    g.chan_decode()
    assert g.chan_status(0)[0] == 0
    assert g.counters()["frames_dropped"] == 0
    g.close()

    # (2) the OFDM stage runs 7 frames ahead of a decoder that is only started afterwards
    mode, block = 1, 65536
    ens = tx.EnsembleTx(mode, [sc], seed=24)
    n_frames = 16
    iq = tx.ofdm_modulate([ens.next_frame_bits() for _ in range(n_frames)], mode)
    u8 = tx.to_u8(tx.impair(iq, 18.0, 0.7e-3, 999, seed=2, tail_samples=3000), 30.0)
    g = gpu_ctx.DabGpu(mode=mode, max_streams=1, frame_slots=8)
    g.msc_configure(0, [sc])
    n = (u8.size // 2 // block) * block
    frames = []
    oracle = None
    decoded_idx = []
    checked = 0
    for off in range(0, n, block):
        g.ofdm_process(u8[None, 2 * off:2 * (off + block)], block_size=block)
        frames += [fr[0] for fr in g.ofdm_pop_frames(0)]
        if len(frames) < 7:
            continue                               # decoder not started yet: frames pile up in the ring
        while True:
            g.chan_decode()
            dec, fi = g.chan_status(0)
            if not dec:
                break
            if oracle is None:
                assert fi == len(frames) - 1 and fi >= 6, "the late decoder must skip to the newest frame"
                assert g.counters()["frames_dropped"] == fi
                oracle = _mk(pyref, sc)            # a fresh MSC_Decoder that sees its first CIF now
            decoded_idx.append(fi)
            out, valid = g.get_msc(0, 0)
            for c in range(4):
                exp = oracle.decode_cif(_cif(frames[fi], c))
                assert bool(valid[c]) == (exp.size > 0), (fi, c)
                if exp.size:
                    assert np.array_equal(out[c], exp), (fi, c)
                    checked += 1
    assert decoded_idx == list(range(decoded_idx[0], decoded_idx[0] + len(decoded_idx))) and len(decoded_idx) >= 6
    assert checked >= 8
    st = g.ofdm_status(0)
    assert st["frames_dropped"] == 0               # every frame was popped in time on the OFDM side
    g.close()


def test_pop_frames_reports_dropped_frames(gpu_ctx, tx):
    """dabgpu_ofdm_pop_frames: a caller that falls more than frame_slots - 1 frames behind gets the newest frame_slots - 1 frames
    and a dropped-frames count in dabgpu_ofdm_status (the reference's observers never lose a frame)."""
    mode, block = 2, 65536
    n_frames = 48
    rng = np.random.default_rng(1)
    P = gpu_ctx.get_params(mode)
    bits = [rng.integers(0, 2, size=P.nb_frame_bits, dtype=np.uint8) for _ in range(n_frames)]
    u8 = tx.to_u8(tx.impair(tx.ofdm_modulate(bits, mode), 20.0, 0.4e-3, 333, seed=3, tail_samples=2000), 30.0)
    g = gpu_ctx.DabGpu(mode=mode, max_streams=1, frame_slots=32)
    n = (u8.size // 2 // block) * block
    for off in range(0, n, block):
        g.ofdm_process(u8[None, 2 * off:2 * (off + block)], block_size=block)
    total = g.ofdm_status(0)["total_frames_read"]
    assert total > 40
    got = g.ofdm_pop_frames(0, max_frames=64)
    assert len(got) == 31
    st = g.ofdm_status(0)
    assert st["frames_dropped"] == total - 31 and st["frames_queued"] == 0
    g.close()


def test_channel_decode_on_streams_without_a_frame_yet(gpu_ctx, tx, monkeypatch):
    """The channel decode runs on its own CUDA stream while the next OFDM stage advances frames_written on the main stream: the
    decode kernels read a snapshot of the counters taken before the fork.  Streams that acquire at different times (leads up to a
    whole frame) are decoded every step from the very first one, in the forked and in the inline mode: same status, same bytes,
    same counters."""
    S, block, n_frames = 6, 65536, 9
    subs = [tx.Subchannel(0, 0, 48, eep_level=2), tx.Subchannel(1, 48, 16, is_uep=True, uep_index=0, dabplus=False)]
    recs = []
    for s in range(S):
        ens = tx.EnsembleTx(1, subs, seed=40 + s)
        iq = tx.ofdm_modulate([ens.next_frame_bits() for _ in range(n_frames)], 1)
        recs.append(tx.to_u8(tx.impair(iq, 17.0, (s - 2) * 0.9e-3, 100 + 36000 * s, seed=s, tail_samples=1000), 30.0))
    n = (min(r.size for r in recs) // 2 // block) * block
    monkeypatch.setenv("DABGPU_CHAN_INLINE", "1")
    inline = gpu_ctx.DabGpu(mode=1, max_streams=S)
    monkeypatch.delenv("DABGPU_CHAN_INLINE")
    forked = gpu_ctx.DabGpu(mode=1, max_streams=S)
    for g in (inline, forked):
        for s in range(S):
            g.msc_configure(s, subs)
    for off in range(0, n, block):
        blk = np.stack([r[2 * off:2 * (off + block)] for r in recs])
        res = []
        for g in (inline, forked):
            g.chan_decode()                       # forked: overlaps the OFDM stage below, which may emit a stream's first frame
            g.ofdm_process(blk, block_size=block)
            item = []
            for s in range(S):
                item.append(np.array(g.chan_status(s)))
                fibs, ok = g.get_fic(s)
                item += [fibs, ok]
                for k in range(len(subs)):
                    out, valid = g.get_msc(s, k)
                    item += [valid] + [out[c] for c in range(4) if valid[c]]
            res.append(item)
        assert len(res[0]) == len(res[1])
        for x, y in zip(res[0], res[1]):
            assert np.array_equal(x, y), off
    ci, cf = inline.counters(), forked.counters()
    assert ci == cf and ci["frames_channel_decoded"] >= S * (n_frames - 3) and ci["frames_dropped"] == 0
    inline.close()
    forked.close()
