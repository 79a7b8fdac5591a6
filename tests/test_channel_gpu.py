"""GPU parity of the integer half of the chain (through the C ABI) against both oracles.

Bar: bit-exact decoded bytes, path errors, CRC flags, RS corrections and superframe events.
"""
import numpy as np
import pytest

from conftest import VIT_LANES_ALWAYS

pytestmark = pytest.mark.gpu


def _mk_trellis_cases(tx, rng, n):
    seg_choices = [tx.FIC_SEGMENTS, tx.eep_segments(48, 2, False), tx.eep_segments(8, 1, False), tx.uep_segments(4),
                   tx.eep_segments(54, 2, True), tx.eep_segments(96, 0, False), tx.eep_segments(40, 3, False), tx.uep_segments(0)]
    softs, segs = [], []
    for trial in range(n):
        sg = seg_choices[trial % len(seg_choices)]
        n_in = int(tx.puncture_mask(sg).sum())
        info = rng.integers(0, 256, size=(sum(b for _, b in sg) // 4 - 6) // 8, dtype=np.uint8)
        enc = tx.channel_encode(info, sg)
        kind = (trial // len(seg_choices)) % 5
        if kind == 0:
            soft = tx.hard_to_soft(enc)                                           # noiseless
        elif kind == 1:
            soft = tx.hard_to_soft(enc, rng, snr_db=float(rng.uniform(-6, 6)))    # noisy
        elif kind == 2:
            soft = rng.integers(-128, 128, size=n_in).astype(np.int8)             # garbage incl. -128
        elif kind == 3:
            soft = (rng.integers(-1, 2, size=n_in) * 127).astype(np.int8)         # tie heavy
        else:
            soft = np.where(rng.random(n_in) < 0.5, -128, 127).astype(np.int8)    # saturating metrics
        softs.append(soft)
        segs.append(sg)
    return softs, segs


def test_viterbi_matches_oracles(gpu_ctx, tx, pyref, vit_flags):
    rng = np.random.default_rng(11)
    softs, segs = _mk_trellis_cases(tx, rng, 160)
    g = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=vit_flags)
    outs, perr = g.viterbi_decode(softs, segs)
    port = pyref.PortViterbi()
    ref = pyref.RefViterbi() if pyref.ref_available() else None
    for i, (soft, sg) in enumerate(zip(softs, segs)):
        exp, _, exp_err = port.decode(soft, sg)
        assert np.array_equal(outs[i], exp), f"trellis {i}: bytes differ from the C restatement"
        assert int(perr[i]) == exp_err, f"trellis {i}: path error {int(perr[i])} != {exp_err}"
        if ref is not None:
            r, _, r_err = ref.decode(soft, sg)
            assert np.array_equal(outs[i], r), f"trellis {i}: bytes differ from the reference build"
            assert int(perr[i]) == r_err
    g.close()


def test_viterbi_lane_short_and_general_branch_error_forms(gpu_ctx, tx, pyref):
    """k_viterbi_lanes takes the short branch-error form (Ei[p] = E[p ^ 7], dot-product errors) when k_vit_prep saw no -128 symbol
    in the call and the general one otherwise (viterbi_lane_core.h: vl_branch).  A call without any -128 -- noiseless, noisy, tie
    heavy and full-range garbage clipped at -127 -- must equal the oracle; the same call with ONE symbol set to -128 switches the
    whole call to the general form and must equal the oracle again, bytes and path errors."""
    rng = np.random.default_rng(21)
    softs, segs = _mk_trellis_cases(tx, rng, 120)
    softs = [np.maximum(s, -127).astype(np.int8) for s in softs]
    assert all(int(s.min()) >= -127 for s in softs)
    port = pyref.PortViterbi()
    ref = pyref.RefViterbi() if pyref.ref_available() else None
    g = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=VIT_LANES_ALWAYS)
    for variant in ("short", "general"):
        if variant == "general":
            softs[37] = softs[37].copy()
            softs[37][softs[37].size // 2] = -128
        outs, perr = g.viterbi_decode(softs, segs)
        for i, (soft, sg) in enumerate(zip(softs, segs)):
            exp, _, exp_err = (ref or port).decode(soft, sg)
            assert np.array_equal(outs[i], exp), f"{variant} form, trellis {i}: bytes differ"
            assert int(perr[i]) == exp_err, f"{variant} form, trellis {i}: path error {int(perr[i])} != {exp_err}"
    g.close()


def test_viterbi_long_trellis_and_renormalisation(gpu_ctx, tx, pyref, vit_flags):
    # 864 CU at EEP 4-A: 27654 steps; noisy input forces several renormalisations
    rng = np.random.default_rng(12)
    sg = tx.eep_segments(864, 3, False)
    n_in = int(tx.puncture_mask(sg).sum())
    info = rng.integers(0, 256, size=(sum(b for _, b in sg) // 4 - 6) // 8, dtype=np.uint8)
    enc = tx.channel_encode(info, sg)
    # without -128 the kernel stays on its packed u16x2 steps: full-scale garbage and full-scale ties drive state 0
    # through the saturation band and a renormalisation every few chunks; moderate noise does the same slowly
    softs = [tx.hard_to_soft(enc, rng, snr_db=-3.0), rng.integers(-128, 128, size=n_in).astype(np.int8),
             rng.integers(-127, 128, size=n_in).astype(np.int8), (rng.integers(-1, 2, size=n_in) * 127).astype(np.int8),
             np.where(rng.random(n_in) < 0.5, -127, 127).astype(np.int8), tx.hard_to_soft(enc, rng, snr_db=8.0)]
    g = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=vit_flags)
    outs, perr = g.viterbi_decode(softs, [sg] * len(softs), descramble=True)
    port = pyref.PortViterbi()
    prbs = pyref.port_scrambler_bytes(outs[0].size)
    for i in range(len(softs)):
        exp, _, exp_err = port.decode(softs[i], sg)
        assert np.array_equal(outs[i], exp ^ prbs)
        assert int(perr[i]) == exp_err
    g.close()


def test_viterbi_lane_plan_edge_cases(gpu_ctx, tx, pyref):
    """The device-side plan of the lane path: many length classes with partial groups in one call, and a trellis longer than
    the classes cover (>= 73728 steps), which must send the whole call back to the warp-per-trellis kernel."""
    rng = np.random.default_rng(13)
    port = pyref.PortViterbi()
    g = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=2)   # DABGPU_FLAG_VIT_LANES_ALWAYS
    # 75 trellises of 15 different lengths (6 .. 9222 steps): every class ends in a partial group
    segs = [[(24, 4 * n)] for n in (6, 7, 38, 262, 270, 774, 1030, 1542, 1548, 2054, 3078, 4102, 6150, 8198, 9222) for _ in range(5)]
    softs = [rng.integers(-127, 128, size=sg[0][1]).astype(np.int8) for sg in segs]
    outs, perr = g.viterbi_decode(softs, segs)
    for i, (soft, sg) in enumerate(zip(softs, segs)):
        exp, _, exp_err = port.decode(soft, sg)
        assert np.array_equal(outs[i], exp) and int(perr[i]) == exp_err, i
    # one oversize trellis next to short ones
    segs = [[(24, 4 * 80006)], [(24, 4 * 774)], [(24, 4 * 1542)]]
    softs = [rng.integers(-127, 128, size=sg[0][1]).astype(np.int8) for sg in segs]
    outs, perr = g.viterbi_decode(softs, segs)
    for i, (soft, sg) in enumerate(zip(softs, segs)):
        exp, _, exp_err = port.decode(soft, sg)
        assert np.array_equal(outs[i], exp) and int(perr[i]) == exp_err, i
    g.close()


def test_viterbi_warp_kernel_longest_first_order(gpu_ctx, tx, pyref):
    """A call of 1 024 .. 6 143 trellises stays on the warp-per-trellis kernel, which then pulls its jobs by length class, longest
    first (launch_viterbi: order_plan).  Mixed lengths incl. UEP row 63 (9 222 steps) and partial classes; bytes and path errors
    of every trellis equal the oracle's, and equal what the same call gives with the order switched off (vit-warps flag)."""
    rng = np.random.default_rng(29)
    port = pyref.PortViterbi()
    kinds = [tx.FIC_SEGMENTS, tx.eep_segments(48, 2, False), tx.uep_segments(0), tx.uep_segments(14), tx.uep_segments(37), tx.eep_segments(8, 1, False),
             tx.eep_segments(54, 2, True)]
    segs = [kinds[i % len(kinds)] for i in range(1190)] + [tx.uep_segments(63)] * 10
    order = rng.permutation(len(segs))
    segs = [segs[i] for i in order]
    softs = [np.clip(rng.normal(0.0, 60.0, size=int(tx.puncture_mask(sg).sum())), -127, 127).astype(np.int8) for sg in segs]
    g = gpu_ctx.DabGpu(mode=1, max_streams=1)            # auto: 1 200 trellises < the lane kernel's threshold
    outs, perr = g.viterbi_decode(softs, segs)
    g.close()
    g2 = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=4)  # DABGPU_FLAG_VIT_LANES_NEVER: no plan at all, jobs in index order
    outs2, perr2 = g2.viterbi_decode(softs, segs)
    g2.close()
    for i in range(len(segs)):
        assert np.array_equal(outs[i], outs2[i]) and int(perr[i]) == int(perr2[i]), i
    for i in list(range(0, len(segs), 7)) + [j for j, sg in enumerate(segs) if sum(n for _, n in sg) // 4 > 9000]:
        exp, _, exp_err = port.decode(softs[i], segs[i])
        assert np.array_equal(outs[i], exp) and int(perr[i]) == exp_err, i


def test_viterbi_bad_arguments(gpu_ctx, tx):
    g = gpu_ctx.DabGpu(mode=1, max_streams=1)
    soft = np.zeros(100, dtype=np.int8)
    with pytest.raises(gpu_ctx.DabGpuError) as e:
        g.viterbi_decode([soft], [tx.FIC_SEGMENTS])      # not enough punctured symbols
    assert e.value.code == gpu_ctx.ERR_INVALID
    with pytest.raises(gpu_ctx.DabGpuError):
        g.viterbi_decode([np.zeros(2304, np.int8)], [[(25, 128), (0, 24)]])   # invalid puncture code
    g.close()


def _subs(tx):
    return [tx.Subchannel(0, 0, 48, eep_level=2), tx.Subchannel(1, 48, 8, eep_level=1, dabplus=False),
            tx.Subchannel(2, 56, 54, eep_level=2, eep_type_b=True, dabplus=False),
            tx.Subchannel(3, 110, 35, is_uep=True, uep_index=4, dabplus=False),
            tx.Subchannel(4, 145, 16, is_uep=True, uep_index=0, dabplus=False),
            tx.Subchannel(5, 200, 96, eep_level=0), tx.Subchannel(6, 300, 40, eep_level=3)]


@pytest.mark.parametrize("snr_db", [None, 4.0])
def test_frame_decode_fic_msc_dabplus(gpu_ctx, tx, pyref, snr_db, vit_flags):
    """Soft-bit frames in -> FIBs, sub-channel bytes, superframe events out; two streams with different seeds."""
    rng = np.random.default_rng(21)
    subs = _subs(tx)
    n_streams, n_frames = 2, 12
    g = gpu_ctx.DabGpu(mode=1, max_streams=n_streams, flags=vit_flags)
    enss = [tx.EnsembleTx(1, subs, seed=100 + s) for s in range(n_streams)]
    use_ref = pyref.ref_available()
    mk_msc = (lambda sc: pyref.RefMsc(sc.start_address, sc.length, sc.is_uep, sc.uep_index, sc.eep_level, sc.eep_type_b)) if use_ref else \
             (lambda sc: pyref.PortMsc(sc.start_address, sc.length, sc.is_uep, sc.uep_index, sc.eep_level, sc.eep_type_b))
    o_msc = [[mk_msc(sc) for sc in subs] for _ in range(n_streams)]
    o_aac = [[(pyref.RefAac() if use_ref else pyref.PortAac()) if sc.dabplus else None for sc in subs] for _ in range(n_streams)]
    o_fic = pyref.RefFic() if use_ref else pyref.PortFic()
    for s in range(n_streams):
        g.msc_configure(s, subs)
    n_checked_bytes = 0
    n_events = 0
    for f in range(n_frames):
        frames = np.stack([tx.hard_to_soft(e.next_frame_bits(), rng, snr_db=snr_db) for e in enss])
        g.softbits_push(frames)
        g.chan_decode()
        for s in range(n_streams):
            decoded, idx = g.chan_status(s)
            assert decoded == 1 and idx == f
            fibs, ok = g.get_fic(s)
            exp_fibs = []
            for c in range(4):
                exp_fibs += o_fic.decode_group(frames[s, c * 2304:(c + 1) * 2304], c)
            got_fibs = [fibs[i, :30].tobytes() for i in range(12) if ok[i]]
            assert got_fibs == exp_fibs
            for k, sc in enumerate(subs):
                out, valid = g.get_msc(s, k)
                log = b""
                for c in range(4):
                    cif = frames[s, 9216 + c * 55296: 9216 + (c + 1) * 55296]
                    exp = o_msc[s][k].decode_cif(cif)
                    assert bool(valid[c]) == (exp.size > 0), (f, s, k, c)
                    if exp.size:
                        assert np.array_equal(out[c], exp), (f, s, k, c)
                        n_checked_bytes += exp.size
                        if sc.dabplus:
                            for ev in o_aac[s][k].process(exp):
                                log += np.array(ev[:5] + (len(ev[5]),), dtype=np.int32).tobytes() + ev[5] + b"\0" * ((-len(ev[5])) % 4)
                if sc.dabplus:
                    got = g.get_dabplus_events(s, k)
                    assert got == log, (f, s, k, pyref.parse_event_log(got)[:3], pyref.parse_event_log(log)[:3])
                    n_events += len(pyref.parse_event_log(log))
    assert n_checked_bytes > 0 and n_events > 0
    c = g.counters()
    assert c["frames_channel_decoded"] == n_streams * n_frames
    g.close()


def test_dabplus_with_byte_errors(gpu_ctx, tx, pyref):
    """Superframes with injected byte errors: <=5 per codeword corrected, >5 uncorrectable, fire-code resync."""
    rng = np.random.default_rng(5)
    sc = tx.Subchannel(0, 0, 48, eep_level=2)
    g = gpu_ctx.DabGpu(mode=1, max_streams=1)
    g.msc_configure(0, [sc])
    ens = tx.EnsembleTx(1, [sc], seed=9)
    # corrupt the payload before channel coding by patching the transmitter's superframe builder
    orig = tx.build_superframe
    counter = {"n": 0}

    def corrupt(*a, **k):
        sf = orig(*a, **k)
        n = counter["n"]
        counter["n"] += 1
        ncw = sf.size // 120
        if n % 4 == 1:      # correctable: up to 5 errors in some codewords
            for cw in range(0, ncw, 2):
                for j in rng.choice(120, size=int(rng.integers(1, 6)), replace=False):
                    sf[cw + j * ncw] ^= rng.integers(1, 256)
        elif n % 4 == 2:    # uncorrectable codeword 3
            for j in rng.choice(120, size=9, replace=False):
                sf[3 + j * ncw] ^= rng.integers(1, 256)
        elif n % 4 == 3:    # broken AU CRC without touching RS: flip payload byte and re-encode parity is not done => RS fixes it
            sf[40] ^= 0x55
        return sf

    tx.build_superframe = corrupt
    try:
        use_ref = pyref.ref_available()
        o_msc = pyref.RefMsc(0, 48) if use_ref else pyref.PortMsc(0, 48)
        o_aac = pyref.RefAac() if use_ref else pyref.PortAac()
        kinds = set()
        for f in range(24):
            frame = tx.hard_to_soft(ens.next_frame_bits())[None, :]
            g.softbits_push(frame)
            g.chan_decode()
            log = b""
            for c in range(4):
                exp = o_msc.decode_cif(frame[0, 9216 + c * 55296: 9216 + (c + 1) * 55296])
                if exp.size:
                    for ev in o_aac.process(exp):
                        kinds.add(ev[0])
                        log += np.array(ev[:5] + (len(ev[5]),), dtype=np.int32).tobytes() + ev[5] + b"\0" * ((-len(ev[5])) % 4)
            assert g.get_dabplus_events(0, 0) == log, f
        assert {pyref.EV_RS_ERROR, pyref.EV_HEADER, pyref.EV_AU} <= kinds
    finally:
        tx.build_superframe = orig
    g.close()


def test_rs_batch_matches_oracle(gpu_ctx, tx, pyref):
    rng = np.random.default_rng(3)
    n = 600
    cws = np.zeros((n, 120), dtype=np.uint8)
    for t in range(n):
        data = rng.integers(0, 256, size=110).tolist()
        cw = np.array(data + tx.rs_encode(data, 10), dtype=np.uint8)
        for p in rng.choice(120, size=t % 9, replace=False):
            cw[p] ^= rng.integers(1, 256)
        if t % 50 == 49:
            cw = rng.integers(0, 256, size=120).astype(np.uint8)
        cws[t] = cw
    g = gpu_ctx.DabGpu(mode=1, max_streams=1)
    counts, fixed, pos = g.rs_decode(cws)
    o = pyref.RefRS() if pyref.ref_available() else pyref.PortRS()
    for t in range(n):
        c, d, p = o.decode(cws[t])
        assert counts[t] == c, t
        assert np.array_equal(fixed[t], d), t
        if c > 0:
            assert np.array_equal(pos[t, :c], p), t
    # packet-mode code RS(204,188): 16 roots, pad 51 (msc_reed_solomon_data_packet_processor.cpp:21-26)
    cws2 = np.zeros((64, 204), dtype=np.uint8)
    for t in range(64):
        data = rng.integers(0, 256, size=188).tolist()
        cw = np.array(data + tx.rs_encode(data, 16), dtype=np.uint8)
        for p in rng.choice(204, size=t % 11, replace=False):
            cw[p] ^= rng.integers(1, 256)
        cws2[t] = cw
    counts2, fixed2, _ = g.rs_decode(cws2, nroots=16, pad=51)
    o2 = pyref.RefRS(16, 51) if pyref.ref_available() else pyref.PortRS(16, 51)
    for t in range(64):
        c, d, _ = o2.decode(cws2[t])
        assert counts2[t] == c and np.array_equal(fixed2[t], d), t
    g.close()


def test_channel_stream_overlap_is_invisible(gpu_ctx, tx, pyref, monkeypatch):
    """dabgpu_chan_decode runs on its own CUDA stream.  Every API that touches its buffers or results joins that stream first:
    a context that keeps the channel decode on the main stream (DABGPU_CHAN_INLINE=1) and an overlapping one must return the
    same bytes when other calls are squeezed between the decode and the getters."""
    rng = np.random.default_rng(41)
    subs = _subs(tx)
    ens = tx.EnsembleTx(1, subs, seed=17)
    frames = [tx.hard_to_soft(ens.next_frame_bits(), rng, snr_db=6.0)[None, :] for _ in range(8)]
    monkeypatch.setenv("DABGPU_CHAN_INLINE", "1")
    inline = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=2)
    monkeypatch.delenv("DABGPU_CHAN_INLINE")
    overl = gpu_ctx.DabGpu(mode=1, max_streams=1, flags=2)
    sg = tx.eep_segments(48, 2, False)
    n_in = int(tx.puncture_mask(sg).sum())
    port = pyref.PortViterbi()
    for g in (inline, overl):
        g.msc_configure(0, subs)
    for f, frame in enumerate(frames):
        res = []
        for g in (inline, overl):
            g.softbits_push(frame)
            g.chan_decode()
            extra = None
            if f % 2 == 0:      # shares the job / plan / scratch buffers with the decode that may still be running
                soft = np.random.default_rng(100 + f).integers(-127, 128, size=n_in).astype(np.int8)
                outs, perr = g.viterbi_decode([soft] * 3, [sg] * 3)
                exp, _, exp_err = port.decode(soft, sg)
                assert all(np.array_equal(o, exp) for o in outs) and all(int(e) == exp_err for e in perr)
                extra = outs[0]
            if f == 5:
                g.msc_configure(0, subs)   # re-declaring the table while a decode may be in flight (unchanged entries keep their state)
            fibs, ok = g.get_fic(0)
            item = [fibs.copy(), ok.copy()]
            for k, sc in enumerate(subs):
                out, valid = g.get_msc(0, k)
                item += [valid.copy()] + [out[c].copy() for c in range(4) if valid[c]]
                if sc.dabplus:
                    item.append(np.frombuffer(g.get_dabplus_events(0, k), dtype=np.uint8).copy())
            res.append(item)
        assert len(res[0]) == len(res[1])
        for a, b in zip(res[0], res[1]):
            assert np.array_equal(a, b), f
    assert inline.counters() == overl.counters()
    inline.close()
    overl.close()
