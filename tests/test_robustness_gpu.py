"""Robustness sweep (BASELINE.json configs[4]): transmission modes I-IV x SNR x carrier
frequency offset (in carrier spacings, incl. > 1 spacing so the coarse search has to move) x timing offset.  The
oracle's behaviour is the ground truth -- including where the reference itself does not lock or (mode III) never
decodes the FIC: same number of frames, same fine-time decisions, soft bits within one quantisation step, identical
FIBs after Viterbi.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SOFT_TOL = 1
BLOCK = {1: 65536, 2: 65536, 3: 8192, 4: 65536}      # Process() partition is part of the reference semantics
BASE_LEAD = {1: 2000, 2: 2000, 3: 5000, 4: 2000}     # leading noise; mode III only locks for some (lead, block) pairs in the reference
N_FRAMES = {1: 4, 2: 10, 3: 10, 4: 6}
FFT = {1: 2048, 2: 512, 3: 256, 4: 1024}

# (mode, snr_db, cfo in carrier spacings, leading samples): a covering subset of the full cross product
CASES = []
_cfos = [0.0, 0.3, -0.3, 2.5, -2.5, 20.2, -20.2]
_snrs = [20.0, 15.0, 10.0, 7.0, 0.0]
for _m in (1, 2, 3, 4):
    _tsym = FFT[_m] + FFT[_m] * 63 // 256
    _leads = [0, 17, _tsym // 2]
    for _i, _cfo in enumerate(_cfos):
        CASES.append((_m, _snrs[(_i + _m) % len(_snrs)], _cfo, _leads[(_i + _m) % 3]))
# mode III cases the reference does lock on (it rejects the fine-time peak for most offsets at this frame size)
CASES += [(3, 20.0, 0.256, 0), (3, 20.0, 0.256, 17), (3, 20.0, -0.3, 0)]


@pytest.mark.parametrize("mode,snr,cfo_carriers,lead", CASES)
def test_sweep_matches_oracle(gpu_ctx, tx, pyref, mode, snr, cfo_carriers, lead):
    use_ref = pyref.ref_available()
    subs = [tx.Subchannel(0, 0, 48, eep_level=2, dabplus=False)]
    ens = tx.EnsembleTx(mode, subs, seed=100 * mode + int(abs(cfo_carriers) * 10))
    frames = [ens.next_frame_bits() for _ in range(N_FRAMES[mode])]
    iq = tx.ofdm_modulate(frames, mode)
    x = tx.impair(iq, snr, cfo_carriers / FFT[mode], BASE_LEAD[mode] + lead, seed=int(snr) + 7, tail_samples=3000)
    u8 = tx.to_u8(x, 30.0)
    block = BLOCK[mode]
    n = (u8.size // 2 // block) * block
    o = pyref.RefOfdm(mode, 1) if use_ref else pyref.PortOfdm(mode)
    g = gpu_ctx.DabGpu(mode=mode, max_streams=1)
    got = []
    for off in range(0, n, block):
        o.process_u8(u8[2 * off:2 * (off + block)])
        g.ofdm_process(u8[None, 2 * off:2 * (off + block)], block_size=block)
        got += g.ofdm_pop_frames(0)
    exp = o.pop_frames()
    st = g.ofdm_status(0)
    est = o.state()
    assert len(got) == len(exp), (len(got), len(exp), st, est)
    assert st["total_frames_desync"] == est["frames_desync"], (st, est)
    o_fic = (pyref.RefFic() if use_ref else pyref.PortFic()) if mode != 3 else None
    P = gpu_ctx.get_params(mode)
    for i, (a, b) in enumerate(zip(got, exp)):
        assert a[3] == b[3], f"frame {i}: fine time offset {a[3]} != {b[3]}"
        d = np.abs(a[0].astype(np.int32) - b[0].astype(np.int32))
        assert d.max() <= SOFT_TOL, f"frame {i}: max soft-bit delta {d.max()}"
        if o_fic is not None:
            # post-Viterbi identity: the FIBs decoded on the GPU from ITS soft bits equal the reference's from the reference's
            groups = a[0][:P.nb_fic_bits].reshape(P.nb_cifs, 2304)
            fibs, ok = g.fic_decode(groups)
            for c in range(P.nb_cifs):
                e = o_fic.decode_group(b[0][c * 2304:(c + 1) * 2304], c)
                assert [fibs[c, k, :30].tobytes() for k in range(3) if ok[c, k]] == e
    g.close()


def test_packet_mode_reed_solomon_204_188(gpu_ctx, tx, pyref):
    """the packet-mode outer code RS(204,188) (16 roots, pad 51; msc_reed_solomon_data_packet_processor.cpp:21-52)
    runs on the same kernel as the DAB+ code."""
    rng = np.random.default_rng(3)
    g = gpu_ctx.DabGpu(mode=1, max_streams=1)
    cws = []
    for n_err in (0, 1, 4, 8, 9, 12):
        data = rng.integers(0, 256, size=188, dtype=np.uint8)
        cw = np.array(list(data) + tx.rs_encode(list(data), 16), dtype=np.uint8)
        pos = rng.choice(204, size=n_err, replace=False)
        cw[pos] ^= rng.integers(1, 256, size=n_err).astype(np.uint8)
        cws.append(cw)
    counts, fixed, pos = g.rs_decode(np.stack(cws), nroots=16, pad=51)
    o = pyref.RefRS(16, 51) if pyref.ref_available() else pyref.PortRS(16, 51)
    for i, cw in enumerate(cws):
        e_cnt, e_data, e_pos = o.decode(cw)
        assert counts[i] == e_cnt
        assert np.array_equal(fixed[i], e_data)
        if e_cnt > 0:
            assert np.array_equal(pos[i, :e_cnt], np.asarray(e_pos)[:e_cnt])
    assert list(counts[:4]) == [0, 1, 4, 8] and counts[5] == -1
    g.close()
