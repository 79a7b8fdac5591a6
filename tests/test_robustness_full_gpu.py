"""Robustness sweep, full cross product (BASELINE.json configs[4]): transmission modes I-IV x SNR {0, 5,
10, 15, 20} dB x carrier frequency offset {0, +-0.3, +-2.5, +-20.2} carrier spacings x timing offset {0, 17, T_sym / 2} = 105
recordings per mode, decoded as ONE batch of 105 streams (the receiver's batched mode), each against the oracle run one stream at
a time.  The oracle's behaviour is the ground truth, including where the reference does not lock, loses lock, or (mode III)
never decodes the FIC.

Soft-bit tolerance.  apply_pll (dsp/apply_pll.cpp:82-116) forms the phase t = n * f in float32; at 20 carrier spacings |t|
reaches 1900 cycles, where one ulp is 1.2e-4 cycle: the reference's derotation carries a pseudo-random phase noise of about
2e-4 rad whose pattern depends on every bit of f = coarse + fine.  The GPU reproduces that arithmetic operation by operation,
so while its f equals the reference's bit for bit the soft bits agree to one quantisation step (a handful of +-1 per frame,
from the FFT).  The fine-frequency loop, however, feeds on the cyclic-prefix phase average, which agrees only to about five
digits (summation order, sine approximation, atan2f): every few frames f lands one or two ulp away, the two noise patterns
decorrelate, 1-2 % of the soft bits move by one step and the few carriers that noise has nearly cancelled (|X| a hundredth of
the typical amplitude, about 1e-4 of them at 10 dB) move by more.  The same happens between two builds of the reference with
different FFT libraries.  Hence: max |delta| <= 1 in every frame where fewer than 1e-3 of the bits differ at all (same f, same noise pattern); otherwise
at most 5e-4 of the bits beyond one step, none beyond SOFT_TOL_WEAK.

Checked per recording: number of frames, fine-time offset and frame count of every frame, desync count, soft bits as above,
and after the channel decoder (FIC + one EEP 3-A sub-channel through the 16-CIF time de-interleaver):
  * every FIB group and logical frame decoded by the GPU from the REFERENCE's soft bits equals the reference's bytes, at every
    SNR (the integer half of the chain is bit-exact by construction);
  * from its OWN soft bits the GPU chain gives the reference's bytes wherever the reference's decoder works with a margin: in
    steady state at SNR >= EXACT_FROM_DB.  Not required to be identical, but reported and bounded: logical frames whose 16-CIF
    history still holds the acquisition frame (demodulated before the fine-frequency loop has converged, up to 0.4 carrier
    spacing off: the reference's own decoder is in its error region there), and the low SNRs, where the Viterbi decoder sits on
    near-ties and a soft bit that differs by the allowed one step can move a survivor.
"""
import numpy as np
import pytest

from conftest import VIT_LANES_ALWAYS

pytestmark = pytest.mark.gpu

SOFT_TOL = 1
SOFT_TOL_WEAK = 16        # weak carriers in frames whose frequency word differs in the last ulp (see above)
WEAK_FRACTION = 5e-4
BLOCK = {1: 65536, 2: 65536, 3: 8192, 4: 65536}      # Process() partition is part of the reference semantics
BASE_LEAD = {1: 2000, 2: 2000, 3: 5000, 4: 2000}
N_FRAMES = {1: 7, 2: 22, 3: 22, 4: 12}               # enough CIFs to fill the 16-deep time de-interleaver and decode a few frames
FFT = {1: 2048, 2: 512, 3: 256, 4: 1024}
SNRS = [0.0, 5.0, 10.0, 15.0, 20.0]
CFOS = [0.0, 0.3, -0.3, 2.5, -2.5, 20.2, -20.2]
EXACT_FROM_DB = 10.0


def _cases(mode):
    tsym = FFT[mode] + FFT[mode] * 63 // 256
    return [(snr, cfo, lead) for snr in SNRS for cfo in CFOS for lead in (0, 17, tsym // 2)]


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_full_cross_product_matches_oracle(gpu_ctx, tx, pyref, mode):
    use_ref = pyref.ref_available()
    cases = _cases(mode)
    S = len(cases)
    sub = tx.Subchannel(0, 0, 48, eep_level=2, dabplus=False)
    P = gpu_ctx.get_params(mode)
    block = BLOCK[mode]
    # one transmission per SNR (the payload does not matter for the sweep), impaired per case
    recs = []
    base = {}
    for i, (snr, cfo, lead) in enumerate(cases):
        if snr not in base:
            ens = tx.EnsembleTx(mode, [sub], seed=1000 * mode + int(snr))
            base[snr] = tx.ofdm_modulate([ens.next_frame_bits() for _ in range(N_FRAMES[mode])], mode)
        x = tx.impair(base[snr], snr, cfo / FFT[mode], BASE_LEAD[mode] + lead, seed=17 * i + mode, tail_samples=3000)
        recs.append(tx.to_u8(x, 30.0))
    n = (min(r.size for r in recs) // 2 // block) * block
    iq = np.stack([r[:2 * n] for r in recs])

    # modes II and IV pin the lane-per-trellis Viterbi (k_chan_deinterleave + k_vit_prep + k_viterbi_lanes at one and two CIFs per
    # frame); modes I and III leave the choice to the device (warp-per-trellis at this batch size); g2 below always uses the default
    g = gpu_ctx.DabGpu(mode=mode, max_streams=S, flags=VIT_LANES_ALWAYS if mode in (2, 4) else 0)
    for s in range(S):
        g.msc_configure(s, [sub])
    got = [[] for _ in range(S)]          # per stream: (soft bits, fine time, fic (fibs, ok), msc (out, valid))
    for off in range(0, n, block):
        g.ofdm_process(iq[:, 2 * off:2 * (off + block)], block_size=block)
        frames = [g.ofdm_pop_frames(s) for s in range(S)]
        todo = max(len(f) for f in frames)
        taken = [0] * S
        for _ in range(todo):             # the channel decoder takes one frame per stream and call, oldest first
            g.chan_decode()
            for s in range(S):
                decoded, _ = g.chan_status(s)
                if decoded and taken[s] < len(frames[s]):
                    f = frames[s][taken[s]]
                    taken[s] += 1
                    got[s].append((f[0], f[3], g.get_fic(s), g.get_msc(s, 0), (f[1], f[2])))
        assert all(taken[s] == len(frames[s]) for s in range(S))

    locked = 0
    report = []
    stats = {"frames": 0, "same_f": 0, "max": 0, "weak": 0.0}
    g2 = gpu_ctx.DabGpu(mode=mode, max_streams=1)     # the integer half alone: reference soft bits in
    for s, (snr, cfo, lead) in enumerate(cases):
        o = pyref.RefOfdm(mode, 1) if use_ref else pyref.PortOfdm(mode)
        for off in range(0, n, block):
            o.process_u8(iq[s, 2 * off:2 * (off + block)])
        exp = o.pop_frames()
        st, est = g.ofdm_status(s), o.state()
        tag = f"mode {mode} snr {snr} cfo {cfo} lead {lead}"
        assert len(got[s]) == len(exp), (tag, len(got[s]), len(exp), st, est)
        assert st["total_frames_desync"] == est["frames_desync"], (tag, st, est)
        locked += 1 if exp else 0
        if not exp:
            continue
        o_fic = (pyref.RefFic() if use_ref else pyref.PortFic()) if mode != 3 else None
        o_msc = (pyref.RefMsc if use_ref else pyref.PortMsc)(sub.start_address, sub.length, sub.is_uep, sub.uep_index, sub.eep_level, sub.eep_type_b)
        g2.msc_configure(0, [])               # an unchanged entry would keep its de-interleaver history: drop it, then add it again
        g2.msc_configure(0, [sub])
        n_bytes = n_same = n_steady = n_steady_same = 0
        for i, ((soft, ft, (fibs, ok), (out, valid), finfo), e) in enumerate(zip(got[s], exp)):
            assert ft == e[3], f"{tag} frame {i}: fine time offset {ft} != {e[3]}"
            d = np.abs(soft.astype(np.int32) - e[0].astype(np.int32))
            same_noise = (d >= 1).mean() < 1e-3     # the two derotations carry the same phase-noise pattern (identical f, see above)
            stats["frames"] += 1
            stats["same_f"] += int(same_noise)
            stats["max"] = max(stats["max"], int(d.max()))
            if same_noise:
                assert d.max() <= SOFT_TOL, f"{tag} frame {i}: max soft-bit delta {d.max()} in a frame with {int((d >= 1).sum())} differing bits"
            else:
                stats["weak"] = max(stats["weak"], float((d > SOFT_TOL).mean()))
                assert d.max() <= SOFT_TOL_WEAK and (d > SOFT_TOL).mean() <= WEAK_FRACTION, \
                    f"{tag} frame {i}: max soft-bit delta {d.max()}, {(d > SOFT_TOL).sum()} bits beyond one step"
            # reference chain on the reference's soft bits
            e_fibs = []
            if o_fic is not None:
                for c in range(P.nb_cifs):
                    e_fibs.append(o_fic.decode_group(e[0][c * P.nb_fib_group_bits:(c + 1) * P.nb_fib_group_bits], c))
            e_msc = [o_msc.decode_cif(e[0][P.nb_fic_bits + c * P.nb_cif_bits:P.nb_fic_bits + (c + 1) * P.nb_cif_bits]) for c in range(P.nb_cifs)]
            # (a) GPU integer half on the same soft bits: always identical
            g2.softbits_push(e[0][None, :])
            g2.chan_decode()
            fibs2, ok2 = g2.get_fic(0)
            out2, valid2 = g2.get_msc(0, 0)
            nf = P.nb_fibs_per_cif
            for c in range(P.nb_cifs):
                if o_fic is not None:
                    assert [fibs2[c * nf + k, :30].tobytes() for k in range(nf) if ok2[c * nf + k]] == e_fibs[c], f"{tag} frame {i}: FIBs differ on equal soft bits"
                assert bool(valid2[c]) == (e_msc[c].size > 0)
                if e_msc[c].size:
                    assert np.array_equal(out2[c], e_msc[c]), f"{tag} frame {i} cif {c}: MSC bytes differ on equal soft bits"
            # (b) the whole GPU chain from the IQ
            for c in range(P.nb_cifs):
                assert bool(valid[c]) == (e_msc[c].size > 0)
                steady = (i * P.nb_cifs + c) >= 16 + P.nb_cifs       # the 16-CIF history no longer holds the acquisition frame
                if e_msc[c].size:
                    n_bytes += e_msc[c].size
                    n_same += int((out[c] == e_msc[c]).sum())
                    if steady:
                        n_steady += 1
                        n_steady_same += int(np.array_equal(out[c], e_msc[c]))
                    if snr >= EXACT_FROM_DB and steady:
                        assert np.array_equal(out[c], e_msc[c]), f"{tag} frame {i} cif {c}: decoded MSC bytes differ"
                if o_fic is not None and snr >= EXACT_FROM_DB and i > 0:
                    assert [fibs[c * nf + k, :30].tobytes() for k in range(nf) if ok[c * nf + k]] == e_fibs[c], f"{tag} frame {i}: FIBs differ"
        if n_bytes:
            report.append((snr, n_same / n_bytes, n_steady, n_steady_same))
    g.close()
    g2.close()
    print(f"mode {mode}: {locked}/{S} recordings locked in the reference; {stats['same_f']}/{stats['frames']} frames with the same phase-noise pattern (< 1e-3 of the bits differ)"
          f", largest soft-bit delta {stats['max']}, largest fraction beyond one step {stats['weak']:.2e}")
    for snr in SNRS:
        rows = [r for r in report if r[0] == snr]
        if rows:
            agree = [r[1] for r in rows]
            print(f"    {snr:4.0f} dB: {len(rows):3d} recordings decoded, byte agreement min {min(agree):.6f} mean {np.mean(agree):.6f}; "
                  f"steady-state logical frames identical: {sum(r[3] for r in rows)}/{sum(r[2] for r in rows)}")
            assert np.mean(agree) > (0.999 if snr >= 10.0 else 0.9), f"mode {mode} {snr} dB: byte agreement {np.mean(agree)}"
    assert locked >= (S // 3 if mode != 3 else 0)      # mode III: the reference rejects the fine-time peak for most offsets
