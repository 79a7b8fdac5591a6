"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libdabref.so, built in place from
/root/reference by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py
The fixtures pin the C restatement (CPU tests) and the CUDA path (GPU tests) on boxes without /root/reference.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref  # noqa: E402

tx = importlib.import_module("sdrplusplus-dab-radio-plugin_b200.synth.dabtx")
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert pyref.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    info = pyref.RefLib.get().build_info()
    rng = np.random.default_rng(2024)

    # ---- Viterbi known-answer vectors (noisy, garbage, tie-heavy, saturating) ---------------------
    rv = pyref.RefViterbi()
    cases = {}
    seg_choices = {"fic": tx.FIC_SEGMENTS, "eep3a48": tx.eep_segments(48, 2, False), "eep2a8": tx.eep_segments(8, 1, False),
                   "uep4": tx.uep_segments(4), "eep3b54": tx.eep_segments(54, 2, True)}
    for name, sg in seg_choices.items():
        n_in = int(tx.puncture_mask(sg).sum())
        info_bytes = rng.integers(0, 256, size=(sum(b for _, b in sg) // 4 - 6) // 8, dtype=np.uint8)
        enc = tx.channel_encode(info_bytes, sg)
        for kind in range(4):
            if kind == 0:
                soft = tx.hard_to_soft(enc, rng, snr_db=1.0)
            elif kind == 1:
                soft = rng.integers(-128, 128, size=n_in).astype(np.int8)
            elif kind == 2:
                soft = (rng.integers(-1, 2, size=n_in) * 127).astype(np.int8)
            else:
                soft = np.where(rng.random(n_in) < 0.5, -128, 127).astype(np.int8)
            out, consumed, err = rv.decode(soft, sg)
            key = f"{name}_{kind}"
            cases[key + "_soft"] = soft
            cases[key + "_segs"] = np.array(sg, dtype=np.int32)
            cases[key + "_out"] = out
            cases[key + "_err"] = np.array([err, consumed], dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, "viterbi_kat.npz"), build_info=np.array(info), **cases)

    # ---- energy dispersal, tables ----------------------------------------------------------------
    tables = {"prbs_bytes": pyref.ref_scrambler_bytes(512)}
    for mode in (1, 2, 3, 4):
        p, prs, cmap, dp = pyref.ref_tables(mode)
        tables[f"ofdm_params_{mode}"] = p
        tables[f"prs_{mode}"] = prs
        tables[f"cmap_{mode}"] = cmap
        tables[f"dab_params_{mode}"] = dp
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **tables)

    # ---- FIC + MSC + DAB+ over 9 frames of soft bits (noisy), small ensemble ----------------------
    subs = [tx.Subchannel(0, 0, 24, eep_level=2), tx.Subchannel(1, 24, 8, eep_level=1, dabplus=False),
            tx.Subchannel(2, 32, 16, is_uep=True, uep_index=0, dabplus=False), tx.Subchannel(3, 48, 27, eep_level=0, eep_type_b=True, dabplus=False)]
    ens = tx.EnsembleTx(1, subs, seed=77, fill_random=False)
    n_frames = 9
    fic = pyref.RefFic()
    msc = [pyref.RefMsc(s.start_address, s.length, s.is_uep, s.uep_index, s.eep_level, s.eep_type_b) for s in subs]
    aac = pyref.RefAac()
    used_bits = max(s.start_address + s.length for s in subs) * 64
    d = {"subs": np.array([[s.start_address, s.length, int(s.is_uep), s.uep_index, s.eep_level, int(s.eep_type_b), int(s.dabplus)] for s in subs], dtype=np.int32),
         "used_bits": np.array([used_bits])}
    fic_soft, msc_soft, fibs_all, msc_out, aac_log = [], [], [], {k: [] for k in range(len(subs))}, []
    for f in range(n_frames):
        soft = tx.hard_to_soft(ens.next_frame_bits(), rng, snr_db=5.0)
        fic_soft.append(soft[:9216])
        m = soft[9216:].reshape(4, 55296)
        msc_soft.append(m[:, :used_bits].copy())    # the rest of the CIF is unused capacity (zeros)
        for c in range(4):
            fibs = fic.decode_group(soft[c * 2304:(c + 1) * 2304], c)
            fibs_all.append(np.frombuffer(b"".join(fibs), dtype=np.uint8) if fibs else np.zeros(0, np.uint8))
            cif = np.zeros(55296, dtype=np.int8)
            cif[:used_bits] = m[c, :used_bits]
            for k in range(len(subs)):
                o = msc[k].decode_cif(cif)
                msc_out[k].append(o)
                if k == 0 and o.size:
                    for ev in aac.process(o):
                        aac_log.append(np.frombuffer(np.array(ev[:5] + (len(ev[5]),), dtype=np.int32).tobytes() + ev[5], dtype=np.uint8))
    d["fic_soft"] = np.stack(fic_soft)
    d["msc_soft"] = np.stack(msc_soft)
    d["fib_counts"] = np.array([x.size // 30 for x in fibs_all], dtype=np.int32)
    d["fibs"] = np.concatenate(fibs_all) if fibs_all else np.zeros(0, np.uint8)
    for k in range(len(subs)):
        d[f"msc_out_{k}_sizes"] = np.array([o.size for o in msc_out[k]], dtype=np.int32)
        d[f"msc_out_{k}"] = np.concatenate(msc_out[k]) if msc_out[k] else np.zeros(0, np.uint8)
    d["aac_event_sizes"] = np.array([x.size for x in aac_log], dtype=np.int32)
    d["aac_events"] = np.concatenate(aac_log) if aac_log else np.zeros(0, np.uint8)
    np.savez_compressed(os.path.join(OUT, "channel_kat.npz"), **d)

    # ---- RS(120,110) ---------------------------------------------------------------------------
    rr = pyref.RefRS()
    cws, outs, counts = [], [], []
    for t in range(64):
        data = rng.integers(0, 256, size=110).tolist()
        cw = np.array(data + tx.rs_encode(data, 10), dtype=np.uint8)
        for p in rng.choice(120, size=t % 8, replace=False):
            cw[p] ^= rng.integers(1, 256)
        c, dd, pos = rr.decode(cw)
        cws.append(cw)
        outs.append(dd)
        counts.append(c)
    np.savez_compressed(os.path.join(OUT, "rs_kat.npz"), cw=np.stack(cws), out=np.stack(outs), counts=np.array(counts, dtype=np.int32))

    # ---- OFDM: Mode II, 6 frames, serialised reference driver -----------------------------------
    ens2 = tx.EnsembleTx(2, [tx.Subchannel(0, 0, 48, eep_level=2)], seed=5)
    frames = [ens2.next_frame_bits() for _ in range(6)]
    iq = tx.ofdm_modulate(frames, 2)
    u8 = tx.to_u8(tx.impair(iq, 14.0, 2.9e-3, 321, seed=11, tail_samples=2000), 30.0)
    o = pyref.RefOfdm(2, 1)
    block = 65536
    for off in range(0, u8.size // 2, block):
        o.process_u8(u8[2 * off:2 * (off + block)])
    fr = o.pop_frames()
    np.savez_compressed(os.path.join(OUT, "ofdm_mode2_kat.npz"), iq_u8=u8, block=np.array([block]),
                        soft=np.stack([f[0] for f in fr]), coarse=np.array([f[1] for f in fr], dtype=np.float32),
                        fine=np.array([f[2] for f in fr], dtype=np.float32), toff=np.array([f[3] for f in fr], dtype=np.int32))
    print("golden fixtures written:", sorted(x for x in os.listdir(OUT) if x.endswith(".npz")))
    for x in os.listdir(OUT):
        if x.endswith(".npz"):
            print(f"  {x}: {os.path.getsize(os.path.join(OUT, x)) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
