"""CPU tests (no GPU): the C restatement against the reference build, against the committed golden vectors that the
reference produced, and against the executable self-checks the reference ships (noiseless round trips of
VIT/examples/run_tests.cpp and run_punctured_decoder.cpp)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


# ---------------------------------------------------------------------------------------------------
# golden vectors produced by the reference build
# ---------------------------------------------------------------------------------------------------
def test_port_viterbi_matches_golden(pyref):
    z = _load("viterbi_kat.npz")
    assert "AVX2" in str(z["build_info"])
    pv = pyref.PortViterbi()
    keys = sorted(k[:-5] for k in z.files if k.endswith("_soft"))
    assert len(keys) == 20
    for k in keys:
        segs = [tuple(int(v) for v in r) for r in z[k + "_segs"]]
        out, consumed, err = pv.decode(z[k + "_soft"], segs)
        assert np.array_equal(out, z[k + "_out"]), k
        assert err == int(z[k + "_err"][0]) and consumed == int(z[k + "_err"][1]), k


def test_port_tables_match_golden(pyref):
    z = _load("tables.npz")
    L = pyref.PortLib.get().L
    assert np.array_equal(pyref.port_scrambler_bytes(512), z["prbs_bytes"])
    for mode in (1, 2, 3, 4):
        p = np.zeros(6, dtype=np.int32)
        assert L.dabo_ofdm_params(mode, p) == 0
        assert np.array_equal(p, z[f"ofdm_params_{mode}"])
        cmap = np.zeros(int(p[5]), dtype=np.int32)
        L.dabo_carrier_map(int(p[4]), int(p[5]), cmap)
        assert np.array_equal(cmap, z[f"cmap_{mode}"])
        prs = np.zeros(2 * int(p[4]), dtype=np.float32)
        L.dabo_prs_fft(mode, prs, int(p[4]))
        assert np.abs(prs.view(np.complex64) - z[f"prs_{mode}"]).max() < 1e-6
    assert L.dabo_ofdm_params(5, np.zeros(6, dtype=np.int32)) != 0     # invalid mode is an error, as in the reference


def test_port_channel_chain_matches_golden(pyref):
    z = _load("channel_kat.npz")
    subs = z["subs"]
    used = int(z["used_bits"][0])
    fic = pyref.PortFic()
    msc = [pyref.PortMsc(int(s[0]), int(s[1]), bool(s[2]), int(s[3]), int(s[4]), bool(s[5])) for s in subs]
    aac = pyref.PortAac()
    fib_counts, fibs = z["fib_counts"], z["fibs"]
    fpos, gi = 0, 0
    out_pos = [0] * len(subs)
    out_idx = 0
    ev_sizes, ev_blob = z["aac_event_sizes"], z["aac_events"]
    ev_i, ev_pos = 0, 0
    for f in range(z["fic_soft"].shape[0]):
        for c in range(4):
            got = fic.decode_group(z["fic_soft"][f, c * 2304:(c + 1) * 2304], c)
            n = int(fib_counts[gi]); gi += 1
            exp = [fibs[fpos + 30 * i: fpos + 30 * (i + 1)].tobytes() for i in range(n)]
            fpos += 30 * n
            assert got == exp, (f, c)
            cif = np.zeros(55296, dtype=np.int8)
            cif[:used] = z["msc_soft"][f, c]
            for k in range(len(subs)):
                o = msc[k].decode_cif(cif)
                sz = int(z[f"msc_out_{k}_sizes"][out_idx])
                assert o.size == sz, (f, c, k)
                assert np.array_equal(o, z[f"msc_out_{k}"][out_pos[k]:out_pos[k] + sz]), (f, c, k)
                out_pos[k] += sz
                if k == 0 and o.size:
                    for ev in aac.process(o):
                        blob = np.frombuffer(np.array(ev[:5] + (len(ev[5]),), dtype=np.int32).tobytes() + ev[5], dtype=np.uint8)
                        n_ev = int(ev_sizes[ev_i]); ev_i += 1
                        assert np.array_equal(blob, ev_blob[ev_pos:ev_pos + n_ev])
                        ev_pos += n_ev
            out_idx += 1
    assert ev_i == len(ev_sizes) and ev_i > 0


def test_port_rs_matches_golden(pyref):
    z = _load("rs_kat.npz")
    pr = pyref.PortRS()
    for cw, out, cnt in zip(z["cw"], z["out"], z["counts"]):
        c, d, _ = pr.decode(cw)
        assert c == int(cnt) and np.array_equal(d, out)


def test_port_ofdm_matches_golden(pyref):
    z = _load("ofdm_mode2_kat.npz")
    o = pyref.PortOfdm(2)
    u8, block = z["iq_u8"], int(z["block"][0])
    for off in range(0, u8.size // 2, block):
        o.process_u8(u8[2 * off:2 * (off + block)])
    fr = o.pop_frames()
    assert len(fr) == z["soft"].shape[0]
    for i, f in enumerate(fr):
        assert f[3] == int(z["toff"][i])
        assert np.abs(f[0].astype(np.int32) - z["soft"][i].astype(np.int32)).max() <= 1    # tolerance: 1 soft-bit LSB
        assert abs(f[1] - float(z["coarse"][i])) < 2e-6 and abs(f[2] - float(z["fine"][i])) < 2e-6


# ---------------------------------------------------------------------------------------------------
# the reference's own executable checks, restated: noiseless encode -> (puncture) -> decode round trips
# (VIT/examples/run_tests.cpp:118-143 code_4 = DAB; run_punctured_decoder.cpp:72-76,193-246 = FIC schedule)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("segs_name", ["unpunctured", "fic", "eep1a", "eep4a", "eep1b", "eep4b", "uep63"])
def test_noiseless_round_trip(pyref, tx, segs_name):
    rng = np.random.default_rng(1)   # the reference uses std::rand() unseeded (seed 1)
    segs = {"unpunctured": [(24, 128 * 16), (0, 24)], "fic": tx.FIC_SEGMENTS, "eep1a": tx.eep_segments(96, 0, False),
            "eep4a": tx.eep_segments(40, 3, False), "eep1b": tx.eep_segments(27, 0, True), "eep4b": tx.eep_segments(30, 3, True),
            "uep63": tx.uep_segments(63)}[segs_name]
    n_bytes = (sum(b for _, b in segs) // 4 - 6) // 8
    pv = pyref.PortViterbi()
    for _ in range(3):
        data = rng.integers(0, 256, size=n_bytes, dtype=np.uint8)
        mother = tx.conv_encode(tx.bytes_to_bits(data))
        soft = tx.hard_to_soft(mother[tx.puncture_mask(segs)])
        out, consumed, err = pv.decode(soft, segs)
        assert np.array_equal(out, data)
        assert consumed == soft.size
        # noiseless: every kept symbol matches (error 0) and every punctured one costs 127 on all paths
        assert err == 127 * int((~tx.puncture_mask(segs)).sum())


def test_puncture_tables_are_the_standard_ones(pyref):
    """EN 300 401 table 13: PI_i keeps 8+i bits of 32; spot rows as printed in the standard (also puncture_codes.h:12-37)."""
    L = pyref.PortLib.get().L
    rows = {1: "11001000100010001000100010001000", 8: "11001100110011001100110011001100", 9: "11101100110011001100110011001100",
            16: "11101110111011101110111011101110", 21: "11111111111111101111111011111110", 24: "1" * 32}
    for pi in range(1, 25):
        cnt = np.zeros(8, dtype=np.uint8)
        L.dabo_pi_counts(pi, cnt)
        assert int(cnt.sum()) == 8 + pi
        if pi in rows:
            bits = "".join("1" * int(c) + "0" * (4 - int(c)) for c in cnt)
            assert bits == rows[pi], pi


# ---------------------------------------------------------------------------------------------------
# restatement vs the reference build itself (only where oracle/_ref exists)
# ---------------------------------------------------------------------------------------------------
def test_port_vs_reference_viterbi_random(pyref, ref_ok, tx):
    rng = np.random.default_rng(99)
    rv, pv = pyref.RefViterbi(), pyref.PortViterbi()
    for trial in range(120):
        sg = [tx.FIC_SEGMENTS, tx.eep_segments(48, 2, False), tx.uep_segments(37), tx.eep_segments(42, 1, True)][trial % 4]
        n_in = int(tx.puncture_mask(sg).sum())
        kind = (trial // 4) % 3
        soft = [rng.integers(-128, 128, size=n_in), rng.integers(-1, 2, size=n_in) * 127, np.where(rng.random(n_in) < 0.5, -128, 127)][kind].astype(np.int8)
        a, b = rv.decode(soft, sg), pv.decode(soft, sg)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]


def test_port_vs_reference_ofdm_modes(pyref, ref_ok, tx):
    for mode, block in ((1, 65536), (4, 32768), (3, 4096)):
        ens = tx.EnsembleTx(mode, tx.default_ensemble(), seed=mode)
        frames = [ens.next_frame_bits() for _ in range(4 if mode == 1 else 8)]
        u8 = tx.to_u8(tx.impair(tx.ofdm_modulate(frames, mode), 16.0, -1.3e-3, 999, seed=mode, tail_samples=2500), 30.0)
        r, p = pyref.RefOfdm(mode, 1), pyref.PortOfdm(mode)
        for off in range(0, u8.size // 2, block):
            r.process_u8(u8[2 * off:2 * (off + block)])
            p.process_u8(u8[2 * off:2 * (off + block)])
        fr, fp = r.pop_frames(), p.pop_frames()
        assert len(fr) == len(fp)
        sr, sp = r.state(), p.state()
        assert (sr["state"], sr["frames_read"], sr["frames_desync"]) == (sp["state"], sp["frames_read"], sp["frames_desync"])
        for a, b in zip(fr, fp):
            assert a[3] == b[3]
            assert np.abs(a[0].astype(np.int32) - b[0].astype(np.int32)).max() <= 1


def test_reference_rejects_invalid_mode(pyref, ref_ok):
    with pytest.raises(ValueError):
        pyref.RefOfdm(7, 1)
