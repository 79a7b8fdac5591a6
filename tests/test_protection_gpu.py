"""GPU parity over EVERY protection profile the reference knows (BASELINE configs[2]: "all EEP/UEP subchannels"): all 64 rows of
UEP_PROTECTION_TABLE (incl. the padded rows 4, 7, 37, 63 ...), EEP {1,2,3,4}-{A,B} at three sizes each, the EEP 2-A n=1 special
row and the 8-CU type-A quirk (subchannel_protection_tables.h:21-170, msc_decoder.cpp:77-154), through dabgpu_msc_configure /
dabgpu_softbits_push / dabgpu_chan_decode with each of the three Viterbi mappings.  Expected bytes: tests/golden/protection_kat.npz,
produced by the reference build's MSC_Decoder; where that build is present the live decoder is asked as well.  Bit-exact."""
import numpy as np
import pytest

import protection_cases as pc
from test_protection_cpu import load_golden

pytestmark = pytest.mark.gpu


class _Sub:
    def __init__(self, row):
        _, self.start_address, self.length, u, self.uep_index, self.eep_level, b, self.n_bytes = [int(v) for v in row]
        self.is_uep, self.eep_type_b, self.dabplus = bool(u), bool(b), False


def test_every_protection_profile_decodes_like_the_reference(gpu_ctx, pyref, vit_flags):
    uep, subs, exp = load_golden()
    n_layouts = int(subs[:, 0].max()) + 1
    rows = [[i for i in range(subs.shape[0]) if subs[i, 0] == li] for li in range(n_layouts)]
    g = gpu_ctx.DabGpu(mode=1, max_streams=n_layouts, flags=vit_flags)
    for li in range(n_layouts):
        g.msc_configure(li, [_Sub(subs[i]) for i in rows[li]])
    live = None
    if pyref.ref_available():
        live = [[pyref.RefMsc(s.start_address, s.length, s.is_uep, s.uep_index, s.eep_level, s.eep_type_b) for s in (_Sub(subs[i]) for i in rows[li])]
                for li in range(n_layouts)]
    got = [[[] for _ in rows[li]] for li in range(n_layouts)]
    P = g.P
    for f in range(pc.N_CIFS // 4):
        frames = np.zeros((n_layouts, P.nb_frame_bits), dtype=np.int8)
        for li in range(n_layouts):
            for c in range(4):
                frames[li, P.nb_fic_bits + c * 55296: P.nb_fic_bits + (c + 1) * 55296] = pc.soft_cif(li, 4 * f + c)
        g.softbits_push(frames)
        g.chan_decode()
        for li in range(n_layouts):
            for k, i in enumerate(rows[li]):
                out, valid = g.get_msc(li, k)
                assert out.shape[1] == int(subs[i, 7]), subs[i].tolist()
                for c in range(4):
                    assert bool(valid[c]) == (4 * f + c >= 15)
                    if valid[c]:
                        got[li][k].append(out[c].copy())
                    if live is not None:
                        e = live[li][k].decode_cif(frames[li, P.nb_fic_bits + c * 55296: P.nb_fic_bits + (c + 1) * 55296])
                        assert (e.size > 0) == bool(valid[c])
                        if e.size:
                            assert np.array_equal(out[c], e), ("live reference", subs[i].tolist(), f, c)
    n = 0
    for li in range(n_layouts):
        for k, i in enumerate(rows[li]):
            assert np.array_equal(np.concatenate(got[li][k]), exp[i]), f"sub-channel row {subs[i].tolist()}"
            n += 1
    assert n == 89
    g.close()
