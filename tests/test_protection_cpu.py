"""Every protection profile of the reference (all 64 UEP rows, EEP 1..4 A/B at three sizes, the 2-A n=1 row and the 8-CU type-A
quirk) through the C restatement of MSC_Decoder, against tests/golden/protection_kat.npz = outputs of the reference build
(tests/golden/make_protection_golden.py).  The GPU twin is tests/test_protection_gpu.py."""
import os

import numpy as np

import protection_cases as pc
from conftest import GOLDEN


def load_golden():
    d = np.load(os.path.join(GOLDEN, "protection_kat.npz"))
    subs, offs, exp = d["subs"], d["offsets"], d["expected"]
    return d["uep_table"], subs, [exp[offs[i]:offs[i + 1]] for i in range(subs.shape[0])]


def test_case_list_matches_the_golden_file():
    uep, subs, exp = load_golden()
    layouts = pc.build_layouts(uep[:, 0])
    flat = [(li, s["start"], s["length"], int(s["is_uep"]), s["uep_index"], s["eep_level"], int(s["eep_type_b"])) for li, lay in enumerate(layouts) for s in lay]
    assert [tuple(r[:7]) for r in subs.tolist()] == flat
    assert sorted(r[4] for r in subs.tolist() if r[3]) == list(range(64)), "every UEP row"
    assert {(r[5], r[6]) for r in subs.tolist() if not r[3]} == {(l, b) for l in range(4) for b in (0, 1)}, "every EEP level and type"
    assert all(e.size == 5 * r[7] for e, r in zip(exp, subs.tolist())), "five decoded CIFs per sub-channel"


def test_c_port_decodes_every_protection_profile_like_the_reference(pyref):
    uep, subs, exp = load_golden()
    n_layouts = int(subs[:, 0].max()) + 1
    for li in range(n_layouts):
        rows = [i for i in range(subs.shape[0]) if subs[i, 0] == li]
        decs = [pyref.PortMsc(int(subs[i, 1]), int(subs[i, 2]), bool(subs[i, 3]), int(subs[i, 4]), int(subs[i, 5]), bool(subs[i, 6])) for i in rows]
        got = [[] for _ in rows]
        for c in range(pc.N_CIFS):
            cif = pc.soft_cif(li, c)
            for k, d in enumerate(decs):
                b = d.decode_cif(cif)
                assert (b.size > 0) == (c >= 15)
                if b.size:
                    got[k].append(b)
        for k, i in enumerate(rows):
            assert np.array_equal(np.concatenate(got[k]), exp[i]), f"sub-channel row {subs[i].tolist()}"
