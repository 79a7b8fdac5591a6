"""Shared by tests/golden/make_protection_golden.py and the protection-profile parity tests: the list of sub-channels that covers
every protection profile of the reference, their packing into CIF layouts, and the seeded soft bits of every CIF."""
import numpy as np

N_CIFS = 20          # CIFs 0..14 fill the time de-interleaver, 15..19 decode
EEP_A_MULT = (12, 8, 6, 4)
EEP_B_MULT = (27, 21, 18, 15)


def profile_list(uep_sizes):
    """uep_sizes: 64 sub-channel sizes in CU as the reference's UEP_PROTECTION_TABLE lists them (column 0)."""
    subs = [dict(is_uep=True, uep_index=i, eep_level=0, eep_type_b=False, length=int(uep_sizes[i])) for i in range(64)]
    for level in range(4):
        for n in (1, 3, 7):     # level 1 (2-A) with n = 1 is the special row; 4-A with n = 2 is the 8-CU quirk
            subs.append(dict(is_uep=False, uep_index=0, eep_level=level, eep_type_b=False, length=EEP_A_MULT[level] * n))
        for n in (1, 2, 5):
            subs.append(dict(is_uep=False, uep_index=0, eep_level=level, eep_type_b=True, length=EEP_B_MULT[level] * n))
    subs.append(dict(is_uep=False, uep_index=0, eep_level=3, eep_type_b=False, length=8))   # type-A, 8 CU, level 4-A: decoded with the 2-A special row
    return subs


def build_layouts(uep_sizes):
    """Greedy packing into CIFs of 864 CU; returns a list of layouts, each a list of sub-channel dicts with `start`."""
    layouts, cur, used = [], [], 0
    for s in profile_list(uep_sizes):
        if used + s["length"] > 864:
            layouts.append(cur)
            cur, used = [], 0
        cur.append(dict(s, start=used))
        used += s["length"]
    if cur:
        layouts.append(cur)
    return layouts


def soft_cif(layout: int, cif: int) -> np.ndarray:
    """55296 soft bits of CIF `cif` of layout `layout`.  Even layouts: garbage over the whole int8 range (ties, -128, saturating
    metrics); odd layouts: noisy hard decisions like a demodulator at low SNR produces."""
    rng = np.random.default_rng(977 * layout + cif + 5)
    if layout % 2 == 0:
        return rng.integers(-128, 128, size=55296, dtype=np.int8)
    hard = rng.integers(0, 2, size=55296) * 2 - 1
    return np.clip(np.rint(hard * 60 + rng.normal(0.0, 45.0, size=55296)), -127, 127).astype(np.int8)
