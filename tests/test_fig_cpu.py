"""Self-configuration from the FIC: host/fic_autoconfig.hpp through the C ABI (dabgpu_autocfg_*, host only,
no GPU needed) against the database the reference builds from the same FIBs.

tests/golden/fig_kat.npz was produced by the reference build (tests/golden/make_fig_golden.py); where oracle/_ref is present the
comparison is repeated live, FIB by FIB.  Bar: identical sub-channel and service-component tables (every field, database order)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

SEQS = ["dabplus18", "mixed", "conflicts", "random"]


def _runnable_from_dump(subs, comps):
    """BasicRadio::UpdateAfterProcessing (basic_radio.cpp:97-141) applied to a database dump."""
    out = []
    for s in subs:
        if not s[8]:
            continue
        comp = next((c for c in comps if c[3] == s[0]), None)
        if comp is None or not comp[9] or comp[5] != 0 or comp[6] not in (0, 63):
            continue
        out.append((int(s[0]), dict(start_address=int(s[1]), length=int(s[2]), is_uep=int(s[3]), uep_prot_index=int(s[4]), eep_prot_level=int(s[5]),
                                    eep_type_b=int(s[6] == 1), is_dabplus=int(comp[6] == 63))))
    return out


@pytest.mark.parametrize("name", SEQS)
def test_autoconfig_matches_reference_golden(dab, name):
    kat = np.load(os.path.join(GOLDEN, "fig_kat.npz"))
    a = dab.FicAutoConfig()
    a.push_fibs(kat[name + "_fibs"])
    subs, comps = a.dump()
    assert np.array_equal(subs, kat[name + "_subs"]), f"{name}: sub-channel table differs from the reference database"
    assert np.array_equal(comps, kat[name + "_comps"]), f"{name}: service component table differs from the reference database"
    got, ids = a.runnable()
    exp = _runnable_from_dump(kat[name + "_subs"], kat[name + "_comps"])
    assert ids == [i for i, _ in exp]
    assert got == [d for _, d in exp]
    a.close()


def test_autoconfig_bench_ensemble_is_recovered(dab, tx):
    """The FIBs of the synthetic transmitter (FIG 0/1 long form + FIG 0/2, repeated every FIB group) give back its sub-channels."""
    subs = tx.default_ensemble()
    ens = tx.EnsembleTx(1, subs, seed=3)
    a = dab.FicAutoConfig()
    for _ in range(3):
        ens.next_frame_bits()
    fibs = np.stack(ens.fibs)                      # 30 data bytes + CRC16 each, as FIC_Decoder would emit them
    assert a.push_fibs(fibs) > 0
    got, ids = a.runnable()
    assert ids == [s.id for s in subs]
    for g, s in zip(got, subs):
        assert (g["start_address"], g["length"], g["is_uep"], g["eep_prot_level"], g["eep_type_b"], g["is_dabplus"]) == \
               (s.start_address, s.length, int(s.is_uep), s.eep_level, int(s.eep_type_b), int(s.dabplus))
    a.close()


def test_autoconfig_crc_flags_and_errors(dab):
    kat = np.load(os.path.join(GOLDEN, "fig_kat.npz"))
    fibs = kat["dabplus18_fibs"]
    a = dab.FicAutoConfig()
    assert a.push_fibs(fibs, crc_ok=np.zeros(len(fibs), dtype=np.uint8)) == 0      # FIBs with a failed CRC are never parsed
    assert a.dump()[0].shape[0] == 0
    assert a.push_fibs(fibs) == len(fibs)
    assert a.push_fibs(fibs) == 0                                                  # repetition changes nothing
    with pytest.raises(dab.DabGpuError):
        a.push_fibs(np.zeros((2, 16), dtype=np.uint8))                             # stride below 30 bytes
    a.close()


def test_autoconfig_matches_reference_live(dab, pyref, ref_ok):
    if not hasattr(pyref.RefLib.get().L, "ref_fig_create"):
        pytest.skip("oracle/_ref built without the FIG chain")
    kat = np.load(os.path.join(GOLDEN, "fig_kat.npz"))
    for name in SEQS:
        a, r = dab.FicAutoConfig(), pyref.RefFig()
        for i, f in enumerate(kat[name + "_fibs"]):
            a.push_fibs(f)
            r.process_fib(f)
            s0, c0 = a.dump()
            s1, c1 = r.dump()
            assert np.array_equal(s0, s1) and np.array_equal(c0, c1), f"{name}: databases diverge after FIB {i}"
        a.close()
