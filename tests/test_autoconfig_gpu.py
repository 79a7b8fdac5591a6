"""A receiver that configures itself from the FIC against one that was told its sub-channels.

Soft-bit frames in -> FIC decoded on the GPU -> FIBs into the host-side FIG parser (dabgpu_autocfg_*) -> dabgpu_msc_configure.
Bar: the self-configured context recovers exactly the transmitter's sub-channel table and, once its time de-interleaver is
full, emits byte-identical sub-channel data and DAB+ events."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_self_configured_receiver_matches_told_receiver(gpu_ctx, tx):
    rng = np.random.default_rng(5)
    subs = [tx.Subchannel(3, 0, 48, eep_level=2), tx.Subchannel(7, 48, 54, eep_level=2, eep_type_b=True, dabplus=False),
            tx.Subchannel(9, 102, 16, is_uep=True, uep_index=0, dabplus=False), tx.Subchannel(12, 118, 24, eep_level=1)]
    ens = tx.EnsembleTx(1, subs, seed=9)
    told = gpu_ctx.DabGpu(mode=1, max_streams=1)
    auto = gpu_ctx.DabGpu(mode=1, max_streams=1)
    told.msc_configure(0, subs)
    cfg = gpu_ctx.FicAutoConfig()
    configured_at = None
    compared = 0
    for f in range(12):
        frame = tx.hard_to_soft(ens.next_frame_bits(), rng, snr_db=6.0)[None, :]
        for g in (told, auto):
            g.softbits_push(frame)
            g.chan_decode()
        fibs, ok = auto.get_fic(0)
        cfg.push_fibs(fibs, crc_ok=ok)
        if cfg.apply(auto, 0):
            assert configured_at is None, "the configuration must be applied once: later FIBs only repeat it"
            configured_at = f
            got, ids = cfg.runnable()
            assert ids == [s.id for s in subs]
            for gsc, s in zip(got, subs):
                assert (gsc["start_address"], gsc["length"], gsc["is_uep"], gsc["is_dabplus"]) == (s.start_address, s.length, int(s.is_uep), int(s.dabplus))
                if s.is_uep:
                    assert gsc["uep_prot_index"] == s.uep_index
                else:
                    assert (gsc["eep_prot_level"], gsc["eep_type_b"]) == (s.eep_level, int(s.eep_type_b))
        if configured_at is not None and f > configured_at:
            for k, sc in enumerate(subs):
                out_a, valid_a = auto.get_msc(0, k)
                out_t, valid_t = told.get_msc(0, k)
                for c in range(4):
                    if valid_a[c]:
                        assert valid_t[c]
                        assert np.array_equal(out_a[c], out_t[c]), (f, k, c)
                        compared += out_a[c].size
    assert configured_at == 0, "FIG 0/1 and 0/2 of this ensemble fit the first frame's FIBs"
    assert compared > 0
    cfg.close()
    told.close()
    auto.close()
