"""BASELINE.json configs[3] per GPU at its full size -- 1 024 Mode I streams, FIC + 18 DAB+ sub-channels each, fed as u8 IQ
through dabgpu_submit / dabgpu_wait -- checked through properties that do not need the oracle to run 1 024 receivers:
  * round trip: every logical frame the chain gives back is byte for byte one the transmitter encoded (channel coding, time
    interleaving, OFDM, AWGN at 15 dB, CFO up to +-20 kHz, arbitrary timing in between), and within one period of the
    transmission every stream gives back every transmitted frame exactly once (a checksum of checksums per stream);
  * every FIB passes its CRC, every sub-channel is flagged valid, no RS / fire code / AU CRC failure in steady state.
The reference chain itself is compared on streams of this workload by bench.py's spot check and, at small sizes, by the other tests.
"""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

S, PERIOD, WARM, BLOCK = 1024, 10, 10, 65536


def _hash(rows, w):
    """order-sensitive 64-bit checksum of each row of bytes (wrapping arithmetic)"""
    return (rows.astype(np.uint64) * w[None, :rows.shape[-1]]).sum(axis=-1, dtype=np.uint64)


def test_full_chain_1024_streams_round_trip(gpu_ctx, tx):
    import torch
    synth = importlib.import_module(gpu_ctx.__name__ + ".synth.gpusynth")
    subs = tx.default_ensemble()
    FS = tx.MODES[1].nb_frame_samples
    made = [tx.periodic_frames(1, subs, seed=4000 + u, period_frames=PERIOD, return_logical=True) for u in range(2)]
    payload = np.stack([m[0] for m in made])
    rng = np.random.default_rng(5)
    w = rng.integers(1, 2 ** 63, size=256, dtype=np.uint64) | np.uint64(1)
    # expected checksums: [unique ensemble][sub-channel] -> sorted hashes of the 40 logical frames of a period
    want = [[np.sort(_hash(m[1][sc.id], w)) for sc in subs] for m in made]
    nb = subs[0].frame_bytes
    assert all(sc.frame_bytes == nb for sc in subs) and made[0][1][0].shape == (PERIOD * 4, nb)

    cyc, _, _ = synth.make_cyclic_streams_u8(S, payload, mode=1, seed0=77, snr_db=15.0, device="cuda:0")
    hb = gpu_ctx.HostBuffer(S * 2 * PERIOD * FS)
    h_np = hb.array.reshape(S, 2 * PERIOD * FS)
    torch.from_numpy(h_np).copy_(cyc)
    del cyc
    torch.cuda.synchronize()

    g = gpu_ctx.DabGpu(mode=1, max_streams=S)
    for s in range(S):
        g.msc_configure(s, subs)
    lay = [g.msc_layout(0, k) for k in range(len(subs))]
    assert all(n == nb for _, n in lay)
    P = g.P
    out = dict(msc_host=torch.empty((S, P.nb_cifs, gpu_ctx.CIF_OUT_STRIDE), dtype=torch.uint8).pin_memory(),
               fic_host=torch.empty((S, P.nb_cifs, gpu_ctx.FIC_GROUP_STRIDE), dtype=torch.uint8).pin_memory(),
               fic_crc_host=torch.empty((S, P.nb_cifs, 4), dtype=torch.uint8).pin_memory(),
               msc_valid_host=torch.empty((S, P.nb_cifs, 64), dtype=torch.uint8).pin_memory(),
               chan_status_host=torch.zeros((S, 2), dtype=torch.int32).pin_memory())
    got = np.zeros((S, len(subs), PERIOD * P.nb_cifs), dtype=np.uint64)
    c_start = None
    for i in range(WARM + PERIOD):
        view = h_np[:, 2 * (i % PERIOD) * FS: 2 * ((i % PERIOD) + 1) * FS]
        t = g.submit(view.ctypes.data, h_np.strides[0], FS, block_size=BLOCK, run_chan_decode=True, **{k: v.data_ptr() for k, v in out.items()})
        g.wait(t)
        if i == WARM - 1:
            c_start = g.counters()
        if i < WARM:
            continue
        j = i - WARM
        assert int(out["chan_status_host"][:, 0].sum()) == S, f"step {i}: not every stream decoded a frame"
        assert bool((out["fic_crc_host"].numpy()[:, :, :3] == 1).all()), f"step {i}: a FIB failed its CRC"
        assert bool((out["msc_valid_host"].numpy()[:, :, :len(subs)] == 1).all()), f"step {i}: a sub-channel CIF is not valid"
        msc = out["msc_host"].numpy()
        for k, (off, n) in enumerate(lay):
            got[:, k, j * P.nb_cifs:(j + 1) * P.nb_cifs] = _hash(msc[:, :, off:off + n].reshape(S * P.nb_cifs, n), w).reshape(S, P.nb_cifs)
    c_end = g.counters()
    got.sort(axis=-1)
    for s in range(S):
        for k in range(len(subs)):
            assert np.array_equal(got[s, k], want[s % 2][k]), f"stream {s} sub-channel {k}: decoded frames are not the transmitted ones"
    d = {k: c_end[k] - c_start[k] for k in c_end}
    assert d["frames_channel_decoded"] == S * PERIOD and d["fibs_crc_ok"] == d["fibs_total"] == S * PERIOD * 12
    assert d["superframes_rs_fail"] == 0 and d["superframes_firecode_fail"] == 0 and d["au_crc_fail"] == 0
    assert d["superframes_ok"] == S * len(subs) * PERIOD * 4 // 5 and d["au_ok"] > 0
    assert d["msc_bytes_decoded"] == S * PERIOD * 4 * len(subs) * nb
    g.close()
    hb.close()
