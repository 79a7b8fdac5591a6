"""Synthetic DAB IQ generation on the GPU with torch (workload generation for bench.py; not timed, not product).

Same signal model as synth/dabtx.py (which the reference decoder validates): PRS + pi/4-DQPSK data symbols with
the frequency interleaver, cyclic prefix, NULL symbol, per-stream CFO / timing offset / AWGN, u8 quantisation
"clamp(trunc(x*s + 127.5))" (examples/app_helpers/app_iq_readers.h:51-63).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import dabtx


def make_streams_u8(n_streams: int, n_frames: int, mode: int = 1, seed0: int = 1, snr_db: float = 15.0,
                    cfo_norm_max: float = 20e3 / 2.048e6, device: str = "cuda", payload_bits: Optional[np.ndarray] = None,
                    rms_lsb: float = 30.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns uint8 [n_streams, 2*(n_frames+1)*frame_samples]: I,Q interleaved, each stream with its own seed,
    CFO uniform in +-cfo_norm_max (cycles/sample), leading noise uniform in [0, frame_samples), SNR snr_db.
    payload_bits: optional [n_unique, n_frames, frame_bits] 0/1 array (coded frames); stream s uses s % n_unique."""
    p = dabtx.MODES[mode]
    K, N, CP, T = p.nb_carriers, p.nb_fft, p.nb_cyclic_prefix, p.nb_symbol_period
    L = p.nb_frame_symbols
    fs = p.nb_frame_samples
    total = (n_frames + 1) * fs
    dev = torch.device(device)
    cmap = torch.from_numpy(dabtx.carrier_map(N, K)).to(dev)
    slot_k = np.concatenate([np.arange(-K // 2, 0), np.arange(1, K // 2 + 1)])
    slot_bin = torch.from_numpy((slot_k % N).astype(np.int64)).to(dev)
    prs = torch.from_numpy(np.exp(1j * np.pi / 2 * dabtx.prs_phase_index(mode)).astype(np.complex64)).to(dev)
    if out is None:
        out = torch.empty((n_streams, 2 * total), dtype=torch.uint8, device=dev)
    pb = None if payload_bits is None else torch.from_numpy(np.ascontiguousarray(payload_bits)).to(dev)
    rng = np.random.default_rng(seed0)
    cfos = rng.uniform(-cfo_norm_max, cfo_norm_max, size=n_streams)
    leads = rng.integers(0, fs, size=n_streams)
    sigma = math.sqrt(0.5 * 10 ** (-snr_db / 10))
    n_idx = torch.arange(total, device=dev, dtype=torch.float64)
    for s in range(n_streams):
        g = torch.Generator(device=dev)
        g.manual_seed(seed0 * 100003 + s)
        if pb is None:
            bits = torch.randint(0, 2, (n_frames, L - 1, 2 * K), generator=g, device=dev, dtype=torch.uint8)
        else:
            bits = pb[s % pb.shape[0]].reshape(n_frames, L - 1, 2 * K)
        b0 = bits[:, :, :K].to(torch.float32)
        b1 = bits[:, :, K:].to(torch.float32)
        q_pair = torch.complex(1 - 2 * b0, 1 - 2 * b1) * (1 / math.sqrt(2.0))
        q = torch.empty_like(q_pair)
        q[:, :, cmap] = q_pair
        cur = prs[None, None, :] * torch.cumprod(q, dim=1)
        cur = cur / cur.abs()    # keep unit modulus over 75 products
        slots = torch.cat([prs[None, None, :].expand(n_frames, 1, K), cur], dim=1)
        spec = torch.zeros((n_frames, L, N), dtype=torch.complex64, device=dev)
        spec[:, :, slot_bin] = slots
        t = torch.fft.ifft(spec, dim=-1) * (N / math.sqrt(K))
        syms = torch.cat([t[..., N - CP:], t], dim=-1).reshape(n_frames, L * T)
        sig = torch.cat([torch.zeros((n_frames, p.nb_null_period), dtype=torch.complex64, device=dev), syms], dim=1).reshape(-1)
        x = torch.zeros(total, dtype=torch.complex64, device=dev)
        lead = int(leads[s])
        x[lead:lead + sig.numel()] = sig[:total - lead]
        ph = torch.remainder(n_idx * float(cfos[s]), 1.0).to(torch.float32) * (2 * math.pi)
        x = x * torch.complex(torch.cos(ph), torch.sin(ph))
        noise = torch.randn((total, 2), generator=g, device=dev, dtype=torch.float32) * sigma
        v = torch.view_as_real(x) + noise
        v = torch.clamp(torch.trunc(v * rms_lsb + 127.5), 0, 255).to(torch.uint8)
        out[s] = v.reshape(-1)
    return out


def make_cyclic_streams_u8(n_streams: int, payload_bits: np.ndarray, mode: int = 1, seed0: int = 1, snr_db: float = 15.0,
                           cfo_norm_max: float = 20e3 / 2.048e6, device: str = "cuda", rms_lsb: float = 30.0):
    """One PERIOD of every stream: uint8 [n_streams, 2*P*frame_samples] that can be repeated back to back for ever.
    payload_bits: [n_unique, P, frame_bits] coded frames whose content is periodic with period P (bench.py: periodic_payload).
    Stream s is the P-frame signal of payload s % n_unique rotated by its timing lead (uniform in [0, frame_samples)), with a
    CFO uniform in +-cfo_norm_max rounded to a whole number of cycles per period (no phase jump where the period repeats) and
    one period of AWGN.  Returns (tensor, cfos, leads)."""
    p = dabtx.MODES[mode]
    K, N, CP, T = p.nb_carriers, p.nb_fft, p.nb_cyclic_prefix, p.nb_symbol_period
    L = p.nb_frame_symbols
    fs = p.nb_frame_samples
    P = int(payload_bits.shape[1])
    total = P * fs
    dev = torch.device(device)
    cmap = torch.from_numpy(dabtx.carrier_map(N, K)).to(dev)
    slot_k = np.concatenate([np.arange(-K // 2, 0), np.arange(1, K // 2 + 1)])
    slot_bin = torch.from_numpy((slot_k % N).astype(np.int64)).to(dev)
    prs = torch.from_numpy(np.exp(1j * np.pi / 2 * dabtx.prs_phase_index(mode)).astype(np.complex64)).to(dev)
    out = torch.empty((n_streams, 2 * total), dtype=torch.uint8, device=dev)
    pb = torch.from_numpy(np.ascontiguousarray(payload_bits)).to(dev)
    rng = np.random.default_rng(seed0)
    cfos = np.round(rng.uniform(-cfo_norm_max, cfo_norm_max, size=n_streams) * total) / total
    leads = rng.integers(0, fs, size=n_streams)
    sigma = math.sqrt(0.5 * 10 ** (-snr_db / 10))
    n_idx = torch.arange(total, device=dev, dtype=torch.float64)
    clean = []
    for u in range(pb.shape[0]):      # the modulated period of every unique ensemble, once
        bits = pb[u].reshape(P, L - 1, 2 * K)
        q_pair = torch.complex(1 - 2 * bits[:, :, :K].to(torch.float32), 1 - 2 * bits[:, :, K:].to(torch.float32)) * (1 / math.sqrt(2.0))
        q = torch.empty_like(q_pair)
        q[:, :, cmap] = q_pair
        cur = prs[None, None, :] * torch.cumprod(q, dim=1)
        cur = cur / cur.abs()
        slots = torch.cat([prs[None, None, :].expand(P, 1, K), cur], dim=1)
        spec = torch.zeros((P, L, N), dtype=torch.complex64, device=dev)
        spec[:, :, slot_bin] = slots
        t = torch.fft.ifft(spec, dim=-1) * (N / math.sqrt(K))
        syms = torch.cat([t[..., N - CP:], t], dim=-1).reshape(P, L * T)
        clean.append(torch.cat([torch.zeros((P, p.nb_null_period), dtype=torch.complex64, device=dev), syms], dim=1).reshape(-1))
    for s in range(n_streams):
        g = torch.Generator(device=dev)
        g.manual_seed(seed0 * 100003 + s)
        x = torch.roll(clean[s % len(clean)], int(leads[s]))
        ph = torch.remainder(n_idx * float(cfos[s]), 1.0).to(torch.float32) * (2 * math.pi)
        x = x * torch.complex(torch.cos(ph), torch.sin(ph))
        noise = torch.randn((total, 2), generator=g, device=dev, dtype=torch.float32) * sigma
        v = torch.view_as_real(x) + noise
        out[s] = torch.clamp(torch.trunc(v * rms_lsb + 127.5), 0, 255).to(torch.uint8).reshape(-1)
    return out, cfos, leads
