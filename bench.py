#!/usr/bin/env python
"""bench.py -- DAB Mode I receive-path throughput on B200 (driver contract: one JSON line on rank 0).

  python bench.py --gpus N --steps K --warmup W [--workload ofdm|full] [--streams S]
  python bench.py --impl reference ...      # the reference's own CPU implementation on the host cores

Workload (BASELINE.json configs[1]): OFDM demodulation (PRS sync, FFT, DQPSK, frequency de-interleave) batched
over 256 Mode I streams per GPU.  One step = every stream advances by one transmission frame (196608 IQ samples,
three Process() blocks of 65536), i.e. 50.3 M samples per GPU per step.  Inputs are synthetic (seeded, per-stream
CFO/timing/noise), resident in HBM for `value`; `e2e` pushes pinned host u8 IQ through the C ABI and reads the
soft-bit frames back every step.  `--workload full` adds FIC/MSC Viterbi + DAB+ RS for a full ensemble.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = "sdrplusplus-dab-radio-plugin_b200"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_FRAME = 393216 + 230400      # u8 IQ in + int8 soft bits out per Mode I frame
FRAME_SAMPLES = 196608
BLOCK = 65536
# dram__bytes_read.sum + dram__bytes_write.sum of one k_ofdm_demod2 launch over 256 frames (ncu --set full, profiles/r1d_ofdm_demod_ncu_full.txt):
# 99.67 MB + 29.47 MB; below the algorithmic 623616 B/frame because part of the soft-bit rows is still in L2 when the kernel ends
NCU_TRAFFIC_PER_FRAME = (99666688 + 29469184) / 256


def _load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region.  NVML is polled in-process every ~2 ms (an nvidia-smi
    subprocess takes longer than the whole timed region); falls back to the nvidia-smi line of B200_PROFILING.md."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.source = "nvml"
        self.ready = threading.Event()   # set after the first sample: the timed region starts only then

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.ready.set()
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            for n, b in bits.items():
                if r & b:
                    self.reasons.add(n)
            time.sleep(0.002)

    def _run_smi(self):
        self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.ready.set()
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_min_mhz": s[0] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "source": self.source}


def cpu_reference(n_threads: int, frames_per_thread: int, target_seconds: float, full: bool = False):
    """Times the reference's own code (oracle/_ref, unmodified sources) or, if that library is absent, the C port, one
    independent receiver per thread on `n_threads` host threads.  OFDM workload: OFDM_Demod only.  Full workload: OFDM_Demod
    -> FIC_Decoder + 18 x MSC_Decoder (EEP 3-A) + 18 x AAC_Frame_Processor, i.e. what the GPU does per stream.
    Returns (MS/s, kind, sample description, wall seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import pyref
    tx = importlib.import_module(PKG + ".synth.dabtx")
    rng = np.random.default_rng(7)
    p = tx.MODES[1]
    subs = tx.default_ensemble()
    use_chain = full and pyref.ref_available() and hasattr(pyref.RefLib.get().L, "ref_time_chain_u8")
    if use_chain:
        ens = tx.EnsembleTx(1, subs, seed=11)
        frames = [ens.next_frame_bits() for _ in range(frames_per_thread)]
    else:
        frames = [rng.integers(0, 2, size=p.nb_frame_bits, dtype=np.uint8) for _ in range(frames_per_thread)]
    iq = tx.ofdm_modulate(frames, 1)
    u8 = tx.to_u8(tx.impair(iq, 15.0, 2.1e-3, 4321, seed=3, tail_samples=2000), 30.0)
    n_samples = u8.size // 2
    what = "OFDM_Demod"
    if use_chain:
        L, kind = pyref.RefLib.get().L, "reference"
        sub_arr = np.array([[sc.start_address, sc.length, int(sc.is_uep), sc.uep_index, sc.eep_level, int(sc.eep_type_b), int(sc.dabplus)]
                            for sc in subs], dtype=np.int32)
        counts = np.zeros(4, dtype=np.int64)
        run = lambda rep: L.ref_time_chain_u8(1, u8, n_samples, BLOCK, rep, sub_arr, len(subs), counts)
        what = f"OFDM_Demod + FIC_Decoder + {len(subs)} x (MSC_Decoder + AAC_Frame_Processor)"
    elif pyref.ref_available():
        L, kind = pyref.RefLib.get().L, "reference"
        run = lambda rep: L.ref_time_ofdm_u8(1, u8, n_samples, BLOCK, rep, None)
    else:
        if not pyref.port_available():
            subprocess.check_call(["make", "-s", "port"], cwd=os.path.join(ROOT, "oracle"))
        L, kind = pyref.PortLib.get().L, "port"
        run = lambda rep: L.dabo_time_ofdm_u8(1, u8, n_samples, BLOCK, rep, None)
    t_one = run(1)     # warm-up + calibration of the bounded sample
    t_one = run(1)
    repeat = max(2, int(target_seconds / max(t_one, 1e-4)))
    fn = lambda: run(repeat)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=fn) for _ in range(n_threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    msps = n_threads * repeat * n_samples / dt / 1e6
    sample = (f"{n_threads} independent Mode I streams x {repeat}x{frames_per_thread} frames u8 IQ through {what} "
              f"(threads=1 as in the plugin, blocks of {BLOCK}), one receiver per host thread, {dt:.1f} s wall")
    return msps, kind, sample, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals, total_dt = [], 0.0
    sample = kind = ""
    for i in range(args.warmup + args.steps):
        msps, kind, sample, dt = cpu_reference(cores, 20, 4.0, full=args.workload == "full")
        if i >= args.warmup:
            vals.append(msps)
            total_dt += dt
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "dab_mode1_iq_msps", "value": v, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_dt / max(len(vals), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("full_chain_mode1_" if args.workload == "full" else "ofdm_demod_mode1_") + f"{args.streams}_streams_per_gpu",
                   "note": "the reference's own CPU code on all host cores, bounded sample per step"},
        "realtime_streams": v / 2.048,
        "cpu_baseline": {"value": v, "unit": "MS/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def channel_leg(pkg, synth, tx, torch, dev, local_rank, stream, S, rank, steps=24, warm=6):
    """Full chain (OFDM -> FIC + 18 x EEP 3-A DAB+ sub-channels -> RS superframes) over the same S streams, `steps` timed steps.
    Reports the Viterbi throughput (decoded information bits / device time of the channel-decode kernels, CUDA events on
    the launching stream) and the whole-chain step time.  Inputs are resident in HBM; results are checked by the counters
    (every FIB CRC must pass on this 15 dB input)."""
    import numpy as np
    subs = tx.default_ensemble()
    n_frames = warm + 2 * steps + 2
    payload = np.zeros((2, n_frames, tx.MODES[1].nb_frame_bits), dtype=np.uint8)
    for u in range(2):
        ens = tx.EnsembleTx(1, subs, seed=5000 * (rank + 1) + u)
        for f in range(n_frames):
            payload[u, f] = ens.next_frame_bits()
    iq = synth.make_streams_u8(S, n_frames, mode=1, seed0=77 + rank, snr_db=15.0, device=str(dev), payload_bits=payload)
    total = iq.shape[1] // 2
    g = pkg.DabGpu(mode=1, max_streams=S, device=local_rank, cuda_stream=stream.cuda_stream)
    g.ofdm_attach_device_input(iq.data_ptr(), total, total)
    for s in range(S):
        g.msc_configure(s, subs)

    def step():
        g.ofdm_advance(FRAME_SAMPLES, block_size=BLOCK)
        g.chan_decode()

    for _ in range(warm):      # >= 5 frames: the time de-interleaver emits nothing before 16 CIFs
        step()
    torch.cuda.synchronize()
    c0 = g.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c1 = g.counters()
    g.profile_enable(True)
    for _ in range(steps):
        step()
    prof = g.profile_read()
    g.profile_enable(False)
    c2 = g.counters()
    g.close()
    bits = lambda a, b: (b["msc_bytes_decoded"] - a["msc_bytes_decoded"]) * 8 + (b["fibs_total"] - a["fibs_total"]) * 256
    chan_ms = prof["viterbi"]["ms"] + prof["chan_misc"]["ms"]
    return {
        "workload": f"full_chain_mode1_{S}_streams_per_gpu (FIC + 18 x EEP 3-A 48 CU DAB+ sub-channels per stream)",
        "steps": steps, "ms_per_step": ms / steps, "iq_msps": S * FRAME_SAMPLES * steps / (ms * 1e-3) / 1e6,
        "realtime_streams": S * FRAME_SAMPLES * steps / (ms * 1e-3) / 1e6 / 2.048,
        "viterbi_mbit_s": bits(c1, c2) / (chan_ms * 1e-3) / 1e6,
        "viterbi_mbit_s_whole_chain": bits(c0, c1) / (ms * 1e-3) / 1e6,
        "viterbi_kernels_ms_per_step": chan_ms / steps,
        "kernel_ms": {k: v["ms"] for k, v in prof.items()},
        "fibs_crc_ok": c1["fibs_crc_ok"] - c0["fibs_crc_ok"], "fibs_total": c1["fibs_total"] - c0["fibs_total"],
        "superframes_ok": c1["superframes_ok"] - c0["superframes_ok"],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ofdm", choices=["ofdm", "full"])
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-channel-leg", action="store_true", help="OFDM workload: skip the short full-chain pass that reports Viterbi Mbit/s")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth.gpusynth")
    tx = importlib.import_module(PKG + ".synth.dabtx")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the DAB path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    S, K, W = args.streams, args.steps, args.warmup
    n_frames = W + 2 * K + 2     # K timed steps for `value`, K more with per-kernel CUDA events for the roofline
    full = args.workload == "full"
    payload = None
    subs = tx.default_ensemble()
    if full:
        # two distinct coded ensembles (FIC + 18 x EEP 3-A DAB+ sub-channels), shared by the streams with different channels
        n_unique = 2
        payload = np.zeros((n_unique, n_frames, tx.MODES[1].nb_frame_bits), dtype=np.uint8)
        for u in range(n_unique):
            ens = tx.EnsembleTx(1, subs, seed=1000 * (rank + 1) + u)
            for f in range(n_frames):
                payload[u, f] = ens.next_frame_bits()
    iq = synth.make_streams_u8(S, n_frames, mode=1, seed0=1 + rank, snr_db=15.0, device=str(dev), payload_bits=payload)
    total_samples = iq.shape[1] // 2
    torch.cuda.synchronize()

    # a non-default stream: the library launches every kernel on it and torch.cuda.Event times that same stream
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    g = pkg.DabGpu(mode=1, max_streams=S, device=local_rank, cuda_stream=stream.cuda_stream)
    g.ofdm_attach_device_input(iq.data_ptr(), total_samples, total_samples)
    if full:
        for s in range(S):
            g.msc_configure(s, subs)

    def step():
        g.ofdm_advance(FRAME_SAMPLES, block_size=BLOCK)
        if full:
            g.chan_decode()

    for _ in range(W):
        step()
    torch.cuda.synchronize()
    c0 = g.counters()
    launches0 = g.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(timeout=20)   # NVML initialisation on a fresh box can take longer than the whole timed region
    sampler.samples.clear()          # keep only what is sampled while the kernels run
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(K):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    c1 = g.counters()
    launches = g.launch_count - launches0
    # second pass with CUDA events around every kernel launch (dabgpu_profile_*): per-kernel times for the roofline.  The
    # library serialises its two stream groups while profiling, so this pass is a little slower than the one above.
    g.profile_enable(True)
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evp0.record(stream)
    for _ in range(K):
        step()
    evp1.record(stream)
    torch.cuda.synchronize()
    ms_prof = evp0.elapsed_time(evp1)
    prof = g.profile_read()
    g.profile_enable(False)
    c2 = g.counters()
    frames_prof = c2["frames_demodulated"] - c1["frames_demodulated"]
    sampler.stop_flag = True
    sampler.join(timeout=2)
    frames_demod = c1["frames_demodulated"] - c0["frames_demodulated"]
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    fr = torch.tensor([float(frames_demod)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fr, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    frames_all = float(fr.item())
    # whole-job throughput: samples actually consumed by all ranks / max time (every stream advances one frame per step)
    value = world * S * FRAME_SAMPLES * K / (ms_max * 1e-3) / 1e6

    # ---- end to end through the C ABI with HOST buffers: dabgpu_submit / dabgpu_wait, two tickets in flight ----
    # every step copies that step's u8 IQ from pinned host memory (H2D), runs the kernels and copies the step's
    # results back (D2H): the soft-bit frames for the OFDM workload, the decoded FIC/MSC bytes for the full chain.
    Ke = min(args.e2e_steps, K)
    n_e = W + Ke + 1
    P = g.P
    host_iq = torch.empty((S, 2 * n_e * FRAME_SAMPLES), dtype=torch.uint8).pin_memory()
    host_iq.copy_(iq[:, :2 * n_e * FRAME_SAMPLES])
    D = pkg.PIPELINE_DEPTH
    if full:
        outs = [dict(msc_host=torch.empty((S, P.nb_cifs, pkg.CIF_OUT_STRIDE), dtype=torch.uint8).pin_memory(),
                     fic_host=torch.empty((S, P.nb_cifs, pkg.FIC_GROUP_STRIDE), dtype=torch.uint8).pin_memory(),
                     fic_crc_host=torch.empty((S, P.nb_cifs, 4), dtype=torch.uint8).pin_memory(),
                     msc_valid_host=torch.empty((S, P.nb_cifs, 64), dtype=torch.uint8).pin_memory(),
                     chan_status_host=torch.zeros((S, 2), dtype=torch.int32).pin_memory()) for _ in range(D)]
    else:
        outs = [dict(frames_host=torch.empty((S, P.nb_frame_bits), dtype=torch.int8).pin_memory(),
                     produced_host=torch.zeros(S, dtype=torch.uint8).pin_memory()) for _ in range(D)]
    d2h_bytes = sum(t.numel() * t.element_size() for t in outs[0].values())
    g2 = pkg.DabGpu(mode=1, max_streams=S, device=local_rank, cuda_stream=stream.cuda_stream)
    if full:
        for s in range(S):
            g2.msc_configure(s, subs)
    h_np = host_iq.numpy()
    tickets = {}
    n_prod = 0

    def e2e_submit(i):
        o = outs[i % D]
        view = h_np[:, 2 * i * FRAME_SAMPLES: 2 * (i + 1) * FRAME_SAMPLES]
        tickets[i] = g2.submit(view.ctypes.data, h_np.strides[0], FRAME_SAMPLES, block_size=BLOCK, run_chan_decode=full,
                               **{k: v.data_ptr() for k, v in o.items()})

    def e2e_retire(i):
        g2.wait(tickets.pop(i))
        o = outs[i % D]
        return int(o["chan_status_host"][:, 0].sum()) if full else int(o["produced_host"].sum())

    for i in range(W):
        e2e_submit(i)
        e2e_retire(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(W, W + Ke):
        if i - D >= W:
            n_prod += e2e_retire(i - D)
        e2e_submit(i)
    for i in range(max(W, W + Ke - D), W + Ke):
        n_prod += e2e_retire(i)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * S * FRAME_SAMPLES * Ke / (float(te.item()) * 1e-3) / 1e6
    g2.close()

    # ---- roofline of the dominant kernel (k_ofdm_demod): algorithmic bytes / CUDA-event time of its launches ----
    peak, peak_src = _load_peaks()
    demod_ms = prof["ofdm_demod"]["ms"]
    demod_launches = prof["ofdm_demod"]["launches"]
    achieved = (frames_prof * ALG_BYTES_PER_FRAME) / (demod_ms * 1e-3) / 1e9 if demod_ms > 0 else 0.0
    roofline = {
        "kernel": "k_ofdm_demod<2048>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": (NCU_TRAFFIC_PER_FRAME * frames_prof / max(demod_launches, 1)) if NCU_TRAFFIC_PER_FRAME else None,
        "peak_source": f"{peak_src} (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback 6.65 TB/s (B200_PROFILING.md)",
        "algorithmic_bytes_per_frame": ALG_BYTES_PER_FRAME, "frames_in_profiled_pass": frames_prof,
        "kernel_ms_total": demod_ms, "kernel_launches": demod_launches, "kernel_share_of_step": demod_ms / ms_prof if ms_prof > 0 else None,
        "profiled_pass_ms_per_step": ms_prof / K,
        "note": "u8 input makes this kernel FP32-issue bound (~100 flop/sample), not HBM bound; see DESIGN.md",
    }
    line = {
        "metric": "dab_mode1_iq_msps", "value": value, "unit": "MS/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": ("full_chain_mode1_" if full else "ofdm_demod_mode1_") + f"{S}_streams_per_gpu",
                   "streams_per_gpu": S, "frame_samples": FRAME_SAMPLES, "process_block": BLOCK, "iq_format": "u8",
                   "cache": f"inputs larger than L2: every step reads {S * FRAME_SAMPLES * 2 / 1e6:.0f} MB of IQ never touched before",
                   "snr_db": 15, "cfo": "uniform +-20 kHz", "timing": "uniform lead in [0, 196608)"},
        "realtime_streams": value / 2.048,
        "frames_demodulated": frames_all,
        "e2e": {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": S * FRAME_SAMPLES * 2, "d2h_bytes_per_step": d2h_bytes,
                "steps": Ke, "frames_returned": n_prod, "api": "dabgpu_submit/dabgpu_wait, 2 tickets in flight, pinned host buffers"},
        "gpu_launches": int(launches),
        "kernel_ms": {k: v["ms"] for k, v in prof.items()},
        "roofline": roofline,
        "clocks": sampler.summary(),
    }
    if not full and not args.no_channel_leg:
        # the metric also names "Viterbi Mbit/s": a short full-chain pass (BASELINE.json configs[2]/[3] on this GPU's streams)
        g.close()
        g = None
        line["channel_decode"] = channel_leg(pkg, synth, tx, torch, dev, local_rank, stream, S, rank)
        line["viterbi_mbit_s"] = line["channel_decode"]["viterbi_mbit_s"]
    if full:
        line["viterbi_mbit_s"] = ((c1["msc_bytes_decoded"] - c0["msc_bytes_decoded"]) * 8 + (c1["fibs_total"] - c0["fibs_total"]) * 256) / (ms * 1e-3) / 1e6
        line["counters"] = {k: c1[k] - c0[k] for k in c1}
    if g is not None:
        g.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, kind, sample, _ = cpu_reference(cores, 20, 12.0, full=full)
        line["cpu_baseline"] = {"value": v, "unit": "MS/s", "cores": cores, "kind": kind, "sample": sample}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
