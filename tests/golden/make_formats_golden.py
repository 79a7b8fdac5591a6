"""Generates tests/golden/formats_kat.npz with the REFERENCE's own readers (app_iq_readers.h, app_viterbi_convert_block.h,
compiled into oracle/_ref/libdabref.so).  Run where /root/reference exists:  python tests/golden/make_formats_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import pyref  # noqa: E402

MODES = ["raw_u8", "raw_s8", "raw_s16l", "raw_s16b", "raw_u16l", "raw_u16b", "raw_s32l", "raw_s32b", "raw_u32l", "raw_u32b",
         "raw_f32l", "raw_f32b", "raw_f64l", "raw_f64b"]


def raw_for(mode, rng):
    if mode in ("raw_u8", "raw_s8"):
        return np.concatenate([np.arange(256, dtype=np.uint8), rng.integers(0, 256, 1024, dtype=np.uint8)])
    if "f32" in mode:
        v = rng.standard_normal(600).astype(np.float32 if mode.endswith("l") else ">f4")
        return np.frombuffer(v.tobytes(), dtype=np.uint8).copy()
    if "f64" in mode:
        v = rng.standard_normal(600).astype(np.float64 if mode.endswith("l") else ">f8")
        return np.frombuffer(v.tobytes(), dtype=np.uint8).copy()
    raw = rng.integers(0, 256, 4096, dtype=np.uint8)
    raw[:16] = [0, 0, 0, 0, 255, 255, 255, 255, 0, 128, 0, 0, 255, 127, 255, 255]    # extremes in either byte order
    return raw


def main():
    assert pyref.ref_available()
    rng = np.random.default_rng(99)
    out = {"build_info": np.array(pyref.RefLib.get().build_info())}
    for m in MODES:
        raw = raw_for(m, rng)
        out[m + "_raw"] = raw
        out[m + "_c32"] = pyref.ref_iq_convert(m, raw)
    bits = rng.integers(-128, 128, 8 * 700).astype(np.int8)
    bits[:24] = [0, -1, 1, 127, -128, -127, 0, 0] * 3
    out["soft_bits"] = bits
    out["hard_bytes"] = pyref.ref_softbits_to_bytes(bits)
    out["soft_again"] = pyref.ref_bytes_to_softbits(out["hard_bytes"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "formats_kat.npz"), **out)
    print({k: v.shape for k, v in out.items() if k != "build_info"})


if __name__ == "__main__":
    main()
