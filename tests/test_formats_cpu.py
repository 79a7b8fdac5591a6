"""Capture file formats: host/capture_formats.hpp through the C ABI against the reference's own
readers.  tests/golden/formats_kat.npz comes from the reference build (tests/golden/make_formats_golden.py); where oracle/_ref is
present the comparison is repeated live on fresh random buffers.  Bar: bit-identical floats and bytes."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_iq_reader_modes_match_reference_golden(dab):
    kat = np.load(os.path.join(GOLDEN, "formats_kat.npz"))
    for mode in dab.IQ_FILE_MODES:
        got = dab.iq_convert(mode, kat[mode + "_raw"].tobytes())
        exp = kat[mode + "_c32"]
        assert got.shape == exp.shape, mode
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), f"{mode}: samples differ from the reference reader (bit pattern)"


def test_raw_u8_reader_is_what_the_gpu_computes(dab):
    """(u8 - 127.5) * (1/127.5): the conversion k_ofdm_ctl / k_ofdm_demod fuse into their loads (app_iq_readers.h:19-87)."""
    raw = np.arange(256, dtype=np.uint8).repeat(2)
    got = dab.iq_convert("raw_u8", raw.tobytes())
    exp = ((np.arange(256, dtype=np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)).astype(np.float32)
    assert np.array_equal(got.real, exp) and np.array_equal(got.imag, exp)


def test_hard_byte_frames_match_reference_golden(dab):
    kat = np.load(os.path.join(GOLDEN, "formats_kat.npz"))
    assert np.array_equal(dab.softbits_to_bytes(kat["soft_bits"]), kat["hard_bytes"])
    assert np.array_equal(dab.bytes_to_softbits(kat["hard_bytes"]), kat["soft_again"])
    with pytest.raises(dab.DabGpuError):
        dab.softbits_to_bytes(np.zeros(13, dtype=np.int8))


def test_unknown_mode_is_rejected(dab):
    for mode in ("wav", "raw_s24l", ""):
        with pytest.raises(dab.DabGpuError):
            dab.iq_convert(mode, b"\0" * 16)


def test_formats_match_reference_live(dab, pyref, ref_ok):
    if not hasattr(pyref.RefLib.get().L, "ref_iq_convert"):
        pytest.skip("oracle/_ref built without the capture-format helpers")
    rng = np.random.default_rng(4)
    for mode in dab.IQ_FILE_MODES:
        if "f32" in mode or "f64" in mode:
            v = (rng.standard_normal(2000) * 10.0 ** rng.integers(-3, 4)).astype({"raw_f32l": "<f4", "raw_f32b": ">f4", "raw_f64l": "<f8", "raw_f64b": ">f8"}[mode])
            raw = np.frombuffer(v.tobytes(), dtype=np.uint8)
        else:
            raw = rng.integers(0, 256, 8192, dtype=np.uint8)
        got, exp = dab.iq_convert(mode, raw.tobytes()), pyref.ref_iq_convert(mode, raw)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), mode
    bits = rng.integers(-128, 128, 8 * 4096).astype(np.int8)
    assert np.array_equal(dab.softbits_to_bytes(bits), pyref.ref_softbits_to_bytes(bits))
    b = rng.integers(0, 256, 4096, dtype=np.uint8)
    assert np.array_equal(dab.bytes_to_softbits(b), pyref.ref_bytes_to_softbits(b))
