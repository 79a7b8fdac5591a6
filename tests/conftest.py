import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG = "sdrplusplus-dab-radio-plugin_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def dab():
    """The product binding (ctypes over csrc/libdabgpu.so)."""
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def tx():
    return importlib.import_module(PKG + ".synth.dabtx")


@pytest.fixture(scope="session")
def pyref():
    """oracle bindings; builds the C restatement if needed (the reference build needs /root/reference)."""
    import subprocess
    mod = importlib.import_module("pyref")
    if not mod.port_available():
        subprocess.check_call(["make", "-s", "port"], cwd=os.path.join(ROOT, "oracle"))
    return mod


@pytest.fixture(scope="session")
def ref_ok(pyref):
    if not pyref.ref_available():
        if os.path.isdir("/root/reference"):
            import subprocess
            subprocess.check_call(["make", "-s", "ref"], cwd=os.path.join(ROOT, "oracle"))
        else:
            pytest.skip("oracle/_ref/libdabref.so not built and /root/reference absent")
    return True


def has_gpu() -> bool:
    try:
        m = importlib.import_module(PKG)
        return m.load_library().dabgpu_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx(dab):
    if not has_gpu():
        pytest.fail("no CUDA device or libdabgpu.so missing: GPU tests cannot run (there is no CPU fallback)")
    return dab


# Viterbi mapping pinned through dabgpu_config.flags: auto (device-side choice by batch size), every call on the
# lane-per-trellis kernel, every call on the warp-per-trellis kernel.  All three must give the reference's bits.
VIT_LANES_ALWAYS, VIT_LANES_NEVER = 2, 4


@pytest.fixture(params=[0, VIT_LANES_ALWAYS, VIT_LANES_NEVER], ids=["vit-auto", "vit-lanes", "vit-warps"])
def vit_flags(request):
    return request.param


GOLDEN = os.path.join(ROOT, "tests", "golden")
