"""CPU check of the lane-per-trellis Viterbi arithmetic (csrc/viterbi_lane_core.h) against the oracle.

The header compiles for the host with plain-C++ emulations of the five GPU instructions it is built from, so the
register layouts, the decision-bit positions, the relative-metric bookkeeping (own renormalisation, the reference's
renormalisation, saturation level) and the traceback are verified bit-exactly here, without a GPU: decoded bytes and
the accumulated path error of 1200 trellises (noisy, garbage, ties, +-full scale, bursts, all-punctured; lengths 6..6150).
"""
import importlib
import os
import subprocess
import sys

from conftest import ROOT


def test_lane_core_matches_oracle(pyref):
    sys.path.insert(0, os.path.join(ROOT, "tests", "host"))
    exe = importlib.import_module("build_checks").build_lane_core_check()
    res = subprocess.run([exe, "1200"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert " 0 mismatches" in res.stdout
