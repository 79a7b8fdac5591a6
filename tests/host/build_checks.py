"""TEST INFRASTRUCTURE: builds tests/host/lane_core_check (g++), the CPU check of the product header csrc/viterbi_lane_core.h
(compiled with its host emulation of the GPU instructions) against the oracle library oracle/_ref/libdaboracle.so.
Kept outside the package: nothing under sdrplusplus-dab-radio-plugin_b200/ may link or load anything under oracle/."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sdrplusplus-dab-radio-plugin_b200", "csrc")
LANE_SRC = os.path.join(HERE, "lane_core_check.cpp")
LANE_BIN = os.path.join(HERE, "_bin", "lane_core_check")


def build_lane_core_check(force: bool = False) -> str:
    ora = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ora, "libdaboracle.so")):
        subprocess.check_call(["make", "-s", "port"], cwd=os.path.join(ROOT, "oracle"))
    deps = [LANE_SRC, os.path.join(CSRC, "viterbi_lane_core.h"), os.path.join(ora, "libdaboracle.so")]
    if not force and os.path.exists(LANE_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(LANE_BIN) for d in deps):
        return LANE_BIN
    os.makedirs(os.path.dirname(LANE_BIN), exist_ok=True)
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas", "-o", LANE_BIN, LANE_SRC,
           "-L" + ora, "-ldaboracle", "-Wl,-rpath," + ora]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building tests/host/lane_core_check")
    return LANE_BIN


if __name__ == "__main__":
    print(build_lane_core_check(force="--force" in sys.argv))
