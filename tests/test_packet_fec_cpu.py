"""Packet-mode FEC, CPU side: the restatement oracle/pyref.py:PortPacketFec against the golden
callback log recorded from the reference's MSC_Reed_Solomon_Data_Packet_Processor (tests/golden/make_packet_fec_golden.py), and,
where oracle/_ref is present, against the reference build on a fresh packet stream.  Bar: identical bytes, flags and order."""
import os

import numpy as np

from conftest import GOLDEN


def _calls(kat):
    off = np.concatenate([[0], np.cumsum(kat["call_len"])])
    loff = np.concatenate([[0], np.cumsum(kat["log_len"])])
    for i in range(kat["call_len"].size):
        yield kat["calls"][off[i]:off[i + 1]], int(kat["used"][i]), kat["logs"][loff[i]:loff[i + 1]].tobytes()


def test_port_packet_fec_matches_golden(pyref):
    kat = np.load(os.path.join(GOLDEN, "packet_fec_kat.npz"))
    port = pyref.PortPacketFec()
    n_corrected = n_plain = 0
    for i, (buf, used, log) in enumerate(_calls(kat)):
        got_used, got = port.read_packet(buf)
        exp = pyref.parse_packet_log(log)
        assert got_used == used, f"call {i}: consumed {got_used} != {used}"
        assert got == exp, f"call {i}: callbacks differ"
        n_corrected += sum(1 for _, ok in exp if ok)
        n_plain += sum(1 for _, ok in exp if not ok)
    assert n_corrected > 150 and n_plain > 100    # both exits of the processor are exercised


def test_golden_corrected_packets_are_the_transmitted_ones(pyref, tx):
    """the first two sets of the golden stream: clean, and 60 byte errors (all rows correctable) -> original packets come out"""
    rng = np.random.default_rng(2024)
    clean = tx.packet_fec_set(rng)
    second = tx.packet_fec_set(rng)
    kat = np.load(os.path.join(GOLDEN, "packet_fec_kat.npz"))
    out = []
    for buf, used, log in _calls(kat):
        out += pyref.parse_packet_log(log)
    data = [p.tobytes() for p in clean[:-9]] + [p.tobytes() for p in second[:-9]]
    assert [p for p, _ in out[:len(data)]] == data and all(ok for _, ok in out[:len(data)])


def test_port_vs_reference_packet_fec_random(pyref, ref_ok, tx):
    rng = np.random.default_rng(77)
    ref, port = pyref.RefPacketFec(), pyref.PortPacketFec()
    for trial in range(8):
        s = tx.packet_fec_set(rng)
        for _ in range(int(rng.integers(0, 120))):
            k = int(rng.integers(0, len(s)))
            s[k][int(rng.integers(0, s[k].size))] ^= int(rng.integers(1, 256))
        if trial % 4 == 3:
            del s[int(rng.integers(0, len(s)))]
        for p in s:
            assert ref.read_packet(p) == port.read_packet(p)
