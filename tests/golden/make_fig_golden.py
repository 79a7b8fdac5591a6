"""Generates tests/golden/fig_kat.npz: FIB sequences and the database the REFERENCE builds from them
(FIG_Processor -> Radio_FIG_Handler -> DAB_Database_Updater, compiled in place into oracle/_ref/libdabref.so).
Run in the container that has /root/reference:  python tests/golden/make_fig_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import pyref  # noqa: E402


def fig0(ext, payload, pd=0):
    body = bytes([(pd << 5) | ext]) + bytes(payload)
    assert len(body) <= 29
    return bytes([(0 << 5) | len(body)]) + body


def fig_raw(ftype, body):
    return bytes([(ftype << 5) | len(body)]) + bytes(body)


def sub_short(sid, start, index, table_switch=0):
    return bytes([(sid << 2) | (start >> 8), start & 255, (table_switch << 6) | index])


def sub_long(sid, start, option, level, size):
    return bytes([(sid << 2) | (start >> 8), start & 255, 0x80 | (option << 4) | (level << 2) | (size >> 8), size & 255])


def service(sid, comps, pd=0):
    hdr = sid.to_bytes(4 if pd else 2, "big") + bytes([len(comps)])
    return hdr + b"".join(comps)


def comp_audio(ascty, sub, primary=1):
    return bytes([(0 << 6) | ascty, (sub << 2) | (primary << 1)])


def comp_data(dscty, sub, primary=1):
    return bytes([(1 << 6) | dscty, (sub << 2) | (primary << 1)])


def comp_packet(scid, primary=1):
    return bytes([(3 << 6) | (scid >> 6), ((scid & 63) << 2) | (primary << 1)])


def packet_def(scid, dscty, sub, addr, caorg=None):
    b = bytes([scid >> 4, ((scid & 15) << 4) | (1 if caorg is not None else 0), dscty, (sub << 2) | (addr >> 8), addr & 255])
    return b + (caorg.to_bytes(2, "big") if caorg is not None else b"")


def fib(*figs, pad=0x00, end=True):
    b = b"".join(figs)
    assert len(b) <= 30
    if len(b) < 30 and end:
        b += b"\xff"
    return np.frombuffer(b + bytes([pad]) * (30 - len(b)), dtype=np.uint8).copy()


def sequences():
    rng = np.random.default_rng(7)
    seqs = {}
    # a plain DAB+ ensemble: 18 x EEP 3-A 48 CU (the bench ensemble), FIG 0/1 and 0/2 spread over several FIBs
    s = []
    for k in range(0, 18, 6):
        s.append(fib(fig0(1, b"".join(sub_long(i, 48 * i, 0, 2, 48) for i in range(k, k + 6)))))
    for k in range(0, 18, 5):
        s.append(fib(fig0(2, b"".join(service(0xD100 + i, [comp_audio(63, i)]) for i in range(k, min(k + 5, 18))))))
    seqs["dabplus18"] = s
    # mixed: UEP short forms, EEP A/B all levels, DAB + DAB+ audio, stream data, packet data with FIG 0/3 and 0/14,
    # a 32-bit service id, secondary components, labels (type 1) and unknown extensions in between
    s = [
        fib(fig0(1, sub_short(1, 0, 5) + sub_short(2, 24, 14) + sub_short(3, 56, 33) + sub_short(4, 140, 34) + sub_long(5, 204, 0, 0, 96)),
            fig0(1, sub_long(6, 300, 1, 3, 27) + sub_long(7, 327, 1, 0, 54))),
        fib(fig_raw(1, bytes([0x01]) + bytes(range(21))),                       # a label FIG: skipped
            fig0(13, bytes(5))),                                                # unsupported extension: skipped
        fib(fig0(2, service(0xE1C0, [comp_audio(0, 1), comp_data(60, 6, primary=0)]) + service(0xE1C1, [comp_audio(63, 2)])),
            fig0(2, service(0xE1C2, [comp_audio(63, 3), comp_audio(63, 4, primary=0)]))),
        fib(fig0(2, service(0xE0D1C3C4, [comp_packet(0x123)], pd=1), pd=1), fig0(2, service(0xE1C5, [comp_data(5, 5)]))),
        fib(fig0(3, packet_def(0x123, 60, 7, 0x155) + packet_def(0x456, 5, 8, 3, caorg=0x1234)), fig0(14, bytes([(7 << 2) | 1, (9 << 2) | 2]))),
        fib(fig0(1, sub_long(8, 400, 0, 1, 8) + sub_long(9, 408, 0, 2, 12) + sub_short(10, 420, 63, table_switch=1))),
        fib(fig0(2, service(0xE1C6, [comp_audio(63, 8)]) + service(0xE1C7, [comp_audio(17, 9)]) + service(0xE1C8, [comp_audio(63, 10)]))),
    ]
    seqs["mixed"] = s
    # conflicts (a later different value must be ignored), a sub-channel switching form, truncated FIGs, invalid types
    s = [
        fib(fig0(1, sub_long(1, 0, 0, 2, 48))), fib(fig0(1, sub_long(1, 10, 1, 1, 60))), fib(fig0(1, sub_short(1, 0, 7))),
        fib(fig0(1, sub_short(2, 100, 20))), fib(fig0(1, sub_long(2, 100, 0, 2, 52))),
        fib(fig0(2, service(0x1001, [comp_audio(63, 1)]))), fib(fig0(2, service(0x1001, [comp_audio(0, 2)]))),
        fib(fig0(2, service(0x1002, [comp_data(60, 2)]))), fib(fig0(2, service(0x1002, [comp_audio(63, 2)]))),
        fib(bytes([0x1F]) + bytes(10), end=False),                   # length byte runs past the FIB
        fib(fig0(1, sub_long(3, 200, 0, 2, 48)[:3])),                # long form cut short
        fib(fig_raw(3, bytes(4)), fig0(1, sub_long(4, 300, 0, 2, 48))),   # invalid FIG type stops the FIB
        fib(fig_raw(7, bytes(2)), fig0(1, sub_long(5, 350, 0, 2, 48))),   # type 7 ends the FIB
        fib(fig0(2, service(0x1003, [bytes([(2 << 6) | 1, 4])]) + service(0x1004, [comp_audio(63, 6)]))),   # reserved TMId
        fib(fig0(1, sub_long(6, 500, 0, 3, 40))), fib(fig0(2, service(0x1004, [comp_audio(63, 6)]))),
        fib(fig_raw(6, bytes(3)), fig0(1, sub_long(7, 600, 1, 2, 54)), fig0(2, service(0x1005, [comp_audio(63, 7)]))),
    ]
    seqs["conflicts"] = s
    # random soup of the handled FIGs
    s = []
    for _ in range(160):
        figs, room = [], 30
        while room > 8:
            kind = rng.integers(0, 5)
            if kind == 0:
                f = fig0(1, b"".join(sub_long(int(rng.integers(0, 24)), int(rng.integers(0, 800)), int(rng.integers(0, 2)), int(rng.integers(0, 4)),
                                              int(rng.integers(4, 120))) for _ in range(int(rng.integers(1, 3)))))
            elif kind == 1:
                f = fig0(1, b"".join(sub_short(int(rng.integers(0, 24)), int(rng.integers(0, 800)), int(rng.integers(0, 64))) for _ in range(int(rng.integers(1, 3)))))
            elif kind == 2:
                comps = [comp_audio(int(rng.choice([0, 63, 63, 5])), int(rng.integers(0, 24)), int(rng.integers(0, 2))) if rng.random() < 0.7 else
                         comp_data(int(rng.choice([5, 24, 60, 63, 7])), int(rng.integers(0, 24)), int(rng.integers(0, 2))) for _ in range(int(rng.integers(1, 3)))]
                f = fig0(2, service(0x2000 + int(rng.integers(0, 12)), comps))
            elif kind == 3:
                f = fig0(2, service(0xE0000000 + int(rng.integers(0, 6)), [comp_packet(int(rng.integers(0, 8)), int(rng.integers(0, 2)))], pd=1), pd=1)
            else:
                f = fig0(3, packet_def(int(rng.integers(0, 8)), int(rng.choice([5, 60, 9])), int(rng.integers(0, 24)), int(rng.integers(0, 1024)))) \
                    if rng.random() < 0.6 else fig0(14, bytes([(int(rng.integers(0, 24)) << 2) | int(rng.integers(0, 4))]))
            if len(f) > room:
                break
            figs.append(f)
            room -= len(f)
        s.append(fib(*figs))
    seqs["random"] = s
    return seqs


def main():
    assert pyref.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    out = {"build_info": np.array(pyref.RefLib.get().build_info())}
    for name, fibs in sequences().items():
        r = pyref.RefFig()
        arr = np.stack(fibs)
        snaps = []
        for f in arr:
            r.process_fib(f)
        subs, comps = r.dump()
        out[name + "_fibs"] = arr
        out[name + "_subs"] = subs
        out[name + "_comps"] = comps
        print(name, arr.shape, "->", subs.shape[0], "sub-channels,", comps.shape[0], "components,", int(subs[:, 8].sum()), "complete")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fig_kat.npz"), **out)


if __name__ == "__main__":
    main()
