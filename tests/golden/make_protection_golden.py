"""Generates tests/golden/protection_kat.npz from the REFERENCE's own MSC_Decoder (oracle/_ref/libdabref.so, built in place from
/root/reference by oracle/Makefile):  python tests/golden/make_protection_golden.py     (build container only)

Every protection profile the reference knows -- all 64 rows of UEP_PROTECTION_TABLE and EEP {1,2,3,4}-{A,B} at three sizes each,
plus the EEP 2-A n=1 special row and the "any type-A sub-channel of 8 CU" quirk (subchannel_protection_tables.h:21-170,
msc_decoder.cpp:77-154) -- is packed into a handful of CIF layouts.  The soft bits are NOT stored (6 MB of noise): the tests
regenerate them from the seeds below with protection_cases.soft_cif(); stored are the sub-channel table, the UEP table as the
reference lists it, and the bytes MSC_Decoder::DecodeCIF returned for the CIFs 15..19 of every sub-channel."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyref  # noqa: E402
import protection_cases as pc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert pyref.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    L = pyref.RefLib.get().L
    uep = np.zeros((64, 12), dtype=np.int32)
    for i in range(64):
        row = np.zeros(12, dtype=np.int32)
        assert L.ref_uep_descriptor(i, row.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int))) == 0
        uep[i] = row
    layouts = pc.build_layouts(uep[:, 0])
    table, expected = [], []
    for li, subs in enumerate(layouts):
        decs = [pyref.RefMsc(s["start"], s["length"], s["is_uep"], s["uep_index"], s["eep_level"], s["eep_type_b"]) for s in subs]
        outs = [[] for _ in subs]
        for c in range(pc.N_CIFS):
            cif = pc.soft_cif(li, c)
            for k, d in enumerate(decs):
                b = d.decode_cif(cif)
                assert (b.size > 0) == (c >= 15), (li, k, c, b.size)
                if b.size:
                    outs[k].append(b)
        for k, s in enumerate(subs):
            table.append([li, s["start"], s["length"], int(s["is_uep"]), s["uep_index"], s["eep_level"], int(s["eep_type_b"]), outs[k][0].size])
            expected.append(np.concatenate(outs[k]))
    offs = np.cumsum([0] + [e.size for e in expected]).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "protection_kat.npz"), uep_table=uep, subs=np.array(table, dtype=np.int32),
                        expected=np.concatenate(expected), offsets=offs,
                        build_info=np.frombuffer(pyref.RefLib.get().build_info().encode(), dtype=np.uint8))
    print(f"{len(table)} sub-channels in {len(layouts)} CIF layouts, {offs[-1]} expected bytes")


if __name__ == "__main__":
    main()
