"""debug: where do GPU and reference soft bits differ by more than one step?  (one recording of the robustness sweep)"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pyref
from test_robustness_full_gpu import _cases, N_FRAMES, BASE_LEAD, FFT, BLOCK
dab = importlib.import_module("sdrplusplus-dab-radio-plugin_b200")
tx = importlib.import_module("sdrplusplus-dab-radio-plugin_b200.synth.dabtx")
mode = int(sys.argv[1]); want = (float(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]))
cases = _cases(mode)
i = cases.index(want)
snr, cfo, lead = want
sub = tx.Subchannel(0, 0, 48, eep_level=2, dabplus=False)
ens = tx.EnsembleTx(mode, [sub], seed=1000 * mode + int(snr))
base = tx.ofdm_modulate([ens.next_frame_bits() for _ in range(N_FRAMES[mode])], mode)
x = tx.impair(base, snr, cfo / FFT[mode], BASE_LEAD[mode] + lead, seed=17 * i + mode, tail_samples=3000)
u8 = tx.to_u8(x, 30.0)
block = BLOCK[mode]
n = (u8.size // 2 // block) * block
o = pyref.RefOfdm(mode, 1)
g = dab.DabGpu(mode=mode, max_streams=1)
got = []
for off in range(0, n, block):
    o.process_u8(u8[2 * off:2 * (off + block)])
    g.ofdm_process(u8[None, 2 * off:2 * (off + block)], block_size=block)
    got += g.ofdm_pop_frames(0)
exp = o.pop_frames()
P = dab.get_params(mode)
K2 = 2 * P.nb_data_carriers
for k, (a, b) in enumerate(zip(got, exp)):
    d = np.abs(a[0].astype(np.int32) - b[0].astype(np.int32))
    bad = np.nonzero(d > 1)[0]
    print(f"frame {k}: coarse {a[1]:.9f}/{b[1]:.9f} fine {a[2]:.9e}/{b[2]:.9e} ft {a[3]}/{b[3]} n(|d|==1) {(d == 1).sum()} n(|d|>1) {bad.size} max {d.max()}")
    for p in bad[:12]:
        sym, pos = divmod(int(p), K2)
        print(f"    bit {p}: symbol {sym} pos {pos} (carrier slot {pos % P.nb_data_carriers}, {'im' if pos >= P.nb_data_carriers else 're'}) gpu {a[0][p]} ref {b[0][p]}  partner gpu {a[0][sym*K2 + (pos + P.nb_data_carriers) % K2]} ref {b[0][sym*K2 + (pos + P.nb_data_carriers) % K2]}")

# channel decode: GPU chain on its own soft bits vs the reference on the reference's
print("--- channel decode")
g3 = dab.DabGpu(mode=mode, max_streams=1); g3.msc_configure(0, [sub])
g4 = dab.DabGpu(mode=mode, max_streams=1); g4.msc_configure(0, [sub])
o_msc = pyref.RefMsc(sub.start_address, sub.length, sub.is_uep, sub.uep_index, sub.eep_level, sub.eep_type_b)
o_msc2 = pyref.RefMsc(sub.start_address, sub.length, sub.is_uep, sub.uep_index, sub.eep_level, sub.eep_type_b)
for k, (a, b) in enumerate(zip(got, exp)):
    g3.softbits_push(a[0][None, :]); g3.chan_decode(); out3, v3 = g3.get_msc(0, 0)
    g4.softbits_push(b[0][None, :]); g4.chan_decode(); out4, v4 = g4.get_msc(0, 0)
    for c in range(P.nb_cifs):
        e = o_msc.decode_cif(b[0][P.nb_fic_bits + c * P.nb_cif_bits:P.nb_fic_bits + (c + 1) * P.nb_cif_bits])
        e2 = o_msc2.decode_cif(a[0][P.nb_fic_bits + c * P.nb_cif_bits:P.nb_fic_bits + (c + 1) * P.nb_cif_bits])
        if e.size:
            print(f"frame {k} cif {c}: gpu(gpu soft) vs ref(ref soft) differing bytes {(out3[c] != e).sum()} ; gpu(ref soft) vs ref(ref soft) {(out4[c] != e).sum()} ; ref(gpu soft) vs ref(ref soft) {(e2 != e).sum()} of {e.size}")
