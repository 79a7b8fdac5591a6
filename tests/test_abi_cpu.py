"""The C-ABI library loads without a GPU and exports every symbol include/dabgpu.h declares; host-side logic."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dabgpu.h")).read()
    return sorted(set(re.findall(r"DABGPU_API\s+[\w\s\*]+?\b(dabgpu_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported(dab):
    names = _declared_symbols()
    assert len(names) >= 25
    L = dab.load_library()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert set(dab.EXPORTS) == set(names), set(dab.EXPORTS) ^ set(names)


def test_header_cites_reference_interfaces():
    src = open(os.path.join(ROOT, "include", "dabgpu.h")).read()
    for ref in ("ofdm_demodulator.h:109-141", "dab_viterbi_decoder.h:22-33", "fic_decoder.cpp:53-117", "msc_decoder.cpp:46-154",
                "aac_frame_processor.cpp:126-362", "reed_solomon_decoder.h:18-26"):
        assert ref in src


def test_params_and_errors_without_gpu(dab):
    p = dab.get_params(1)
    assert (p.nb_frame_bits, p.nb_fic_bits, p.nb_cif_bits, p.nb_frame_samples) == (230400, 9216, 55296, 196608)
    assert dab.get_params(3).nb_fib_group_bits == 3072
    with pytest.raises(dab.DabGpuError) as e:
        dab.get_params(5)                       # "Invalid transmission mode" (dab_ofdm_params_ref.cpp:53-54)
    assert e.value.code == dab.ERR_INVALID and "Invalid transmission mode" in str(e.value)
    L = dab.load_library()
    if L.dabgpu_device_count() == 0:
        with pytest.raises(dab.DabGpuError) as e2:
            dab.DabGpu(mode=1)
        assert e2.value.code == dab.ERR_CUDA and "no CPU fallback" in str(e2.value)


def test_product_never_references_the_oracle():
    """Nothing under the package (the product) may import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "sdrplusplus-dab-radio-plugin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyref" not in txt and "libdaboracle" not in txt and "libdabref" not in txt and "dab_oracle" not in txt, f


def _shard_worker(rank, world, port, n_streams, q):
    import torch.distributed as dist
    import importlib
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("sdrplusplus-dab-radio-plugin_b200.shard")
    mine = shard.streams_for_rank(n_streams, rank, world)
    owned = torch.zeros(n_streams, dtype=torch.int32)
    owned[mine] = 1
    dist.all_reduce(owned)                                        # reporting only: the data path has no collective
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, owned.tolist(), float(t.item()), len(mine)))
    dist.destroy_process_group()


def test_stream_sharding_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_streams, world = 1025, 2
    procs = [ctx.Process(target=_shard_worker, args=(r, world, 29581, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, owned, mx, n in res:
        assert owned == [1] * n_streams          # every stream owned exactly once
        assert mx == 513
    assert sorted(r[3] for r in res) == [512, 513]


def test_cpp_adapters_build_and_fail_loudly_without_gpu(dab):
    """host/dab_adapters.hpp (the reference's class surfaces over the C ABI) compiles with g++ -std=c++17 and links against
    libdabgpu.so; without a CUDA device the adapter constructors throw instead of computing anything on the CPU."""
    import importlib
    import subprocess
    import tempfile
    b = importlib.import_module("sdrplusplus-dab-radio-plugin_b200.build")
    b.build()
    exe = b.build_adapter_check()
    assert os.path.exists(exe)
    hdr = open(os.path.join(ROOT, "sdrplusplus-dab-radio-plugin_b200", "host", "dab_adapters.hpp")).read()
    for cls in ("class OFDM_Demod", "class DAB_Viterbi_Decoder", "class FIC_Decoder", "class MSC_Decoder", "class Reed_Solomon_Decoder",
                "class AAC_Frame_Processor", "class BasicRadio", "class Radio_Block"):
        assert cls in hdr
    if dab.load_library().dabgpu_device_count() == 0:
        with tempfile.TemporaryDirectory() as d:
            fin = os.path.join(d, "in.bin")
            open(fin, "wb").write(b"\x00" * 64)
            res = subprocess.run([exe, "rs", fin, os.path.join(d, "out.bin")], capture_output=True, text=True, timeout=60)
            assert res.returncode == 3 and "no CPU fallback" in res.stderr


def test_bench_reference_arm_line_and_no_cpu_fallback():
    """bench.py's contract as far as it can be checked without a GPU: `--impl reference` prints ONE JSON line with the keys the driver
    reads (same metric / unit / config as our arm, `impl`, `cpu_baseline` with kind / cores / sample and the per-core split, a zero-copy
    `e2e`), and our own arm refuses to run without a CUDA device instead of computing anything on the CPU."""
    import json
    import subprocess
    import sys
    bench = os.path.join(ROOT, "bench.py")
    res = subprocess.run([sys.executable, bench, "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-800:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "dab_mode1_iq_msps" and d["unit"] == "MS/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "full_chain_mode1_1024_streams_per_gpu" and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "fft_shim" in cb["sample"] or cb["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if cb["kind"] == "reference" and cb.get("per_core"):
        split = cb["per_core"]["ms_per_frame"]
        assert abs(split["ofdm"] + split["fic"] + split["msc"] + split["dabplus_rs"] - split["total"]) < 0.05 * split["total"]
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([sys.executable, bench, "--steps", "1"], capture_output=True, text=True, timeout=300)
        assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
