"""CPU: the C-ABI library loads and exports every symbol include/sml_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sml_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sml_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import sml_b200  # builds the library if needed
    lib = ctypes.CDLL(os.path.join(ROOT, "subspace-multimodal-learning_b200", "libsml_b200.so"))
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sml_b200._lib.lib.sml_version() >= 100


def test_gemm_desc_layout_matches_header():
    """ctypes mirror of sml_gemm_desc: field order / count follow the header."""
    import sml_b200
    src = open(os.path.join(ROOT, "include", "sml_b200.h")).read()
    body = re.search(r"typedef struct sml_gemm_desc \{(.*?)\} sml_gemm_desc;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = re.sub(r"^(const\s+)?(void|float|int32_t|int64_t)\s*\*?", "", decl)
        fields += [n.strip().lstrip("*").strip() for n in names.split(",")]
    assert fields == [f[0] for f in sml_b200._lib.GemmDesc._fields_]


def test_no_cpu_fallback():
    import pytest
    import torch
    import sml_b200
    with pytest.raises(RuntimeError, match="CUDA-only"):
        sml_b200.NystromAttention(dim=64, dim_head=8, heads=8, num_landmarks=16)(torch.randn(1, 20, 64))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "subspace-multimodal-learning_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def test_bench_reference_arm_does_not_load_the_product():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours) must not map libsml_b200.so or import sml_b200,
    and its JSON line carries the same config / steps / warm-up as our arm would"""
    import json
    import subprocess
    import sys
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--patches', '300', '--steps', '2', '--warmup', '1'];"
            "runpy.run_path(%r, run_name='__main__');"
            "assert 'sml_b200' not in sys.modules, 'product package imported by the reference arm';"
            "assert 'libsml_b200' not in open('/proc/self/maps').read(), 'libsml_b200.so mapped in the reference arm'"
            % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"]["slides_per_step_per_gpu"] == 8 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_a3_chunking_covers_every_tile():
    """host logic of the fused A3 path (nystrom.py:_a3_chunks / _a3_cfg): every chunk gets at least one key tile, the chunks
    cover the padded length, and the 64-bit batch stride survives the int32 pair"""
    import sml_b200
    from sml_b200.nystrom import _a3_cfg, _a3_chunks
    for b, h, npad in ((8, 8, 16640), (1, 8, 100864), (2, 8, 1280), (1, 1, 256), (64, 8, 16640)):
        for tile, slots, rows in ((64, 3 * 148, 4), (32, 148, 1)):
            n = _a3_chunks(b, h, npad, tile, slots, rows_per_cta=rows)
            ntot = npad // tile
            tpc = -(-ntot // n)
            assert n >= 1 and (n - 1) * tpc < ntot <= n * tpc, (b, h, npad, tile, n)
    cfg = _a3_cfg(64, 8, 256, 64, 100864 * 16, 3)          # batch stride 3 h d n' > 2^31
    bs = 100864 * 16 * 3 * 8 * 64
    assert (cfg[6] & 0xffffffff) | (cfg[7] << 32) == bs
