"""The C-ABI library loads and exports every symbol include/lfbm5d_cuda.h declares; without a GPU it fails loudly."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lfbm5d_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lfbm[35]d_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header():
    import lfbm5d_b200 as L
    lib = L.load_library()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(L.EXPORTS) == syms


def test_host_library_exports_and_noise(oracle):
    """include/lfbm5d_host_c.h: every declared symbol is exported; the product's mt19937ar noise and PSNR equal the oracle's
    restatement of utilities.cpp:154-185 / :412-435 (itself pinned to the reference) bit for bit."""
    import numpy as np
    import lfbm5d_b200 as L
    src = open(os.path.join(ROOT, "include", "lfbm5d_host_c.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(lfio_[A-Za-z0-9_]+)\s*\(", src)))
    h = L.load_host_library()
    assert syms == sorted(L.HOST_EXPORTS)
    for s in syms:
        assert hasattr(h, s), "missing export " + s
    clean = np.random.RandomState(3).rand(4, 3, 20, 24).astype(np.float32) * 255
    a, b = L.add_noise(clean, 25.0), oracle.add_noise(clean, 25.0)
    assert np.array_equal(a, b)
    assert L.psnr(a, clean) == oracle.psnr(a, clean)


def test_params_struct_layout():
    import lfbm5d_b200 as L
    assert C.sizeof(L.Params) == 20 * 4
    p = L.make_params(10.0, 2.7, 17, 17, 1, 1024, 1024, 3, 8, 18, 6, 16, 4, L.ID, L.SADCT, L.HAAR)
    assert (p.awidth, p.N, p.k, p.tau_4D, p.color_space, p.ang_major) == (17, 8, 16, 6, 2, 11)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import lfbm5d_b200 as L
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        L.LFBM5D(0)


def test_product_does_not_touch_oracle():
    """The product path must not import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "lfbm5d_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "common.cuh" or "oracle DCT mode" in txt, f
    out = os.popen("ldd %s" % os.path.join(pkg, "_lib", "liblfbm5d_cuda.so")).read()
    assert "oracle" not in out and "libref" not in out
