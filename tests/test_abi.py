"""The C-ABI shared library: loads, exports every symbol include/*.h declares,
and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

import centrifuger_b200 as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_in(header):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cfr_[a-z0-9_]+)\s*\(", hdr)))


def _declared_symbols():
    """every entry point declared by every header under include/"""
    out = []
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if h.endswith(".h"):
            out += _declared_in(h)
    return sorted(set(out))


def test_header_and_binding_agree():
    assert _declared_in("centrifuger_b200.h") == sorted(cb.ABI_SYMBOLS)
    assert _declared_in("centrifuger_b200_build.h") == sorted(cb.BUILD_ABI_SYMBOLS)
    assert _declared_symbols() == sorted(cb.ABI_SYMBOLS + cb.BUILD_ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = cb.load_library()
    for s in _declared_symbols():
        assert getattr(L, s) is not None, s


def test_default_params():
    L = cb.load_library()
    p = cb.Params()
    L.cfr_default_params(C.byref(p))
    assert (p.max_result, p.min_hit_len, p.max_result_per_hit_factor, p.dust) == (1, 0, 40, 1)
    assert p.consider_secondary_hit_len == 2000 and abs(p.consider_secondary_score_factor - 0.995) < 1e-12


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_open_without_gpu_fails_loudly(tiny_dir):
    with pytest.raises(cb.CfrError) as e:
        cb.Classifier(os.path.join(tiny_dir, "idx"))
    assert e.value.code == -5  # CFR_ERR_CUDA
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/"""
    pkg = os.path.join(ROOT, "centrifuger_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt, os.path.join(dp, f)
