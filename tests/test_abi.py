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


def _encode_reference(s1, o1, s2, o2):
    """numpy restatement of the device's k_encode over the batch buffer (mate 1 from 0, mate 2 from the next
    multiple of 32): 2-bit codes, N bits (padding reads as N)"""
    import numpy as np
    len1 = int(o1[-1] - o1[0])
    len2 = int(o2[-1] - o2[0]) if o2 is not None else 0
    pos2 = (len1 + 31) & ~31
    n_words = (pos2 + len2) // 32 + 2
    buf = np.full(n_words * 32, ord("N"), dtype=np.uint8)
    buf[:len1] = s1[int(o1[0]):int(o1[0]) + len1]
    if len2:
        buf[pos2:pos2 + len2] = s2[int(o2[0]):int(o2[0]) + len2]
    lut = np.full(256, 4, dtype=np.int64)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    x = lut[buf].reshape(-1, 32)
    nmask = ((x > 3).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
    codes = (np.where(x > 3, 0, x).astype(np.uint64) << (2 * np.arange(32, dtype=np.uint64))).sum(axis=1).astype(np.uint64)
    return codes, nmask, pos2


@pytest.mark.parametrize("threads", [1, 5])
def test_pack_reads_matches_the_encode_stage(threads):
    """cfr_pack_reads (the host side of cfr_submit_packed) writes what k_encode writes: ragged and empty reads,
    lowercase / IUPAC / N bytes, single-end, an empty batch, a batch that starts in the middle of its buffers"""
    import numpy as np
    rng = np.random.default_rng(5 + threads)
    alpha = list(b"ACGTNacgtRY-") + [0, 255, 0x40, 0x42, 0x55, 0xc1]  # neighbours of the four letters, bytes with the high bit
    p = [.22, .22, .22, .22, .02, .01, .01, .01, .01, .01, .01, .01] + [.005] * 6
    for n, hi, two in ((300, 200, True), (1, 1, True), (0, 1, True), (40000, 160, False), (500, 40, True)):
        r1 = [bytes(rng.choice(alpha, p=p, size=int(L)).astype(np.uint8)) for L in rng.integers(0, hi, size=n)]
        r2 = [bytes(rng.choice(alpha, p=p, size=int(L)).astype(np.uint8)) for L in rng.integers(0, hi, size=n)]
        s1, o1 = cb.pack_reads([b"GATTACA"] + r1)  # the batch proper starts at offset 7 of the caller's buffer
        s2, o2 = cb.pack_reads(r2) if two else (None, None)
        o1 = o1[1:]
        pk, (codes, nmask, p1, p2) = cb.pack_batch(s1, o1, s2, o2, threads=threads)
        ec, em, pos2 = _encode_reference(s1, o1, s2, o2)
        assert pk.n_words == len(ec) == len(codes) and pk.n_reads == n
        assert (codes == ec).all() and (nmask == em).all()
        assert (p1 == o1 - o1[0]).all()
        if two:
            assert (p2 == o2 + np.uint64(pos2)).all()
        else:
            assert pk.off2 is None


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_patched_reference_binary_links_and_fails_loudly_without_a_gpu(tiny_dir):
    """INTEGRATION.md, Option 2, applied for real (oracle/patch_reference.py): the reference's own CentrifugerClass.cpp with its
    batch loop calling libcfrb200.so builds against the unmodified reference headers; without a device it exits with the
    library's error instead of classifying on the CPU"""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "centrifuger_patched")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/centrifuger_patched not built (needs /root/reference)")
    r = subprocess.run([exe, "-x", os.path.join(tiny_dir, "idx"), "-u", os.path.join(tiny_dir, "se_100.fq")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0
    assert "no CUDA device" in r.stderr.decode() and "no CPU path" in r.stderr.decode()
