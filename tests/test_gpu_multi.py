"""More than one GPU behind the drop-in boundary: `centrifuger-b200 --gpus G` replicates the index per GPU, sends
batch j to GPU j mod G, keeps the rows in input order (CentrifugerClass.cpp:674-691) and sums the per-taxon
counters of the replicas over NCCL once at the end (cfr_counts_allreduce_local).  Needs >= 2 devices
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a single-GPU box."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "centrifuger_b200", "centrifuger-b200")


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(args):
    p = subprocess.run([CLI] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    return p.stdout, p.stderr.decode()


@pytest.mark.skipif(_device_count() < 2, reason="needs at least two GPUs")
def test_cli_on_several_gpus_writes_the_same_rows(small_dir):
    idx = os.path.join(small_dir, "idx")
    f1, f2 = os.path.join(small_dir, "pe_150_1.fq"), os.path.join(small_dir, "pe_150_2.fq")
    base = ["-x", idx, "-1", f1, "-2", f2, "-k", "5", "--batch", "1500"]  # 20 000 pairs -> 14 batches
    one, _ = _run(base)
    ref = os.path.join(ROOT, "oracle", "_ref", "centrifuger")
    if os.path.exists(ref):
        want = subprocess.run([ref, "-x", idx, "-1", f1, "-2", f2, "-k", "5", "-t", "4"], stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL, check=True).stdout
        assert one == want
    for g in sorted({2, _device_count()}):
        out, err = _run(base + ["--gpus", str(g)])
        assert out == one, "rows differ with --gpus %d" % g
        assert "Reduced the per-taxon counters of %d GPUs over NCCL: 20000 reads" % g in err, err
        assert "WARNING" not in err


@pytest.mark.skipif(_device_count() < 2, reason="needs at least two GPUs")
def test_counts_allreduce_local_sums_the_replicas(small_dir):
    import ctypes as C

    import numpy as np

    import centrifuger_b200 as cb
    from oracle_binding import read_fastx
    idx = os.path.join(small_dir, "idx")
    _, r1 = read_fastx(os.path.join(small_dir, "se_100.fq"))
    a = cb.Classifier(idx, device=0)
    b = cb.Classifier(idx, device=1)
    a.classify(r1[:3000])
    b.classify(r1[3000:8000])
    whole = cb.Classifier(idx, device=0)
    whole.classify(r1[:8000])
    want = whole.taxon_counts()
    n = len(want)
    out = np.zeros(n, dtype=np.uint64)
    hs = (C.c_void_p * 2)(a.h, b.h)
    st = a.L.cfr_counts_allreduce_local(hs, 2, out.ctypes.data_as(C.c_void_p), C.c_uint64(n))
    assert st == 0, a.L.cfr_last_error()
    assert np.array_equal(out, want)
    assert int(a.taxon_counts()[a.node_cnt + 1]) == 3000  # the live counters stay per replica
    for c in (a, b, whole):
        c.close()
