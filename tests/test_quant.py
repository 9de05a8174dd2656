"""Quantification against the reference's `centrifuger-quant` (oracle/_ref/centrifuger-quant,
the unmodified CentrifugerQuant.cpp / Quantifier.hpp): the abundance report must be byte-identical in all four
output formats, with and without --min-score / --min-length.

CPU tests: the host quantifier through the drop-in `centrifuger-b200-quant` on classification TSVs written by the
reference classifier.  `-m gpu` tests: no TSV at all -- the reads are classified on the GPU, the assignments are
coalesced on the device batch by batch (k_quant_keys + radix sorts), and cfr_quant_report prints the report."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
OURS = os.path.join(ROOT, "centrifuger_b200", "centrifuger-b200-quant")

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "centrifuger-quant")), reason="oracle/_ref/centrifuger-quant not built")


def _run(cmd):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout


def _classify_ref(idx, files, extra, out):
    with open(out, "wb") as f:
        f.write(_run([os.path.join(REF, "centrifuger"), "-x", idx, "-t", "4"] + files + extra))


@needs_ref
@pytest.mark.parametrize("k", [1, 3, 5])
def test_quant_cli_equals_centrifuger_quant(tiny_dir, small_dir, tmp_path, k):
    import centrifuger_b200.build as b
    b.build()
    cases = [(os.path.join(tiny_dir, "idx"), ["-1", os.path.join(tiny_dir, "pe_100_1.fq"), "-2", os.path.join(tiny_dir, "pe_100_2.fq")]),
             (os.path.join(tiny_dir, "idx_b8"), ["-u", os.path.join(tiny_dir, "se_100.fq")]),
             (os.path.join(small_dir, "idx"), ["-1", os.path.join(small_dir, "pe_150_1.fq"), "-2", os.path.join(small_dir, "pe_150_2.fq")])]
    for idx, files in cases:
        tsv = str(tmp_path / "c.tsv")
        _classify_ref(idx, files, ["-k", str(k)], tsv)
        for extra in ([], ["--min-score", "300"], ["--min-length", "60"]):
            for fmt in (0, 1, 2, 3):
                args = ["-x", idx, "-c", tsv, "--output-format", str(fmt)] + extra
                assert _run([OURS] + args) == _run([os.path.join(REF, "centrifuger-quant")] + args), (idx, k, extra, fmt)


@needs_ref
def test_quant_cli_reads_gzip_and_stdin(tiny_dir, tmp_path):
    import gzip
    idx = os.path.join(tiny_dir, "idx")
    tsv = str(tmp_path / "c.tsv")
    _classify_ref(idx, ["-u", os.path.join(tiny_dir, "se_100.fq")], ["-k", "2"], tsv)
    want = _run([os.path.join(REF, "centrifuger-quant"), "-x", idx, "-c", tsv])
    with gzip.open(tsv + ".gz", "wb") as f:
        f.write(open(tsv, "rb").read())
    assert _run([OURS, "-x", idx, "-c", tsv + ".gz"]) == want
    p = subprocess.run([OURS, "-x", idx, "-c", "-"], input=open(tsv, "rb").read(), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                       check=True)
    assert p.stdout == want


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("k", [1, 5])
def test_device_coalescing_report_equals_centrifuger_quant(small_dir, tmp_path, k):
    import centrifuger_b200 as cb
    from oracle_binding import read_fastx
    idx = os.path.join(small_dir, "idx")
    f1, f2 = os.path.join(small_dir, "pe_150_1.fq"), os.path.join(small_dir, "pe_150_2.fq")
    _, r1 = read_fastx(f1)
    _, r2 = read_fastx(f2)
    tsv = str(tmp_path / "c.tsv")
    _classify_ref(idx, ["-1", f1, "-2", f2], ["-k", str(k)], tsv)
    for extra, kw in (([], {}), (["--min-score", "400", "--min-length", "70"], dict(min_score=400, min_hit_length=70))):
        clf = cb.Classifier(idx, k=k, max_batch_reads=3000)  # several device chunks per call
        clf.quant_enable(**kw)
        for lo in range(0, len(r1), 7000):  # and several calls
            clf.classify(r1[lo:lo + 7000], r2[lo:lo + 7000])
        st = clf.quant_stats()
        assert st["batches"] >= 3 and 0 < st["distinct_records"] < len(r1) / 4
        for fmt in (0, 1, 2, 3):
            out = str(tmp_path / ("q%d.txt" % fmt))
            clf.quant_report(idx, out, fmt)
            want = _run([os.path.join(REF, "centrifuger-quant"), "-x", idx, "-c", tsv, "--output-format", str(fmt)] + extra)
            assert open(out, "rb").read() == want, (k, extra, fmt)
        clf.close()


@pytest.mark.gpu
@needs_ref
def test_cli_quant_report_flag(small_dir, tmp_path):
    idx = os.path.join(small_dir, "idx")
    f1 = os.path.join(small_dir, "se_100.fq")
    cli = os.path.join(ROOT, "centrifuger_b200", "centrifuger-b200")
    rep = str(tmp_path / "rep.txt")
    tsv = subprocess.run([cli, "-x", idx, "-u", f1, "-k", "3", "--quant-report", rep, "--quant-format", "3", "--batch", "4000"],
                         check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    t = str(tmp_path / "c.tsv")
    open(t, "wb").write(tsv)
    want = _run([os.path.join(REF, "centrifuger-quant"), "-x", idx, "-c", t, "--output-format", "3"])
    assert open(rep, "rb").read() == want
