"""Host-side ingest of the drop-in CLI (no GPU): FASTA/FASTQ parsing with kseq semantics,
gz input, CRLF, multi-line FASTA, /1 /2 suffix stripping, interleaved input, multiple files."""
import gzip
import os
import subprocess

import pytest

import centrifuger_b200 as cb
from oracle_binding import read_fastx

EXE = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")


def _dry(args):
    r = subprocess.run([EXE, "--dry-run"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    return [tuple(l.split("\t")) for l in r.stdout.decode().split("\n") if l]


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(EXE):
        from centrifuger_b200 import build
        build.build()


def test_fastq_pairs_match_reference_reader(tiny_dir):
    f1, f2 = os.path.join(tiny_dir, "pe_100_1.fq"), os.path.join(tiny_dir, "pe_100_2.fq")
    ids, r1 = read_fastx(f1)
    _, r2 = read_fastx(f2)
    got = _dry(["-1", f1, "-2", f2])
    assert got == [(i, a.decode(), b.decode()) for i, a, b in zip(ids, r1, r2)]


def test_formats(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">s1 comment here\nACGT\nACG\n>s2/1\r\nTTTT\r\n>s3\n\n>s4\nGG")
    assert _dry(["-u", str(fa)]) == [("s1", "ACGTACG", ""), ("s2", "TTTT", ""), ("s3", "", ""), ("s4", "GG", "")]
    fq = tmp_path / "q.fq.gz"
    with gzip.open(fq, "wb") as f:
        f.write(b"@r1/1 x\nACGTN\n+\n@@@+I\n@r2\nAC\n+r2\n+@\n")
    assert _dry(["-u", str(fq)]) == [("r1", "ACGTN", ""), ("r2", "AC", "")]
    # two files back to back, and interleaved input
    assert [x[0] for x in _dry(["-u", str(fa), "-u", str(fq)])] == ["s1", "s2", "s3", "s4", "r1", "r2"]
    inter = tmp_path / "i.fq"
    inter.write_bytes(b"@p1/1\nAAAA\n+\nIIII\n@p1/2\nCCCC\n+\nIIII\n@p2/1\nGG\n+\nII\n@p2/2\nTT\n+\nII\n")
    assert _dry(["-i", str(inter)]) == [("p1", "AAAA", "CCCC"), ("p2", "GG", "TT")]


def test_mate_count_mismatch_is_an_error(tmp_path):
    a = tmp_path / "a.fq"
    b = tmp_path / "b.fq"
    a.write_bytes(b"@x\nAC\n+\nII\n@y\nAC\n+\nII\n")
    b.write_bytes(b"@x\nAC\n+\nII\n")
    r = subprocess.run([EXE, "--dry-run", "-1", str(a), "-2", str(b)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"different number of reads" in r.stderr


def test_usage_and_version_and_rejections():
    r = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and b"-x FILE: index prefix" in r.stderr  # CentrifugerClass.cpp:347-351
    assert subprocess.run([EXE, "-v"], stdout=subprocess.PIPE).stdout.decode().strip() == "Centrifuger v1.1.3-r347"
    r = subprocess.run([EXE, "-x", "nope", "-u", "x.fq", "--merge-readpair"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"not supported" in r.stderr
    r = subprocess.run([EXE, "-u", "x.fq"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"Need to use -x" in r.stderr
