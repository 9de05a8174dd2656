"""The CUDA path against the UNMODIFIED reference binary (oracle/_ref/centrifuger -t <cores>) at the sizes of
BASELINE.json's configs -- not layout-vs-layout, and not the single-threaded oracle, which is too slow here:

  configs[1]   100 Mbp / 50 taxa index (built by the reference builder), 200 k x 100 bp single-end reads
  configs[2]   2 Gbp / 500 sequences, 50 k x 2x150 bp pairs, -k 5      } indexes generated and built on the GPU by
  configs[3]   20 Gbp / 5000 sequences, 20 k x 2x150 bp pairs, -k 5    } this repo's builder (seconds); the reference
                                                                        } binary loads the very same files

The rows compared are the reference's own TSV: the C-ABI results are formatted by cfr_format_tsv.  The 20 Gbp
case is the one that runs the 64-bit BWT positions, the 40-bit sector counters and the pair-line superblocks
with real high bits (n = 2 * 10^10 > 2^32 rows).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import centrifuger_b200 as cb

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = os.path.join(ROOT, "oracle", "_ref", "centrifuger")


def _reference_tsv(idx, k, files, tmp):
    out = os.path.join(tmp, "ref.tsv")
    cores = os.cpu_count() or 1
    with open(out, "wb") as fo:
        subprocess.run([REF, "-x", idx, "-k", str(k), "-t", str(cores)] + files, check=True, stdout=fo,
                       stderr=subprocess.DEVNULL)
    return open(out).read()


def _write_fastq(path, arr, suffix):
    import bench
    n, rl = arr.shape
    bench.write_fastq_sample(np.ascontiguousarray(arr).reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(rl), n, path,
                             suffix)


def _ours_tsv(clf, r1, r2):
    n, rl = r1.shape
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(rl)
    s1 = np.ascontiguousarray(r1).reshape(-1)
    s2 = np.ascontiguousarray(r2).reshape(-1) if r2 is not None else None
    res, ids = clf.classify_packed(s1, off, s2, off.copy() if r2 is not None else None)
    ids = ids.reshape(-1, clf.k)
    rows = [cb.TSV_HEADER]
    for i in range(n):
        rows.append(clf.format_tsv("r%d" % i, res[i], ids[i]))
    return "".join(rows)


def _first_difference(a, b):
    la, lb = a.splitlines(), b.splitlines()
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return "row %d:\n  ours %s\n  ref  %s" % (i, x, y)
    return "row counts %d vs %d" % (len(la), len(lb))


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/centrifuger not built")
def test_configs1_200k_reads_equal_the_reference_binary(tmp_path):
    import gen_data
    import make_data
    d = make_data.ensure("c2", log=lambda *a: None)
    if d is None:
        pytest.skip("data/c2 not available")
    genomes = make_data.genomes_of("c2")
    r1 = gen_data.make_reads_se_fast(genomes, 200_000, 100, seed=31)
    f1 = str(tmp_path / "r.fq")
    _write_fastq(f1, r1, "")
    idx = os.path.join(d, "idx")
    ref = _reference_tsv(idx, 1, ["-u", f1], str(tmp_path))
    clf = cb.Classifier(idx, k=1)
    ours = _ours_tsv(clf, r1, None)
    clf.close()
    assert ours == ref, _first_difference(ours, ref)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/centrifuger not built")
@pytest.mark.parametrize("name,species,pairs", [("configs2", 100, 50_000), ("configs3", 1000, 20_000)])
def test_gpu_built_collections_equal_the_reference_binary(tmp_path, name, species, pairs):
    from centrifuger_b200 import builder as B
    strains, glen = 5, 4_000_000
    big = species >= 1000
    # the 20 Gbp index (7.3 GB of files) is kept under data/ so that bench.py and the other tests of a run reuse it
    if big:
        import make_data
        d = make_data.ensure("c4", log=lambda *a: None)
        if d is None:
            pytest.skip("no GPU builder")
        prefix = os.path.join(d, "idx")
    else:
        prefix = str(tmp_path / "idx")
        B.build_synthetic(prefix, species, strains, glen)
    src = B.SyntheticReads(species, strains, glen)
    r1, r2, gi = src.pairs(pairs, 150, seed=41)
    f1, f2 = str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq")
    _write_fastq(f1, r1, "/1")
    _write_fastq(f2, r2, "/2")
    ref = _reference_tsv(prefix, 5, ["-1", f1, "-2", f2], str(tmp_path))
    clf = cb.Classifier(prefix, k=5)
    assert clf.info(22) == (64 if big else 32)
    ours = _ours_tsv(clf, r1, r2)
    n_rows = clf.n
    clf.close()
    assert n_rows == species * strains * glen
    assert ours == ref, _first_difference(ours, ref)
    # and the reads go where they came from: the species of a pair's source genome is among its assignments
    hit = 0
    lines = [ln.split("\t") for ln in ours.splitlines()[1:]]
    by_read = {}
    for f in lines:
        by_read.setdefault(f[0], []).append(f[1])
    for i in range(pairs):
        want = "seq%d_" % (int(gi[i]) // strains)
        if any(s.startswith(want) or s in ("species", "genus", "family") for s in by_read.get("r%d" % i, [])):
            hit += 1
    assert hit > 0.95 * pairs
