"""The index builder against the reference's centrifuger-build.

Parity bar: the .1.cfr the builder writes is BYTE-IDENTICAL to the file the unmodified reference
builder wrote for the same genomes and options -- the committed tiny indexes (four option sets:
default, --rbbwt-b 1, --rbbwt-b 8, --offrate 3 / --ftabchars 5), the 10 Mbp `small` collection, and on
the GPU box the 100 Mbp configs[1] collection.  The taxonomy file (.2.cfr) is compared the same way.

CPU tests drive the builder's host twin (tests/buildsim_binding.py: the same passes as host loops);
`-m gpu` tests drive the sm_100a build through the C ABI.
"""
import filecmp
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import gen_data  # noqa: E402
import make_data  # noqa: E402
from centrifuger_b200 import builder as B  # noqa: E402

TINY_VARIANTS = {"idx": dict(ftabchars=6), "idx_b1": dict(ftabchars=6, rbbwt_b=1),
                 "idx_b8": dict(ftabchars=6, rbbwt_b=8), "idx_off3": dict(ftabchars=5, offrate=3)}


def _write_collection(d, name):
    genomes, nodes, names = gen_data.make_genomes(seed=1, **make_data.DATASETS[name]["genomes"])
    gen_data.write_reference(str(d), genomes, nodes, names)
    return str(d)


@pytest.fixture(scope="module")
def sim():
    import buildsim_binding
    return buildsim_binding.load()


@pytest.fixture(scope="module")
def tiny_collection(tmp_path_factory):
    return _write_collection(tmp_path_factory.mktemp("tinyref"), "tiny")


def _build(lib, ref_dir, out, **kw):
    return B.build_index(os.path.join(ref_dir, "ref.fa"), os.path.join(ref_dir, "nodes.dmp"),
                         os.path.join(ref_dir, "names.dmp"), os.path.join(ref_dir, "seqid.map"), out, lib=lib, **kw)


@pytest.mark.parametrize("variant", sorted(TINY_VARIANTS))
def test_tiny_index_is_byte_identical_to_the_reference_builders(sim, tiny_collection, tiny_dir, tmp_path, variant):
    out = str(tmp_path / "b")
    _build(sim, tiny_collection, out, **TINY_VARIANTS[variant])
    ref = os.path.join(tiny_dir, variant)
    assert filecmp.cmp(out + ".1.cfr", ref + ".1.cfr", shallow=False)
    assert filecmp.cmp(out + ".2.cfr", ref + ".2.cfr", shallow=False)
    assert filecmp.cmp(out + ".3.cfr", ref + ".3.cfr", shallow=False)


@pytest.mark.parametrize("batch_rows", [100_000, 7_000, 1_031])
def test_batched_sort_gives_the_same_file(sim, tiny_collection, tiny_dir, tmp_path, batch_rows):
    """the suffix array is sorted in batches cut by sampled splitter keys: any batch size, same file"""
    out = str(tmp_path / "b")
    st = _build(sim, tiny_collection, out, max_batch_rows=batch_rows, **TINY_VARIANTS["idx"])
    assert st.batches > 1
    assert filecmp.cmp(out + ".1.cfr", os.path.join(tiny_dir, "idx.1.cfr"), shallow=False)


def test_text_end_and_long_repeats(sim, tmp_path):
    """Suffixes that run into the end of the text (the end sorts below every base), a text ending in a run
    of A (the zero padding must not be taken for bases), repeats far longer than one 31-base key, a
    sequence shorter than ftabchars + 1 (dropped, Builder.hpp:145-151): checked against a plain
    suffix sort of the same text."""
    rng = np.random.default_rng(5)
    unit = rng.integers(0, 4, size=700, dtype=np.uint8)
    seqs = [np.concatenate([unit, unit, rng.integers(0, 4, size=300, dtype=np.uint8), unit]),
            np.concatenate([rng.integers(0, 4, size=500, dtype=np.uint8), np.zeros(90, dtype=np.uint8)]),
            np.concatenate([unit[:400], np.zeros(45, dtype=np.uint8)])]
    text = np.concatenate(seqs)
    n = len(text)
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    ids = np.arange(len(seqs), dtype=np.uint64)
    L = B._bind(sim)
    for batch in (0, 400):
        p = B._params(L, 2, 4, 0, 0, False, batch)
        out = str(tmp_path / ("t%d.1.cfr" % batch))
        assert L.cfr_build_fm_index(text.ctypes.data, n, lens.ctypes.data, ids.ctypes.data, len(lens), p, out.encode(),
                                    None) == 0, L.cfr_build_last_error()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import cfr_dump
        keep = {}
        fields = dict(cfr_dump.dump(out[:-6], keep))
        # plain suffix array: the end of the text is the smallest symbol
        tb = bytes(text + 1)
        sa = sorted(range(n), key=lambda i: tb[i:])
        bwt = np.array([text[i - 1] if i else text[n - 1] for i in sa], dtype=np.uint8)
        assert fields["firstISA"] == sa.index(0)
        assert fields["C"] == tuple(int(x) for x in np.concatenate([[0], np.cumsum(np.bincount(text, minlength=4))]))
        # decode the run-block BWT again through the dumped arrays
        b = fields["bwt.b"]
        typ = np.unpackbits(np.frombuffer(keep["bwt.useRunBlock"][0], dtype=np.uint8), bitorder="little")

        def wt_symbols(name, cnt):
            if cnt == 0:
                return np.zeros(0, dtype=np.uint8)
            root = np.unpackbits(np.frombuffer(keep[name + ".node0.v"][0], dtype=np.uint8), bitorder="little")[:cnt]
            lo = [np.unpackbits(np.frombuffer(keep[name + ".node%d.v" % k][0], dtype=np.uint8), bitorder="little")
                  if (name + ".node%d.v" % k) in keep else np.zeros(0, dtype=np.uint8) for k in (1, 2)]
            out_s, at = np.zeros(cnt, dtype=np.uint8), [0, 0]
            for i in range(cnt):
                h = int(root[i])
                out_s[i] = 2 * h + lo[h][at[h]]
                at[h] += 1
            return out_s
        plain = wt_symbols("bwt.waveletSeq", fields["bwt.waveletSeq.n"])
        runs = wt_symbols("bwt.runBlockSeq", fields["bwt.runBlockSeq.n"])
        got, ip, ir = [], 0, 0
        for k in range(fields["bwt.blockCnt"]):
            size = min(b, n - k * b)
            if typ[k]:
                got += [runs[ir]] * size
                ir += 1
            else:
                got += list(plain[ip:ip + size])
                ip += size
        assert np.array_equal(np.array(got, dtype=np.uint8), bwt)


def test_small_collection_is_byte_identical(sim, small_dir, tmp_path):
    ref = _write_collection(tmp_path, "small")
    out = str(tmp_path / "b")
    _build(sim, ref, out, max_batch_rows=3_000_000)
    for k in (1, 2, 3):
        assert filecmp.cmp("%s.%d.cfr" % (out, k), os.path.join(small_dir, "idx.%d.cfr" % k), shallow=False), k


def test_synthetic_generator_is_position_addressable(sim):
    L = B._bind(sim)
    a = B.synth_bases(3, 2, 1000, 500, lib=sim)
    b = B.synth_bases(3, 2, 1100, 300, lib=sim)
    assert np.array_equal(a[100:400], b)
    base = B.synth_bases(3, 0, 0, 200000, div_ppm=0, lib=sim)
    strain = B.synth_bases(3, 1, 0, 200000, div_ppm=10000, lib=sim)
    diff = float(np.mean(base != strain))
    assert 0.008 < diff < 0.012
    assert np.bincount(base, minlength=4).min() > 48000
    g = np.array([7, 16], dtype=np.uint64)
    o = np.array([5, 99], dtype=np.uint64)
    out = np.zeros((2, 40), dtype=np.uint8)
    L.cfr_synth_fragments(g.ctypes.data, o.ctypes.data, 2, 40, 5, 10000, 1, out.ctypes.data)
    assert np.array_equal(out[0], B.synth_bases(1, 2, 5, 40, lib=sim))
    assert np.array_equal(out[1], B.synth_bases(3, 1, 99, 40, lib=sim))


# ---------------------------------------------------------------------------- on the GPU
@pytest.mark.gpu
@pytest.mark.parametrize("variant", sorted(TINY_VARIANTS))
def test_gpu_tiny_index_is_byte_identical(tiny_collection, tiny_dir, tmp_path, variant):
    out = str(tmp_path / "b")
    _build(None, tiny_collection, out, **TINY_VARIANTS[variant])
    assert filecmp.cmp(out + ".1.cfr", os.path.join(tiny_dir, variant + ".1.cfr"), shallow=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name,batch_rows", [("small", 0), ("small", 1_500_000), ("c2", 0), ("c2", 30_000_000)])
def test_gpu_collection_is_byte_identical(tmp_path, name, batch_rows):
    d = make_data.ensure(name, log=lambda *a: None)
    if d is None:
        pytest.skip("data/%s not available" % name)
    ref = _write_collection(tmp_path, name)
    out = str(tmp_path / "b")
    st = _build(None, ref, out, max_batch_rows=batch_rows)
    assert (st.batches > 1) == (batch_rows > 0)
    for k in (1, 2, 3):
        assert filecmp.cmp("%s.%d.cfr" % (out, k), os.path.join(d, "idx.%d.cfr" % k), shallow=False), k


@pytest.mark.gpu
def test_gpu_synthetic_collection_classifies_its_own_reads(tmp_path):
    """A synthetic collection built on the device (the generator of the 20 / 140 Gbp workloads, at 40 Mbp):
    the reference binary loads the index, and reads drawn from the generator go to their own species."""
    import subprocess
    import centrifuger_b200 as cb
    sp, st, gl = 40, 5, 200_000
    prefix = str(tmp_path / "syn")
    B.build_synthetic(prefix, sp, st, gl, max_batch_rows=9_000_000)
    src = B.SyntheticReads(sp, st, gl)
    r1, r2, gi = src.pairs(4000, 150, seed=3)
    clf = cb.Classifier(prefix, k=5)
    res, ids = clf.classify([bytes(x) for x in r1], [bytes(x) for x in r2])
    ok = 0
    for i in range(len(gi)):
        n_as = int(res["n_assign"][i])
        if n_as == 0:
            continue
        tids = [int(ids[i][j]) if res["by_rank"][i] else clf.seq_taxid(int(ids[i][j])) for j in range(min(n_as, 5))]
        names = [clf.rank_name(t) for t in tids]  # noqa: F841
        want_species = int(gi[i]) // st
        got_species = {(int(ids[i][j]) // st) for j in range(min(n_as, 5))} if not res["by_rank"][i] else None
        if got_species is not None and want_species in got_species:
            ok += 1
    assert ok > 0.9 * len(gi)
    # the unmodified reference classifier agrees on the same index (TSV of 500 pairs)
    exe = os.path.join(ROOT, "oracle", "_ref", "centrifuger")
    if os.path.exists(exe):
        f1, f2 = str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq")
        gen_data.write_fastq(f1, r1[:500], suffix="/1")
        gen_data.write_fastq(f2, r2[:500], suffix="/2")
        ref_tsv = subprocess.run([exe, "-x", prefix, "-1", f1, "-2", f2, "-k", "5", "-t", "4"], check=True,
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        ours = clf.classify_tsv(["r%d" % i for i in range(500)], [bytes(x) for x in r1[:500]],
                                [bytes(x) for x in r2[:500]])
        assert ours == ref_tsv
    clf.close()
