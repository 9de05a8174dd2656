"""Host logic: the product's per-task stage functions (cfr_core.cuh /
cfr_pipeline.cuh, compiled for the host by tests/hostsim) and its .cfr parser
against the oracle.  No GPU needed; the CUDA path is checked by test_gpu_*.py."""
import os
import random

import pytest

from hostsim_binding import HostSim, hostsim_dust, hostsim_dust_screen, result_tuples
from oracle_binding import Oracle, dust_mask, read_fastx


def _compare(idx, files, layout, arena_rows=0, limit=None, check_counters=True, **kw):
    _, r1 = read_fastx(files[0])
    r2 = read_fastx(files[1])[1] if len(files) == 2 else None
    if limit:
        r1 = r1[:limit]
        r2 = r2[:limit] if r2 else None
    o = Oracle(idx, **kw)
    hs = HostSim(idx, layout=layout, **kw)
    try:
        o.reset_counters()
        exp = []
        for i in range(len(r1)):
            t = o.result_tuple(o.query(r1[i], r2[i] if r2 else None))
            exp.append(t[:7])
        res, ids, cnt = hs.classify(r1, r2, arena_rows=arena_rows)
        got = result_tuples(res, ids, hs.stride)
        assert got == exp
        oc = o.counters()
        for k in ("n_rank", "n_access", "n_search", "n_locate", "n_lf", "n_extend"):
            if check_counters:
                assert oc[k] == cnt[k], k
            elif k in ("n_rank", "n_extend", "n_lf"):
                assert cnt[k] <= oc[k], k  # the load-time tables only remove BackwardExtend / LF steps
    finally:
        hs.close()
        o.close()


@pytest.mark.parametrize("variant", ["idx", "idx_b1", "idx_b8", "idx_off3"])
def test_pair_layout_two_steps_equal_two_literal_extends(tiny_dir, variant):
    """The pair lines (two BackwardExtend steps per 128-byte line, cfr_core.cuh) against the literal
    steps: every BWT row as a single-row range and as the low end of short ranges, 20 000 random
    ranges (wide, narrow, around firstISA), every symbol pair and the one-symbol form; results, stop
    tests and operation counters must agree."""
    hs = HostSim(os.path.join(tiny_dir, variant), layout=4)
    try:
        assert hs.pair_check(20000, 7) == 0
    finally:
        hs.close()


@pytest.mark.parametrize("layout", [1, 2, 3, 4])  # 3 = occ sectors walked with 64-bit positions, 4 = pair lines in the search
@pytest.mark.parametrize("variant", ["idx", "idx_b1", "idx_b8", "idx_off3"])
def test_tiny_all_read_sets(tiny_dir, layout, variant):
    idx = os.path.join(tiny_dir, variant)
    sets = (["se_100.fq"], ["pe_100_1.fq", "pe_100_2.fq"], ["edge.fq"], ["edge_1.fq", "edge_2.fq"])
    for files in sets:
        fs = [os.path.join(tiny_dir, f) for f in files]
        for kw in (dict(), dict(k=5), dict(k=3, hitk_factor=2), dict(dust=False, min_hit_len=16),
                   dict(k=2, hitk_factor=0)):
            _compare(idx, fs, layout, **kw)


@pytest.mark.parametrize("layout", [1, 2, 4])
def test_unlimited_results_k0(tiny_dir, layout):
    """-k 0 (and negative -k): every best-scoring sequence is reported, nothing is reduced by rank, every row of a
    hit is resolved (Classifier.hpp:620-623, :784-785)"""
    idx = os.path.join(tiny_dir, "idx")
    for files in (["se_100.fq"], ["pe_100_1.fq", "pe_100_2.fq"], ["edge.fq"], ["se_com.fq"]):
        fs = [os.path.join(tiny_dir, f) for f in files]
        for kw in (dict(k=0), dict(k=-3, hitk_factor=2), dict(k=0, dust=False, min_hit_len=16)):
            _compare(idx, fs, layout, **kw)


@pytest.mark.parametrize("layout", [1, 2])
def test_small_arena_deferral(tiny_dir, layout):
    """a locate arena much smaller than the batch needs forces the multi-pass path"""
    idx = os.path.join(tiny_dir, "idx")
    fs = [os.path.join(tiny_dir, "pe_100_1.fq"), os.path.join(tiny_dir, "pe_100_2.fq")]
    _compare(idx, fs, layout, arena_rows=300, k=5)


@pytest.mark.parametrize("layout", [1, 2, 3])
@pytest.mark.parametrize("width", [7, 9])
def test_wide_lookup_table_same_answers(tiny_dir, layout, width, monkeypatch):
    """the wide lookup table (one probe instead of the W-mer probe + the first extends) changes no result:
    reads with Ns, reads shorter than the table width, searches that end inside the table"""
    monkeypatch.setenv("HOSTSIM_WIDE_LOOKUP", str(width))
    for variant in ("idx", "idx_off3", "idx_b1"):
        idx = os.path.join(tiny_dir, variant)
        for files in (["se_100.fq"], ["pe_100_1.fq", "pe_100_2.fq"], ["edge.fq"], ["edge_1.fq", "edge_2.fq"]):
            fs = [os.path.join(tiny_dir, f) for f in files]
            for kw in (dict(), dict(k=5), dict(dust=False, min_hit_len=16)):
                _compare(idx, fs, layout, check_counters=False, **kw)


@pytest.mark.parametrize("layout", [1, 2, 3])
@pytest.mark.parametrize("shift", [0, 1, 2, 3])
def test_dense_locate_table_same_answers(tiny_dir, layout, shift, monkeypatch):
    """the dense locate table (answers of the reference's own walk, stored for every 2^shift-th row)
    changes no result; idx_off3 samples every 8th row, the others every 16th"""
    monkeypatch.setenv("HOSTSIM_DENSE_LOCATE", str(shift))
    monkeypatch.setenv("HOSTSIM_DENSE16", str(shift & 1))  # 16-bit entries at the odd spacings
    for variant in ("idx", "idx_off3", "idx_b8"):
        idx = os.path.join(tiny_dir, variant)
        for files in (["se_100.fq"], ["pe_100_1.fq", "pe_100_2.fq"], ["edge_1.fq", "edge_2.fq"]):
            fs = [os.path.join(tiny_dir, f) for f in files]
            for kw in (dict(), dict(k=5), dict(k=3, hitk_factor=2)):
                _compare(idx, fs, layout, check_counters=False, **kw)


def test_reduce_taxids_fuzz(tiny_dir, example_idx):
    """Taxonomy::ReduceTaxIds / LCA on random id sets (siblings, mixed depths, root-level and
    out-of-range ids, duplicates) against the oracle"""
    rng = random.Random(29)
    for idx in (os.path.join(tiny_dir, "idx"), example_idx):
        o = Oracle(idx)
        hs = HostSim(idx)
        nodes = o.scalar(10)
        for it in range(3000):
            cnt = rng.randint(2, 9)
            mode = rng.random()
            if mode < 0.5:
                ids = [rng.randrange(nodes) for _ in range(cnt)]
            elif mode < 0.8:  # few distinct values: ties, siblings, duplicates
                pool = [rng.randrange(nodes) for _ in range(3)]
                ids = [rng.choice(pool) for _ in range(cnt)]
            else:  # now and then an id the taxonomy does not know
                ids = [rng.randrange(nodes + 2) for _ in range(cnt)]
            for k in (1, 2, 5):
                if cnt <= k:
                    continue
                assert hs.reduce_taxids(ids, k) == o.reduce_taxids(ids, k), (ids, k)
        hs.close()
        o.close()


def test_expand_taxids_fuzz(tiny_dir, example_idx):
    """the child lists of ReduceTaxIds / LCA (--expand-taxid) on random id sets against the oracle"""
    rng = random.Random(31)
    for idx in (os.path.join(tiny_dir, "idx"), example_idx):
        o = Oracle(idx)
        hs = HostSim(idx)
        nodes = o.scalar(10)
        non_empty = 0
        for it in range(3000):
            cnt = rng.randint(2, 9)
            mode = rng.random()
            if mode < 0.5:
                ids = [rng.randrange(nodes) for _ in range(cnt)]
            elif mode < 0.8:
                pool = [rng.randrange(nodes) for _ in range(3)]
                ids = [rng.choice(pool) for _ in range(cnt)]
            else:
                ids = [rng.randrange(nodes + 2) for _ in range(cnt)]
            for k in (1, 2, 3, 5):
                if cnt <= k:
                    continue
                exp_ids, exp_lists = o.reduce_taxids_expanded(ids, k)
                got_ids, got_lists = hs.expand_taxids(ids, k)
                assert got_ids == exp_ids, (ids, k)
                # lists the classifier would not print count as empty on both sides
                assert got_lists == (exp_lists if exp_lists else [[] for _ in exp_ids]), (ids, k)
                non_empty += any(len(l) for l in got_lists)
        assert non_empty > 1000
        hs.close()
        o.close()


@pytest.mark.parametrize("layout", [1, 2])
def test_expand_taxid_pipeline(tiny_dir, layout):
    """--expand-taxid through the whole pipeline: results unchanged, child lists equal to the oracle's"""
    idx = os.path.join(tiny_dir, "idx")
    for files in (["se_100.fq"], ["pe_100_1.fq", "pe_100_2.fq"], ["edge_1.fq", "edge_2.fq"]):
        fs = [os.path.join(tiny_dir, f) for f in files]
        _, r1 = read_fastx(fs[0])
        r2 = read_fastx(fs[1])[1] if len(fs) == 2 else None
        for kw, arena in ((dict(), 0), (dict(k=2), 0), (dict(k=3), 300), (dict(k=4, hitk_factor=0), 0)):
            o = Oracle(idx, **kw)
            hs = HostSim(idx, layout=layout, **kw)
            k = hs.p.max_result
            res, ids, lists = hs.classify_expanded(r1, r2, arena_rows=arena)
            got = result_tuples(res, ids, k)
            seen = 0
            for i in range(len(r1)):
                ores, child, cnt = o.query_expanded(r1[i], r2[i] if r2 else None)
                assert got[i] == o.result_tuple(ores)[:7]
                exp, at = [], 0
                for j in range(min(ores.n, k)):
                    exp.append([int(x) for x in child[at:at + cnt[j]]])
                    at += cnt[j]
                assert lists[i] == exp, (files, kw, i)
                seen += any(len(l) for l in exp)
            assert seen > 0
            hs.close()
            o.close()


@pytest.mark.parametrize("layout", [1, 2, 3])
def test_long_reads_and_consider_secondary(tiny_dir, layout):
    """2 - 9 kbp reads (hundreds of hits per strand, hit lengths past the 2000-base bar) and the
    near-tie rule with lowered bars on the short read sets"""
    from conftest import golden_path
    long_fa = [golden_path("tiny", "long.fa")]
    idx = os.path.join(tiny_dir, "idx")
    for kw in (dict(), dict(k=5), dict(k=2, secondary_len=1000, secondary_factor=0.1),
               dict(dust=False, secondary_len=3000, secondary_factor=0.05)):
        _compare(idx, long_fa, layout, **kw)
    _compare(idx, long_fa, layout, arena_rows=700, k=2, secondary_len=1000, secondary_factor=0.1)
    for files, kw in ((["se_100.fq"], dict(secondary_len=50, secondary_factor=0.9)),
                      (["se_100.fq"], dict(k=3, secondary_len=60, secondary_factor=0.8)),
                      (["pe_100_1.fq", "pe_100_2.fq"], dict(secondary_len=100, secondary_factor=0.9)),
                      (["pe_100_1.fq", "pe_100_2.fq"], dict(k=5, secondary_len=50, secondary_factor=0.5)),
                      (["edge_1.fq", "edge_2.fq"], dict(k=2, secondary_len=30, secondary_factor=0.7))):
        _compare(idx, [os.path.join(tiny_dir, f) for f in files], layout, **kw)


def test_example(example_idx):
    from conftest import golden_path
    fs = [golden_path("example", "example_1.fq"), golden_path("example", "example_2.fq")]
    for layout in (1, 2, 3):
        _compare(example_idx, fs, layout)
        _compare(example_idx, fs[:1], layout, k=5, dust=False)


def test_dust_compressed_interval_table():
    """the 64-slot replacement of the reference's perfect-interval vector is exact"""
    rng = random.Random(5)
    tests = [b"A" * 100, b"AC" * 50, b"ACG" * 40, b"AAAC" * 30, b"N" * 10 + b"ACGT" * 10 + b"A" * 80,
             b"acgt" * 30, b"AAGG" * 60, b"ACACG" * 80, b"A" * 70 + b"N" * 70 + b"A" * 70, b"AC", b"", b"ACG",
             b"A" * 30 + b"N" * 66 + b"C" * 30]
    for _ in range(1500):
        L = rng.choice([30, 64, 65, 100, 150, 300, 700])
        mode = rng.random()
        if mode < 0.3:
            s = bytes(rng.choice(b"ACGT") for _ in range(L))
        elif mode < 0.6:
            per = rng.randint(1, 8)
            unit = bytes(rng.choice(b"ACGT") for _ in range(per))
            s = bytes(unit[i % per] if rng.random() > 0.05 else rng.choice(b"ACGTN") for i in range(L))
        else:
            s = bytes(rng.choice(b"AAAAAAACGTN") for _ in range(L))
        tests.append(s)
    for s in tests:
        assert hostsim_dust(s) == dust_mask(s), s


def test_dust_screen_is_safe():
    """a read the register-only screen clears is never masked by the reference's SDUST, and the
    screen clears the bulk of random reads (that is its point)"""
    rng = random.Random(17)
    tests = [b"A" * 7, b"A" * 6, b"ACACACACACAC", b"ACACACACACA", b"ACGACGACGACGACGAC", b"ACGACGACGACGACGA",
             b"ACGT" * 5 + b"A", b"ACGTA" * 5 + b"AC", b"ACGTAC" * 5 + b"AC", b"ACGTACG" * 5 + b"AC", b"", b"AC",
             b"ACGTTGCATGCATGACGATCGATCGTAGCTAGCTAGCTGATCGATGCATCGA" * 3]
    n_random = n_cleared = 0
    for it in range(6000):
        L = rng.choice([7, 8, 12, 30, 63, 64, 65, 66, 95, 96, 97, 100, 128, 150, 151, 300])
        mode = rng.random()
        if mode < 0.5:
            s = bytes(rng.choice(b"ACGT") for _ in range(L))
            n_random += 1
            n_cleared += 0 if hostsim_dust_screen(s) else 1
        elif mode < 0.8:  # periodic with noise: right at the masking threshold
            per = rng.randint(1, 13)
            unit = bytes(rng.choice(b"ACGT") for _ in range(per))
            s = bytes(unit[i % per] if rng.random() > 0.04 else rng.choice(b"ACGT") for i in range(L))
        elif mode < 0.9:  # a short repeat planted in a random read
            s = bytearray(rng.choice(b"ACGT") for _ in range(L))
            per = rng.randint(1, 6)
            unit = bytes(rng.choice(b"ACGT") for _ in range(per))
            rl = rng.randint(5, 30)
            p0 = rng.randrange(max(1, L - rl))
            for i in range(p0, min(L, p0 + rl)):
                s[i] = unit[(i - p0) % per]
            s = bytes(s)
        else:
            s = bytes(rng.choice(b"AAAAAACGT") for _ in range(L))
        tests.append(s)
    for s in tests:
        if not hostsim_dust_screen(s):
            assert dust_mask(s) == s, s
    assert n_cleared > 0.7 * n_random, (n_cleared, n_random)


@pytest.mark.parametrize("layout", [1, 2])
def test_rank_access_locate_primitives(tiny_dir, layout):
    rng = random.Random(11)
    for v in ("idx", "idx_b1", "idx_b8"):
        idx = os.path.join(tiny_dir, v)
        o = Oracle(idx)
        hs = HostSim(idx, layout=layout)
        n = o.n
        pos = [0, 1, n - 1, n // 2] + [rng.randrange(n) for _ in range(400)]
        for p in pos:
            assert hs.bwt_access(p) == o.bwt_access(p)
            for c in "ACGT":
                assert hs.bwt_rank(c, p, 1) == o.bwt_rank(c, p, 1)
                assert hs.bwt_rank(c, p, 0) == o.bwt_rank(c, p, 0)
            assert hs.locate(p) == o.locate(p)[0]
        hs.close()
        o.close()


def test_truncated_index_files_are_refused(tiny_dir, tmp_path):
    """the product's .cfr parser (cfr_format.cpp, shared with the library) refuses cut-off files instead of
    reading past their end"""
    import shutil
    rng = random.Random(3)
    for which in (".1.cfr", ".2.cfr"):
        size = os.path.getsize(os.path.join(tiny_dir, "idx" + which))
        # (the last byte of .1.cfr alone may be missing: indexes written before the end-marker flag existed, FMIndex.hpp:178-181)
        for cut in [0, 7, 40, size - 9] + [rng.randrange(1, size - 1) for _ in range(8)]:
            for ext in (".1.cfr", ".2.cfr", ".3.cfr", ".4.cfr"):
                shutil.copyfile(os.path.join(tiny_dir, "idx" + ext), str(tmp_path / ("t" + ext)))
            with open(str(tmp_path / ("t" + which)), "r+b") as f:
                f.truncate(cut)
            with pytest.raises(RuntimeError):
                HostSim(str(tmp_path / "t"))
    HostSim(os.path.join(tiny_dir, "idx")).close()


def test_device_sdust_against_reference_header():
    """the product's SDUST (64-slot interval table, register-only screen in front) against the UNMODIFIED
    Dustmasker.hpp (oracle/_ref/dust_ref) on generated reads -- no oracle in between"""
    import subprocess
    from oracle_binding import REF_DIR
    exe = os.path.join(REF_DIR, "dust_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dust_ref not built")
    rng = random.Random(71)
    reads = [b"ACG", b"A" * 300, b"ACGT" * 80, b"A" * 70 + b"N" * 70 + b"C" * 70]
    for it in range(5000):
        L = rng.choice([3, 30, 63, 64, 65, 100, 150, 151, 300, 700, 1500])
        mode = rng.random()
        if mode < 0.3:
            s = bytearray(rng.choice(b"ACGT") for _ in range(L))
        elif mode < 0.7:
            per = rng.randint(1, 12)
            unit = bytes(rng.choice(b"ACGT") for _ in range(per))
            noise = rng.choice([0, 0.02, 0.05, 0.15])
            s = bytearray(unit[i % per] if rng.random() >= noise else rng.choice(b"ACGTN") for i in range(L))
        else:
            s = bytearray(rng.choice(b"ACGT") for _ in range(L))
            for _ in range(rng.randint(1, 4)):
                a = rng.randrange(L)
                n = rng.randint(3, 90)
                unit = bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(1, 4))) if rng.random() < 0.7 else b"N"
                s[a:a + n] = (unit * 90)[:n]
            s = s[:L]
        reads.append(bytes(s))
    out = subprocess.run([exe], input=b"".join(r + b"\n" for r in reads), stdout=subprocess.PIPE, check=True).stdout.split(b"\n")
    masked = 0
    for r, exp in zip(reads, out):
        assert hostsim_dust(r) == exp, r
        if not hostsim_dust_screen(r):
            assert exp == r, r  # a read the screen clears is one the reference leaves alone
        masked += exp != r
    assert masked > 1500


def test_device_taxonomy_reduction_against_reference_header(tiny_dir):
    """tax_reduce / tax_lca / tax_expand (what the scoring kernel runs) directly against the UNMODIFIED
    Taxonomy.hpp (oracle/_ref/taxonomy_ref): promoted ids and child lists on 5000 random id sets"""
    import subprocess
    from oracle_binding import REF_DIR
    exe = os.path.join(REF_DIR, "taxonomy_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/taxonomy_ref not built")
    rng = random.Random(83)
    hs = HostSim(os.path.join(tiny_dir, "idx"))
    o = Oracle(os.path.join(tiny_dir, "idx"))
    nodes = o.scalar(10)
    o.close()
    queries = []
    for it in range(5000):
        cnt = rng.randint(2, 10)
        mode = rng.random()
        if mode < 0.5:
            ids = [rng.randrange(nodes) for _ in range(cnt)]
        elif mode < 0.8:
            pool = [rng.randrange(nodes) for _ in range(3)]
            ids = [rng.choice(pool) for _ in range(cnt)]
        else:
            ids = [rng.randrange(nodes + 2) for _ in range(cnt)]
        k = rng.choice([1, 1, 2, 3, 5])
        if cnt > k:  # the kernel reduces only when there are more candidates than -k
            queries.append((k, ids))
    text = "".join("%d %s\n" % (k, " ".join(map(str, ids))) for k, ids in queries)
    lines = subprocess.run([exe, os.path.join(tiny_dir, "idx.2.cfr")], input=text.encode(), stdout=subprocess.PIPE,
                           check=True).stdout.decode().split("\n")
    for (k, ids), line in zip(queries, lines):
        left, right = line.split("|")
        ref_ids = [int(x) for x in left.split()]
        ref_lists = [[int(x) for x in l.split(",") if x] for l in right.split(";")] if right else []
        if len(ref_lists) != len(ref_ids):
            ref_lists = []  # Classifier.hpp:823: nothing is printed then
        got_ids, got_lists = hs.expand_taxids(ids, k)
        assert got_ids == ref_ids, (k, ids)
        assert got_lists == (ref_lists if ref_lists else [[] for _ in ref_ids]), (k, ids, line)
    hs.close()


@pytest.mark.parametrize("layout", [1, 2])
def test_device_fm_primitives_against_reference_headers(tiny_dir, layout):
    """rank / access / locate of the product's BWT layouts (run-block as stored, and the occ sectors it is
    transcoded into) directly against the UNMODIFIED compactds headers (oracle/_ref/fm_ref)"""
    import subprocess
    from oracle_binding import REF_DIR
    exe = os.path.join(REF_DIR, "fm_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/fm_ref not built")
    rng = random.Random(97)
    for variant in ("idx", "idx_b1", "idx_b8", "idx_off3"):
        hs = HostSim(os.path.join(tiny_dir, variant), layout=layout)
        o = Oracle(os.path.join(tiny_dir, variant))
        n = o.n
        o.close()
        queries = []
        for p in [0, 1, n - 2, n - 1] + [rng.randrange(n) for _ in range(800)]:
            queries += [("R", rng.choice("ACGT"), p, rng.randrange(2)), ("A", p)]
        queries += [("L", rng.randrange(n)) for _ in range(600)]
        text = "".join(" ".join(map(str, q)) + "\n" for q in queries)
        out = subprocess.run([exe, os.path.join(tiny_dir, variant + ".1.cfr")], input=text.encode(), stdout=subprocess.PIPE,
                             check=True).stdout.decode().split("\n")
        for q, line in zip(queries, out):
            if q[0] == "R":
                assert hs.bwt_rank(q[1], q[2], q[3]) == int(line), (variant, q)
            elif q[0] == "A":
                assert hs.bwt_access(q[1]) == line, (variant, q)
            else:
                assert hs.locate(q[1]) == int(line.split()[0]), (variant, q)
        hs.close()


@pytest.mark.parametrize("layout,k,hitk,sec", [(1, 1, 40, None), (2, 1, 40, None), (2, 3, 2, None), (3, 5, 0, None),
                                               (2, 1, 40, (50, 0.9)), (1, 2, 1, (100, 0.5))])
def test_device_pipeline_against_reference_header(tiny_dir, layout, k, hitk, sec):
    """the whole per-read path of the product (search, row plan, locate, scoring, reduction) directly against
    Classifier::Query of the UNMODIFIED Classifier.hpp (oracle/_ref/classifier_ref), DUST off: golden read
    sets plus generated reads, single and paired.  The oracle library only translates ids to names here."""
    import subprocess
    import sys
    from oracle_binding import REF_DIR
    exe = os.path.join(REF_DIR, "classifier_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/classifier_ref not built")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz"))
    from fuzz_hostsim import make_read
    rng = random.Random(101 + k)
    idx = os.path.join(tiny_dir, "idx")
    _, se = read_fastx(os.path.join(tiny_dir, "se_100.fq"))
    _, p1 = read_fastx(os.path.join(tiny_dir, "pe_100_1.fq"))
    _, p2 = read_fastx(os.path.join(tiny_dir, "pe_100_2.fq"))
    genomes = [b"".join(se[i:i + 40]) for i in range(0, 240, 40)]
    for paired in (False, True):
        r1 = list(se) if not paired else list(p1)
        r2 = None if not paired else list(p2)
        for _ in range(400):
            L = rng.choice([rng.randrange(24, 160), 100, 150, rng.randrange(300, 1200)])
            r1.append(make_read(rng, genomes, L))
            if paired:
                r2.append(make_read(rng, genomes, max(24, L + rng.randrange(-20, 20))))
        keep = [i for i in range(len(r1)) if r1[i] and (not paired or r2[i])]
        r1 = [r1[i] for i in keep]
        r2 = [r2[i] for i in keep] if paired else None
        text = b"".join(r1[i] + b"\t" + (r2[i] if paired else b"-") + b"\n" for i in range(len(r1)))
        extra = [str(hitk)] + ([str(sec[0]), str(sec[1])] if sec else [])
        kw = dict(hitk_factor=hitk)
        if sec:
            kw.update(secondary_len=sec[0], secondary_factor=sec[1])
        out = subprocess.run([exe, idx, "0", str(k)] + extra, input=text, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                             check=True).stdout.decode().split("\n")
        hs = HostSim(idx, layout=layout, k=k, dust=False, **kw)
        names = Oracle(idx, k=k, dust=False)
        res, ids, _ = hs.classify(r1, r2)
        classified = 0
        for i in range(len(r1)):
            n = int(res["n_assign"][i])
            parts = []
            for j in range(min(n, k)):
                a = int(ids[i][j])
                if int(res["by_rank"][i]):
                    parts.append("%s:%d" % (names.L.cfr_oracle_rank_name(names.h, a).decode(), names.L.cfr_oracle_orig_taxid(names.h, a)))
                else:
                    parts.append("%s:%d" % (names.L.cfr_oracle_seq_name(names.h, a).decode(),
                                            names.L.cfr_oracle_orig_taxid(names.h, names.L.cfr_oracle_seqid_to_taxid(names.h, a))))
            got = "%d %d %d %d %d %s" % (int(res["score"][i]), int(res["secondary_score"][i]), int(res["hit_length"][i]),
                                         int(res["query_length"][i]), n, ";".join(parts))
            assert got == out[i], (paired, i, r1[i])
            classified += n > 0
        assert classified > 300
        hs.close()
        names.close()


def test_positions_beyond_2_pow_32():
    """the arithmetic that only a collection of more than 2^32 rows exercises, stated here with Python integers: 40-bit
    sector counters (occ_pack / occ_base: the three high bytes share a word), FixedSizeElemArray::Read past bit 2^32
    (FixedSizeElemArray.hpp:102), the row plan of a wide hit (Classifier.hpp:620-666), and one FMIndex::BackwardExtend /
    LF step (FMIndex.hpp:352-386) on sectors whose counts and C[] offsets have high bits.  (Real indexes of 2 * 10^10 and
    1.4 * 10^11 rows are compared with the reference binary on the GPU box: tests/test_gpu_reference_at_scale.py,
    tests/cli_bench.py c5.)"""
    import ctypes as C

    import numpy as np

    from hostsim_binding import lib
    L = lib()
    rng = random.Random(77)
    M40 = (1 << 40) - 1
    for _ in range(20000):
        a, c, g = (rng.randrange(0, 1 << rng.choice([8, 31, 32, 33, 39, 40])) for _ in range(3))
        a, c, g = a & M40, c & M40, g & M40
        sec = (a + c + g + 63) // 64 + rng.randrange(0, 1 << 20)
        want = [a, c, g, sec * 64 - (a + c + g)]
        for sym in range(4):
            assert L.hostsim_occ_roundtrip(a, c, g, sec, sym) == want[sym], (a, c, g, sec, sym)
    for _ in range(5000):  # the 32-bit walker below 2^32 rows
        a, c, g = (rng.randrange(0, 1 << 30) for _ in range(3))
        sec = (a + c + g + 63) // 64 + rng.randrange(0, 1 << 10)
        want = [a, c, g, sec * 64 - (a + c + g)]
        for sym in range(4):
            assert L.hostsim_occ_roundtrip32(a, c, g, sec, sym) == want[sym] & 0xffffffff

    # sampled SA: 17-bit elements, element indexes around 2^32 / 17 .. 2^33 / 17 bits (the array is lazily committed)
    bits = 17
    n_elem = (1 << 33) // bits + 64
    words = np.zeros((n_elem * bits + 63) // 64 + 2, dtype=np.uint64)
    probes = [rng.randrange((1 << 32) // bits - 40, (1 << 32) // bits + 40) for _ in range(200)] + \
             [rng.randrange(0, n_elem) for _ in range(300)]
    vals = {}
    for i in sorted(set(probes)):
        v = rng.randrange(1, 1 << bits)
        vals[i] = v
        s = i * bits
        w, r = s >> 6, s & 63
        words[w] |= np.uint64((v << r) & 0xffffffffffffffff)
        if r + bits > 64:
            words[w + 1] |= np.uint64(v >> (64 - r))
    for i, v in vals.items():
        assert L.hostsim_sa_read(words.ctypes.data, bits, i) == v, i

    # row plans of hits with ranges up to 2^37 rows, at row numbers up to 2^38
    out = (C.c_uint64 * 5)()
    for _ in range(3000):
        k, hitk = rng.choice([1, 5, 40]), rng.choice([0, 2, 40])
        sp = rng.randrange(0, 1 << 38)
        rng_rows = rng.choice([1, 7, k * max(hitk, 1), k * max(hitk, 1) + 1, rng.randrange(1, 1 << 37)])
        ep = sp + rng_rows - 1
        mx = k * hitk
        if rng_rows <= mx or hitk <= 0 or k <= 0:
            step, fwd, total = 0, rng_rows, rng_rows
        else:
            step = rng_rows // mx + (1 if rng_rows % mx else 0)
            fwd = (rng_rows - 1) // step + 1
            down = 1 if fwd >= mx else mx - fwd
            total = fwd + min(down, (rng_rows - 1) // step + 1)
        t0, t1 = rng.randrange(0, total), total - 1

        def row(t):
            if step == 0:
                return sp + t
            return sp + t * step if t < fwd else ep - (t - fwd) * step
        L.hostsim_plan_rows(sp, ep, k, hitk, t0, t1, out)
        assert list(out) == [step, fwd, total, row(t0), row(t1)], (sp, ep, k, hitk)

    # one extend / LF step on two sectors at sector 2^28 + ... of a virtual index of 2^36 rows
    n = 1 << 36
    C5 = (C.c_uint64 * 5)(0, n // 4 + 12345, n // 2 + 777, 3 * n // 4 + 99, n)
    res = (C.c_uint64 * 5)()
    for _ in range(3000):
        sec0 = (1 << 28) + rng.randrange(0, 1 << 27)
        lo = [rng.getrandbits(64) for _ in range(2)]
        hi = [rng.getrandbits(64) for _ in range(2)]
        before = sec0 * 64
        cnt0 = [rng.randrange(0, before // 4) for _ in range(3)]
        sym = [[((lo[s] >> i) & 1) | (((hi[s] >> i) & 1) << 1) for i in range(64)] for s in range(2)]
        flat = sym[0] + sym[1]

        def rank(c, p, inclusive):  # Sequence::Rank over the virtual BWT: rows before sec0 counted by cnt0
            upto = p - before + (1 if inclusive else 0)
            base = cnt0[c] if c < 3 else before - sum(cnt0)
            return base + sum(1 for x in flat[:upto] if x == c)
        last_code = rng.randrange(4)
        first_isa = before + rng.randrange(0, 128) if rng.random() < 0.7 else rng.randrange(0, n)

        def fm_rank(c, p, inclusive):
            r = rank(c, p, inclusive)
            if c == last_code and (p < first_isa or (not inclusive and p == first_isa)):
                r += 1
            return r
        c = rng.randrange(4)
        sp = before + rng.randrange(0, 127)
        ep = sp if rng.random() < 0.3 else before + rng.randrange(sp - before, 127)
        nsp = C5[c] + fm_rank(c, sp, 0)
        nep = (C5[c] + fm_rank(c, ep, 1) - 1) if sp != ep else nsp + (0 if flat[ep - before] == c else -1)
        lf = C5[flat[sp - before]] + fm_rank(flat[sp - before], sp, 1) - 1
        cnt = (C.c_uint64 * 3)(*cnt0)
        L.hostsim_extend_high(sec0, n, C5, first_isa, last_code, lo[0], hi[0], lo[1], hi[1], cnt, c, sp, ep, res)
        m = (1 << 64) - 1
        assert [res[0], res[1]] == [nsp & m, nep & m] and [res[2], res[3]] == [nsp & m, nep & m], (sec0, c, sp, ep)
        assert res[4] == lf & m
