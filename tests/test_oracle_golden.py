"""The oracle (oracle/cfr_oracle.c) against the golden vectors.

Pins the CPU restatement to (1) the reference's own fixture
example/example_class.out and (2) outputs of the unmodified reference binary
committed under tests/golden (see tests/golden/make_golden.py).
"""
import hashlib
import os

import pytest

from conftest import golden_path
from oracle_binding import Oracle, dust_mask, read_fastx


def _args_to_kw(args):
    kw = dict(dust=True)
    it = iter(args)
    for a in it:
        if a == "--no-dust":
            kw["dust"] = False
        elif a == "-k":
            kw["k"] = int(next(it))
        elif a == "--hitk-factor":
            kw["hitk_factor"] = int(next(it))
        elif a == "--min-hitlen":
            kw["min_hit_len"] = int(next(it))
        elif a == "--consider-secondary":
            ln, fac = next(it).split(",")
            kw["secondary_len"], kw["secondary_factor"] = int(ln), float(fac)
    return kw


def _run(idx, files, args):
    ids, r1 = read_fastx(files[0])
    r2 = read_fastx(files[1])[1] if len(files) == 2 else None
    o = Oracle(idx, **_args_to_kw(args))
    try:
        return o.classify_tsv(ids, r1, r2)
    finally:
        o.close()


def test_example_class_out(example_idx):
    """README.md:195-209 fixture; matches the reference only with --no-dust."""
    files = [golden_path("example", "example_1.fq"), golden_path("example", "example_2.fq")]
    got = _run(example_idx, files, ["--no-dust"])
    exp = open(golden_path("example", "example_class.out")).read()
    assert got == exp
    assert hashlib.md5(got.encode()).hexdigest() == "fd328aa27f2bf4c83ac4eaacde30920c"


@pytest.mark.parametrize("case", ["pe_default", "pe_nodust", "pe_k5", "se_default", "se_k5_nodust"])
def test_example_reference_outputs(example_idx, manifest, case):
    m = manifest["example"][case]
    files = [golden_path("example", f) for f in m["files"]]
    got = _run(example_idx, files, m["args"])
    assert got == open(golden_path("example", case + ".tsv")).read()
    assert hashlib.md5(got.encode()).hexdigest() == m["md5"]


def test_tiny_reference_outputs(tiny_dir, manifest):
    """48 (index variant x read set x option) cases produced by the reference binary."""
    assert len(manifest["tiny"]) >= 40
    for name, m in sorted(manifest["tiny"].items()):
        files = [os.path.join(tiny_dir, f) for f in m["files"]]
        got = _run(os.path.join(tiny_dir, m["index"]), files, m["args"])
        exp = open(golden_path("tiny", "expected", name + ".tsv")).read()
        assert got == exp, name
        assert hashlib.md5(got.encode()).hexdigest() == m["md5"], name


def test_tiny_expand_taxid_outputs(tiny_dir, manifest):
    """--expand-taxid (the expandedTaxIDs column) as written by the reference binary, -k 1..5"""
    assert len(manifest["expanded"]) >= 18
    with_lists = 0
    for name, m in sorted(manifest["expanded"].items()):
        files = [os.path.join(tiny_dir, f) for f in m["files"]]
        ids, r1 = read_fastx(files[0])
        r2 = read_fastx(files[1])[1] if len(files) == 2 else None
        o = Oracle(os.path.join(tiny_dir, "idx"), **_args_to_kw(m["args"]))
        got = o.classify_tsv_expanded(ids, r1, r2)
        o.close()
        exp = open(golden_path("tiny", "expanded", name + ".tsv")).read()
        assert got == exp, name
        assert hashlib.md5(got.encode()).hexdigest() == m["md5"], name
        with_lists += m["rows_with_lists"]
    assert with_lists > 400


def test_unlimited_results_k0(tiny_dir, manifest):
    """-k 0 and negative -k: every row of a hit is resolved and every best-scoring sequence is reported, never
    reduced by rank (Classifier.hpp:620-623, :784-785); goldens of the reference binary (tests/golden/make_golden_k0.py)"""
    for name, m in sorted(manifest["k0"].items()):
        files = [golden_path("tiny", f) for f in m["files"]]
        args = [a for a in m["args"] if a != "--expand-taxid"]
        ids, r1 = read_fastx(files[0])
        r2 = read_fastx(files[1])[1] if len(files) == 2 else None
        o = Oracle(os.path.join(tiny_dir, m["index"]), **_args_to_kw(args))
        got = o.classify_tsv_expanded(ids, r1, r2) if "--expand-taxid" in m["args"] else o.classify_tsv(ids, r1, r2)
        o.close()
        assert got == open(golden_path("tiny", "k0", name + ".tsv")).read(), name
        assert hashlib.md5(got.encode()).hexdigest() == m["md5"], name


def test_long_reads_and_consider_secondary(tiny_dir, manifest):
    """reads of 2 - 9 kbp, and the near-tie rule (Classifier.hpp:763-781, `2nd >= (size_t)(factor * best)`
    once the second hit length passes the bar) with the bar lowered so that short reads reach it"""
    for name, m in sorted(manifest["long"].items()):
        files = [golden_path("tiny", f) for f in m["files"]]
        args = [a for a in m["args"] if a != "--expand-taxid"]
        ids, r1 = read_fastx(files[0])
        r2 = read_fastx(files[1])[1] if len(files) == 2 else None
        o = Oracle(os.path.join(tiny_dir, "idx"), **_args_to_kw(args))
        got = o.classify_tsv_expanded(ids, r1, r2) if "--expand-taxid" in m["args"] else o.classify_tsv(ids, r1, r2)
        o.close()
        assert got == open(golden_path("tiny", "long", name + ".tsv")).read(), name
        assert hashlib.md5(got.encode()).hexdigest() == m["md5"], name


def test_reduce_taxids_against_reference_header(tiny_dir):
    """Taxonomy::ReduceTaxIds with its child lists: the oracle against the UNMODIFIED reference header
    (oracle/_ref/taxonomy_ref, built from /root/reference/Taxonomy.hpp) on 6000 random id sets"""
    import random
    import subprocess
    from oracle_binding import REF_DIR
    exe = os.path.join(REF_DIR, "taxonomy_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/taxonomy_ref not built")
    rng = random.Random(41)
    o = Oracle(os.path.join(tiny_dir, "idx"))
    nodes = o.scalar(10)
    queries = []
    for it in range(6000):
        cnt = rng.randint(1, 10)
        mode = rng.random()
        if mode < 0.5:
            ids = [rng.randrange(nodes) for _ in range(cnt)]
        elif mode < 0.8:
            pool = [rng.randrange(nodes) for _ in range(3)]
            ids = [rng.choice(pool) for _ in range(cnt)]
        else:
            ids = [rng.randrange(nodes + 2) for _ in range(cnt)]
        queries.append((rng.choice([1, 1, 2, 3, 5]), ids))
    text = "".join("%d %s\n" % (k, " ".join(map(str, ids))) for k, ids in queries)
    r = subprocess.run([exe, os.path.join(tiny_dir, "idx.2.cfr")], input=text.encode(), stdout=subprocess.PIPE, check=True)
    lines = r.stdout.decode().split("\n")
    with_lists = 0
    for (k, ids), line in zip(queries, lines):
        left, right = line.split("|")
        ref_ids = [int(x) for x in left.split()]
        ref_lists = [[int(x) for x in l.split(",") if x] for l in right.split(";")] if right or len(ref_ids) == 0 else []
        if right == "":
            ref_lists = []  # no list at all, or one empty list: both print as an empty column
        got_ids, got_lists = o.reduce_taxids_expanded(ids, k)
        assert got_ids == ref_ids, (k, ids)
        if len(ref_lists) != len(ref_ids):
            ref_lists = []  # Classifier.hpp:823 prints nothing then
        norm = lambda ls: ls if any(ls) else []
        assert norm(got_lists) == norm(ref_lists), (k, ids, line)
        with_lists += bool(norm(ref_lists))
    assert with_lists > 1500
    o.close()


def test_dust_mask_against_reference_header():
    """SDUST (Dustmasker.hpp) as ClassifyReads_Thread applies it: the oracle's masker against the UNMODIFIED
    reference header (oracle/_ref/dust_ref) on 12000 generated reads rich in low-complexity stretches, N runs
    longer than the window, lowercase and non-ACGT bytes"""
    import random
    import subprocess
    from oracle_binding import REF_DIR
    exe = os.path.join(REF_DIR, "dust_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dust_ref not built")
    rng = random.Random(53)
    reads = [b"", b"A", b"AC", b"ACG", b"A" * 300, b"N" * 200, b"ACGT" * 80, b"A" * 70 + b"N" * 70 + b"C" * 70]
    for it in range(12000):
        L = rng.choice([3, 7, 30, 63, 64, 65, 100, 128, 150, 151, 250, 300, 700, 2000]) if rng.random() < 0.8 else rng.randrange(1, 400)
        mode = rng.random()
        if mode < 0.25:
            s = bytearray(rng.choice(b"ACGT") for _ in range(L))
        elif mode < 0.6:
            per = rng.randint(1, 12)
            unit = bytes(rng.choice(b"ACGT") for _ in range(per))
            noise = rng.choice([0, 0.02, 0.05, 0.15])
            s = bytearray(unit[i % per] if rng.random() >= noise else rng.choice(b"ACGTN") for i in range(L))
        elif mode < 0.8:  # random with planted repeats and N runs
            s = bytearray(rng.choice(b"ACGT") for _ in range(L))
            for _ in range(rng.randint(1, 4)):
                a = rng.randrange(L)
                n = rng.randint(3, 90)
                unit = bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(1, 4))) if rng.random() < 0.7 else b"N"
                s[a:a + n] = (unit * 90)[:n]
            s = s[:L]
        else:
            s = bytearray(rng.choice(b"AAAAACGTNacgtXR-") for _ in range(L))
        reads.append(bytes(s))
    text = b"".join(r + b"\n" for r in reads)
    out = subprocess.run([exe], input=text, stdout=subprocess.PIPE, check=True).stdout.split(b"\n")
    masked = 0
    for r, exp in zip(reads, out):
        assert dust_mask(r) == exp, r
        masked += exp != r
    assert masked > 4000


@pytest.mark.parametrize("variant", ["idx", "idx_b1", "idx_b8", "idx_off3"])
def test_fm_index_against_reference_headers(tiny_dir, variant):
    """rank / access / FM rank / backward search / locate walk: the oracle against the UNMODIFIED
    compactds headers (oracle/_ref/fm_ref loads the same .1.cfr) on random queries"""
    import random
    import subprocess
    from oracle_binding import REF_DIR
    exe = os.path.join(REF_DIR, "fm_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/fm_ref not built")
    rng = random.Random(61)
    o = Oracle(os.path.join(tiny_dir, variant))
    n = o.n
    _, reads = read_fastx(os.path.join(tiny_dir, "se_100.fq"))
    queries = []
    for p in [0, 1, n - 2, n - 1] + [rng.randrange(n) for _ in range(1500)]:
        c = rng.choice("ACGT")
        incl = rng.randrange(2)
        queries += [("R", c, p, incl), ("A", p), ("F", c, p, incl)]
    for _ in range(1200):
        queries.append(("L", rng.randrange(n)))
    for _ in range(1500):
        r = rng.choice(reads)
        a = rng.randrange(len(r) - 12)
        s = bytearray(r[a:a + rng.randrange(11, len(r) - a + 1)])
        if rng.random() < 0.3:
            s[rng.randrange(len(s))] = rng.choice(b"ACGTN")
        if rng.random() < 0.1:
            s = bytearray(rng.choice(b"ACGT") for _ in range(rng.randrange(10, 40)))
        queries.append(("S", bytes(s).decode()))
    text = "".join(" ".join(map(str, q)) + "\n" for q in queries)
    out = subprocess.run([exe, os.path.join(tiny_dir, variant + ".1.cfr")], input=text.encode(), stdout=subprocess.PIPE,
                         check=True).stdout.decode().split("\n")
    for q, line in zip(queries, out):
        if q[0] == "R":
            assert o.bwt_rank(q[1], q[2], q[3]) == int(line), q
        elif q[0] == "A":
            assert o.bwt_access(q[1]) == line, q
        elif q[0] == "F":
            assert o.fm_rank(q[1], q[2], q[3]) == int(line), q
        elif q[0] == "L":
            assert list(o.locate(q[1])) == [int(x) for x in line.split()], q
        else:
            assert list(o.backward_search(q[1].encode())) == [int(x) for x in line.split()], q
    o.close()


@pytest.mark.parametrize("variant,mhl", [("idx", 0), ("idx_b8", 16), ("idx_off3", 30)])
def test_search_stage_against_reference_header(tiny_dir, variant, mhl):
    """GetHitsFromRead / AdjustHitBoundaryFromStrandHits / SearchForwardAndReverse (Classifier.hpp:274-583):
    the oracle's hit lists (sp, ep, length, offset, strand) against the UNMODIFIED Classifier.hpp
    (oracle/_ref/classifier_ref) on the golden read sets plus generated reads, single and paired"""
    import random
    import subprocess
    import sys
    from oracle_binding import REF_DIR
    exe = os.path.join(REF_DIR, "classifier_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/classifier_ref not built")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz"))
    from fuzz_hostsim import make_read
    rng = random.Random(89)
    _, se = read_fastx(os.path.join(tiny_dir, "se_100.fq"))
    _, p1 = read_fastx(os.path.join(tiny_dir, "pe_100_1.fq"))
    _, p2 = read_fastx(os.path.join(tiny_dir, "pe_100_2.fq"))
    _, e1 = read_fastx(os.path.join(tiny_dir, "edge_1.fq"))
    _, e2 = read_fastx(os.path.join(tiny_dir, "edge_2.fq"))
    genomes = [b"".join(se[i:i + 40]) for i in range(0, 240, 40)]  # stitched reads: chimeric "genomes" to cut fuzz reads from
    pairs = [(r, None) for r in se] + list(zip(p1, p2)) + list(zip(e1, e2))
    for _ in range(600):
        L = rng.choice([rng.randrange(24, 160), 100, 150, rng.randrange(300, 1500)])
        r1 = make_read(rng, genomes, L)
        r2 = make_read(rng, genomes, max(24, L + rng.randrange(-20, 20))) if rng.random() < 0.5 else None
        pairs.append((r1, r2))
    pairs = [(a, b) for a, b in pairs if a and b != b"" and b"\t" not in a]
    text = b"".join(a + b"\t" + (b if b else b"-") + b"\n" for a, b in pairs)
    args = [exe, os.path.join(tiny_dir, variant)] + ([str(mhl)] if mhl else [])
    out = subprocess.run(args, input=text, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode().split("\n")
    o = Oracle(os.path.join(tiny_dir, variant), min_hit_len=mhl, dust=False)
    with_hits = 0
    for (a, b), line in zip(pairs, out):
        exp = [tuple(int(x) for x in h.split(",")) for h in line.split(";") if h]
        got = [(sp, ep, l, off, strand) for sp, ep, l, off, strand in o.search(a, b, cap=4096)]
        assert got == exp, (a, b)
        with_hits += bool(exp)
    assert with_hits > 500
    o.close()


def test_index_header_facts(example_idx):
    """the example index header as parsed."""
    o = Oracle(example_idx)
    assert o.scalar(0) == 6901256 and o.scalar(1) == 2 and o.scalar(2) == 3450628
    assert o.scalar(3) == 4901542 and chr(o.scalar(4)) == "C"
    assert o.scalar(5) == 16 and o.scalar(6) == 1 and o.scalar(7) == 431329
    assert o.scalar(8) == 10 and o.scalar(9) == 1 and o.scalar(14) == 23
    assert [o.scalar(18 + i) for i in range(5)] == [0, 2131485, 3452524, 4776192, 6901256]
    o.close()


def test_rank_access_consistency(tiny_dir):
    """Rank must be the prefix sum of Access (the check compactds/test.cpp:940-1002 performs)."""
    for v in ("idx", "idx_b1", "idx_b8"):
        o = Oracle(os.path.join(tiny_dir, v))
        n = o.n
        cnt = {c: 0 for c in "ACGT"}
        step = 1
        for i in range(0, min(n, 20000), step):
            c = o.bwt_access(i)
            for x in "ACGT":
                assert o.bwt_rank(x, i, 0) == cnt[x]
            cnt[c] += 1
            assert o.bwt_rank(c, i, 1) == cnt[c]
        o.close()


def test_dust_known_answers():
    assert dust_mask(b"A" * 100) == b"N" * 100
    assert dust_mask(b"AC") == b"AC"
    s = b"ACGTTGCAAGCTTGACCATGGTACCGATCGATTAGCCGTA"
    assert dust_mask(s) == s  # high complexity: untouched
    m = dust_mask(b"ACGTTGCAAGCTTGACCATGGTACCGATCG" + b"T" * 40 + b"ACGTTGCAAGCTTGACCATGGTACCGATCG")
    assert m.count(b"N") >= 36 and m[:20] == b"ACGTTGCAAGCTTGACCATG"
