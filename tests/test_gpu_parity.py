"""Parity of the CUDA path (through the C ABI, centrifuger_b200.Classifier) with the
oracle and with the committed reference goldens.  Bit-exact: all work on this
path is integer / byte / index arithmetic."""
import hashlib
import os
import random

import numpy as np
import pytest

import centrifuger_b200 as cb
from conftest import golden_path
from oracle_binding import Oracle, dust_mask, read_fastx

pytestmark = pytest.mark.gpu

LAYOUTS = [cb.LAYOUT_RUNBLOCK, cb.LAYOUT_OCCLINE]

# kernel variants the library picks by index size or environment: the occ sectors walked with 32-bit
# positions (default below 2^32 rows) or 64-bit positions, one 256-bit sector load or two 128-bit
# loads, SDUST with or without the register-only screen
VARIANTS = {"default": {}, "nolanes_pairs": {"CFR_B200_LANES": "0", "CFR_B200_PAIRS": "1"}, "pos64": {"CFR_B200_POS64": "1"}, "pos64_ld128": {"CFR_B200_POS64": "1", "CFR_B200_OCC_LOAD": "0"},
            "ld128": {"CFR_B200_OCC_LOAD": "0"}, "noscreen": {"CFR_B200_DUST_SCREEN": "0"},
            "wide12": {"CFR_B200_WIDE_LOOKUP": "12"}, "wide11_pos64": {"CFR_B200_WIDE_LOOKUP": "11", "CFR_B200_POS64": "1"},
            "literal": {"CFR_B200_DENSE_LOCATE": "-1"}, "dense1_pos64": {"CFR_B200_DENSE_LOCATE": "1", "CFR_B200_POS64": "1"},
            "dense3": {"CFR_B200_DENSE_LOCATE": "3", "CFR_B200_DENSE16": "0"}, "dense0_32bit": {"CFR_B200_DENSE16": "0"}, "dense2_ld128": {"CFR_B200_DENSE_LOCATE": "2", "CFR_B200_OCC_LOAD": "0"},
            # the pair lines (two BackwardExtend steps per 128-byte line, four lanes per strand task) in k_search
            "pairs": {"CFR_B200_PAIRS": "1"}, "pairs_pos64_literal": {"CFR_B200_PAIRS": "1", "CFR_B200_POS64": "1",
                                                                      "CFR_B200_DENSE_LOCATE": "-1"},
            "pairs_wide12_sb10": {"CFR_B200_PAIRS": "1", "CFR_B200_WIDE_LOOKUP": "12", "CFR_B200_PAIR_SEARCH_BLOCKS": "10"},
            # pair lines staged by LDG + STS rounds instead of the cp.async rounds
            "pairs_ldg_dense1": {"CFR_B200_PAIRS": "1", "CFR_B200_PAIR_FETCH": "1", "CFR_B200_DENSE_LOCATE": "1"},
            # ... or by one cp.async round per request kind (the default is one wait per iteration for everything)
            "pairs_rounds_wide11": {"CFR_B200_PAIRS": "1", "CFR_B200_PAIR_FETCH": "2", "CFR_B200_WIDE_LOOKUP": "11"},
            "pairs_wide11_pos64": {"CFR_B200_PAIRS": "1", "CFR_B200_WIDE_LOOKUP": "11", "CFR_B200_POS64": "1"},
            # ... or by one bulk (TMA) copy per lane counted on an mbarrier per warp
            "pairs_tma_wide12": {"CFR_B200_PAIRS": "1", "CFR_B200_PAIR_FETCH": "4", "CFR_B200_WIDE_LOOKUP": "12"}}


@pytest.fixture(params=sorted(VARIANTS))
def variant_env(request, monkeypatch):
    for k in ("CFR_B200_POS64", "CFR_B200_OCC_LOAD", "CFR_B200_DUST_SCREEN", "CFR_B200_WIDE_LOOKUP", "CFR_B200_DENSE_LOCATE",
              "CFR_B200_PAIRS", "CFR_B200_PAIR_SEARCH_BLOCKS", "CFR_B200_PAIR_FETCH", "CFR_B200_DENSE16", "CFR_B200_LANES"):
        monkeypatch.delenv(k, raising=False)
    for k, v in VARIANTS[request.param].items():
        monkeypatch.setenv(k, v)  # read by cfr_open
    return request.param


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
    return kw


def _tuples(res, ids, k):
    out = []
    for i in range(len(res)):
        n = int(res["n_assign"][i])
        out.append((int(res["score"][i]), int(res["secondary_score"][i]), int(res["hit_length"][i]),
                    int(res["query_length"][i]), n, int(res["by_rank"][i]),
                    tuple(int(x) for x in ids[i][:min(n, k)])))
    return out


def _oracle_tuples(o, r1, r2):
    return [o.result_tuple(o.query(r1[i], r2[i] if r2 else None))[:7] for i in range(len(r1))]


# ---------------------------------------------------------------- primitives
@pytest.mark.parametrize("layout", LAYOUTS)
def test_rank_access_locate(tiny_dir, layout):
    rng = random.Random(3)
    for v in ("idx", "idx_b1", "idx_b8", "idx_off3"):
        idx = os.path.join(tiny_dir, v)
        o = Oracle(idx)
        g = cb.Classifier(idx, layout=layout)
        assert g.layout == layout
        n = o.n
        pos = [0, 1, 2, n - 1, n - 2, n // 2] + [rng.randrange(n) for _ in range(3000)]
        codes, ps, inc, exp = [], [], [], []
        for p in pos:
            for c in range(4):
                for incl in (0, 1):
                    codes.append(c)
                    ps.append(p)
                    inc.append(incl)
                    exp.append(o.bwt_rank("ACGT"[c], p, incl))
        got = g.debug_rank(codes, ps, inc)
        assert got.tolist() == exp
        acc = g.debug_access(pos)
        assert ["ACGT"[a] for a in acc] == [o.bwt_access(p) for p in pos]
        loc = g.debug_locate(pos)
        assert loc.tolist() == [o.locate(p)[0] for p in pos]
        g.close()
        o.close()


def test_dust_kernel(tiny_dir):
    rng = random.Random(5)
    reads = [b"A" * 100, b"AC" * 50, b"ACG" * 40, b"AAAC" * 30, b"N" * 10 + b"ACGT" * 10 + b"A" * 80, b"acgt" * 30,
             b"AAGG" * 60, b"ACACG" * 80, b"A" * 70 + b"N" * 70 + b"A" * 70, b"AC", b"", b"ACG",
             b"A" * 30 + b"N" * 66 + b"C" * 30]
    for _ in range(3000):
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
        reads.append(s)
    g = cb.Classifier(os.path.join(tiny_dir, "idx"))
    m1, m2 = g.debug_dust(reads, reads[::-1])
    assert m1 == [dust_mask(r) for r in reads]
    assert m2 == [dust_mask(r) for r in reads[::-1]]
    g.close()


# ---------------------------------------------------------------- goldens
@pytest.mark.parametrize("layout", LAYOUTS)
def test_tiny_goldens_tsv(tiny_dir, manifest, layout):
    """every committed reference TSV is reproduced byte for byte"""
    for name, m in sorted(manifest["tiny"].items()):
        files = [os.path.join(tiny_dir, f) for f in m["files"]]
        ids, r1 = read_fastx(files[0])
        r2 = read_fastx(files[1])[1] if len(files) == 2 else None
        g = cb.Classifier(os.path.join(tiny_dir, m["index"]), layout=layout, **_args_to_kw(m["args"]))
        got = g.classify_tsv(ids, r1, r2)
        g.close()
        assert got == open(golden_path("tiny", "expected", name + ".tsv")).read(), name
        assert hashlib.md5(got.encode()).hexdigest() == m["md5"], name


@pytest.mark.parametrize("layout", LAYOUTS)
def test_example_goldens(example_idx, manifest, layout):
    """BASELINE.json config 1: the bundled example, bit-exact vs example_class.out"""
    for case, m in sorted(manifest["example"].items()):
        files = [golden_path("example", f) for f in m["files"]]
        ids, r1 = read_fastx(files[0])
        r2 = read_fastx(files[1])[1] if len(files) == 2 else None
        g = cb.Classifier(example_idx, layout=layout, **_args_to_kw(m["args"]))
        got = g.classify_tsv(ids, r1, r2)
        g.close()
        assert hashlib.md5(got.encode()).hexdigest() == m["md5"], case
        if case == "pe_nodust":
            assert got == open(golden_path("example", "example_class.out")).read()


# ---------------------------------------------------------------- vs oracle
@pytest.mark.parametrize("layout", LAYOUTS)
def test_small_vs_oracle(small_dir, layout, monkeypatch):
    # operation counts equal the oracle's when every rank / LF step of the reference really runs:
    # the load-time tables that skip steps (dense locate table, wide lookup) are switched off here
    monkeypatch.setenv("CFR_B200_DENSE_LOCATE", "-1")
    monkeypatch.setenv("CFR_B200_WIDE_LOOKUP", "0")
    idx = os.path.join(small_dir, "idx")
    sets = {"se": ["se_100.fq"], "pe": ["pe_150_1.fq", "pe_150_2.fq"], "edge": ["edge_1.fq", "edge_2.fq"]}
    for name, files in sets.items():
        fs = [os.path.join(small_dir, f) for f in files]
        _, r1 = read_fastx(fs[0])
        r2 = read_fastx(fs[1])[1] if len(fs) == 2 else None
        r1 = r1[:6000]
        r2 = r2[:6000] if r2 else None
        for kw in (dict(), dict(k=5), dict(k=2, hitk_factor=0, dust=False)):
            o = Oracle(idx, **kw)
            o.reset_counters()
            exp = _oracle_tuples(o, r1, r2)
            oc = o.counters()
            o.close()
            g = cb.Classifier(idx, layout=layout, **kw)
            g.reset_counters()
            res, ids = g.classify(r1, r2)
            assert _tuples(res, ids, g.k) == exp, (name, kw)
            c = g.counters()
            for key in ("n_rank", "n_access", "n_search", "n_locate", "n_lf", "n_extend"):
                assert c[key] == oc[key], (name, kw, key)
            assert c["n_reads"] == len(r1)
            tc = g.taxon_counts()
            assert int(tc[g.node_cnt + 1]) == len(r1)
            assert int(tc[g.node_cnt + 2]) == int((res["n_assign"] > 0).sum())
            assert int(tc[:g.node_cnt + 1].sum()) == int(np.minimum(res["n_assign"], g.k).sum())
            g.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_deferral_chunking_and_resident_api(small_dir, layout):
    idx = os.path.join(small_dir, "idx")
    _, r1 = read_fastx(os.path.join(small_dir, "pe_150_1.fq"))
    _, r2 = read_fastx(os.path.join(small_dir, "pe_150_2.fq"))
    r1, r2 = r1[:5000], r2[:5000]
    g = cb.Classifier(idx, layout=layout, k=5)
    res0, ids0 = g.classify(r1, r2)
    g.close()
    base = _tuples(res0, ids0, 5)
    # a tiny arena forces the multi-pass path
    g = cb.Classifier(idx, layout=layout, k=5, arena_rows=1500)
    res, ids = g.classify(r1, r2)
    assert _tuples(res, ids, 5) == base
    g.close()
    # small chunks AND a tiny arena: the copy/compute pipeline with follow-up passes in its drain step
    g = cb.Classifier(idx, layout=layout, k=5, max_batch_reads=601, arena_rows=900)
    res, ids = g.classify(r1, r2)
    assert _tuples(res, ids, 5) == base
    g.close()
    # small device chunks
    g = cb.Classifier(idx, layout=layout, k=5, max_batch_reads=777)
    res, ids = g.classify(r1, r2)
    assert _tuples(res, ids, 5) == base
    # resident API
    s1, o1 = cb.pack_reads(r1)
    s2, o2 = cb.pack_reads(r2)
    b = g.upload(s1, o1, s2, o2)
    g.classify_resident(b)
    res, ids = g.fetch(b)
    assert _tuples(res, ids.reshape(-1, 5), 5) == base
    g.classify_resident(b)  # idempotent: classifying a resident batch twice changes nothing
    res, ids = g.fetch(b)
    assert _tuples(res, ids.reshape(-1, 5), 5) == base
    b.free()
    g.close()


def test_edge_inputs(tiny_dir):
    idx = os.path.join(tiny_dir, "idx")
    g = cb.Classifier(idx)
    o = Oracle(idx)
    res, ids = g.classify([], None)
    assert len(res) == 0
    reads = [b"", b"A", b"ACGTACGTAC", b"N" * 300, b"ACGT" * 200]
    res, ids = g.classify(reads)
    assert _tuples(res, ids, 1) == _oracle_tuples(o, reads, None)
    res, ids = g.classify(reads, reads[::-1])
    assert _tuples(res, ids, 1) == _oracle_tuples(o, reads, reads[::-1])
    g.close()
    o.close()
    with pytest.raises(cb.CfrError):
        cb.Classifier(os.path.join(tiny_dir, "does_not_exist"))


def test_arena_regrows_for_a_read_that_exceeds_it(tiny_dir):
    """a read whose rows exceed the whole arena is classified all the same (the reference has no such limit): the arena
    grows to the read's need; with --expand-taxid the lists of the reads scored before survive the regrowth"""
    idx = os.path.join(tiny_dir, "idx")
    _, r1 = read_fastx(os.path.join(tiny_dir, "edge.fq"))
    _, com = read_fastx(golden_path("tiny", "se_com.fq"))
    for kw in (dict(k=5), dict(k=0), dict(k=2, hitk_factor=0)):
        o = Oracle(idx, **kw)
        exp = _oracle_tuples(o, r1 + com[:80], None)
        o.close()
        g = cb.Classifier(idx, arena_rows=8, **kw)
        res, ids = g.classify(r1 + com[:80])
        assert _tuples(res, ids, g.k) == exp, kw
        g.close()
    big = cb.Classifier(idx, k=1, expand_taxid=True)
    eres, eids, elists = big.classify_expanded(com)
    big.close()
    g = cb.Classifier(idx, k=1, expand_taxid=True, arena_rows=5)
    res, ids, lists = g.classify_expanded(com)
    g.close()
    assert np.array_equal(res, eres) and np.array_equal(ids, eids) and lists == elists


# ---------------------------------------------------------------- drop-in CLI
def test_cli_binary_reproduces_reference_tsv(tiny_dir, manifest):
    """centrifuger-b200 (host C++ over the C ABI) prints the reference's TSV byte for byte"""
    import subprocess
    exe = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")
    assert os.path.exists(exe), "run python -m centrifuger_b200.build"
    for name in ("idx__pe__default", "idx__se__k5", "idx__edgepe__k3_hitk2", "idx_b8__edge__default",
                 "idx__edge__mhl16_nodust"):
        m = manifest["tiny"][name]
        files = [os.path.join(tiny_dir, f) for f in m["files"]]
        cmd = [exe, "-x", os.path.join(tiny_dir, m["index"]), "-t", "4", "--batch", "97"]
        cmd += ["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]]
        r = subprocess.run(cmd + m["args"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout.decode() == open(golden_path("tiny", "expected", name + ".tsv")).read(), name
        err = r.stderr.decode()
        assert "Finishes loading index." in err and "can be classified." in err
    # no arguments: usage on stderr, exit 0 (the reference's CI depends on it)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and b"-x FILE: index prefix" in r.stderr
    r = subprocess.run([exe, "-v"], stdout=subprocess.PIPE)
    assert r.stdout.decode().strip() == "Centrifuger v1.1.3-r347"


def test_masked_reads_come_back(tiny_dir):
    """cfr_submit_batch_masked returns the reads as Classifier::Query saw them = the oracle's DUST mask"""
    idx = os.path.join(tiny_dir, "idx")
    _, r1 = read_fastx(os.path.join(tiny_dir, "edge_1.fq"))
    _, r2 = read_fastx(os.path.join(tiny_dir, "edge_2.fq"))
    r1 = r1 + [b"A" * 100, b"AC" * 60, b"ACGTTGCA" * 12 + b"A" * 30]
    r2 = r2 + [b"ACGTTGCA" * 12 + b"T" * 30, b"acgtNNNN" * 10, b"G" * 7]
    g = cb.Classifier(idx, k=5)
    res, ids, m1, m2 = g.classify_masked(r1, r2)
    exp = g.classify(r1, r2)
    assert np.array_equal(res, exp[0]) and np.array_equal(ids, exp[1])
    assert m1 == [dust_mask(x) for x in r1] and m2 == [dust_mask(x) for x in r2]
    g.close()
    g = cb.Classifier(idx, dust=False)
    _, _, m1, m2 = g.classify_masked(r1, r2)
    assert m1 == list(r1) and m2 == list(r2)
    g.close()


def test_cli_un_cl_read_outputs(tiny_dir, manifest, tmp_path):
    """--un / --cl: the files of unclassified / classified reads (DUST-masked as classified, FASTQ or
    FASTA like the input, read ids as printed) decompress to what the reference binary writes"""
    import gzip
    import subprocess
    exe = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")
    for name, m in manifest["reads_out"].items():
        od = tmp_path / name
        od.mkdir()
        files = [golden_path("tiny", f) for f in m["files"]]
        cmd = [exe, "-x", os.path.join(tiny_dir, "idx"), "--batch", "53"]
        cmd += ["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]]
        cmd += m["args"] + ["--un", str(od / "un"), "--cl", str(od / "cl")]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        assert hashlib.md5(r.stdout).hexdigest() == m["tsv_md5"], name
        got = {f: hashlib.md5(gzip.open(str(od / f), "rb").read()).hexdigest() for f in sorted(os.listdir(str(od)))}
        assert got == m["outputs"], name


def test_kernel_variants_vs_oracle(small_dir, variant_env):
    """every kernel variant gives the oracle's answers and operation counts (occ-sector layout)"""
    idx = os.path.join(small_dir, "idx")
    _, r1 = read_fastx(os.path.join(small_dir, "pe_150_1.fq"))
    _, r2 = read_fastx(os.path.join(small_dir, "pe_150_2.fq"))
    _, e1 = read_fastx(os.path.join(small_dir, "edge_1.fq"))
    _, e2 = read_fastx(os.path.join(small_dir, "edge_2.fq"))
    r1, r2 = r1[:4000] + e1, r2[:4000] + e2
    for kw in (dict(), dict(k=5)):
        o = Oracle(idx, **kw)
        o.reset_counters()
        exp = _oracle_tuples(o, r1, r2)
        oc = o.counters()
        o.close()
        g = cb.Classifier(idx, layout=cb.LAYOUT_OCCLINE, **kw)
        g.reset_counters()
        res, ids = g.classify(r1, r2)
        assert _tuples(res, ids, g.k) == exp, (variant_env, kw)
        c = g.counters()
        assert c["n_search"] == oc["n_search"] and c["n_locate"] == oc["n_locate"], (variant_env, kw)
        if variant_env in ("literal", "pairs_pos64_literal"):  # no load-time table skips a step: every count equals the oracle's
            for key in ("n_rank", "n_access", "n_lf", "n_extend"):
                assert c[key] == oc[key], (variant_env, kw, key)
        else:  # the dense locate table (default) shortens LF walks, the wide lookup table skips extends
            assert c["n_lf"] < oc["n_lf"] and c["n_rank"] < oc["n_rank"], (variant_env, kw)
            assert (c["n_extend"] < oc["n_extend"]) == ("wide" in variant_env), (variant_env, kw)
        if variant_env.startswith("pairs"):
            assert g.info(23) > 0  # the pair lines were built and k_search walked them
        g.close()


def test_streaming_submit_wait(small_dir):
    """cfr_submit_batch / cfr_wait_batch: several batches in flight give the same answers as the
    synchronous call, in any wait order"""
    idx = os.path.join(small_dir, "idx")
    _, r1 = read_fastx(os.path.join(small_dir, "pe_150_1.fq"))
    _, r2 = read_fastx(os.path.join(small_dir, "pe_150_2.fq"))
    g = cb.Classifier(idx, k=5, arena_rows=6000)
    parts = [(r1[i:i + 1500], r2[i:i + 1500]) for i in range(0, 7500, 1500)]
    exp = [g.classify(a, b) for a, b in parts]
    jobs = []
    for a, b in parts:
        s1, o1 = cb.pack_reads(a)
        s2, o2 = cb.pack_reads(b)
        jobs.append(g.submit(s1, o1, s2, o2))
    for (tk, res, ids, keep), (eres, eids) in zip(reversed(jobs), reversed(exp)):
        g.wait(tk)
        assert np.array_equal(res, eres) and np.array_equal(ids.reshape(-1, 5), eids)
    g.close()


def test_packed_submit_same_results(tiny_dir, small_dir):
    """cfr_submit_packed (bases packed on the host by cfr_pack_reads: 2-bit codes + N bits, no k_encode pass) gives what
    cfr_submit_batch gives for the same reads, and the oracle's answers: paired and single-end, the edge-case reads
    (empty reads, non-ACGT bytes, low complexity), with and without DUST, with --expand-taxid lists"""
    idx = os.path.join(small_dir, "idx")
    _, r1 = read_fastx(os.path.join(small_dir, "pe_150_1.fq"))
    _, r2 = read_fastx(os.path.join(small_dir, "pe_150_2.fq"))
    _, e1 = read_fastx(os.path.join(small_dir, "edge_1.fq"))
    _, e2 = read_fastx(os.path.join(small_dir, "edge_2.fq"))
    r1, r2 = r1[:4000] + e1 + [b"", b"acgtn" * 30], r2[:4000] + e2 + [b"ACGT" * 40, b""]
    for dust in (True, False):
        g = cb.Classifier(idx, k=5, dust=dust, arena_rows=9000)
        o = Oracle(idx, k=5, dust=dust)
        for a, b in ((r1, r2), (r2, None), ([], None)):
            s1, o1 = cb.pack_reads(a)
            s2, o2 = cb.pack_reads(b) if b is not None else (None, None)
            pk, keep = cb.pack_batch(s1, o1, s2, o2, threads=3)
            h0 = g.info(13)
            tk, res, ids = g.submit_packed(pk)
            g.wait(tk)
            if len(a):
                # bytes over the host link: 12 per 32 bases, plus the offsets of reads that differ in length
                assert 0 <= g.info(13) - h0 - int(pk.n_words) * 12 <= 16 * (len(a) + 1)
                eres, eids = g.classify(a, b)
                assert np.array_equal(res, eres) and np.array_equal(ids.reshape(-1, 5), eids)
                assert _tuples(res, ids.reshape(-1, 5), 5) == _oracle_tuples(o, a, b)
        g.close()
        o.close()
    tidx = os.path.join(tiny_dir, "idx")
    _, c1 = read_fastx(golden_path("tiny", "se_com.fq"))
    g = cb.Classifier(tidx, k=1, expand_taxid=True)
    eres, eids, elists = g.classify_expanded(c1)
    s1, o1 = cb.pack_reads(c1)
    pk, keep = cb.pack_batch(s1, o1)
    tk, res, ids = g.submit_packed(pk)
    g.wait(tk)
    lists = g._expansion_lists(res, lambda c, o_, i, cap, n: g.L.cfr_fetch_expanded(g.h, tk, c, o_, i, cap, n))
    assert np.array_equal(res, eres) and lists == elists and sum(len(x) for l in lists for x in l) > 0
    g.close()


def _oracle_expansion(o, r1, r2, k):
    tup, lists = [], []
    for i in range(len(r1)):
        res, child, cnt = o.query_expanded(r1[i], r2[i] if r2 else None)
        tup.append(o.result_tuple(res)[:7])
        row, at = [], 0
        for j in range(min(res.n, k)):
            row.append([int(x) for x in child[at:at + cnt[j]]])
            at += cnt[j]
        lists.append(row)
    return tup, lists


@pytest.mark.parametrize("layout", LAYOUTS)
def test_expand_taxid_lists(tiny_dir, small_dir, layout):
    """_classifierParam.outputExpandedResult (--expand-taxid): the ids promoted into each reported id,
    streaming and resident forms, also with a locate arena small enough to need several passes"""
    cases = [(os.path.join(tiny_dir, "idx"), os.path.join(tiny_dir, "pe_100_1.fq"), os.path.join(tiny_dir, "pe_100_2.fq"), 10**9),
             (os.path.join(tiny_dir, "idx"), os.path.join(tiny_dir, "se_100.fq"), None, 10**9),
             (os.path.join(small_dir, "idx"), os.path.join(small_dir, "pe_150_1.fq"), os.path.join(small_dir, "pe_150_2.fq"), 3000)]
    n_lists = 0
    for idx, f1, f2, limit in cases:
        _, r1 = read_fastx(f1)
        r2 = read_fastx(f2)[1][:limit] if f2 else None
        r1 = r1[:limit]
        for kw, arena in ((dict(), 0), (dict(k=2), 0), (dict(k=3), 1200), (dict(k=4, hitk_factor=0), 0)):
            o = Oracle(idx, **kw)
            exp_t, exp_l = _oracle_expansion(o, r1, r2, kw.get("k", 1))
            o.close()
            n_lists += sum(1 for row in exp_l if any(row))
            g = cb.Classifier(idx, layout=layout, expand_taxid=True, arena_rows=arena, **kw)
            res, ids, lists = g.classify_expanded(r1, r2)
            assert _tuples(res, ids, g.k) == exp_t, (idx, kw)
            assert lists == exp_l, (idx, kw)
            # resident form
            s1, o1 = cb.pack_reads(r1)
            s2, o2 = cb.pack_reads(r2) if r2 else (None, None)
            b = g.upload(s1, o1, s2, o2)
            g.classify_resident(b)
            res, ids = g.fetch(b)
            assert g.fetch_expanded(b, res) == exp_l, (idx, kw)
            b.free()
            g.close()
    assert n_lists > 300
    # a handle opened without the flag refuses the fetch
    g = cb.Classifier(os.path.join(tiny_dir, "idx"))
    with pytest.raises(cb.CfrError):
        g.classify_expanded([b"ACGT" * 30])
    g.close()


def test_cli_expand_taxid(tiny_dir, manifest):
    """--expand-taxid: the TSV with its expandedTaxIDs column, byte for byte what the reference binary prints"""
    import subprocess
    exe = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")
    for name, m in sorted(manifest["expanded"].items()):
        files = [golden_path("tiny", f) for f in m["files"]]
        for batch in (("47", "1048576") if name.startswith("pe__k") else ("47",)):  # each launch loads the index
            cmd = [exe, "-x", os.path.join(tiny_dir, "idx"), "--expand-taxid", "--batch", batch] + m["args"]
            cmd += ["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            assert r.returncode == 0, r.stderr.decode()
            assert r.stdout.decode() == open(golden_path("tiny", "expanded", name + ".tsv")).read(), (name, batch)
            assert hashlib.md5(r.stdout).hexdigest() == m["md5"], name


def test_unlimited_results_k0(tiny_dir, manifest, monkeypatch):
    """-k 0 / negative -k (Classifier.hpp:620-623, :784-785): the CLI's TSV is byte for byte the reference binary's
    (also with the pair lines in the search kernel and a small arena), the C ABI agrees with the oracle, and a
    read with more best-scoring sequences than unlimited_cap keeps is an error, never a shortened list"""
    import subprocess
    exe = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")
    for pairs in ("0", "1"):
        monkeypatch.setenv("CFR_B200_PAIRS", pairs)
        for name, m in sorted(manifest["k0"].items()):
            files = [golden_path("tiny", f) for f in m["files"]]
            cmd = [exe, "-x", os.path.join(tiny_dir, m["index"])] + m["args"] + (["--batch", "37"] if pairs == "1" else [])
            cmd += ["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            assert r.returncode == 0, r.stderr.decode()
            assert r.stdout.decode() == open(golden_path("tiny", "k0", name + ".tsv")).read(), (name, pairs)
            assert hashlib.md5(r.stdout).hexdigest() == m["md5"], name
    monkeypatch.delenv("CFR_B200_PAIRS")
    idx = os.path.join(tiny_dir, "idx")
    _, r1 = read_fastx(golden_path("tiny", "pe_100_1.fq"))
    _, r2 = read_fastx(golden_path("tiny", "pe_100_2.fq"))
    _, com = read_fastx(golden_path("tiny", "se_com.fq"))
    for layout in LAYOUTS:
        for kw, reads in ((dict(k=0), (r1, r2)), (dict(k=-2, hitk_factor=3), (r1, r2)), (dict(k=0, dust=False), (com, None))):
            o = Oracle(idx, **kw)
            exp = _oracle_tuples(o, *reads)
            o.close()
            assert max(t[4] for t in exp) > 5
            for cap, arena in ((0, 0), (16, 300)):
                g = cb.Classifier(idx, layout=layout, unlimited_cap=cap, arena_rows=arena, **kw)
                assert g.k == (cap or 64)
                res, ids = g.classify(*reads)
                g.close()
                assert _tuples(res, ids, g.k) == exp, (layout, kw, cap)
    g = cb.Classifier(idx, k=0, unlimited_cap=4)
    with pytest.raises(cb.CfrError) as e:
        g.classify(com)
    assert e.value.code == -7 and "unlimited_cap" in str(e.value)
    g.close()


def test_patched_reference_binary(tiny_dir, manifest):
    """INTEGRATION.md's Option 2 applied for real: oracle/_ref/centrifuger_patched is the REFERENCE's binary (its argv
    handling, kseq ingest and ResultWriter, compiled from /root/reference by oracle/patch_reference.py) with the batch
    fan-out of CentrifugerClass.cpp:681-688 replaced by one call into libcfrb200.so.  Its TSV is the unmodified binary's."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "centrifuger_patched")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/centrifuger_patched not built")
    n = 0
    for section, sub in (("tiny", "expected"), ("expanded", "expanded"), ("k0", "k0"), ("long", "long")):
        for name, m in sorted(manifest[section].items()):
            if section == "tiny" and (m["index"] != "idx" or name.split("__")[2] not in ("default", "k5", "k2_hitk0", "mhl16_nodust")):
                continue
            files = [golden_path("tiny", f) for f in m["files"]]
            cmd = [exe, "-x", os.path.join(tiny_dir, m.get("index", "idx"))] + m["args"]
            if section == "expanded" and "--expand-taxid" not in m["args"]:
                cmd += ["--expand-taxid"]
            cmd += ["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            assert r.returncode == 0, (name, r.stderr.decode()[-400:])
            assert r.stdout.decode() == open(golden_path("tiny", sub, name + ".tsv")).read(), (section, name)
            n += 1
    assert n >= 40


def test_cli_long_reads_and_consider_secondary(tiny_dir, manifest):
    """reads of 2 - 9 kbp and the near-tie rule of --consider-secondary (Classifier.hpp:763-781; bars lowered
    so short reads reach it): the CLI's TSV is byte for byte the reference binary's"""
    import subprocess
    exe = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")
    for name, m in sorted(manifest["long"].items()):
        files = [golden_path("tiny", f) for f in m["files"]]
        for extra in (([], ["--batch", "13"]) if name.startswith("long__k") else (["--batch", "13"],)):
            cmd = [exe, "-x", os.path.join(tiny_dir, "idx")] + m["args"] + extra
            cmd += ["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            assert r.returncode == 0, r.stderr.decode()
            assert r.stdout.decode() == open(golden_path("tiny", "long", name + ".tsv")).read(), (name, extra)
            assert hashlib.md5(r.stdout).hexdigest() == m["md5"], name


@pytest.mark.parametrize("layout", LAYOUTS)
def test_long_reads_vs_oracle(tiny_dir, layout):
    """long reads through the C ABI with operation counts, also with an arena that needs several passes"""
    idx = os.path.join(tiny_dir, "idx")
    _, r1 = read_fastx(golden_path("tiny", "long.fa"))
    for kw, arena in ((dict(), 0), (dict(k=2, secondary_len=1000, secondary_factor=0.1), 700), (dict(k=5, dust=False), 0)):
        o = Oracle(idx, **kw)
        exp = _oracle_tuples(o, r1, None)
        o.close()
        g = cb.Classifier(idx, layout=layout, arena_rows=arena, **kw)
        res, ids = g.classify(r1)
        assert _tuples(res, ids, g.k) == exp, kw
        g.close()


def test_cli_sample_sheet(tiny_dir, manifest, tmp_path):
    """--sample-sheet: every output file byte for byte what the reference binary writes (a file named by two
    rows is appended to), nothing on stdout"""
    import subprocess
    exe = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")
    for case, m in manifest["sample_sheet"].items():
        od = tmp_path / case
        od.mkdir()
        sheet = tmp_path / (case + ".sheet")
        with open(str(sheet), "w") as f:
            for row in m["rows"]:
                r1, r2, o = row[:3]
                bc, um = (row[3], row[4]) if len(row) > 3 else (".", ".")
                f.write("%s %s %s %s %s\n" % (golden_path("tiny", r1), r2 if r2 == "." else golden_path("tiny", r2),
                                              bc if bc == "." else golden_path("tiny", bc), um if um == "." else golden_path("tiny", um),
                                              str(od / (o + ".tsv"))))
        r = subprocess.run([exe, "-x", os.path.join(tiny_dir, "idx"), "--batch", "41", "--sample-sheet", str(sheet)] + m["args"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0 and r.stdout == b"", r.stderr.decode()
        got = {o: hashlib.md5(open(str(od / o), "rb").read()).hexdigest() for o in sorted(os.listdir(str(od)))}
        assert got == m["outputs"], case


def test_gpu_fuzz_rounds():
    """a few rounds of tests/fuzz/fuzz_gpu.py: generated reads, random options, kernel variants, arena and chunk sizes"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz", "fuzz_gpu.py"), "6", "303"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    assert r.returncode == 0 and b"ok:" in r.stdout, r.stdout.decode()[-3000:]


def test_cli_read_format_barcode_umi(tiny_dir, manifest, tmp_path):
    """--read-format / --barcode / --UMI with real classification: TSV (barcode / UMI columns) and the
    --un / --cl files (reads as cut and masked, _bc / _um files) are the reference binary's"""
    import gzip
    import subprocess
    exe = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")
    for name, m in sorted(manifest["barcode"].items()):
        files = [golden_path("tiny", f) for f in m["files"]]
        args = [golden_path("tiny", o[1:]) if o.startswith("@") else o for o in m["args"]]
        od = tmp_path / name
        od.mkdir()
        cmd = [exe, "-x", os.path.join(tiny_dir, "idx"), "--batch", "61"] + args
        cmd += ["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]]
        cmd += ["--un", str(od / "un"), "--cl", str(od / "cl")]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr.decode()
        assert r.stdout.decode() == open(golden_path("tiny", "barcode", name + "__real.tsv")).read(), name
        got = {f: hashlib.md5(gzip.open(str(od / f), "rb").read()).hexdigest() for f in sorted(os.listdir(str(od)))}
        assert got == m["real"]["outputs"], name
