"""Size-independent properties of the CUDA path at BASELINE configs[1] scale (100 Mbp / 50-taxa
index, hundreds of thousands of reads), where the oracle is too slow to be the checker:
layout equivalence, shard/chunk invariance, counter checksums, idempotence, and the domain
round trip "an error-free read is assigned to its source sequence or an ancestor taxon"."""
import os
import sys

import numpy as np
import pytest

import centrifuger_b200 as cb

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def c2():
    import gen_data
    import make_data
    d = make_data.ensure("c2", log=lambda *a: None)
    if d is None:
        pytest.skip("data/c2 not available")
    genomes = make_data.genomes_of("c2")
    cat = gen_data.concat_genomes(genomes)
    return os.path.join(d, "idx"), genomes, cat


def _pack(arr):
    n, rl = arr.shape
    return np.ascontiguousarray(arr).reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(rl))


def _eq(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_layouts_shards_chunks_agree_at_scale(c2):
    import gen_data
    idx, genomes, cat = c2
    n = 300_000
    s1, o1 = _pack(gen_data.make_reads_se_fast(genomes, n, 100, seed=21, cat=cat))
    occ = cb.Classifier(idx, layout=cb.LAYOUT_OCCLINE)
    rb = cb.Classifier(idx, layout=cb.LAYOUT_RUNBLOCK)
    a = occ.classify_packed(s1, o1)
    assert _eq(a, rb.classify_packed(s1, o1))                      # two layouts, same answers
    assert _eq(a, occ.classify_packed(s1, o1))                     # idempotent
    occ.taxon_counts_reset()
    # shards: classifying three slices separately == classifying the whole batch
    parts_r, parts_i = [], []
    from centrifuger_b200 import distributed as D
    for r in range(3):
        ss, oo, lo, hi = D.shard_packed(s1, o1, r, 3)
        res, ids = occ.classify_packed(ss, oo)
        parts_r.append(res)
        parts_i.append(ids)
    assert _eq(a, (np.concatenate(parts_r), np.concatenate(parts_i)))
    tc = occ.taxon_counts()                                         # counters are additive over shards
    assert int(tc[occ.node_cnt + 1]) == n
    assert int(tc[occ.node_cnt + 2]) == int((a[0]["n_assign"] > 0).sum())
    assert int(tc[:occ.node_cnt + 1].sum()) == int(np.minimum(a[0]["n_assign"], 1).sum())
    # small device chunks + tiny arena (pipeline + follow-up passes) == one big chunk
    small = cb.Classifier(idx, layout=cb.LAYOUT_OCCLINE, max_batch_reads=37_111, arena_rows=200_000)
    assert _eq(a, small.classify_packed(s1, o1))
    for c in (occ, rb, small):
        c.close()


def test_paired_k5_layouts_agree(c2):
    import gen_data
    idx, genomes, cat = c2
    n = 100_000
    r1, r2 = gen_data.make_reads_pe_fast(genomes, n, 150, seed=23, cat=cat)
    s1, o1 = _pack(r1)
    s2, o2 = _pack(r2)
    occ = cb.Classifier(idx, layout=cb.LAYOUT_OCCLINE, k=5)
    rb = cb.Classifier(idx, layout=cb.LAYOUT_RUNBLOCK, k=5)
    a = occ.classify_packed(s1, o1, s2, o2)
    assert _eq(a, rb.classify_packed(s1, o1, s2, o2))
    assert int((a[0]["query_length"] == 300).sum()) == n
    occ.close()
    rb.close()


def test_error_free_reads_hit_their_source(c2):
    """domain round trip: a 100-bp substring of genome s scores (100-15)^2 with hitLength 100 and is
    assigned to s itself or, when several sequences share the 100-mer, to a common ancestor taxon"""
    import gen_data
    idx, genomes, cat = c2
    rng = np.random.default_rng(5)
    allg, starts, lens = cat
    n = 50_000
    gi = rng.integers(0, len(lens), size=n)
    pos = starts[gi] + (rng.random(n) * (lens[gi] - 100)).astype(np.int64)
    reads = gen_data.ACGT[allg[pos[:, None] + np.arange(100)[None, :]]]
    flip = rng.random(n) < 0.5
    reads[flip] = gen_data.revcomp(reads[flip])
    s1, o1 = _pack(reads)
    g = cb.Classifier(idx, dust=False)
    res, ids = g.classify_packed(s1, o1)
    assert (res["score"] == 85 * 85).all() and (res["hit_length"] == 100).all() and (res["n_assign"] == 1).all()
    # taxonomy from the generator
    _, nodes, _ = gen_data.make_genomes(seed=1, **__import__("make_data").DATASETS["c2"]["genomes"])
    name_to_seqid = {g.seq_name(i): i for i in range(len(genomes))}
    bad = 0
    for i in range(n):
        src_name, src_tax = genomes[gi[i]][0], genomes[gi[i]][1]
        if res["by_rank"][i] == 0:
            sid = int(ids[i])
            if sid != name_to_seqid[src_name]:
                # another strain carries the identical 100-mer and the best set has one member only if
                # the source itself scored lower -- impossible for an error-free read
                bad += 1
        else:
            t = g.orig_taxid(int(ids[i]))
            x = src_tax
            ok = False
            while True:
                if x == t:
                    ok = True
                    break
                if nodes[x][0] == x:
                    break
                x = nodes[x][0]
            bad += 0 if ok else 1
    assert bad == 0
    g.close()


def test_load_time_tables_change_nothing_at_scale(c2, monkeypatch):
    """the literal path (every LF step and BackwardExtend of the reference) and the default path (dense
    locate table; wide lookup table forced on) give identical records for 200k single reads and 60k
    pairs, while running fewer steps"""
    import gen_data
    idx, genomes, cat = c2
    s1, o1 = _pack(gen_data.make_reads_se_fast(genomes, 200_000, 100, seed=31, cat=cat))
    r1, r2 = gen_data.make_reads_pe_fast(genomes, 60_000, 150, seed=33, cat=cat)
    p1, q1 = _pack(r1)
    p2, q2 = _pack(r2)
    out = {}
    for name, env in (("literal", {"CFR_B200_DENSE_LOCATE": "-1", "CFR_B200_WIDE_LOOKUP": "0"}),
                      ("tables", {"CFR_B200_WIDE_LOOKUP": "12"}),
                      ("tables_pos64", {"CFR_B200_WIDE_LOOKUP": "11", "CFR_B200_DENSE_LOCATE": "2", "CFR_B200_POS64": "1"})):
        for k in ("CFR_B200_DENSE_LOCATE", "CFR_B200_WIDE_LOOKUP", "CFR_B200_POS64"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        g = cb.Classifier(idx, k=5)
        g.reset_counters()
        a = g.classify_packed(s1, o1)
        b = g.classify_packed(p1, q1, p2, q2)
        out[name] = (a, b, g.counters())
        g.close()
    for name in ("tables", "tables_pos64"):
        assert _eq(out["literal"][0], out[name][0]) and _eq(out["literal"][1], out[name][1]), name
        assert out[name][2]["n_lf"] < out["literal"][2]["n_lf"] and out[name][2]["n_extend"] < out["literal"][2]["n_extend"]
        assert out[name][2]["n_locate"] == out["literal"][2]["n_locate"] and out[name][2]["n_search"] == out["literal"][2]["n_search"]
