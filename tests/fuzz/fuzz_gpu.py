#!/usr/bin/env python
"""Differential fuzz on the GPU: the CUDA path through the C ABI against the oracle on generated reads
(see fuzz_hostsim.py) with random options, kernel variants (CFR_B200_* knobs), locate arena sizes and
device chunk sizes; results and --expand-taxid lists must be equal.  Test infrastructure only.
usage: fuzz_gpu.py [rounds] [seed]"""
import gzip
import os
import random
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import centrifuger_b200 as cb  # noqa: E402
from fuzz_hostsim import make_read  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

KNOBS = {"CFR_B200_POS64": [None, None, "1"], "CFR_B200_OCC_LOAD": [None, None, "0"], "CFR_B200_DUST_SCREEN": [None, None, "0"],
         "CFR_B200_WIDE_LOOKUP": [None, None, "11", "12"], "CFR_B200_DENSE_LOCATE": [None, None, "-1", "1", "3"],
         "CFR_B200_QUORUM": [None, None, "4", "24"], "CFR_B200_SEARCH_BLOCKS": [None, None, "8", "12"]}


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    import gen_data
    import make_data
    gs, _, _ = gen_data.make_genomes(seed=1, **make_data.DATASETS["tiny"]["genomes"])
    genomes = [gen_data.ACGT[g[2]].tobytes() for g in gs]
    d = tempfile.mkdtemp(prefix="cfr_fuzz_gpu_")
    tg = os.path.join(ROOT, "tests", "golden", "tiny")
    for f in os.listdir(tg):
        if f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    total = 0
    for it in range(rounds):
        variant = rng.choice(["idx", "idx", "idx_b1", "idx_b8", "idx_off3"])
        kw = dict(k=rng.choice([1, 1, 2, 3, 5]), hitk_factor=rng.choice([40, 40, 2, 0, 1]),
                  min_hit_len=rng.choice([0, 0, 16, 20, 30]), dust=rng.random() < 0.7)
        if rng.random() < 0.4:
            kw.update(secondary_len=rng.choice([30, 60, 100, 500]), secondary_factor=rng.choice([0.5, 0.8, 0.9, 0.995]))
        env = {}
        for key, choices in KNOBS.items():
            v = rng.choice(choices)
            os.environ.pop(key, None)
            if v is not None:
                os.environ[key] = env[key] = v
        layout = rng.choice([cb.LAYOUT_RUNBLOCK, cb.LAYOUT_OCCLINE, cb.LAYOUT_OCCLINE])
        paired = rng.random() < 0.5
        n = rng.randrange(200, 3000)
        lens = [rng.choice([rng.randrange(1, 40), rng.randrange(40, 160), rng.randrange(100, 320), 100, 150,
                            rng.randrange(300, 2500)]) for _ in range(n)]
        r1 = [make_read(rng, genomes, L) for L in lens]
        r2 = [make_read(rng, genomes, max(1, L + rng.randrange(-20, 20))) for L in lens] if paired else None
        if rng.random() < 0.2:
            r1[rng.randrange(n)] = b""
        arena = rng.choice([0, 0, 0, 3000, 20000])
        chunk = rng.choice([0, 0, 577, 1 << 20])
        idx = os.path.join(d, variant)
        o = Oracle(idx, **kw)
        g = cb.Classifier(idx, layout=layout, expand_taxid=True, arena_rows=arena, max_batch_reads=chunk, **kw)
        k = kw["k"]
        where = (it, variant, kw, env, "layout", layout, "arena", arena, "chunk", chunk)
        try:
            if chunk == 577:  # the chunked copy/compute pipeline (no lists there): results only
                res, ids = g.classify(r1, r2)
                lists = None
            else:
                res, ids, lists = g.classify_expanded(r1, r2)
        except cb.CfrError as e:
            if arena and "arena_rows" in str(e):  # a single read larger than the arena is a legitimate, loud refusal
                print("round", it, "refused:", e)
                g.close()
                o.close()
                continue
            print("ERROR", where)
            raise
        for i in range(n):
            ores, child, cnt = o.query_expanded(r1[i], r2[i] if r2 else None)
            exp = o.result_tuple(ores)[:7]
            m = int(res["n_assign"][i])
            got = (int(res["score"][i]), int(res["secondary_score"][i]), int(res["hit_length"][i]), int(res["query_length"][i]),
                   m, int(res["by_rank"][i]), tuple(int(x) for x in ids[i][:min(m, k)]))
            if got != exp:
                print("MISMATCH", where, "read", i, "\n r1", r1[i], "\n r2", r2[i] if r2 else None, "\n got", got, "\n exp", exp)
                sys.exit(1)
            if lists is not None:
                el, at = [], 0
                for j in range(min(ores.n, k)):
                    el.append([int(x) for x in child[at:at + cnt[j]]])
                    at += cnt[j]
                if lists[i] != el:
                    print("LIST MISMATCH", where, "read", i, lists[i], el)
                    sys.exit(1)
        total += n
        g.close()
        o.close()
    shutil.rmtree(d)
    print("ok:", rounds, "rounds,", total, "reads/pairs, seed", seed)


if __name__ == "__main__":
    main()
