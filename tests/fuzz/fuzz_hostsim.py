#!/usr/bin/env python
"""Differential fuzz on the CPU: the product's stage functions (compiled for the host by tests/hostsim)
against the oracle on generated reads over the tiny golden index -- genome fragments with
substitutions / indels / N runs / low-complexity inserts / chimeras / lowercase, random lengths, random
options (k, hitk, min-hitlen, dust, consider-secondary, arena size, layout, expand-taxid).
Test infrastructure only.   usage: fuzz_hostsim.py [rounds] [seed] [tiny|small]"""
import gzip
import os
import random
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hostsim_binding import HostSim, result_tuples  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

COMP = {65: 84, 67: 71, 71: 67, 84: 65}


def rc(s):
    return bytes(COMP.get(c, 78) for c in reversed(s))


def make_read(rng, genomes, L):
    g = rng.choice(genomes)
    p = rng.randrange(0, len(g) - L)
    s = bytearray(g[p:p + L])
    mode = rng.random()
    rate = rng.choice([0, 0.005, 0.01, 0.03, 0.1])
    for i in range(L):
        if rng.random() < rate:
            s[i] = rng.choice(b"ACGT")
    if mode < 0.15 and L > 40:
        a = rng.randrange(0, L - 20)
        s[a:a + rng.randrange(1, 20)] = b"N" * rng.randrange(1, 20)
    elif mode < 0.3 and L > 60:
        a = rng.randrange(0, L - 40)
        unit = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 5)))
        n = rng.randrange(10, 60)
        s[a:a + n] = (unit * 60)[:n]
    elif mode < 0.4:
        g2 = rng.choice(genomes)
        p2 = rng.randrange(0, len(g2) - L)
        s[L // 2:] = g2[p2:p2 + L - L // 2]
    elif mode < 0.45:
        s = bytearray(bytes(s).lower())
    elif mode < 0.5 and L > 30:
        a = rng.randrange(0, L - 10)
        del s[a:a + rng.randrange(1, 4)]
    elif mode < 0.55:
        s = bytearray(rng.choice(b"ACGT") for _ in range(L))
    s = bytes(s)
    return rc(s) if rng.random() < 0.5 else s


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    import gen_data
    import make_data
    dataset = sys.argv[3] if len(sys.argv) > 3 else "tiny"  # "small": data/small (10 Mbp, 100 sequences; make_data builds it)
    gs, _, _ = gen_data.make_genomes(seed=1, **make_data.DATASETS[dataset]["genomes"])
    genomes = [gen_data.ACGT[g[2]].tobytes() for g in gs]
    d = tempfile.mkdtemp(prefix="cfr_fuzz_")
    tg = os.path.join(ROOT, "tests", "golden", "tiny")
    for f in os.listdir(tg):
        if f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    if dataset != "tiny":
        big = make_data.ensure(dataset, log=lambda *a: None)
        assert big, "data/%s is not available" % dataset
    total = 0
    for it in range(rounds):
        variant = rng.choice(["idx", "idx", "idx_b1", "idx_b8", "idx_off3"]) if dataset == "tiny" else "idx"
        kw = dict(k=rng.choice([1, 1, 2, 3, 5, 0, -1]), hitk_factor=rng.choice([40, 40, 2, 0, 1]),
                  min_hit_len=rng.choice([0, 0, 16, 20, 30]), dust=rng.random() < 0.7)
        if rng.random() < 0.4:
            kw.update(secondary_len=rng.choice([30, 60, 100, 500]), secondary_factor=rng.choice([0.5, 0.8, 0.9, 0.995]))
        layout = rng.choice([1, 2, 3, 4])
        paired = rng.random() < 0.5
        n = rng.randrange(50, 400)
        lens = [rng.choice([rng.randrange(1, 40), rng.randrange(40, 160), rng.randrange(100, 320), 100, 150,
                            rng.randrange(300, 2500)]) for _ in range(n)]
        r1 = [make_read(rng, genomes, L) for L in lens]
        r2 = [make_read(rng, genomes, max(1, L + rng.randrange(-20, 20))) for L in lens] if paired else None
        if rng.random() < 0.2:
            r1[rng.randrange(n)] = b""
        arena = rng.choice([0, 0, 0, 500, 2000])
        idx = os.path.join(d if dataset == "tiny" else big, variant)
        # the load-time tables and the SDUST screen of the product, as hostsim exposes them
        for key, choices in (("HOSTSIM_WIDE_LOOKUP", [None, None, "7", "9"]), ("HOSTSIM_DENSE_LOCATE", [None, "0", "1", "2", "3"]),
                             ("HOSTSIM_NO_DUST_SCREEN", [None, None, "1"]), ("HOSTSIM_DENSE16", [None, "1"])):
            v = rng.choice(choices)
            if v is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = v
        o = Oracle(idx, **kw)
        hs = HostSim(idx, layout=layout, **kw)
        k = hs.stride  # id slots per read: -k, or 64 for -k <= 0 (every best-scoring sequence)
        try:
            res, ids, lists = hs.classify_expanded(r1, r2, arena_rows=arena)
        except RuntimeError as e:
            if arena:  # a single read larger than the arena is a legitimate refusal
                print("round", it, "refused:", e, kw, "arena", arena)
                hs.close()
                o.close()
                continue
            raise
        got = result_tuples(res, ids, k)
        for i in range(n):
            ores, child, cnt = o.query_expanded(r1[i], r2[i] if r2 else None)
            exp = o.result_tuple(ores)[:7]
            if got[i] != exp:
                print("MISMATCH round", it, "read", i, variant, kw, "layout", layout, "arena", arena)
                print(" r1", r1[i], "\n r2", r2[i] if r2 else None, "\n got", got[i], "\n exp", exp)
                sys.exit(1)
            el, at = [], 0
            for j in range(min(ores.n, k)):
                el.append([int(x) for x in child[at:at + cnt[j]]])
                at += cnt[j]
            if lists[i] != el:
                print("LIST MISMATCH round", it, "read", i, variant, kw, lists[i], el)
                sys.exit(1)
        total += n
        hs.close()
        o.close()
    shutil.rmtree(d)
    print("ok:", rounds, "rounds,", total, "reads/pairs, seed", seed)


if __name__ == "__main__":
    main()
