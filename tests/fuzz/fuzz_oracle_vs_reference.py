#!/usr/bin/env python
"""Pins the oracle beyond the committed goldens: generated reads (see fuzz_hostsim.py) are classified by
the UNMODIFIED reference binary (oracle/_ref/centrifuger) and by the oracle with the same random
options (-k, --hitk-factor, --min-hitlen, --no-dust, --consider-secondary, --expand-taxid); the two
TSVs must be byte-identical.  Build container only (needs oracle/_ref).  Test infrastructure only.
usage: fuzz_oracle_vs_reference.py [rounds] [seed] [tiny|small]"""
import gzip
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from fuzz_hostsim import make_read  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "centrifuger")


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    import gen_data
    import make_data
    dataset = sys.argv[3] if len(sys.argv) > 3 else "tiny"  # "small": data/small (10 Mbp, 100 sequences; make_data builds it)
    gs, _, _ = gen_data.make_genomes(seed=1, **make_data.DATASETS[dataset]["genomes"])
    genomes = [gen_data.ACGT[g[2]].tobytes() for g in gs]
    d = tempfile.mkdtemp(prefix="cfr_fuzz_ref_")
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
        kw = dict(k=rng.choice([1, 1, 2, 3, 5]), hitk_factor=rng.choice([40, 40, 2, 0, 1]),
                  min_hit_len=rng.choice([0, 0, 16, 20, 30]), dust=rng.random() < 0.7)
        args = ["-k", str(kw["k"]), "--hitk-factor", str(kw["hitk_factor"])]
        if kw["min_hit_len"]:
            args += ["--min-hitlen", str(kw["min_hit_len"])]
        if not kw["dust"]:
            args += ["--no-dust"]
        if rng.random() < 0.4:
            kw.update(secondary_len=rng.choice([30, 60, 100, 500]), secondary_factor=rng.choice([0.5, 0.8, 0.9, 0.995]))
            args += ["--consider-secondary", "%d,%s" % (kw["secondary_len"], kw["secondary_factor"])]
        expand = rng.random() < 0.5
        paired = rng.random() < 0.5
        n = rng.randrange(50, 300)
        lens = [rng.choice([rng.randrange(1, 40), rng.randrange(40, 160), rng.randrange(100, 320), 100, 150,
                            rng.randrange(300, 2500)]) for _ in range(n)]
        r1 = [make_read(rng, genomes, L) for L in lens]
        r2 = [make_read(rng, genomes, max(1, L + rng.randrange(-20, 20))) for L in lens] if paired else None
        ids = ["f%d" % i for i in range(n)]
        for m, reads in ((1, r1), (2, r2)):
            if reads is None:
                continue
            with open(os.path.join(d, "r_%d.fa" % m), "wb") as f:
                for i, s in zip(ids, reads):
                    f.write(b">%s\n%s\n" % (i.encode(), s))
        cmd = [REF, "-x", os.path.join(d if dataset == "tiny" else big, variant), "-t", "1"] + args + (["--expand-taxid"] if expand else [])
        cmd += ["-1", os.path.join(d, "r_1.fa"), "-2", os.path.join(d, "r_2.fa")] if paired else ["-u", os.path.join(d, "r_1.fa")]
        exp = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        o = Oracle(os.path.join(d if dataset == "tiny" else big, variant), **kw)
        got = o.classify_tsv_expanded(ids, r1, r2) if expand else o.classify_tsv(ids, r1, r2)
        o.close()
        if got != exp:
            gl, el = got.split("\n"), exp.split("\n")
            for a, b in zip(gl, el):
                if a != b:
                    print("MISMATCH round", it, variant, args, "expand", expand, "\n oracle   ", a, "\n reference", b)
                    break
            sys.exit(1)
        total += n
    shutil.rmtree(d)
    print("ok:", rounds, "rounds,", total, "reads/pairs, seed", seed)


if __name__ == "__main__":
    main()
