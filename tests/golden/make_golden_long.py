#!/usr/bin/env python
"""Golden fixtures for long reads and the near-tie rule of --consider-secondary (Classifier.hpp:763-781,
the one floating-point operation on the path): TSVs of the unmodified reference binary for
 * tiny/long.fa -- 36 reads of 2.1 - 9 kbp from the tiny collection (1 % errors, a few with N runs
   and low-complexity inserts), so hit lengths pass the default 2000-base bar, and
 * the short read sets with the bar lowered (--consider-secondary 50,0.9 etc.).
Adds the "long" section to MANIFEST.json.

    python tests/golden/make_golden_long.py     (build container: needs oracle/_ref)
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "centrifuger")
sys.path.insert(0, os.path.join(ROOT, "tools"))

CASES = {
    "long__default": (["long.fa"], []),
    "long__k5": (["long.fa"], ["-k", "5"]),
    "long__k2_sec1000_0.1": (["long.fa"], ["-k", "2", "--consider-secondary", "1000,0.1"]),
    "long__nodust_sec3000_0.05": (["long.fa"], ["--no-dust", "--consider-secondary", "3000,0.05"]),
    "se__sec50_0.9": (["se_100.fq"], ["--consider-secondary", "50,0.9"]),
    "se__k3_sec60_0.8": (["se_100.fq"], ["-k", "3", "--consider-secondary", "60,0.8"]),
    "pe__sec100_0.9": (["pe_100_1.fq", "pe_100_2.fq"], ["--consider-secondary", "100,0.9"]),
    "pe__k5_sec50_0.5": (["pe_100_1.fq", "pe_100_2.fq"], ["-k", "5", "--consider-secondary", "50,0.5"]),
    "edgepe__k2_sec30_0.7": (["edge_1.fq", "edge_2.fq"], ["-k", "2", "--consider-secondary", "30,0.7"]),
    "pe__expand_sec100_0.9": (["pe_100_1.fq", "pe_100_2.fq"], ["--expand-taxid", "--consider-secondary", "100,0.9"]),
}


def write_long_reads(path):
    import gen_data
    import make_data
    genomes, _, _ = gen_data.make_genomes(seed=1, **make_data.DATASETS["tiny"]["genomes"])
    rng = np.random.default_rng(91)
    comp = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}
    with open(path, "wb") as f:
        for i in range(36):
            g = gen_data.ACGT[genomes[int(rng.integers(len(genomes)))][2]].tobytes()
            L = int(rng.integers(2100, 9001))
            p = int(rng.integers(0, len(g) - L))
            s = bytearray(g[p:p + L])
            for q in np.nonzero(rng.random(L) < 0.01)[0]:
                s[int(q)] = int(rng.choice(list(b"ACGT")))
            if i % 7 == 3:  # an N run and a low-complexity insert
                a = int(rng.integers(100, L - 400))
                s[a:a + 40] = b"N" * 40
                s[a + 200:a + 320] = (b"AT" * 60)
            if i % 9 == 4:  # a chimera: the second half comes from another sequence
                g2 = gen_data.ACGT[genomes[int(rng.integers(len(genomes)))][2]].tobytes()
                p2 = int(rng.integers(0, len(g2) - L // 2))
                s[L // 2:] = g2[p2:p2 + L - L // 2]
            s = bytes(s)
            if rng.random() < 0.5:
                s = bytes(comp[c] for c in reversed(s))
            f.write(b">long%d some description\n" % i)
            for a in range(0, L, 80):
                f.write(s[a:a + 80] + b"\n")


def main():
    tg = os.path.join(HERE, "tiny")
    write_long_reads(os.path.join(tg, "long.fa"))
    d = tempfile.mkdtemp(prefix="cfr_golden_long_")
    for f in os.listdir(tg):
        if f.startswith("idx.") and f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    os.makedirs(os.path.join(tg, "long"), exist_ok=True)
    section = {}
    for name, (files, extra) in CASES.items():
        paths = [os.path.join(tg, f) for f in files]
        cmd = [REF, "-x", os.path.join(d, "idx"), "-t", "1"] + extra
        cmd += ["-u", paths[0]] if len(paths) == 1 else ["-1", paths[0], "-2", paths[1]]
        out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
        with open(os.path.join(tg, "long", name + ".tsv"), "wb") as f:
            f.write(out)
        section[name] = {"files": files, "args": extra, "md5": hashlib.md5(out).hexdigest()}
    mp = os.path.join(HERE, "MANIFEST.json")
    manifest = json.load(open(mp))
    manifest["long"] = section
    with open(mp, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    shutil.rmtree(d)
    print("wrote", len(section), "cases")


if __name__ == "__main__":
    main()
