#!/usr/bin/env python
"""Differential fuzz of --merge-readpair through the CLI against the UNMODIFIED reference binary: pairs with
inserts from far shorter than a read (read-through into an adapter) to longer than two reads, read lengths
30 - 250, errors, Ns, tandem repeats, differing mate lengths, with and without qualities.  Both programs run
with every read unclassified (reference: --min-hitlen 5000 --no-dust; ours: --dry-run-output): the query
length column shows what was merged and the --un files what is written for merged pairs; both must be
byte-identical.  Build container only.  Test infrastructure only.   usage: fuzz_merge_cli.py [rounds] [seed]"""
import gzip
import hashlib
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "centrifuger")
EXE = os.path.join(ROOT, "centrifuger_b200", "centrifuger-b200")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def rc(s):
    return "".join(COMP[c] for c in reversed(s))


def md5s(d):
    return {f: hashlib.md5(gzip.open(os.path.join(d, f), "rb").read()).hexdigest() for f in sorted(os.listdir(d))}


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    d = tempfile.mkdtemp(prefix="cfr_fuzz_merge_")
    tg = os.path.join(ROOT, "tests", "golden", "tiny")
    for f in os.listdir(tg):
        if f.startswith("idx.") and f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    for it in range(rounds):
        fastq = rng.random() < 0.8
        n = rng.randrange(20, 200)
        adapter = "".join(rng.choice("ACGT") for _ in range(300))
        f1, f2 = os.path.join(d, "m%d_1.f" % it), os.path.join(d, "m%d_2.f" % it)
        with open(f1, "w") as a, open(f2, "w") as b:
            for i in range(n):
                L1 = rng.choice([100, 100, 150, rng.randrange(30, 251)])
                L2 = L1 if rng.random() < 0.7 else rng.randrange(30, 251)
                ins = rng.randrange(10, 2 * max(L1, L2) + 60)
                if rng.random() < 0.15:
                    unit = "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 7)))
                    frag = (unit * 600)[:ins]
                else:
                    frag = "".join(rng.choice("ACGT") for _ in range(ins))
                r = [(frag + adapter)[:L1], (rc(frag) + adapter)[:L2]]
                rate = rng.choice([0, 0.005, 0.02, 0.08])
                for m in range(2):
                    s = list(r[m])
                    for q in range(len(s)):
                        if rng.random() < rate:
                            s[q] = rng.choice("ACGTN")
                    r[m] = "".join(s)
                for m, fh in ((0, a), (1, b)):
                    if fastq:
                        fh.write("@p%d/%d\n%s\n+\n%s\n" % (i, m + 1, r[m], "".join(rng.choice("#5?FI") for _ in r[m])))
                    else:
                        fh.write(">p%d/%d\n%s\n" % (i, m + 1, r[m]))
        outs = []
        for who, cmd in (("ref", [REF, "-x", os.path.join(d, "idx"), "-t", "1", "--min-hitlen", "5000", "--no-dust"]),
                         ("our", [EXE, "--dry-run-output", "--batch", str(rng.choice([11, 1 << 20]))])):
            od = os.path.join(d, "%s_%d" % (who, it))
            os.makedirs(od)
            r = subprocess.run(cmd + ["--merge-readpair", "-1", f1, "-2", f2, "--un", os.path.join(od, "un")],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            outs.append((r.returncode, r.stdout, md5s(od)))
        if outs[0] != outs[1]:
            print("MISMATCH round", it, f1, f2, "\n ref rc", outs[0][0], outs[0][2], "\n our rc", outs[1][0], outs[1][2])
            x, y = outs[0][1].decode().split("\n"), outs[1][1].decode().split("\n")
            for p, q in zip(x, y):
                if p != q:
                    print(" ref:", p, "\n our:", q)
                    break
            print(" kept in", d)
            sys.exit(1)
    shutil.rmtree(d)
    print("ok:", rounds, "rounds, seed", seed)


if __name__ == "__main__":
    main()
