#!/usr/bin/env python
"""Differential fuzz of the CLI's FASTA/FASTQ ingest against the UNMODIFIED reference binary: generated read
files with the irregularities kseq.h tolerates (multi-line records, CRLF, blank lines, '@' / '+' / '>' at
the start of quality lines, comments, missing final newline, lowercase, empty sequences, gzip) are run
through `oracle/_ref/centrifuger --min-hitlen 5000 --no-dust --un P` (every read unclassified: the outputs
depend on the parser only) and through `centrifuger-b200 --dry-run-output --un P`; TSV and read files must be
byte-identical.  Build container only.  Test infrastructure only.   usage: fuzz_cli_parser.py [rounds] [seed]"""
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


def record(rng, i, fastq, paired_tag):
    L = rng.choice([0, 1, 5, 30, 100, 151, 400]) if rng.random() < 0.3 else rng.randrange(20, 200)
    seq = "".join(rng.choice("ACGTNacgtn" if rng.random() < 0.1 else "ACGT") for _ in range(L))
    name = "q%d" % i + (paired_tag if rng.random() < 0.7 else "")
    comment = rng.choice(["", "", " a comment", "\tBC:Z:ACGT x=1", " 1:N:0:ATCACG"])
    nl = "\r\n" if rng.random() < 0.1 else "\n"
    width = rng.choice([0, 0, 60, 17]) if L else 0
    lines = [seq[a:a + width] for a in range(0, L, width)] if width else [seq]
    out = ("@" if fastq else ">") + name + comment + nl + nl.join(lines) + nl
    if fastq:
        q = "".join(rng.choice("#5?FI@+>") for _ in range(L))
        if width and rng.random() < 0.5:
            q = nl.join(q[a:a + width] for a in range(0, L, width))
        out += "+" + (name if rng.random() < 0.2 else "") + nl + q + nl
    if rng.random() < 0.05:
        out += nl
    return out


def md5s(d):
    return {f: hashlib.md5(gzip.open(os.path.join(d, f), "rb").read()).hexdigest() for f in sorted(os.listdir(d))}


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    d = tempfile.mkdtemp(prefix="cfr_fuzz_cli_")
    tg = os.path.join(ROOT, "tests", "golden", "tiny")
    for f in os.listdir(tg):
        if f.startswith("idx.") and f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    for it in range(rounds):
        paired = rng.random() < 0.5
        fastq = rng.random() < 0.7
        n = rng.randrange(1, 80)
        files = []
        for m in range(2 if paired else 1):
            text = "".join(record(rng, i, fastq if rng.random() < 0.9 else not fastq, "/%d" % (m + 1)) for i in range(n))
            if rng.random() < 0.2:
                text = text.rstrip("\r\n")
            gz = rng.random() < 0.3
            path = os.path.join(d, "r%d_%d.%s" % (it, m, "fq.gz" if gz else "fq"))
            with (gzip.open(path, "wb") if gz else open(path, "wb")) as f:
                f.write(text.encode())
            files.append(path)
        inputs = ["-1", files[0], "-2", files[1]] if paired else ["-u", files[0]]
        if paired and rng.random() < 0.3:  # the same pairs as one interleaved file
            recs = [[record(rng, i, fastq, "/%d" % (m + 1)) for m in range(2)] for i in range(n)]
            path = os.path.join(d, "r%d_il.fq" % it)
            with open(path, "wb") as f:
                f.write("".join(a + b for a, b in recs).encode())
            inputs = ["-i", path]
        outs = []
        for who, cmd in (("ref", [REF, "-x", os.path.join(d, "idx"), "-t", "1", "--min-hitlen", "5000", "--no-dust"]),
                         ("our", [EXE, "--dry-run-output", "--batch", str(rng.choice([1, 7, 1 << 20]))])):
            od = os.path.join(d, "%s_%d" % (who, it))
            os.makedirs(od)
            r = subprocess.run(cmd + inputs + ["--un", os.path.join(od, "un")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            outs.append((r.returncode, r.stdout, md5s(od)))
        if outs[0][0] != 0 and outs[1][0] != 0:
            continue  # both refuse the input (mate files of different length): what was printed before does not matter
        if outs[0] != outs[1]:
            print("MISMATCH round", it, files, "\n ref rc", outs[0][0], outs[0][2], "\n our rc", outs[1][0], outs[1][2])
            a, b = outs[0][1].decode().split("\n"), outs[1][1].decode().split("\n")
            for x, y in zip(a, b):
                if x != y:
                    print(" ref:", x, "\n our:", y)
                    break
            print(" kept in", d)
            sys.exit(1)
    shutil.rmtree(d)
    print("ok:", rounds, "rounds, seed", seed)


if __name__ == "__main__":
    main()
