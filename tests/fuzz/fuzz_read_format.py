#!/usr/bin/env python
"""Differential fuzz of --read-format / --barcode / --UMI / --barcode-whitelist / --barcode-translate against
the UNMODIFIED reference binary: random descriptions (several stretches per category, negative ends, ends
past the record, '-' strands, header-comment fields by number and by prefix, barcodes cut out of read 1) over
the tiny golden read sets, both programs run with every read unclassified (reference: --min-hitlen 5000
--no-dust; ours: --dry-run-output); TSV and --un files must be byte-identical.  Build container only.
Test infrastructure only.   usage: fuzz_read_format.py [rounds] [seed]"""
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
TG = os.path.join(ROOT, "tests", "golden", "tiny")


def stretch(rng, cat, maxlen, short=False):
    start = rng.randrange(0, max(1, maxlen // 3))
    if short:  # several stretches of one category: keep their sum below the record length (the reference's buffer)
        start = rng.randrange(0, max(1, maxlen - 12))
        end = start + rng.randrange(0, max(1, maxlen // 4))
    else:
        end = rng.choice([-1, -1, -rng.randrange(2, 8), rng.randrange(start, maxlen + 20)])
    s = "%s:%d:%d" % (cat, start, end)
    r = rng.random()
    if r < 0.2:
        s += ":-"
    elif r < 0.3:
        s += ":+"
    return s


def md5s(d):
    return {f: hashlib.md5(gzip.open(os.path.join(d, f), "rb").read()).hexdigest() for f in sorted(os.listdir(d))}


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    d = tempfile.mkdtemp(prefix="cfr_fuzz_fmt_")
    for f in os.listdir(TG):
        if f.startswith("idx.") and f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(TG, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    crashed = 0
    for it in range(rounds):
        paired = rng.random() < 0.4
        reads = ["pe_100_1.fq", "pe_100_2.fq"] if paired else [rng.choice(["se_100.fq", "se_com.fq"])]
        items, opts = [], []
        for cat in (("r1", "r2") if paired else ("r1",)):
            cnt = rng.randrange(0, 4)
            for _ in range(cnt):
                items.append(stretch(rng, cat, 100, short=cnt > 1))
        bc_mode = rng.choice(["none", "file", "inline", "header"]) if not paired or rng.random() < 0.5 else "none"
        if reads[0] != "se_com.fq" and bc_mode == "header":
            bc_mode = "file"
        if paired and bc_mode == "file":
            bc_mode = "inline"  # bc.fq has one record per read of se_100.fq
        if bc_mode == "file":
            opts += ["--barcode", os.path.join(TG, "bc.fq")]
            cnt = rng.randrange(1, 3)
            for _ in range(cnt):
                items.append(stretch(rng, "bc", 28, short=cnt > 1))
            if rng.random() < 0.5:
                opts += ["--UMI", os.path.join(TG, "bc.fq")]
                items.append(stretch(rng, "um", 28))
        elif bc_mode == "inline":
            items.append(stretch(rng, "bc", 40))
            if rng.random() < 0.5:
                items.append(stretch(rng, "um", 40))
        elif bc_mode == "header":
            items.append("bc:hd:%s:%d:%d" % (rng.choice(["0", "1", "2", "CB", "UB:Z"]), rng.randrange(0, 7), rng.choice([-1, -2, 12, 30])))
            if rng.random() < 0.5:
                items.append("um:hd:%s:%d:-1" % (rng.choice(["1", "UB", "XX"]), rng.randrange(0, 6)))
        whitelist = bc_mode == "file" and rng.random() < 0.4 and items.count(next(i for i in items if i.startswith("bc"))) == 1
        if whitelist and sum(1 for i in items if i.startswith("bc")) == 1:
            items = [i for i in items if not i.startswith("bc")] + ["bc:0:15" + rng.choice(["", ":+"])]
            opts += ["--barcode-whitelist", os.path.join(TG, "bc_whitelist.txt")]
            if rng.random() < 0.5:
                opts += ["--barcode-translate", os.path.join(TG, "bc_translate.tsv")]
        rng.shuffle(items)
        if items:
            opts += ["--read-format", rng.choice([",", ";"]).join(items)]
        inputs = ["-1", os.path.join(TG, reads[0]), "-2", os.path.join(TG, reads[1])] if paired else ["-u", os.path.join(TG, reads[0])]
        outs = []
        for who, cmd in (("ref", [REF, "-x", os.path.join(d, "idx"), "-t", "1", "--min-hitlen", "5000", "--no-dust"]),
                         ("our", [EXE, "--dry-run-output", "--batch", str(rng.choice([13, 1 << 20]))])):
            od = os.path.join(d, "%s_%d" % (who, it))
            os.makedirs(od)
            r = subprocess.run(cmd + opts + inputs + ["--un", os.path.join(od, "un")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            outs.append((r.returncode, r.stdout, md5s(od)))
        if outs[0][0] < 0:
            crashed += 1  # stretches longer than the record overrun the reference's buffer: nothing to compare with
            continue
        if outs[0][0] != 0 and outs[1][0] != 0:
            continue
        if outs[0] != outs[1]:
            print("MISMATCH round", it, opts, inputs, "\n ref rc", outs[0][0], outs[0][2], "\n our rc", outs[1][0], outs[1][2])
            a, b = outs[0][1].decode().split("\n"), outs[1][1].decode().split("\n")
            for x, y in zip(a, b):
                if x != y:
                    print(" ref:", x, "\n our:", y)
                    break
            print(" kept in", d)
            sys.exit(1)
    shutil.rmtree(d)
    print("ok:", rounds, "rounds (%d skipped: the reference crashed), seed" % crashed, seed)


if __name__ == "__main__":
    main()
