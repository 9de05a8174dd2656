#!/usr/bin/env python
"""Differential fuzz of the CLI's block-parallel FASTQ reader (csrc/cfr_cli_bulk.hpp) against its serial reader (which
tests/fuzz/fuzz_cli_parser.py pins to the unmodified reference binary): mostly regular four-line FASTQ -- quality lines
that start with '@' or '+', ids with /1 /2 and comments, a last line without a newline -- with an occasional
irregularity in the middle of a file (CRLF, blank line, multi-line record, empty read, FASTA record, a cut-off record),
parsed with byte ranges of a few hundred bytes on 1 - 8 threads, so that range boundaries fall on every kind of line
and the hand-over to the serial reader happens mid-file.  Both mates, several files per mate, with and without
qualities (--un).  No GPU needed (--dry-run-pipeline / --dry-run-output).  Test infrastructure only.
usage: fuzz_bulk_ingest.py [rounds] [seed]"""
import hashlib
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EXE = os.path.join(ROOT, "centrifuger_b200", "centrifuger-b200")


def make_file(rng, path, n, tag, first_id, irregular):
    out = []
    for i in range(n):
        L = rng.randrange(1, 120)
        seq = "".join(rng.choice("ACGTN") for _ in range(L))
        q = "".join(rng.choice("@+>#5?FI") for _ in range(L))
        name = "r%d" % (first_id + i) + (tag if rng.random() < 0.8 else "") + rng.choice(["", " c", "\tx y"])
        rec = "@%s\n%s\n+%s\n%s\n" % (name, seq, name if rng.random() < 0.1 else "", q)
        if irregular and rng.random() < irregular:
            kind = rng.randrange(7)
            if kind == 0:
                rec = rec.replace("\n", "\r\n")
            elif kind == 1:
                rec += "\n"
            elif kind == 2 and L > 4:
                h = L // 2
                rec = "@%s\n%s\n%s\n+\n%s\n%s\n" % (name, seq[:h], seq[h:], q[:h], q[h:])
            elif kind == 3:
                rec = "@%s\n\n+\n\n" % name
            elif kind == 4:
                rec = ">%s\n%s\n" % (name, seq)
            elif kind == 5:
                rec = "@%s\n%s\n+\n%s\n" % (name, "A" * 1500, "I" * 1500)
            else:
                rec = "@%s\n%s%s\n+\n%s%s\n" % (name, rng.choice("+>@"), seq, "I", q)
        out.append(rec)
    data = "".join(out)
    if rng.random() < 0.3:
        data = data.rstrip("\n")
    with open(path, "w", newline="") as f:
        f.write(data)


def run(args, env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([EXE] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e)
    return r.returncode, hashlib.md5(r.stdout).hexdigest(), r.stderr.decode()


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = random.Random(seed)
    d = tempfile.mkdtemp(prefix="cfr_fuzz_bulk_")
    fast_rounds = 0
    try:
        for it in range(rounds):
            paired = rng.random() < 0.6
            nfiles = rng.choice([1, 1, 2, 3])
            irregular = rng.choice([0, 0, 0.02, 0.2])
            counts = [rng.randrange(1, 400) for _ in range(nfiles)]
            mism = paired and rng.random() < 0.1
            f1, f2, at = [], [], 0
            for k, n in enumerate(counts):
                p1 = os.path.join(d, "a%d_%d_1.fq" % (it, k))
                make_file(rng, p1, n, "/1", at, irregular)
                f1.append(p1)
                if paired:
                    p2 = os.path.join(d, "a%d_%d_2.fq" % (it, k))
                    make_file(rng, p2, n + (rng.choice([-1, 1, 3]) if mism and k == nfiles - 1 else 0), "/2", at, irregular)
                    f2.append(p2)
                at += n
            args = (["-1", ",".join(f1), "-2", ",".join(f2)] if paired else ["-u", ",".join(f1)]) + ["--batch", str(rng.choice([7, 64, 1000]))]
            # file lists are given as repeated options (the reference's way)
            args = []
            for p in f1:
                args += ["-1" if paired else "-u", p]
            for p in f2:
                args += ["-2", p]
            args += ["--batch", str(rng.choice([7, 64, 1000]))]
            modes = [["--dry-run-pipeline"], ["--dry-run-output", "--un", os.path.join(d, "un")]]
            for mode in modes:
                # a batch also ends at a byte budget (256 MB of bases per mate; here a few hundred to a few thousand bytes)
                cap = {"CFR_B200_MAX_BATCH_BASES": str(rng.choice([150, 700, 4000]))} if rng.random() < 0.4 else {}
                serial = run(mode + args, dict(cap, CFR_B200_BULK_INGEST="0"))
                un_serial = {f: hashlib.md5(open(os.path.join(d, f), "rb").read()).hexdigest() for f in os.listdir(d) if f.startswith("un")}
                env = {"CFR_B200_BULK_INGEST": "1", "CFR_B200_INGEST_BLOCK": str(rng.choice([16, 100, 333, 1000, 5000, 1 << 20])),
                       "CFR_B200_INGEST_THREADS": str(rng.randrange(1, 9)), "CFR_B200_STAGE_REPORT": "1"}
                env.update(cap)
                bulk = run(mode + args, env)
                un_bulk = {f: hashlib.md5(open(os.path.join(d, f), "rb").read()).hexdigest() for f in os.listdir(d) if f.startswith("un")}
                if serial[:2] != bulk[:2] or un_serial != un_bulk:
                    print("MISMATCH round", it, mode[0], env, args)
                    keep = tempfile.mkdtemp(prefix="cfr_fuzz_bulk_keep_")
                    for p in f1 + f2:
                        shutil.copy(p, keep)
                    print("inputs kept in", keep)
                    return 1
                if '"block_parallel_ingest": true' in bulk[2]:
                    fast_rounds += 1
            for p in f1 + f2:
                os.remove(p)
        print("ok: fuzz_bulk_ingest %d rounds (%d runs used the block-parallel path)" % (rounds, fast_rounds))
        return 0
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
