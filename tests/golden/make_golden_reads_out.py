#!/usr/bin/env python
"""Golden fixtures for the --un / --cl read outputs (ResultWriter.hpp:118-172, :244-262): md5 of the
DECOMPRESSED files the unmodified reference binary writes, for a few cases over the tiny index.
Adds the "reads_out" section to MANIFEST.json and the FASTA read set tiny/se_100.fa.

    python tests/golden/make_golden_reads_out.py     (build container: needs oracle/_ref)
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "centrifuger")

sys.path.insert(0, os.path.join(ROOT, "tools"))

CASES = {
    "pe_default": (["pe_100_1.fq", "pe_100_2.fq"], []),
    "edgepe_k5": (["edge_1.fq", "edge_2.fq"], ["-k", "5"]),
    "edge_se_default": (["edge.fq"], []),
    "edge_se_nodust": (["edge.fq"], ["--no-dust"]),
    "fasta_se_default": (["se_100.fa"], []),
    # --merge-readpair (ReadPairMerger.hpp): pairs whose inserts are shorter than two reads
    "overlap_merge": (["ov_1.fq", "ov_2.fq"], ["--merge-readpair"]),
    "overlap_merge_k5_nodust": (["ov_1.fq", "ov_2.fq"], ["--merge-readpair", "-k", "5", "--no-dust"]),
    "overlap_nomerge": (["ov_1.fq", "ov_2.fq"], []),
}


def write_overlapping_pairs(tg):
    """200 pairs of 100-bp reads with inserts of 40..260 bp from the tiny collection (read-through, overlap,
    no overlap), 1 % errors, mixed qualities, a few low-complexity fragments"""
    import numpy as np
    import gen_data
    import make_data
    genomes, _, _ = gen_data.make_genomes(seed=1, **make_data.DATASETS["tiny"]["genomes"])
    rng = np.random.default_rng(77)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    adapter = bytes(rng.choice(list(b"ACGT"), size=100).tolist())
    with open(os.path.join(tg, "ov_1.fq"), "wb") as f1, open(os.path.join(tg, "ov_2.fq"), "wb") as f2:
        for i in range(200):
            g = gen_data.ACGT[genomes[int(rng.integers(len(genomes)))][2]].tobytes()
            ins = int(rng.integers(40, 261))
            p = int(rng.integers(0, len(g) - ins))
            frag = g[p:p + ins]
            if i % 23 == 0:
                frag = (b"AC" * 200)[:ins]
            rc = bytes(comp[c] for c in reversed(frag))
            r = []
            for s in ((frag + adapter)[:100], (rc + adapter)[:100]):
                s = bytearray(s)
                for q in range(100):
                    if rng.random() < 0.01:
                        s[q] = int(rng.choice(list(b"ACGTN")))
                r.append(bytes(s))
            qs = [bytes(rng.choice(list(b"#5?FI"), size=100).tolist()) for _ in range(2)]
            f1.write(b"@ov%d/1\n%s\n+\n%s\n" % (i, r[0], qs[0]))
            f2.write(b"@ov%d/2\n%s\n+\n%s\n" % (i, r[1], qs[1]))


def unpack_index(dst):
    src = os.path.join(HERE, "tiny")
    for f in os.listdir(src):
        if f.startswith("idx.") and f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(src, f), "rb") as fi, open(os.path.join(dst, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    return os.path.join(dst, "idx")


def output_md5s(prefix_dir):
    out = {}
    for f in sorted(os.listdir(prefix_dir)):
        if f.endswith(".gz"):
            out[f] = hashlib.md5(gzip.open(os.path.join(prefix_dir, f), "rb").read()).hexdigest()
    return out


def main():
    tg = os.path.join(HERE, "tiny")
    # a FASTA version of the first 120 single-end reads (records without qualities are written as FASTA)
    with open(os.path.join(tg, "se_100.fq")) as fi, open(os.path.join(tg, "se_100.fa"), "w") as fo:
        lines = fi.read().splitlines()
        for i in range(0, min(len(lines), 4 * 120), 4):
            fo.write(">" + lines[i][1:] + "\n" + lines[i + 1] + "\n")
    write_overlapping_pairs(tg)
    manifest = json.load(open(os.path.join(HERE, "MANIFEST.json")))
    manifest["reads_out"] = {}
    d = tempfile.mkdtemp(prefix="cfr_golden_")
    try:
        idx = unpack_index(d)
        for name, (files, extra) in CASES.items():
            od = os.path.join(d, name)
            os.makedirs(od)
            fs = [os.path.join(tg, f) for f in files]
            cmd = [REF, "-x", idx, "-t", "1"] + (["-u", fs[0]] if len(fs) == 1 else ["-1", fs[0], "-2", fs[1]]) + extra + \
                  ["--un", os.path.join(od, "un"), "--cl", os.path.join(od, "cl")]
            tsv = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
            manifest["reads_out"][name] = {"files": files, "args": extra, "tsv_md5": hashlib.md5(tsv).hexdigest(),
                                           "outputs": output_md5s(od)}
            print(name, manifest["reads_out"][name]["outputs"])
    finally:
        shutil.rmtree(d, ignore_errors=True)
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
