#!/usr/bin/env python
"""Golden fixtures for -k 0 / negative -k ("report every best-scoring sequence": Classifier.hpp:620-623 resolves
every row of a hit, :784-785 never reduces by rank): TSVs of the unmodified reference binary on the tiny index.
Adds the "k0" section to MANIFEST.json.

    python tests/golden/make_golden_k0.py     (build container: needs oracle/_ref)
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "centrifuger")

CASES = {
    "se__k0": ("idx", ["se_100.fq"], ["-k", "0"]),
    "pe__k0": ("idx", ["pe_100_1.fq", "pe_100_2.fq"], ["-k", "0"]),
    "com__k0": ("idx", ["se_com.fq"], ["-k", "0"]),
    "com_b8__k0": ("idx_b8", ["se_com.fq"], ["-k", "0"]),
    "edge__k0_nodust": ("idx", ["edge.fq"], ["-k", "0", "--no-dust"]),
    "edgepe__k0": ("idx", ["edge_1.fq", "edge_2.fq"], ["-k", "0"]),
    "pe__kneg3_hitk2": ("idx", ["pe_100_1.fq", "pe_100_2.fq"], ["-k", "-3", "--hitk-factor", "2"]),
    "com__k0_mhl16_nodust": ("idx", ["se_com.fq"], ["-k", "0", "--min-hitlen", "16", "--no-dust"]),
    "com__k0_expand": ("idx", ["se_com.fq"], ["-k", "0", "--expand-taxid"]),
    "long__k0": ("idx", ["long.fa"], ["-k", "0"]),
}


def main():
    tg = os.path.join(HERE, "tiny")
    d = tempfile.mkdtemp(prefix="cfr_golden_k0_")
    for f in os.listdir(tg):
        if f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    os.makedirs(os.path.join(tg, "k0"), exist_ok=True)
    section = {}
    for name, (idx, files, extra) in CASES.items():
        paths = [os.path.join(tg, f) for f in files]
        cmd = [REF, "-x", os.path.join(d, idx), "-t", "1"] + extra
        cmd += ["-u", paths[0]] if len(paths) == 1 else ["-1", paths[0], "-2", paths[1]]
        out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
        with open(os.path.join(tg, "k0", name + ".tsv"), "wb") as f:
            f.write(out)
        most = max([int(l.split(b"\t")[7]) for l in out.splitlines()[1:]] or [0])
        section[name] = {"index": idx, "files": files, "args": extra, "md5": hashlib.md5(out).hexdigest(), "most_assignments": most}
    mp = os.path.join(HERE, "MANIFEST.json")
    manifest = json.load(open(mp))
    manifest["k0"] = section
    with open(mp, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    shutil.rmtree(d)
    print("wrote", len(section), "cases:", {k: v["most_assignments"] for k, v in section.items()})


if __name__ == "__main__":
    main()
