#!/usr/bin/env python
"""Golden fixtures for --expand-taxid (Classifier.hpp:792-838, Taxonomy.hpp:767-831, :855-882,
:938-971): the TSV of the unmodified reference binary, with its expandedTaxIDs column, for a grid of
-k values over the tiny index.  Adds the "expanded" section to MANIFEST.json.

    python tests/golden/make_golden_expanded.py     (build container: needs oracle/_ref)
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

READ_SETS = {"se": ["se_100.fq"], "pe": ["pe_100_1.fq", "pe_100_2.fq"], "edgepe": ["edge_1.fq", "edge_2.fq"]}
OPTIONS = {"k1": [], "k2": ["-k", "2"], "k3": ["-k", "3"], "k5_hitk2": ["-k", "5", "--hitk-factor", "2"],
           "k1_mhl16_nodust": ["--min-hitlen", "16", "--no-dust"], "k4_hitk0": ["-k", "4", "--hitk-factor", "0"]}


def main():
    tg = os.path.join(HERE, "tiny")
    d = tempfile.mkdtemp(prefix="cfr_golden_exp_")
    for f in os.listdir(tg):
        if f.startswith("idx.") and f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    os.makedirs(os.path.join(tg, "expanded"), exist_ok=True)
    section = {}
    for rs, files in READ_SETS.items():
        for on, extra in OPTIONS.items():
            paths = [os.path.join(tg, f) for f in files]
            cmd = [REF, "-x", os.path.join(d, "idx"), "-t", "1", "--expand-taxid"] + extra
            cmd += ["-u", paths[0]] if len(paths) == 1 else ["-1", paths[0], "-2", paths[1]]
            out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
            name = "%s__%s" % (rs, on)
            with open(os.path.join(tg, "expanded", name + ".tsv"), "wb") as f:
                f.write(out)
            filled = sum(1 for l in out.decode().split("\n")[1:] if l and l.split("\t")[8])
            section[name] = {"files": files, "args": extra, "md5": hashlib.md5(out).hexdigest(), "rows_with_lists": filled}
    mp = os.path.join(HERE, "MANIFEST.json")
    manifest = json.load(open(mp))
    manifest["expanded"] = section
    with open(mp, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    shutil.rmtree(d)
    print("wrote", len(section), "cases;", {k: v["rows_with_lists"] for k, v in section.items()})


if __name__ == "__main__":
    main()
