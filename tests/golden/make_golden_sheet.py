#!/usr/bin/env python
"""Golden fixtures for --sample-sheet (CentrifugerClass.cpp:467-516, ResultWriter.hpp:75-107): the TSV
files the unmodified reference binary writes for two sheets over the tiny index -- a paired-end one
whose first output file is used again by the third row (appended to, no second header) and a
single-end one.  Adds the "sample_sheet" section to MANIFEST.json.

    python tests/golden/make_golden_sheet.py     (build container: needs oracle/_ref)
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

# case -> (rows of (read1, read2 or ".", output name), extra args)
SHEETS = {
    "pe": ([("pe_100_1.fq", "pe_100_2.fq", "a"), ("edge_1.fq", "edge_2.fq", "b"), ("ov_1.fq", "ov_2.fq", "a")], []),
    "se_k3": ([("se_100.fq", ".", "a"), ("edge.fq", ".", "b"), ("se_100.fa", ".", "c")], ["-k", "3"]),
    # rows with barcode and UMI files (4th and 5th element; the technical reads of bc.fq serve both)
    "se_bc": ([("se_100.fq", ".", "a", "bc.fq", "bc.fq"), ("se_com.fq", ".", "b", "bc.fq", "bc.fq")],
              ["--read-format", "bc:0:15,um:16:25"]),
}


def write_sheet(path, rows, src, outdir):
    with open(path, "w") as f:
        for row in rows:
            r1, r2, o = row[:3]
            bc, um = (row[3], row[4]) if len(row) > 3 else (".", ".")
            f.write("%s %s %s %s %s\n" % (os.path.join(src, r1), r2 if r2 == "." else os.path.join(src, r2),
                                          bc if bc == "." else os.path.join(src, bc), um if um == "." else os.path.join(src, um),
                                          os.path.join(outdir, o + ".tsv")))


def main():
    tg = os.path.join(HERE, "tiny")
    d = tempfile.mkdtemp(prefix="cfr_golden_sheet_")
    for f in os.listdir(tg):
        if f.startswith("idx.") and f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    os.makedirs(os.path.join(tg, "sheet"), exist_ok=True)
    section = {}
    for case, (rows, extra) in SHEETS.items():
        od = os.path.join(d, case)
        os.makedirs(od)
        sheet = os.path.join(d, case + ".sheet")
        write_sheet(sheet, rows, tg, od)
        r = subprocess.run([REF, "-x", os.path.join(d, "idx"), "-t", "1", "--sample-sheet", sheet] + extra,
                           check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
        assert r.stdout == b""
        outs = {}
        for o in sorted(os.listdir(od)):
            data = open(os.path.join(od, o), "rb").read()
            with open(os.path.join(tg, "sheet", "%s__%s" % (case, o)), "wb") as f:
                f.write(data)
            outs[o] = hashlib.md5(data).hexdigest()
        section[case] = {"rows": [list(r) for r in rows], "args": extra, "outputs": outs}
    mp = os.path.join(HERE, "MANIFEST.json")
    manifest = json.load(open(mp))
    manifest["sample_sheet"] = section
    with open(mp, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    shutil.rmtree(d)
    print("wrote", {k: sorted(v["outputs"]) for k, v in section.items()})


if __name__ == "__main__":
    main()
