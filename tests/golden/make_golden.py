#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/.

Run in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

* example/  : the reference's own fixture (README.md:195-209): example_1.fq,
              example_2.fq and example_class.out copied verbatim, plus the
              outputs of the unmodified reference binary for the default
              (DUST on), -k 5 and single-end modes.  The example index itself
              (19 MB) is rebuilt from /root/reference/example/ref.fa by
              tools/make_data.py and is not committed.
* tiny/     : a 360 kbp / 18-sequence synthetic collection (tools/gen_data.py,
              seed 1) indexed by the reference builder in four variants
              (auto block size, --rbbwt-b 1, --rbbwt-b 8, --offrate 3), the
              index files gzip'ed, its reads, and the reference binary's TSV
              for a grid of options.  MANIFEST.json lists every case.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_data  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "centrifuger")
REF_EXAMPLE = "/root/reference/example"

OPTION_GRID = {
    "default": [],
    "nodust": ["--no-dust"],
    "k5": ["-k", "5"],
    "k3_hitk2": ["-k", "3", "--hitk-factor", "2"],
    "mhl16_nodust": ["--min-hitlen", "16", "--no-dust"],
    "k2_hitk0": ["-k", "2", "--hitk-factor", "0"],
}


def run_ref(idx, files, extra):
    cmd = [REF, "-x", idx, "-t", "1"] + (["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]])
    return subprocess.run(cmd + extra, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout


def md5(b):
    return hashlib.md5(b).hexdigest()


def main():
    manifest = {"example": {}, "tiny": {}}
    # ---- example ----
    ex = os.path.join(HERE, "example")
    os.makedirs(ex, exist_ok=True)
    for f in ("example_1.fq", "example_2.fq", "example_class.out"):
        shutil.copyfile(os.path.join(REF_EXAMPLE, f), os.path.join(ex, f))
    d = make_data.ensure("example")
    idx = os.path.join(d, "cfr_ref_idx")
    pe = [os.path.join(ex, "example_1.fq"), os.path.join(ex, "example_2.fq")]
    cases = {"pe_default": (pe, []), "pe_nodust": (pe, ["--no-dust"]), "pe_k5": (pe, ["-k", "5"]),
             "se_default": (pe[:1], []), "se_k5_nodust": (pe[:1], ["-k", "5", "--no-dust"])}
    for name, (files, extra) in cases.items():
        out = run_ref(idx, files, extra)
        with open(os.path.join(ex, name + ".tsv"), "wb") as f:
            f.write(out)
        manifest["example"][name] = {"files": [os.path.basename(x) for x in files], "args": extra, "md5": md5(out)}
    assert open(os.path.join(ex, "pe_nodust.tsv"), "rb").read() == open(os.path.join(ex, "example_class.out"), "rb").read()
    # ---- tiny ----
    td = make_data.ensure("tiny")
    tg = os.path.join(HERE, "tiny")
    os.makedirs(os.path.join(tg, "expected"), exist_ok=True)
    for v in make_data.DATASETS["tiny"]["variants"]:
        for k in (1, 2, 3, 4):
            src = os.path.join(td, "%s.%d.cfr" % (v, k))
            with open(src, "rb") as fi, gzip.GzipFile(os.path.join(tg, "%s.%d.cfr.gz" % (v, k)), "wb", mtime=0) as fo:
                fo.write(fi.read())
    for f in ("se_100.fq", "pe_100_1.fq", "pe_100_2.fq", "edge.fq", "edge_1.fq", "edge_2.fq"):
        shutil.copyfile(os.path.join(td, f), os.path.join(tg, f))
    read_sets = {"se": ["se_100.fq"], "pe": ["pe_100_1.fq", "pe_100_2.fq"], "edge": ["edge.fq"],
                 "edgepe": ["edge_1.fq", "edge_2.fq"]}
    for v in make_data.DATASETS["tiny"]["variants"]:
        for rs, files in read_sets.items():
            for on, extra in OPTION_GRID.items():
                if v != "idx" and on not in ("default", "k5"):
                    continue
                out = run_ref(os.path.join(td, v), [os.path.join(td, f) for f in files], extra)
                name = "%s__%s__%s" % (v, rs, on)
                with open(os.path.join(tg, "expected", name + ".tsv"), "wb") as f:
                    f.write(out)
                manifest["tiny"][name] = {"index": v, "files": files, "args": extra, "md5": md5(out)}
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", len(manifest["example"]), "example cases and", len(manifest["tiny"]), "tiny cases")


if __name__ == "__main__":
    main()
