#!/usr/bin/env python
"""Golden fixtures for --read-format / --barcode / --UMI (ReadFormatter.hpp, CentrifugerClass.cpp:128-227,
ResultWriter.hpp:158-168, :222-273): outputs of the unmodified reference binary over the tiny index.
Each case is run twice: as given (TSV + --un/--cl files, for the GPU test) and with
`--min-hitlen 5000 --no-dust` so that every read is unclassified -- those outputs depend on the host
plumbing only and are reproduced on the CPU by `centrifuger-b200 --dry-run-output`.
Writes tiny/bc.fq (one technical read per read of se_100.fq: 16-base barcode + 10-base UMI + 2 filler bases) and
tiny/se_com.fq (se_100.fq with CB:Z: / UB:Z: header comments); adds "barcode" to MANIFEST.json.

    python tests/golden/make_golden_barcode.py     (build container: needs oracle/_ref)
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "centrifuger")

# case -> (read files, [(option, file or literal)...]); files are relative to tiny/
CASES = {
    "se_bc_um_files": (["se_100.fq"], ["--barcode", "@bc.fq", "--UMI", "@bc.fq", "--read-format", "bc:0:15,um:16:25"]),
    "se_bc_rc_trim": (["se_100.fq"], ["--barcode", "@bc.fq", "--read-format", "bc:0:15:-,r1:10:-11"]),
    "pe_inline": (["pe_100_1.fq", "pe_100_2.fq"], ["--read-format", "r1:26:-1,bc:0:15,um:16:25"]),
    "se_header": (["se_com.fq"], ["--read-format", "bc:hd:0:5:-1,um:hd:UB:5:-1"]),
    "pe_segments_strand": (["pe_100_1.fq", "pe_100_2.fq"], ["--read-format", "r1:0:49,r1:60:-1,r2:0:-1:-", "-k", "3"]),
    "se_um_only": (["se_100.fq"], ["--UMI", "@bc.fq", "--read-format", "um:-12:-3;r1:5:80:+"]),
    "fa_inline_um": (["se_100.fa"], ["--read-format", "r1:5:80:+,um:0:4"]),
    "pe_unsorted_segments": (["pe_100_1.fq", "pe_100_2.fq"], ["--read-format", "r1:50:-1,r1:0:30,r2:20:90,bc:hd:1:0:-1"]),
    # BarcodeCorrector / BarcodeTranslator: bc_whitelist.txt lacks one of the seven barcodes (-> "N") and holds
    # close variants of another (several one-substitution candidates: counts, then base qualities decide)
    "se_whitelist": (["se_100.fq"], ["--barcode", "@bc.fq", "--read-format", "bc:0:15", "--barcode-whitelist", "@bc_whitelist.txt"]),
    "se_whitelist_translate": (["se_100.fq"], ["--barcode", "@bc.fq", "--UMI", "@bc.fq", "--read-format", "bc:0:15,um:16:25",
                                               "--barcode-whitelist", "@bc_whitelist.txt", "--barcode-translate", "@bc_translate.tsv"]),
    "se_translate_only": (["se_100.fq"], ["--barcode", "@bc.fq", "--read-format", "bc:0:15", "--barcode-translate", "@bc_translate_all.tsv"]),
}


def write_inputs(tg):
    rng = np.random.default_rng(123)
    recs = open(os.path.join(tg, "se_100.fq")).read().split("\n")
    with open(os.path.join(tg, "bc.fq"), "w") as fb, open(os.path.join(tg, "se_com.fq"), "w") as fc:
        pool = ["".join(rng.choice(list("ACGT"), size=16)) for _ in range(7)]
        seen = set()
        for i in range(len(recs) // 4):
            name = recs[4 * i][1:].split()[0]
            bc = pool[int(rng.integers(len(pool)))]
            if i % 11 == 0:
                bc = bc[:5] + "N" + bc[6:]
            umi = "".join(rng.choice(list("ACGT"), size=10))
            q = "".join(rng.choice(list("#5?FI"), size=28))
            seen.add(bc)
            fb.write("@%s\n%s%sTT\n+\n%s\n" % (name, bc, umi, q))
            comment = "CB:Z:%s\tUB:Z:%s" % (bc, umi) if i % 13 else ("UB:Z:%s" % umi if i % 2 else "")
            fc.write("@%s%s\n%s\n+\n%s\n" % (name, (" " + comment) if comment else "", recs[4 * i + 1], recs[4 * i + 3]))
    listed = pool[:6]  # pool[6] is not on the list
    for b in "ACGT":   # close variants of pool[0] at the position where every 11th barcode carries an N
        v = pool[0][:5] + b + pool[0][6:]
        if v not in listed:
            listed.append(v)
    listed.append(pool[1][:9] + ("A" if pool[1][9] != "A" else "C") + pool[1][10:])
    with open(os.path.join(tg, "bc_whitelist.txt"), "w") as f:
        f.write("".join(b + "\n" for b in listed))
    with open(os.path.join(tg, "bc_translate.tsv"), "w") as f:
        f.write("".join("CELL%03d\t%s\n" % (i, b) for i, b in enumerate(listed)))
    with open(os.path.join(tg, "bc_translate_all.tsv"), "w") as f:  # without a whitelist every barcode as read must be listed
        f.write("".join("RAW%03d,%s\n" % (i, b) for i, b in enumerate(sorted(seen))))


def main():
    tg = os.path.join(HERE, "tiny")
    write_inputs(tg)
    d = tempfile.mkdtemp(prefix="cfr_golden_bc_")
    for f in os.listdir(tg):
        if f.startswith("idx.") and f.endswith(".cfr.gz"):
            with gzip.open(os.path.join(tg, f), "rb") as fi, open(os.path.join(d, f[:-3]), "wb") as fo:
                shutil.copyfileobj(fi, fo)
    os.makedirs(os.path.join(tg, "barcode"), exist_ok=True)
    section = {}
    for name, (files, opts) in CASES.items():
        args = [os.path.join(tg, o[1:]) if o.startswith("@") else o for o in opts]
        entry = {"files": files, "args": opts}
        for mode, extra in (("real", []), ("unclassified", ["--min-hitlen", "5000", "--no-dust"])):
            od = os.path.join(d, name + "_" + mode)
            os.makedirs(od)
            paths = [os.path.join(tg, f) for f in files]
            cmd = [REF, "-x", os.path.join(d, "idx"), "-t", "1"] + args + extra
            cmd += ["-u", paths[0]] if len(paths) == 1 else ["-1", paths[0], "-2", paths[1]]
            cmd += ["--un", os.path.join(od, "un"), "--cl", os.path.join(od, "cl")]
            out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
            outs = {f: hashlib.md5(gzip.open(os.path.join(od, f), "rb").read()).hexdigest() for f in sorted(os.listdir(od))}
            with open(os.path.join(tg, "barcode", "%s__%s.tsv" % (name, mode)), "wb") as f:
                f.write(out)
            entry[mode] = {"tsv_md5": hashlib.md5(out).hexdigest(), "outputs": outs}
        section[name] = entry
    mp = os.path.join(HERE, "MANIFEST.json")
    manifest = json.load(open(mp))
    manifest["barcode"] = section
    with open(mp, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    shutil.rmtree(d)
    print("wrote", len(section), "cases")


if __name__ == "__main__":
    main()
