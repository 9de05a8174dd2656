"""Host-side ingest of the drop-in CLI (no GPU): FASTA/FASTQ parsing with kseq semantics,
gz input, CRLF, multi-line FASTA, /1 /2 suffix stripping, interleaved input, multiple files."""
import gzip
import os
import subprocess

import pytest

import centrifuger_b200 as cb
from oracle_binding import read_fastx

EXE = os.path.join(os.path.dirname(cb.LIB_PATH), "centrifuger-b200")


def _dry(args):
    r = subprocess.run([EXE, "--dry-run"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    return [tuple(l.split("\t")) for l in r.stdout.decode().split("\n") if l]


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(EXE):
        from centrifuger_b200 import build
        build.build()


def test_fastq_pairs_match_reference_reader(tiny_dir):
    f1, f2 = os.path.join(tiny_dir, "pe_100_1.fq"), os.path.join(tiny_dir, "pe_100_2.fq")
    ids, r1 = read_fastx(f1)
    _, r2 = read_fastx(f2)
    got = _dry(["-1", f1, "-2", f2])
    assert got == [(i, a.decode(), b.decode()) for i, a, b in zip(ids, r1, r2)]


def test_formats(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">s1 comment here\nACGT\nACG\n>s2/1\r\nTTTT\r\n>s3\n\n>s4\nGG")
    assert _dry(["-u", str(fa)]) == [("s1", "ACGTACG", ""), ("s2", "TTTT", ""), ("s3", "", ""), ("s4", "GG", "")]
    fq = tmp_path / "q.fq.gz"
    with gzip.open(fq, "wb") as f:
        f.write(b"@r1/1 x\nACGTN\n+\n@@@+I\n@r2\nAC\n+r2\n+@\n")
    assert _dry(["-u", str(fq)]) == [("r1", "ACGTN", ""), ("r2", "AC", "")]
    # two files back to back, and interleaved input
    assert [x[0] for x in _dry(["-u", str(fa), "-u", str(fq)])] == ["s1", "s2", "s3", "s4", "r1", "r2"]
    inter = tmp_path / "i.fq"
    inter.write_bytes(b"@p1/1\nAAAA\n+\nIIII\n@p1/2\nCCCC\n+\nIIII\n@p2/1\nGG\n+\nII\n@p2/2\nTT\n+\nII\n")
    assert _dry(["-i", str(inter)]) == [("p1", "AAAA", "CCCC"), ("p2", "GG", "TT")]


def test_mate_count_mismatch_is_an_error(tmp_path):
    a = tmp_path / "a.fq"
    b = tmp_path / "b.fq"
    a.write_bytes(b"@x\nAC\n+\nII\n@y\nAC\n+\nII\n")
    b.write_bytes(b"@x\nAC\n+\nII\n")
    r = subprocess.run([EXE, "--dry-run", "-1", str(a), "-2", str(b)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"different number of reads" in r.stderr


def test_usage_and_version_and_rejections():
    r = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and b"-x FILE: index prefix" in r.stderr  # CentrifugerClass.cpp:347-351
    assert subprocess.run([EXE, "-v"], stdout=subprocess.PIPE).stdout.decode().strip() == "Centrifuger v1.1.3-r347"
    r = subprocess.run([EXE, "-x", "nope", "-u", "x.fq", "--no-such-option"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"-x FILE: index prefix" in r.stderr  # unknown option: usage, failure (CentrifugerClass.cpp:541-545)
    r = subprocess.run([EXE, "-u", "x.fq"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"Need to use -x" in r.stderr


def _rc(s):
    return bytes({65: 84, 67: 71, 71: 67, 84: 65}.get(c, 78) for c in reversed(s))


def test_merge_readpair_port_matches_reference_merger(tmp_path):
    """--merge-readpair: the CLI's port of ReadPairMerger::Merge against the UNMODIFIED reference header
    (oracle/_ref/merge_ref, built by oracle/Makefile) on pairs with every kind of overlap: read-through
    (insert shorter than the reads), plain overlaps with mismatches and quality conflicts, tandem
    repeats near the minimum overlap, no overlap, Ns; FASTQ and FASTA."""
    import random
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "merge_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/merge_ref not built")
    rng = random.Random(41)
    pairs = []
    for it in range(1500):
        rl = rng.choice([50, 75, 100, 150])
        mode = rng.random()
        if mode < 0.25:
            ins = rng.randint(20, rl)                       # read-through: adapters behind the insert
        elif mode < 0.75:
            ins = rng.randint(rl, 2 * rl + 10)              # overlap of 0..rl bases
        else:
            ins = rng.randint(2 * rl + 10, 3 * rl)          # no overlap
        if rng.random() < 0.15:                             # tandem repeat fragment
            unit = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 6)))
            frag = (unit * (ins // len(unit) + 2))[:ins]
        else:
            frag = bytes(rng.choice(b"ACGT") for _ in range(ins))
        adapter = bytes(rng.choice(b"ACGT") for _ in range(rl))
        r1 = bytearray((frag + adapter)[:rl])
        r2 = bytearray((_rc(frag) + adapter)[:rl])
        for r in (r1, r2):
            for p in range(len(r)):
                x = rng.random()
                if x < 0.02:
                    r[p] = rng.choice(b"ACGT")
                elif x < 0.025:
                    r[p] = ord("N")
        q1 = bytes(rng.choice(b"#5?FI") for _ in range(len(r1)))
        q2 = bytes(rng.choice(b"#5?FI") for _ in range(len(r2)))
        pairs.append((bytes(r1), q1, bytes(r2), q2))
    for fastq in (True, False):
        f1, f2 = tmp_path / ("m_1.f%s" % ("q" if fastq else "a")), tmp_path / ("m_2.f%s" % ("q" if fastq else "a"))
        with open(f1, "wb") as a, open(f2, "wb") as b:
            for i, (r1, q1, r2, q2) in enumerate(pairs):
                if fastq:
                    a.write(b"@p%d/1\n%s\n+\n%s\n" % (i, r1, q1))
                    b.write(b"@p%d/2\n%s\n+\n%s\n" % (i, r2, q2))
                else:
                    a.write(b">p%d/1\n%s\n" % (i, r1))
                    b.write(b">p%d/2\n%s\n" % (i, r2))
        got = subprocess.run([EXE, "--dry-run", "--merge-readpair", "-1", str(f1), "-2", str(f2)], stdout=subprocess.PIPE,
                             check=True).stdout
        inp = b"".join(b"%s\t%s\t%s\t%s\n" % (r1, q1 if fastq else b"-", r2, q2 if fastq else b"-") for r1, q1, r2, q2 in pairs)
        exp = subprocess.run([ref], input=inp, stdout=subprocess.PIPE, check=True).stdout
        assert got == exp
        codes = [l.split(b"\t")[0] for l in exp.splitlines()]
        assert codes.count(b"1") > 100 and codes.count(b"2") > 100 and codes.count(b"0") > 100


def _pipe(args):
    r = subprocess.run([EXE, "--dry-run-pipeline"] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    return r.returncode, [tuple(l.split("\t")) for l in r.stdout.decode().split("\n") if l], r.stderr.decode()


def test_threaded_ingest_stage_matches_serial_parse(tiny_dir, tmp_path):
    """the batches the ingest stage hands to the GPU stage (mate 2 parsed by a trailing second thread,
    batches cut at --batch reads) hold exactly the records of the serial parse, in order"""
    f1, f2 = os.path.join(tiny_dir, "pe_100_1.fq"), os.path.join(tiny_dir, "pe_100_2.fq")
    ref = _dry(["-1", f1, "-2", f2])
    for batch in ("1", "7", "64", "300", "1048576"):
        rc, got, _ = _pipe(["--batch", batch, "-1", f1, "-2", f2])
        assert rc == 0 and got == ref, batch
    se = os.path.join(tiny_dir, "edge.fq")
    rc, got, _ = _pipe(["--batch", "5", "-u", se])
    assert rc == 0 and got == [(a, b, "") for a, b, _ in _dry(["-u", se])]
    # interleaved input and gz input
    il = tmp_path / "il.fq.gz"
    with gzip.open(il, "wb") as f:
        for (i, a, b) in ref[:50]:
            f.write(("@%s/1\n%s\n+\n%s\n@%s/2\n%s\n+\n%s\n" % (i, a, "I" * len(a), i, b, "I" * len(b))).encode())
    rc, got, _ = _pipe(["--batch", "9", "-i", str(il)])
    assert rc == 0 and got == ref[:50]
    # mate files of different length: an error, whichever file is the short one
    short = tmp_path / "short.fq"
    short.write_bytes(b"".join(b"@%s\n%s\n+\n%s\n" % (i.encode(), b.encode(), b"I" * len(b)) for i, _, b in ref[:120]))
    rc, _, err = _pipe(["--batch", "50", "-1", f1, "-2", str(short)])
    assert rc != 0 and "different number of reads" in err
    rc, _, err = _pipe(["--batch", "50", "-1", str(short), "-2", f2])
    assert rc != 0 and "different number of reads" in err


def test_merge_readpair_in_the_ingest_stage(tiny_dir):
    """--merge-readpair inside the pipeline: a merged pair reaches the GPU stage as (merged read, empty mate)"""
    f1, f2 = os.path.join(tiny_dir, "ov_1.fq"), os.path.join(tiny_dir, "ov_2.fq")
    codes = _dry(["--merge-readpair", "-1", f1, "-2", f2])   # (code, merged read, merged qualities)
    plain = _dry(["-1", f1, "-2", f2])
    rc, got, _ = _pipe(["--merge-readpair", "--batch", "33", "-1", f1, "-2", f2])
    assert rc == 0 and len(got) == len(plain)
    merged = 0
    for c, p, g in zip(codes, plain, got):
        if c[0] != "0":
            assert g == (p[0], c[1], ""), p[0]
            merged += 1
        else:
            assert g == p
    assert merged > 50


def test_batches_end_early_when_a_long_read_arrives(tiny_dir, tmp_path):
    """hit-table budget: reads x (longest read / 24 + 1) per batch is bounded, so a long read among short
    ones cuts the batch; the records and their order do not change"""
    from conftest import golden_path
    _, short = read_fastx(os.path.join(tiny_dir, "se_100.fq"))
    _, long_ = read_fastx(golden_path("tiny", "long.fa"))
    mixed = tmp_path / "mixed.fa"
    with open(mixed, "wb") as f:
        for i, s in enumerate(short[:60] + long_[:3] + short[60:120] + long_[3:5] + short[120:]):
            f.write(b">m%d\n%s\n" % (i, s))
    ref = _dry(["-u", str(mixed)])
    assert len(ref) == len(short) + 5
    for budget in ("1", "200", "5000", "100000"):
        env = dict(os.environ, CFR_B200_SLOT_BUDGET=budget)
        r = subprocess.run([EXE, "--dry-run-pipeline", "-u", str(mixed)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           timeout=120, env=env)
        got = [tuple(l.split("\t")) for l in r.stdout.decode().split("\n") if l]
        assert r.returncode == 0 and got == [(a, b, "") for a, b, _ in ref], budget


def _write_sheet(path, rows, outdir):
    from conftest import golden_path
    with open(path, "w") as f:
        for row in rows:
            r1, r2, o = row[:3]
            bc, um = (row[3], row[4]) if len(row) > 3 else (".", ".")
            f.write("%s %s %s %s %s\n" % (golden_path("tiny", r1), r2 if r2 == "." else golden_path("tiny", r2),
                                          bc if bc == "." else golden_path("tiny", bc), um if um == "." else golden_path("tiny", um),
                                          os.path.join(str(outdir), o + ".tsv")))


def test_sample_sheet_routes_reads_to_their_output_files(manifest, tmp_path):
    """--sample-sheet on the ingest and output stages (no GPU: every read is printed as unclassified): each
    input file's rows land in its row's output file, a file named twice is appended to without a second
    header, nothing goes to stdout -- the read ids per file are those of the reference binary's outputs"""
    from conftest import golden_path
    for case, m in manifest["sample_sheet"].items():
        od = tmp_path / case
        od.mkdir()
        sheet = tmp_path / (case + ".sheet")
        _write_sheet(sheet, m["rows"], od)
        for batch in ("1048576", "37"):
            for f in os.listdir(str(od)):
                os.remove(str(od / f))
            r = subprocess.run([EXE, "--dry-run-output", "--batch", batch, "--sample-sheet", str(sheet)] + m["args"],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
            assert r.returncode == 0 and r.stdout == b"", r.stderr.decode()
            assert sorted(os.listdir(str(od))) == sorted(m["outputs"])
            for o in m["outputs"]:
                exp = open(golden_path("tiny", "sheet", "%s__%s" % (case, o))).read().split("\n")
                got = open(str(od / o)).read().split("\n")
                assert got[0] == exp[0] and sum(1 for l in got if l.startswith("readID")) == 1
                exp_reads = []
                for l in exp[1:]:
                    if l and (not exp_reads or exp_reads[-1][0] != l.split("\t")[0]):  # -k > 1: several rows per read
                        exp_reads.append((l.split("\t")[0], l.split("\t")[6]))
                assert [(l.split("\t")[0], l.split("\t")[6]) for l in got[1:] if l] == exp_reads, (case, o, batch)
                if "barcode" in exp[0]:  # the barcode / UMI columns of every read
                    cols = lambda ls: sorted(set((l.split("\t")[0],) + tuple(l.split("\t")[8:10]) for l in ls[1:] if l))
                    assert cols(got) == cols(exp), (case, o)
    # rows that ask for what this build does not do are refused
    bad = tmp_path / "bad.sheet"
    bad.write_text("%s . . . %s\n%s %s . . %s\n" % (golden_path("tiny", "se_100.fq"), tmp_path / "x.tsv", golden_path("tiny", "pe_100_1.fq"),
                                                   golden_path("tiny", "pe_100_2.fq"), tmp_path / "y.tsv"))
    r = subprocess.run([EXE, "--dry-run-output", "--sample-sheet", str(bad)], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"not supported" in r.stderr
    r = subprocess.run([EXE, "--dry-run-output", "--sample-sheet", str(tmp_path / "missing.sheet")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"Cannot open the sample sheet" in r.stderr


def test_glob_in_file_names(tiny_dir):
    """a '*' in a read file name stands for the matching files in glob order (ReadFiles.hpp:135-172)"""
    ref = _dry(["-1", os.path.join(tiny_dir, "edge_1.fq"), "-1", os.path.join(tiny_dir, "pe_100_1.fq"),
                "-2", os.path.join(tiny_dir, "edge_2.fq"), "-2", os.path.join(tiny_dir, "pe_100_2.fq")])
    got = _dry(["-1", os.path.join(tiny_dir, "*e*_1.fq"), "-2", os.path.join(tiny_dir, "*e*_2.fq")])
    assert got == ref and len(ref) > 300


def test_read_format_barcode_umi_plumbing(manifest, tmp_path):
    """--read-format / --barcode / --UMI on the ingest and output stages: with every read unclassified the TSV
    (barcode / UMI columns, query lengths after the read stretches are cut) and the --un files (reads and
    qualities as cut, _bc / _um FASTA files) are byte for byte the reference binary's"""
    import hashlib
    from conftest import golden_path
    for name, m in sorted(manifest["barcode"].items()):
        files = [golden_path("tiny", f) for f in m["files"]]
        args = [golden_path("tiny", o[1:]) if o.startswith("@") else o for o in m["args"]]
        for batch in ("1048576", "29"):
            od = tmp_path / (name + "_" + batch)
            od.mkdir()
            cmd = [EXE, "--dry-run-output", "--batch", batch] + args
            cmd += ["-u", files[0]] if len(files) == 1 else ["-1", files[0], "-2", files[1]]
            cmd += ["--un", str(od / "un"), "--cl", str(od / "cl")]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
            assert r.returncode == 0, r.stderr.decode()
            exp = open(golden_path("tiny", "barcode", name + "__unclassified.tsv")).read()
            assert r.stdout.decode() == exp, (name, batch)
            got = {f: hashlib.md5(gzip.open(str(od / f), "rb").read()).hexdigest() for f in sorted(os.listdir(str(od)))}
            assert got == m["unclassified"]["outputs"], (name, batch)
    r = subprocess.run([EXE, "--dry-run-output", "-u", "x.fq", "--read-format", "r3:0:1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"Format description error in r3:0:1" in r.stderr  # ReadFormatter.hpp:207-211
    short = tmp_path / "short_bc.fq"
    short.write_text("@a\nACGT\n+\nIIII\n")
    r = subprocess.run([EXE, "--dry-run-output", "-u", golden_path("tiny", "se_100.fq"), "--barcode", str(short)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"barcode file and read file have different number of reads" in r.stderr


def test_barcode_whitelist_and_translate_refusals(tmp_path):
    from conftest import golden_path
    se, wl = golden_path("tiny", "se_100.fq"), golden_path("tiny", "bc_whitelist.txt")
    # a whitelist needs a barcode file to learn the barcode frequencies from (CentrifugerClass.cpp:565-574)
    r = subprocess.run([EXE, "--dry-run-output", "-u", se, "--read-format", "bc:0:15", "--barcode-whitelist", wl],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"Barcode whitelist has to be used with --barcode option" in r.stderr
    # a barcode the translation table does not know ends the run (BarcodeTranslator.hpp:66-72)
    r = subprocess.run([EXE, "--dry-run-output", "-u", se, "--barcode", golden_path("tiny", "bc.fq"), "--read-format", "bc:0:15",
                        "--barcode-translate", golden_path("tiny", "bc_translate.tsv")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"does not exist in the translation table." in r.stderr
