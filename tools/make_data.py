#!/usr/bin/env python
"""Build the datasets the tests and bench.py use, under data/ (git-ignored,
shipped to the GPU box with the snapshot).

Indexes are produced by the UNMODIFIED reference builder compiled into
oracle/_ref/centrifuger-build (index construction is out of scope for this
repo -- SURVEY.md 8(f) N3 -- the reference builder is used as a tool).
Everything is seeded and idempotent: a dataset is rebuilt only if missing.
"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_data  # noqa: E402

DATA = os.path.join(ROOT, "data")
REF_BIN = os.path.join(ROOT, "oracle", "_ref")
REF_EXAMPLE = "/root/reference/example"

DATASETS = {
    # name: genome kwargs, builder variants {idx name: extra args}, reads
    "tiny": dict(genomes=dict(species=6, strains=3, length=20000, conserved=300),
                 variants={"idx": ["--ftabchars", "6"],
                           "idx_b1": ["--ftabchars", "6", "--rbbwt-b", "1"],
                           "idx_b8": ["--ftabchars", "6", "--rbbwt-b", "8"],
                           "idx_off3": ["--ftabchars", "5", "--offrate", "3"]},
                 se=(300, 100), pe=(300, 100), edge=True),
    "small": dict(genomes=dict(species=20, strains=5, length=100000, conserved=500),
                  variants={"idx": []}, se=(20000, 100), pe=(20000, 150), edge=True),
    "c2": dict(genomes=dict(species=10, strains=5, length=2000000, conserved=0),
               variants={"idx": []}, se=None, pe=None, edge=False),
    # the largest index whose files fit the 512 MiB gpurun snapshot: its occ lines (350 MB) do not fit L2
    "m700": dict(genomes=dict(species=35, strains=5, length=4000000, conserved=0),
                 variants={"idx": []}, se=None, pe=None, edge=False),
    "c3": dict(genomes=dict(species=100, strains=5, length=4000000, conserved=0),
               variants={"idx": []}, se=None, pe=None, edge=False),
}


# Synthetic collections generated AND indexed on the GPU by this repo's builder (centrifuger_b200/builder.py,
# csrc/cfr_build.cu): the text of a 20 - 140 Gbp collection never exists on the host, and the reference
# builder would need hours and more memory than the box has (FMBuilder.hpp:328-329: n/2 bytes of samples
# as size_t on top of the text and BWT).  The builder's parity with the reference builder is pinned on the
# collections above (tests/test_builder.py: byte-identical files).
SYNTHETIC = {
    "s400": dict(species=20, strains=5, genome_len=4_000_000),      # 400 Mbp: quick checks
    "s2g": dict(species=100, strains=5, genome_len=4_000_000),      # 2 Gbp: the shape of BASELINE configs[2]
    "c4": dict(species=1000, strains=5, genome_len=4_000_000),      # BASELINE configs[3]: 20 Gbp
    "c5": dict(species=7000, strains=5, genome_len=4_000_000),      # BASELINE configs[4]: 140 Gbp
}


def ensure_synthetic(name, log=print, device=0):
    """data/<name>/idx.{1,2,3,4}.cfr built on the GPU; returns the directory (None without a GPU)."""
    d = os.path.join(DATA, name)
    prefix = os.path.join(d, "idx")
    if index_ready(prefix) and os.path.exists(prefix + ".ok"):
        return d
    try:
        import torch
        if not torch.cuda.is_available():
            return None
    except Exception:
        return None
    sys.path.insert(0, ROOT)
    from centrifuger_b200 import builder
    os.makedirs(d, exist_ok=True)
    spec = SYNTHETIC[name]
    t = time.time()
    st = builder.build_synthetic(prefix, spec["species"], spec["strains"], spec["genome_len"], device=device,
                                 verbose=bool(os.environ.get("CFR_BUILD_VERBOSE")))
    with open(prefix + ".ok", "w") as f:
        f.write("batches %d sort %.1f derive %.1f runblock %.1f write %.1f total %.1f\n" %
                (st.batches, st.sort_seconds, st.derive_seconds, st.runblock_seconds, st.write_seconds, time.time() - t))
    log("built data/%s on the GPU in %.1fs (%d batches: sort %.1fs, derive %.1fs, run blocks %.1fs, file %.1fs)" %
        (name, time.time() - t, st.batches, st.sort_seconds, st.derive_seconds, st.runblock_seconds, st.write_seconds))
    return d


def have_builder():
    return os.path.exists(os.path.join(REF_BIN, "centrifuger-build"))


def index_ready(prefix):
    return all(os.path.exists("%s.%d.cfr" % (prefix, i)) for i in (1, 2))


def build_index(refdir, prefix, extra, threads=None):
    threads = threads or min(16, os.cpu_count() or 1)
    cmd = [os.path.join(REF_BIN, "centrifuger-build"), "-r", os.path.join(refdir, "ref.fa"),
           "--taxonomy-tree", os.path.join(refdir, "nodes.dmp"),
           "--name-table", os.path.join(refdir, "names.dmp"),
           "--conversion-table", os.path.join(refdir, "seqid.map"),
           "-o", prefix, "-t", str(threads)] + list(extra)
    t = time.time()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.time() - t


def genomes_of(name):
    g, _, _ = gen_data.make_genomes(seed=1, **DATASETS[name]["genomes"])
    return g


def ensure(name, log=print):
    """Make sure data/<name>/ exists; returns its directory (or None if it cannot be built)."""
    d = os.path.join(DATA, name)
    if name in SYNTHETIC:
        return ensure_synthetic(name, log=log)
    if name == "example":
        prefix = os.path.join(d, "cfr_ref_idx")
        if index_ready(prefix):
            return d
        if not (have_builder() and os.path.isdir(REF_EXAMPLE)):
            return None
        os.makedirs(d, exist_ok=True)
        cmd = [os.path.join(REF_BIN, "centrifuger-build"), "-r", REF_EXAMPLE + "/ref.fa",
               "--taxonomy-tree", REF_EXAMPLE + "/nodes.dmp", "--name-table", REF_EXAMPLE + "/names.dmp",
               "--conversion-table", REF_EXAMPLE + "/ref_seqid.map", "-o", prefix]
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        log("built data/example")
        return d
    spec = DATASETS[name]
    need_idx = [v for v in spec["variants"] if not index_ready(os.path.join(d, v))]
    need_reads = []
    if spec["se"] and not os.path.exists(os.path.join(d, "se_%d.fq" % spec["se"][1])):
        need_reads.append("se")
    if spec["pe"] and not os.path.exists(os.path.join(d, "pe_%d_1.fq" % spec["pe"][1])):
        need_reads.append("pe")
    if spec["edge"] and not os.path.exists(os.path.join(d, "edge.fq")):
        need_reads.append("edge")
    if not need_idx and not need_reads:
        return d
    if need_idx and not have_builder():
        return None
    os.makedirs(d, exist_ok=True)
    genomes, nodes, names = gen_data.make_genomes(seed=1, **spec["genomes"])
    if need_idx:
        gen_data.write_reference(d, genomes, nodes, names)
        for v in need_idx:
            dt = build_index(d, os.path.join(d, v), spec["variants"][v])
            log("built data/%s/%s in %.1fs" % (name, v, dt))
        os.remove(os.path.join(d, "ref.fa"))  # regenerable from the seed; keeps the snapshot small
    if "se" in need_reads:
        n, rl = spec["se"]
        gen_data.write_fastq(os.path.join(d, "se_%d.fq" % rl), gen_data.make_reads_se(genomes, n, rl))
    if "pe" in need_reads:
        n, rl = spec["pe"]
        r1, r2 = gen_data.make_reads_pe(genomes, n, rl)
        gen_data.write_fastq(os.path.join(d, "pe_%d_1.fq" % rl), r1, suffix="/1")
        gen_data.write_fastq(os.path.join(d, "pe_%d_2.fq" % rl), r2, suffix="/2")
    if "edge" in need_reads:
        e = gen_data.edge_case_reads(genomes)
        gen_data.write_fastq(os.path.join(d, "edge.fq"), e, prefix="e")
        gen_data.write_fastq(os.path.join(d, "edge_1.fq"), e, prefix="e", suffix="/1")
        gen_data.write_fastq(os.path.join(d, "edge_2.fq"), e[1:] + e[:1], prefix="e", suffix="/2")
    return d


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["example", "tiny", "small", "c2"]):
        t0 = time.time()
        r = ensure(nm)
        print("%s -> %s (%.1fs)" % (nm, r, time.time() - t0))
