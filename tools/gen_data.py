#!/usr/bin/env python
"""Deterministic synthetic genomes / taxonomy / reads.

Genomes: S species x T strains; a strain = the species base sequence (uniform
ACGT) with `div` substitutions; optional conserved blocks shared genus-wide and
collection-wide so that reads tie across many sequences (exercises the >40*k
row stride of Classifier.hpp:635-666 and Taxonomy LCA / ReduceTaxIds).
Taxonomy: root(1,"no rank") -> superkingdom(2) -> family(5000+f) ->
genus(1000+g) -> species(10+s) -> strain(100000+...) in NCBI nodes/names.dmp
form, plus a seqid->taxid conversion table, exactly the three inputs the
reference's `centrifuger-build` takes (CentrifugerBuild.cpp:10-17).
Reads: uniform over sequences / positions / strands, substitution errors,
qualities 'I'; paired-end: insert ~N(300,30), mate 2 reverse-complemented.
"""
import argparse
import os

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.full(256, ord("N"), dtype=np.uint8)
for a, b in zip(b"ACGT", b"TGCA"):
    COMP[a] = b


def mutate(rng, seq, rate):
    out = seq.copy()
    n = len(seq)
    k = rng.binomial(n, rate)
    if k:
        pos = rng.choice(n, size=k, replace=False)
        out[pos] = (out[pos] + rng.integers(1, 4, size=k)) % 4
    return out


def make_genomes(species, strains, length, div=0.01, seed=1, conserved=0, genus_size=2,
                 family_size=2):
    """Returns list of (name, taxid, codes[uint8 0..3]) and the taxonomy tables."""
    rng = np.random.default_rng(seed)
    universal = rng.integers(0, 4, size=conserved, dtype=np.uint8) if conserved else None
    n_genus = (species + genus_size - 1) // genus_size
    genus_block = [rng.integers(0, 4, size=conserved, dtype=np.uint8) if conserved else None
                   for _ in range(n_genus)]
    genomes = []
    nodes = {1: (1, "no rank"), 2: (1, "superkingdom")}
    names = {1: "root", 2: "Bacteria"}
    for s in range(species):
        g = s // genus_size
        f = g // family_size
        sp_tax, ge_tax, fa_tax = 10 + s, 1000 + g, 5000 + f
        nodes[fa_tax] = (2, "family")
        names[fa_tax] = "Family%d" % f
        # every 3rd genus hangs off an unranked clade to exercise "no rank" skipping
        if g % 3 == 2:
            clade = 7000 + g
            nodes[clade] = (fa_tax, "no rank")
            names[clade] = "Clade%d" % g
            nodes[ge_tax] = (clade, "genus")
        else:
            nodes[ge_tax] = (fa_tax, "genus")
        names[ge_tax] = "Genus%d" % g
        nodes[sp_tax] = (ge_tax, "species")
        names[sp_tax] = "Species%d" % s
        base = rng.integers(0, 4, size=length, dtype=np.uint8)
        if conserved:
            p1 = length // 4
            base[p1:p1 + conserved] = universal
            p2 = length // 2
            base[p2:p2 + conserved] = genus_block[g]
        for t in range(strains):
            st_tax = 100000 + s * 100 + t
            rank = "strain" if t % 3 != 2 else "no rank"
            nodes[st_tax] = (sp_tax, rank)
            names[st_tax] = "Species%d strain%d" % (s, t)
            seq = base if t == 0 else mutate(rng, base, div)
            genomes.append(("seq_%d_%d" % (s, t), st_tax, seq))
    return genomes, nodes, names


def write_reference(outdir, genomes, nodes, names):
    os.makedirs(outdir, exist_ok=True)
    with open(os.path.join(outdir, "ref.fa"), "wb") as f:
        for name, _, codes in genomes:
            f.write(b">" + name.encode() + b"\n")
            s = ACGT[codes].tobytes()
            for i in range(0, len(s), 80):
                f.write(s[i:i + 80] + b"\n")
    with open(os.path.join(outdir, "nodes.dmp"), "w") as f:
        for tid in sorted(nodes):
            f.write("%d\t|\t%d\t|\t%s\t|\n" % (tid, nodes[tid][0], nodes[tid][1]))
    with open(os.path.join(outdir, "names.dmp"), "w") as f:
        for tid in sorted(names):
            f.write("%d\t|\t%s\t|\t\t|\tscientific name\t|\n" % (tid, names[tid]))
    with open(os.path.join(outdir, "seqid.map"), "w") as f:
        for name, tax, _ in genomes:
            f.write("%s\t%d\n" % (name, tax))


def _sample(rng, genomes, n, rlen, err):
    lens = np.array([len(g[2]) for g in genomes])
    gi = rng.integers(0, len(genomes), size=n)
    pos = (rng.random(n) * (lens[gi] - rlen)).astype(np.int64)
    out = np.empty((n, rlen), dtype=np.uint8)
    for i in range(n):
        out[i] = genomes[gi[i]][2][pos[i]:pos[i] + rlen]
    errs = rng.random((n, rlen)) < err
    out = np.where(errs, (out + rng.integers(1, 4, size=(n, rlen))) % 4, out).astype(np.uint8)
    return ACGT[out]  # ascii


def revcomp(a):
    return COMP[a[..., ::-1]]


def make_reads_se(genomes, n, rlen, err=0.01, seed=7):
    rng = np.random.default_rng(seed)
    r = _sample(rng, genomes, n, rlen, err)
    flip = rng.random(n) < 0.5
    r[flip] = revcomp(r[flip])
    return r


def make_reads_pe(genomes, n, rlen, err=0.01, seed=11, insert_mu=300, insert_sd=30):
    rng = np.random.default_rng(seed)
    lens = np.array([len(g[2]) for g in genomes])
    gi = rng.integers(0, len(genomes), size=n)
    ins = np.clip(rng.normal(insert_mu, insert_sd, size=n).astype(np.int64), rlen, None)
    ins = np.minimum(ins, lens[gi] - 1)
    pos = (rng.random(n) * (lens[gi] - ins)).astype(np.int64)
    r1 = np.empty((n, rlen), dtype=np.uint8)
    r2 = np.empty((n, rlen), dtype=np.uint8)
    for i in range(n):
        g = genomes[gi[i]][2]
        frag = g[pos[i]:pos[i] + ins[i]]
        r1[i] = frag[:rlen]
        r2[i] = frag[len(frag) - rlen:]
    for r in (r1, r2):
        e = rng.random((n, rlen)) < err
        r[...] = np.where(e, (r + rng.integers(1, 4, size=(n, rlen))) % 4, r)
    a1 = ACGT[r1]
    a2 = revcomp(ACGT[r2])
    flip = rng.random(n) < 0.5
    a1f = np.where(flip[:, None], a2, a1)
    a2f = np.where(flip[:, None], a1, a2)
    return a1f, a2f


def concat_genomes(genomes):
    """(codes of all genomes back to back, start offset of each, lengths)"""
    lens = np.array([len(g[2]) for g in genomes], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return np.concatenate([g[2] for g in genomes]), starts, lens


def make_reads_se_fast(genomes, n, rlen, err=0.01, seed=7, cat=None):
    """Vectorised variant of make_reads_se for bench-sized batches (same model,
    different random stream).  Returns a (n, rlen) uint8 ASCII array."""
    rng = np.random.default_rng(seed)
    allg, starts, lens = cat if cat is not None else concat_genomes(genomes)
    gi = rng.integers(0, len(lens), size=n)
    pos = starts[gi] + (rng.random(n) * (lens[gi] - rlen)).astype(np.int64)
    out = allg[pos[:, None] + np.arange(rlen, dtype=np.int64)[None, :]]
    e = rng.random((n, rlen)) < err
    out = np.where(e, (out + rng.integers(1, 4, size=(n, rlen), dtype=np.uint8)) % 4, out).astype(np.uint8)
    out = ACGT[out]
    flip = rng.random(n) < 0.5
    out[flip] = revcomp(out[flip])
    return out


def make_reads_pe_fast(genomes, n, rlen, err=0.01, seed=11, insert_mu=300, insert_sd=30, cat=None):
    rng = np.random.default_rng(seed)
    allg, starts, lens = cat if cat is not None else concat_genomes(genomes)
    gi = rng.integers(0, len(lens), size=n)
    ins = np.clip(rng.normal(insert_mu, insert_sd, size=n).astype(np.int64), rlen, None)
    ins = np.minimum(ins, lens[gi] - 1)
    pos = starts[gi] + (rng.random(n) * (lens[gi] - ins)).astype(np.int64)
    ar = np.arange(rlen, dtype=np.int64)[None, :]
    r1 = allg[pos[:, None] + ar]
    r2 = allg[(pos + ins - rlen)[:, None] + ar]
    outs = []
    for r in (r1, r2):
        e = rng.random((n, rlen)) < err
        outs.append(np.where(e, (r + rng.integers(1, 4, size=(n, rlen), dtype=np.uint8)) % 4, r).astype(np.uint8))
    a1 = ACGT[outs[0]]
    a2 = revcomp(ACGT[outs[1]])
    flip = rng.random(n) < 0.5
    return np.where(flip[:, None], a2, a1), np.where(flip[:, None], a1, a2)


def write_fastq(path, reads, prefix="r", suffix=""):
    """reads: 2-D uint8 ascii array or list of bytes."""
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            s = r.tobytes() if hasattr(r, "tobytes") else bytes(r)
            f.write(b"@%s%d%s\n%s\n+\n%s\n" % (prefix.encode(), i, suffix.encode(), s, b"I" * len(s)))


def edge_case_reads(genomes, seed=3):
    """Hand-made reads exercising the parity traps of SURVEY.md appendix B."""
    rng = np.random.default_rng(seed)
    g0 = ACGT[genomes[0][2]].tobytes()
    g1 = ACGT[genomes[-1][2]].tobytes()
    L = len(g0)
    rd = lambda n: ACGT[rng.integers(0, 4, size=n)].tobytes()
    rc = lambda s: COMP[np.frombuffer(s, dtype=np.uint8)[::-1]].tobytes()
    reads = []
    reads.append(g0[100:200])                                   # perfect forward
    reads.append(rc(g0[300:400]))                               # perfect reverse
    reads.append(g0[500:540] + b"N" + g0[541:600])              # N in the middle
    reads.append(b"N" * 100)                                    # all N
    reads.append(g0[700:722])                                   # 22 bp: shorter than min-hitlen
    reads.append(g0[700:723])                                   # exactly 23
    reads.append(g0[700:724])
    reads.append(g0[800:809])                                   # shorter than the lookup width
    reads.append(b"A")                                          # 1 bp
    reads.append(g0[900:1000].lower())                          # lowercase: never matches
    reads.append(g0[1000:1050] + g0[1050:1100].lower())         # half lowercase
    reads.append(rd(100))                                       # random: unclassified
    reads.append(g0[1200:1250] + g1[1300:1350])                 # chimera of two genomes
    reads.append(g0[1400:1450] + rc(g1[1500:1550]))             # chimera, opposite strands
    reads.append(b"A" * 100)                                    # homopolymer (DUST)
    reads.append(b"AC" * 50)                                    # dinucleotide repeat (DUST)
    reads.append(g0[1600:1650] + b"ACG" * 17)                   # half low complexity
    reads.append(g0[1700:1730] + b"T" * 40 + g0[1770:1800])     # low complexity island
    reads.append(b"NNNNN" + g0[1800:1895])                      # leading N
    reads.append(g0[1900:1995] + b"NNNNN")                      # trailing N
    reads.append(g0[2000:2100][:50] + b"R" + g0[2051:2100])     # IUPAC code
    reads.append(g0[L // 4 + 10: L // 4 + 110])                 # inside the universal block (if any)
    reads.append(rc(g0[L // 2 + 10: L // 2 + 110]))             # inside the genus block (if any)
    reads.append(g0[L // 4 - 50: L // 4 + 50])                  # straddles unique / universal
    reads.append(g0[0:100])                                     # genome start (boundary rows)
    reads.append(g0[L - 100:L])                                 # genome end
    reads.append(g0[L - 50:L] + g1[0:50])                       # fake junction
    reads.append(g0[3000:3250])                                 # 250 bp long read
    reads.append(g0[3000:3100] + rd(1) + g0[3101:3200] + rd(1) + g0[3201:3300])  # adjacent unique hits
    m = bytearray(g0[4000:4100])
    for p in (24, 49, 74):
        m[p] = ord("ACGT"[("ACGT".index(chr(m[p])) + 1) % 4])
    reads.append(bytes(m))                                      # 3 SNPs -> 4 short hits
    return reads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--species", type=int, default=10)
    ap.add_argument("--strains", type=int, default=5)
    ap.add_argument("--length", type=int, default=2_000_000)
    ap.add_argument("--div", type=float, default=0.01)
    ap.add_argument("--conserved", type=int, default=0)
    ap.add_argument("--se", type=int, default=0)
    ap.add_argument("--pe", type=int, default=0)
    ap.add_argument("--rlen", type=int, default=100)
    ap.add_argument("--edge", action="store_true")
    ap.add_argument("--reads-only", action="store_true")
    a = ap.parse_args()
    genomes, nodes, names = make_genomes(a.species, a.strains, a.length, a.div, 1, a.conserved)
    if not a.reads_only:
        write_reference(a.out, genomes, nodes, names)
    os.makedirs(a.out, exist_ok=True)
    if a.se:
        write_fastq(os.path.join(a.out, "se_%d.fq" % a.rlen), make_reads_se(genomes, a.se, a.rlen))
    if a.pe:
        r1, r2 = make_reads_pe(genomes, a.pe, a.rlen)
        write_fastq(os.path.join(a.out, "pe_%d_1.fq" % a.rlen), r1, suffix="/1")
        write_fastq(os.path.join(a.out, "pe_%d_2.fq" % a.rlen), r2, suffix="/2")
    if a.edge:
        e = edge_case_reads(genomes)
        write_fastq(os.path.join(a.out, "edge.fq"), e, prefix="e")
        # pair every edge read with the next one, reverse-complemented or not
        e2 = e[1:] + e[:1]
        write_fastq(os.path.join(a.out, "edge_1.fq"), e, prefix="e", suffix="/1")
        write_fastq(os.path.join(a.out, "edge_2.fq"), e2, prefix="e", suffix="/2")


if __name__ == "__main__":
    main()
