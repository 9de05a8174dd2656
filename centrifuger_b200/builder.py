"""Index construction: host side above the builder's C ABI (include/centrifuger_b200_build.h).

Mirror of the reference's `Builder<Sequence_RunBlock>::Build` / `Save` (Builder.hpp:87-312) for
nucleotide collections with the default options of `centrifuger-build`
(CentrifugerBuild.cpp:10-27): `-r ref.fa --taxonomy-tree nodes.dmp --name-table names.dmp
--conversion-table seqid.map -o PREFIX [--offrate N] [--ftabchars N] [--rbbwt-b N]`.

    <prefix>.1.cfr   FM index       -- written by the GPU builder (csrc/cfr_build.cu)
    <prefix>.2.cfr   taxonomy       -- Taxonomy::Init + Save (Taxonomy.hpp:476-485, :1238-1257)
    <prefix>.3.cfr   sequence lengths {seqId u64, length u64}, ascending seqId (Builder.hpp:292-300)
    <prefix>.4.cfr   version / sample rate / sequence type / build date (Builder.hpp:265-277)

Not covered (the reference options outside this repo's scope): protein collections, --subset-tax,
--concat-same-tax, file-level conversion tables, checkpoints.
"""
import ctypes as C
import os
import struct
import time

import numpy as np

RANKS = ["no rank", "strain", "species", "genus", "family", "order", "class", "phylum", "kingdom", "domain",
         "forma", "infraclass", "infraorder", "parvorder", "subclass", "subfamily", "subgenus", "subkingdom",
         "suborder", "subphylum", "subspecies", "subtribe", "superclass", "superfamily", "superkingdom",
         "superorder", "superphylum", "tribe", "varietas", "life", "acellular root"]  # Taxonomy.hpp:25-58
RANK_ID = {r: i for i, r in enumerate(RANKS)}
VERSION = "v1.1.3-r347"  # defs.h:8 (the files are the reference's format, written for its binaries too)


class BuildParams(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("precompute_width", C.c_int32), ("rbbwt_b", C.c_uint64),
                ("max_batch_rows", C.c_uint64), ("device", C.c_int32), ("verbose", C.c_int32)]


class BuildStats(C.Structure):
    _fields_ = [("batches", C.c_uint64), ("max_tie_depth_bases", C.c_uint64), ("sort_seconds", C.c_double),
                ("derive_seconds", C.c_double), ("runblock_seconds", C.c_double), ("write_seconds", C.c_double)]


class BuildError(RuntimeError):
    pass


def _lib(lib=None):
    if lib is not None:
        return lib
    from . import load_library
    return load_library()


def _bind(L):
    L.cfr_build_last_error.restype = C.c_char_p
    L.cfr_build_default_params.argtypes = [C.POINTER(BuildParams)]
    L.cfr_build_fm_index.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                     C.POINTER(BuildParams), C.c_char_p, C.POINTER(BuildStats)]
    L.cfr_build_synthetic_fm_index.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                               C.POINTER(BuildParams), C.c_char_p, C.POINTER(BuildStats)]
    L.cfr_synth_bases.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
    L.cfr_synth_fragments.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                      C.c_void_p]
    return L


def _params(L, offrate, ftabchars, rbbwt_b, device, verbose, max_batch_rows):
    p = BuildParams()
    L.cfr_build_default_params(C.byref(p))
    p.sample_rate = 1 << offrate
    p.precompute_width = ftabchars
    p.rbbwt_b = rbbwt_b
    p.device = device
    p.verbose = 1 if verbose else 0
    p.max_batch_rows = max_batch_rows
    return p


# ---------------------------------------------------------------------------- taxonomy (.2.cfr)
class Taxonomy:
    """Taxonomy::Init(nodesFile, namesFile, seqIdFile, false) restated (Taxonomy.hpp:476-485)."""

    def __init__(self, nodes, names, seq_to_tax):
        """nodes: {taxid: (parent taxid, rank string)}, names: {taxid: scientific name},
        seq_to_tax: ordered list of (sequence name, taxid) as the conversion table lists them."""
        # ReadTaxonomyTree (:146-240): keep the nodes on a path from a present taxid to the top; compact ids are
        # the ranks of the original ids in ascending order (std::map iteration)
        selected = set()
        for _, t in seq_to_tax:
            p = t
            if p not in nodes:
                continue
            while p not in selected:
                selected.add(p)
                p = nodes[p][0]
                if p not in nodes:
                    break
        self.orig = sorted(selected)
        self.compact = {t: i for i, t in enumerate(self.orig)}
        n = len(self.orig)
        self.parent, self.rank, self.leaf = [0] * n, [0] * n, [1] * n
        for i, t in enumerate(self.orig):
            self.rank[i] = RANK_ID.get(nodes[t][1], 0)
        for i, t in enumerate(self.orig):
            par = nodes[t][0]
            if par in self.compact:
                self.parent[i] = self.compact[par]
                self.leaf[self.compact[par]] = 0
            else:
                self.parent[i] = i
        # ReadTaxonomyName (:243-273): blanks inside a name become '_'
        self.name = ["" for _ in range(n)]
        for t, nm in names.items():
            if t in self.compact:
                self.name[self.compact[t]] = "_".join(nm.split())
        # ReadSeqNameFile (:303-368): sequence ids in order of first appearance
        self.seq_name, self.seq_index, self.seq_tax = [], {}, []
        for s, t in seq_to_tax:
            if s in self.seq_index:
                raise BuildError("sequence %s is listed twice in the conversion table (unsupported here)" % s)
            self.seq_index[s] = len(self.seq_name)
            self.seq_name.append(s)
            self.seq_tax.append(self.compact[t] if t in self.compact else n)

    def save(self, path):
        """Taxonomy::Save (:1238-1257)"""
        n = len(self.orig)
        with open(path, "wb") as f:
            f.write(struct.pack("<QQQ", n, len(self.seq_name), 0))
            for i in range(n):  # TaxonomyNode: u64 parentTid, u8 rank, u8 leaf, 6 bytes of padding (:61-82)
                f.write(struct.pack("<QBB6x", self.parent[i], self.rank[i], self.leaf[i]))
            f.write(struct.pack("<Q", n))  # MapID<uint64_t>::Save (MapID.hpp:76-81)
            f.write(np.asarray(self.orig, dtype="<u8").tobytes())
            for nm in self.name:
                b = nm.encode()
                f.write(struct.pack("<Q", len(b)) + b)
            f.write(np.asarray(self.seq_tax, dtype="<u8").tobytes())
            for nm in self.seq_name:
                b = nm.encode()
                f.write(struct.pack("<Q", len(b)) + b)


def read_nodes_dmp(path):
    nodes = {}
    for line in open(path):
        if not line.strip() or line[0] == "#":
            continue
        f = [x.strip() for x in line.split("|")]
        t = int(f[0])
        if t not in nodes:
            nodes[t] = (int(f[1]), " ".join(f[2].split()))
    return nodes


def read_names_dmp(path):
    names = {}
    for line in open(path):
        if "scientific name" not in line or line[0] == "#":
            continue
        f = [x.strip() for x in line.split("|")]
        names[int(f[0])] = f[1]
    return names


def read_seqid_map(path):
    out = []
    for line in open(path):
        if not line.strip() or line[0] == "#":
            continue
        a = line.split()
        out.append((a[0], int(a[1])))
    return out


def read_fasta_codes(path):
    """-> list of (record id, uint8 codes 0..3): what SequenceCompactor::Compact keeps of each record
    (SequenceCompactor.hpp:59-84: characters outside "ACGT" -- lower case included -- are dropped)."""
    import gzip
    lut = np.full(256, 255, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    recs, name, parts = [], None, []
    op = gzip.open if path.endswith(".gz") else open

    def flush():
        if name is not None:
            raw = np.frombuffer(b"".join(parts), dtype=np.uint8)
            codes = lut[raw]
            recs.append((name, codes[codes != 255]))

    with op(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                flush()
                name = line[1:].split()[0].decode() if len(line) > 1 and line[1:].split() else ""
                parts = []
            else:
                parts.append(line.strip())
    flush()
    return recs


def _write_meta(prefix, seq_len, sample_rate):
    with open(prefix + ".3.cfr", "wb") as f:
        for sid in sorted(seq_len):
            f.write(struct.pack("<QQ", sid, seq_len[sid]))
    with open(prefix + ".4.cfr", "w") as f:
        f.write("version\tCentrifuger %s\nSA_sample_rate\t%d\nsequence_type\tnucleotide\nbuild_date\t%s" %
                (VERSION, sample_rate, time.strftime("%c")))


def build_index(fasta_paths, nodes_dmp, names_dmp, seqid_map, out_prefix, offrate=4, ftabchars=10, rbbwt_b=0,
                device=0, verbose=False, max_batch_rows=0, lib=None):
    """Builder::Build + Save for a nucleotide collection.  Returns the builder's statistics."""
    L = _bind(_lib(lib))
    tax = Taxonomy(read_nodes_dmp(nodes_dmp), read_names_dmp(names_dmp), read_seqid_map(seqid_map))
    codes, lens, ids, seq_len = [], [], [], {}
    for path in ([fasta_paths] if isinstance(fasta_paths, str) else fasta_paths):
        for name, c in read_fasta_codes(path):
            if name not in tax.seq_index:
                raise BuildError("sequence %s has no entry in the conversion table (unsupported here)" % name)
            sid = tax.seq_index[name]
            if sid in seq_len:  # Builder.hpp:128-129: a sequence id seen before is skipped
                continue
            if len(c) < ftabchars + 1:  # Builder.hpp:145-151
                continue
            seq_len[sid] = len(c)
            codes.append(c)
            lens.append(len(c))
            ids.append(sid)
    if not codes:
        raise BuildError("found 0 genomes in the input or after filtering")
    text = np.ascontiguousarray(np.concatenate(codes))
    la, ia = np.asarray(lens, dtype=np.uint64), np.asarray(ids, dtype=np.uint64)
    p = _params(L, offrate, ftabchars, rbbwt_b, device, verbose, max_batch_rows)
    st = BuildStats()
    rc = L.cfr_build_fm_index(text.ctypes.data, len(text), la.ctypes.data, ia.ctypes.data, len(la), C.byref(p),
                              (out_prefix + ".1.cfr").encode(), C.byref(st))
    if rc != 0:
        raise BuildError("cfr_build_fm_index: %s" % L.cfr_build_last_error().decode())
    tax.save(out_prefix + ".2.cfr")
    _write_meta(out_prefix, seq_len, 1 << offrate)
    return st


# ---------------------------------------------------------------------------- synthetic collections
def synthetic_taxonomy(species, strains, genus_size=2, family_size=2):
    """root(1) -> superkingdom(2) -> family -> genus -> species -> strain, one sequence per strain; genome g =
    species * strains + strain has sequence id g (the order cfr_build_synthetic_fm_index lays the text out in)."""
    nodes = {1: (1, "no rank"), 2: (1, "superkingdom")}
    names = {1: "root", 2: "Bacteria"}
    seqs = []
    for s in range(species):
        g = s // genus_size
        f = g // family_size
        sp_t, ge_t, fa_t = 10_000_000 + s, 20_000_000 + g, 30_000_000 + f
        nodes[fa_t] = (2, "family")
        names[fa_t] = "Family%d" % f
        nodes[ge_t] = (fa_t, "genus")
        names[ge_t] = "Genus%d" % g
        nodes[sp_t] = (ge_t, "species")
        names[sp_t] = "Species%d" % s
        for t in range(strains):
            st_t = 100_000_000 + s * strains + t
            nodes[st_t] = (sp_t, "strain")
            names[st_t] = "Species%d strain %d" % (s, t)
            seqs.append(("seq%d_%d" % (s, t), st_t))
    return nodes, names, seqs


def build_synthetic(out_prefix, species, strains, genome_len, div_ppm=10000, seed=1, offrate=4, ftabchars=10,
                    rbbwt_b=0, device=0, verbose=False, max_batch_rows=0, lib=None):
    """A synthetic collection generated and indexed on the device (the 20 / 140 Gbp bench workloads)."""
    L = _bind(_lib(lib))
    p = _params(L, offrate, ftabchars, rbbwt_b, device, verbose, max_batch_rows)
    st = BuildStats()
    rc = L.cfr_build_synthetic_fm_index(species, strains, genome_len, div_ppm, seed, C.byref(p),
                                        (out_prefix + ".1.cfr").encode(), C.byref(st))
    if rc != 0:
        raise BuildError("cfr_build_synthetic_fm_index: %s" % L.cfr_build_last_error().decode())
    nodes, names, seqs = synthetic_taxonomy(species, strains)
    Taxonomy(nodes, names, seqs).save(out_prefix + ".2.cfr")
    _write_meta(out_prefix, {g: genome_len for g in range(species * strains)}, 1 << offrate)
    return st


def synth_bases(species_index, strain_index, offset, count, div_ppm=10000, seed=1, lib=None):
    """codes 0..3 of one stretch of a synthetic genome (host evaluation of the device generator)"""
    L = _bind(_lib(lib))
    out = np.empty(count, dtype=np.uint8)
    L.cfr_synth_bases(species_index, strain_index, offset, count, div_ppm, seed, out.ctypes.data)
    return out


class SyntheticReads:
    """Seeded reads from a synthetic collection: uniform over genomes / positions / strands, substitution
    errors, paired-end insert ~ N(300, 30) -- the model of tools/gen_data.py, drawn from the generator
    itself because the text of a 20 - 140 Gbp collection never exists on the host."""

    def __init__(self, species, strains, genome_len, div_ppm=10000, seed=1, lib=None):
        self.sp, self.st, self.gl, self.div, self.seed = species, strains, genome_len, div_ppm, seed
        self.L = _bind(_lib(lib))

    def _fragments(self, gi, pos, length):
        out = np.empty((len(gi), length), dtype=np.uint8)
        g = np.ascontiguousarray(gi, dtype=np.uint64)
        o = np.ascontiguousarray(pos, dtype=np.uint64)
        self.L.cfr_synth_fragments(g.ctypes.data, o.ctypes.data, len(g), length, self.st, self.div, self.seed,
                                   out.ctypes.data)
        return out

    def pairs(self, n, rlen, seed, err=0.01, insert_mu=300, insert_sd=30):
        """-> (r1, r2) ASCII arrays (n, rlen) and the genome index of each pair"""
        from numpy.random import default_rng
        rng = default_rng(seed)
        acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
        comp = np.array([3, 2, 1, 0], dtype=np.uint8)
        gi = rng.integers(0, self.sp * self.st, size=n)
        ins = np.clip(rng.normal(insert_mu, insert_sd, size=n).astype(np.int64), rlen, self.gl - 1)
        pos = (rng.random(n) * (self.gl - ins)).astype(np.int64)
        maxins = int(ins.max())
        frag = self._fragments(gi, pos, maxins)
        ar = np.arange(rlen)[None, :]
        r1 = frag[np.arange(n)[:, None], ar]
        r2 = frag[np.arange(n)[:, None], (ins - rlen)[:, None] + ar]
        outs = []
        for r in (r1, r2):
            e = rng.random((n, rlen)) < err
            outs.append(np.where(e, (r + rng.integers(1, 4, size=(n, rlen), dtype=np.uint8)) % 4, r).astype(np.uint8))
        a1 = acgt[outs[0]]
        a2 = acgt[comp[outs[1]][:, ::-1]]
        flip = rng.random(n) < 0.5
        return np.where(flip[:, None], a2, a1), np.where(flip[:, None], a1, a2), gi
