"""ctypes binding to oracle/libcfr_oracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference classifier
(oracle/cfr_oracle.c).  Nothing in centrifuger_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libcfr_oracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


class Param(C.Structure):
    _fields_ = [("max_result", C.c_int32), ("min_hit_len", C.c_int32),
                ("max_result_per_hit_factor", C.c_int32), ("pad_", C.c_int32),
                ("consider_secondary_hit_len", C.c_uint64),
                ("consider_secondary_score_factor", C.c_double)]


class Hit(C.Structure):
    _fields_ = [("sp", C.c_uint64), ("ep", C.c_uint64), ("l", C.c_int32),
                ("strand", C.c_int32), ("offset", C.c_int32), ("pad_", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("score", C.c_uint64), ("secondary_score", C.c_uint64),
                ("hit_length", C.c_int32), ("query_length", C.c_int32),
                ("n", C.c_int32), ("by_rank", C.c_int32),
                ("ids", C.c_uint64 * 64), ("tax_ids", C.c_uint64 * 64)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in
                ("n_rank", "n_access", "n_search", "n_locate", "n_lf", "n_extend")]


def build_oracle():
    """(Re)build the C restatement if missing or stale."""
    src = os.path.join(ORACLE_DIR, "cfr_oracle.c")
    if (not os.path.exists(LIB_PATH)
            or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.cfr_oracle_open.restype = C.c_void_p
        L.cfr_oracle_open.argtypes = [C.c_char_p]
        L.cfr_oracle_close.argtypes = [C.c_void_p]
        L.cfr_oracle_default_param.argtypes = [C.POINTER(Param)]
        L.cfr_oracle_scalar.restype = C.c_uint64
        L.cfr_oracle_scalar.argtypes = [C.c_void_p, C.c_int]
        L.cfr_oracle_bwt_rank.restype = C.c_uint64
        L.cfr_oracle_bwt_rank.argtypes = [C.c_void_p, C.c_char, C.c_uint64, C.c_int]
        L.cfr_oracle_bwt_access.restype = C.c_char
        L.cfr_oracle_bwt_access.argtypes = [C.c_void_p, C.c_uint64]
        L.cfr_oracle_fm_rank.restype = C.c_uint64
        L.cfr_oracle_fm_rank.argtypes = [C.c_void_p, C.c_char, C.c_uint64, C.c_int]
        L.cfr_oracle_backward_search.restype = C.c_uint64
        L.cfr_oracle_backward_search.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64,
                                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.cfr_oracle_locate.restype = C.c_uint64
        L.cfr_oracle_locate.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.cfr_oracle_infer_min_hit_len.restype = C.c_int
        L.cfr_oracle_infer_min_hit_len.argtypes = [C.c_void_p]
        L.cfr_oracle_search.restype = C.c_int
        L.cfr_oracle_search.argtypes = [C.c_void_p, C.POINTER(Param), C.c_char_p, C.c_char_p,
                                        C.POINTER(Hit), C.c_int]
        L.cfr_oracle_query.argtypes = [C.c_void_p, C.POINTER(Param), C.c_char_p, C.c_char_p,
                                       C.POINTER(Result)]
        L.cfr_oracle_seqid_to_taxid.restype = C.c_uint64
        L.cfr_oracle_seqid_to_taxid.argtypes = [C.c_void_p, C.c_uint64]
        L.cfr_oracle_orig_taxid.restype = C.c_uint64
        L.cfr_oracle_orig_taxid.argtypes = [C.c_void_p, C.c_uint64]
        L.cfr_oracle_seq_name.restype = C.c_char_p
        L.cfr_oracle_seq_name.argtypes = [C.c_void_p, C.c_uint64]
        L.cfr_oracle_rank_name.restype = C.c_char_p
        L.cfr_oracle_rank_name.argtypes = [C.c_void_p, C.c_uint64]
        L.cfr_oracle_reduce_taxids.restype = C.c_int
        L.cfr_oracle_reduce_taxids.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int,
                                               C.POINTER(C.c_uint64), C.c_int]
        L.cfr_oracle_reduce_taxids_expanded.restype = C.c_int
        L.cfr_oracle_reduce_taxids_expanded.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int,
                                                        C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_uint64),
                                                        C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int)]
        L.cfr_oracle_query_expanded.restype = C.c_int
        L.cfr_oracle_query_expanded.argtypes = [C.c_void_p, C.POINTER(Param), C.c_char_p, C.c_char_p,
                                                C.POINTER(Result), C.POINTER(C.c_uint64), C.c_int,
                                                C.POINTER(C.c_int32)]
        L.cfr_oracle_format_tsv_expanded.restype = C.c_int
        L.cfr_oracle_format_tsv_expanded.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Result),
                                                     C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.c_char_p,
                                                     C.c_size_t]
        L.cfr_oracle_dust_mask.restype = C.c_int
        L.cfr_oracle_dust_mask.argtypes = [C.c_char_p, C.c_size_t]
        L.cfr_oracle_format_tsv.restype = C.c_int
        L.cfr_oracle_format_tsv.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Result), C.c_char_p,
                                            C.c_size_t]
        L.cfr_oracle_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.cfr_oracle_reset_counters.argtypes = [C.c_void_p]
        _lib = L
    return _lib


TSV_HEADER = "readID\tseqID\ttaxID\tscore\t2ndBestScore\thitLength\tqueryLength\tnumMatches\n"


def dust_mask(seq: bytes) -> bytes:
    buf = C.create_string_buffer(seq, len(seq) + 1)
    lib().cfr_oracle_dust_mask(buf, len(seq))
    return buf.raw[:len(seq)]


class Oracle:
    """Mirror of the reference `Classifier` (Init/Query) over the C restatement."""

    def __init__(self, idx_prefix, k=1, min_hit_len=0, hitk_factor=40,
                 secondary_len=2000, secondary_factor=0.995, dust=True):
        self.L = lib()
        self.h = self.L.cfr_oracle_open(idx_prefix.encode())
        if not self.h:
            raise FileNotFoundError(idx_prefix + ".1.cfr")
        self.p = Param()
        self.L.cfr_oracle_default_param(C.byref(self.p))
        self.p.max_result = k
        self.p.min_hit_len = min_hit_len
        self.p.max_result_per_hit_factor = hitk_factor
        self.p.consider_secondary_hit_len = secondary_len
        self.p.consider_secondary_score_factor = secondary_factor
        self.dust = dust

    def close(self):
        if self.h:
            self.L.cfr_oracle_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def scalar(self, which):
        return self.L.cfr_oracle_scalar(self.h, which)

    @property
    def n(self):
        return self.scalar(0)

    def min_hit_len(self):
        return self.p.min_hit_len if self.p.min_hit_len > 0 else self.L.cfr_oracle_infer_min_hit_len(self.h)

    def bwt_rank(self, c, i, inclusive=1):
        return self.L.cfr_oracle_bwt_rank(self.h, c.encode(), i, inclusive)

    def bwt_access(self, i):
        return self.L.cfr_oracle_bwt_access(self.h, i).decode()

    def fm_rank(self, c, i, inclusive=1):
        return self.L.cfr_oracle_fm_rank(self.h, c.encode(), i, inclusive)

    def backward_search(self, s: bytes, m=None):
        sp, ep = C.c_uint64(0), C.c_uint64(0)
        l = self.L.cfr_oracle_backward_search(self.h, s, len(s) if m is None else m,
                                              C.byref(sp), C.byref(ep))
        return l, sp.value, ep.value

    def locate(self, row):
        st = C.c_uint64(0)
        r = self.L.cfr_oracle_locate(self.h, row, C.byref(st))
        return r, st.value

    def _mask(self, r):
        if r is None:
            return None
        return dust_mask(r) if self.dust else r

    def search(self, r1: bytes, r2: bytes = None, cap=256):
        hits = (Hit * cap)()
        n = self.L.cfr_oracle_search(self.h, C.byref(self.p), self._mask(r1), self._mask(r2), hits, cap)
        return [(h.sp, h.ep, h.l, h.offset, h.strand) for h in hits[:min(n, cap)]]

    def query(self, r1: bytes, r2: bytes = None) -> Result:
        res = Result()
        self.L.cfr_oracle_query(self.h, C.byref(self.p), self._mask(r1), self._mask(r2), C.byref(res))
        return res

    def result_tuple(self, res: Result):
        n = min(res.n, 64)
        return (res.score, res.secondary_score, res.hit_length, res.query_length, res.n,
                res.by_rank, tuple(res.ids[:n]), tuple(res.tax_ids[:n]))

    def format_tsv(self, read_id: str, res: Result) -> str:
        buf = C.create_string_buffer(1 << 16)
        w = self.L.cfr_oracle_format_tsv(self.h, read_id.encode(), C.byref(res), buf, len(buf))
        assert w >= 0
        return buf.raw[:w].decode()

    def classify_tsv(self, ids, reads1, reads2=None, header=True) -> str:
        out = [TSV_HEADER] if header else []
        for i, rid in enumerate(ids):
            res = self.query(reads1[i], reads2[i] if reads2 is not None else None)
            out.append(self.format_tsv(rid, res))
        return "".join(out)

    def reduce_taxids(self, tax_ids, k):
        arr = (C.c_uint64 * len(tax_ids))(*tax_ids)
        out = (C.c_uint64 * (len(tax_ids) + 1))()
        n = self.L.cfr_oracle_reduce_taxids(self.h, arr, len(tax_ids), k, out, len(tax_ids) + 1)
        return list(out[:n])

    def reduce_taxids_expanded(self, tax_ids, k):
        """(promoted ids, child lists) -- the lists are [] when the classifier would print none"""
        n = len(tax_ids)
        arr = (C.c_uint64 * n)(*tax_ids)
        out = (C.c_uint64 * (n + 1))()
        child = (C.c_uint64 * (n + 1))()
        cnt = (C.c_int32 * (n + 1))()
        n_lists = C.c_int(0)
        m = self.L.cfr_oracle_reduce_taxids_expanded(self.h, arr, n, k, out, n + 1, child, n + 1, cnt,
                                                     C.byref(n_lists))
        lists, at = [], 0
        for i in range(n_lists.value):
            lists.append(list(child[at:at + cnt[i]]))
            at += cnt[i]
        return list(out[:m]), (lists if n_lists.value == m else [])

    def query_expanded(self, r1: bytes, r2: bytes = None, cap=4096):
        """(Result, child lists per reported id) with outputExpandedResult set"""
        res = Result()
        child = (C.c_uint64 * cap)()
        cnt = (C.c_int32 * 64)()
        tot = self.L.cfr_oracle_query_expanded(self.h, C.byref(self.p), self._mask(r1), self._mask(r2),
                                               C.byref(res), child, cap, cnt)
        assert tot <= cap
        return res, child, cnt

    def classify_tsv_expanded(self, ids, reads1, reads2=None) -> str:
        out = [TSV_HEADER[:-1] + "\texpandedTaxIDs\n"]
        buf = C.create_string_buffer(1 << 18)
        for i, rid in enumerate(ids):
            res, child, cnt = self.query_expanded(reads1[i], reads2[i] if reads2 is not None else None)
            w = self.L.cfr_oracle_format_tsv_expanded(self.h, rid.encode(), C.byref(res), child, cnt, buf, len(buf))
            assert w >= 0
            out.append(buf.raw[:w].decode())
        return "".join(out)

    def counters(self):
        c = Counters()
        self.L.cfr_oracle_get_counters(self.h, C.byref(c))
        return {k: getattr(c, k) for k, _ in Counters._fields_}

    def reset_counters(self):
        self.L.cfr_oracle_reset_counters(self.h)


def read_fastx(path):
    """Minimal FASTA/FASTQ reader following the reference's kseq conventions
    (ReadFiles.hpp:82-90, :311-312): id = first token, trailing /1 or /2 removed.
    Returns (ids, seqs) with seqs as bytes."""
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    ids, seqs = [], []
    with op(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if not ln:
            i += 1
            continue
        if ln[:1] == b"@":
            name = ln[1:].split()[0] if ln[1:].split() else b""
            seq = lines[i + 1].rstrip(b"\r")
            i += 4
        elif ln[:1] == b">":
            name = ln[1:].split()[0] if ln[1:].split() else b""
            i += 1
            parts = []
            while i < len(lines) and lines[i][:1] not in (b">", b"@"):
                parts.append(lines[i].strip())
                i += 1
            seq = b"".join(parts)
        else:
            i += 1
            continue
        if len(name) >= 2 and name[-2:] in (b"/1", b"/2"):
            name = name[:-2]
        ids.append(name.decode())
        seqs.append(seq)
    return ids, seqs


def run_reference(idx_prefix, r1=None, r2=None, single=None, extra=(), threads=1):
    """Run the unmodified reference binary (oracle/_ref/centrifuger); returns stdout text."""
    exe = os.path.join(REF_DIR, "centrifuger")
    cmd = [exe, "-x", idx_prefix, "-t", str(threads)]
    if single is not None:
        cmd += ["-u", single]
    else:
        cmd += ["-1", r1, "-2", r2]
    cmd += list(extra)
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
