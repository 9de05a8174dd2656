"""ctypes binding to tests/hostsim/libhostsim.so -- TEST INFRASTRUCTURE ONLY.

The host simulation compiles the product's per-task stage functions
(cfr_core.cuh / cfr_pipeline.cuh) with g++ and runs them sequentially, so the
classification logic can be checked against the oracle in a container without a
GPU.  It is not part of the product and the product never falls back to it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
LIB = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
CSRC = os.path.join(ROOT, "centrifuger_b200", "csrc")


class Params(C.Structure):
    _fields_ = [("max_result", C.c_int32), ("min_hit_len", C.c_int32),
                ("max_result_per_hit_factor", C.c_int32), ("dust", C.c_int32),
                ("consider_secondary_hit_len", C.c_uint64),
                ("consider_secondary_score_factor", C.c_double),
                ("layout", C.c_int32), ("max_batch_reads", C.c_int32),
                ("arena_rows", C.c_uint64), ("expand_taxid", C.c_int32), ("unlimited_cap", C.c_int32)]


class ReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("seq1", C.c_void_p), ("off1", C.c_void_p),
                ("seq2", C.c_void_p), ("off2", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("score", C.c_uint64), ("secondary_score", C.c_uint64),
                ("hit_length", C.c_int32), ("query_length", C.c_int32),
                ("n_assign", C.c_int32), ("by_rank", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in
                ("n_rank", "n_access", "n_search", "n_locate", "n_lf", "n_extend",
                 "n_bases", "n_reads", "n_launches")]


RESULT_DTYPE = np.dtype([("score", "<u8"), ("secondary_score", "<u8"), ("hit_length", "<i4"),
                         ("query_length", "<i4"), ("n_assign", "<i4"), ("by_rank", "<i4")])


def pack_reads(reads):
    """list of bytes -> (uint8 array, uint64 offsets[n+1])"""
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        off[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    buf = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if reads else np.zeros(0, np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    return buf, off


def make_batch(reads1, reads2=None):
    s1, o1 = pack_reads(reads1)
    keep = [s1, o1]
    b = ReadBatch()
    b.n_reads = len(reads1)
    b.seq1 = s1.ctypes.data
    b.off1 = o1.ctypes.data
    if reads2 is not None:
        s2, o2 = pack_reads(reads2)
        keep += [s2, o2]
        b.seq2 = s2.ctypes.data
        b.off2 = o2.ctypes.data
    return b, keep


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in
                    ("cfr_core.cuh", "cfr_pipeline.cuh", "cfr_types.h", "cfr_format.cpp", "cfr_format.hpp")]
    if (not os.path.exists(LIB)) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-DCFR_HOSTSIM", "-shared", "-fPIC", "-o", LIB,
                               SRC, os.path.join(CSRC, "cfr_format.cpp")])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hostsim_open.restype = C.c_void_p
        L.hostsim_open.argtypes = [C.c_char_p, C.POINTER(Params)]
        L.hostsim_close.argtypes = [C.c_void_p]
        L.hostsim_min_hit_len.argtypes = [C.c_void_p]
        L.hostsim_bwt_rank.restype = C.c_uint64
        L.hostsim_bwt_rank.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int]
        L.hostsim_bwt_access.argtypes = [C.c_void_p, C.c_uint64]
        L.hostsim_locate.restype = C.c_uint64
        L.hostsim_locate.argtypes = [C.c_void_p, C.c_uint64]
        L.hostsim_occ_roundtrip.restype = C.c_uint64
        L.hostsim_occ_roundtrip.argtypes = [C.c_uint64] * 4 + [C.c_int]
        L.hostsim_occ_roundtrip32.restype = C.c_uint32
        L.hostsim_occ_roundtrip32.argtypes = [C.c_uint64] * 4 + [C.c_int]
        L.hostsim_sa_read.restype = C.c_uint64
        L.hostsim_sa_read.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.hostsim_plan_rows.restype = None
        L.hostsim_plan_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        L.hostsim_extend_high.restype = None
        L.hostsim_extend_high.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_uint64, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]
        L.hostsim_pair_check.restype = C.c_uint64
        L.hostsim_pair_check.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.hostsim_reduce_taxids.restype = C.c_int
        L.hostsim_reduce_taxids.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.hostsim_expand_taxids.restype = C.c_int
        L.hostsim_expand_taxids.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_int, C.POINTER(C.c_uint64),
                                            C.POINTER(C.c_uint32)]
        L.hostsim_classify_expanded.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(ReadBatch), C.c_void_p,
                                                C.c_void_p, C.POINTER(Counters), C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_uint64, C.POINTER(C.c_uint64)]
        L.hostsim_dust.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
        L.hostsim_dust_screen.argtypes = [C.c_char_p, C.c_int]
        L.hostsim_dust_screen.restype = C.c_int
        L.hostsim_classify.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(ReadBatch), C.c_void_p,
                                       C.c_void_p, C.POINTER(Counters)]
        _lib = L
    return _lib


def default_params(k=1, min_hit_len=0, hitk_factor=40, dust=True, secondary_len=2000,
                   secondary_factor=0.995, layout=1):
    p = Params()
    p.max_result = k
    p.min_hit_len = min_hit_len
    p.max_result_per_hit_factor = hitk_factor
    p.dust = 1 if dust else 0
    p.consider_secondary_hit_len = secondary_len
    p.consider_secondary_score_factor = secondary_factor
    p.layout = layout
    return p


def hostsim_dust(seq: bytes) -> bytes:
    out = C.create_string_buffer(len(seq) + 1)
    lib().hostsim_dust(seq, len(seq), out)
    return out.raw[:len(seq)]


def hostsim_dust_screen(seq: bytes) -> bool:
    """True when the register-only screen sends the read to the full SDUST"""
    return bool(lib().hostsim_dust_screen(seq, len(seq)))


class HostSim:
    def __init__(self, prefix, **kw):
        self.L = lib()
        self.p = default_params(**kw)
        self.stride = self.p.max_result if self.p.max_result > 0 else (self.p.unlimited_cap or 64)
        self.h = self.L.hostsim_open(prefix.encode(), C.byref(self.p))
        if not self.h:
            raise RuntimeError("hostsim_open failed for " + prefix)

    def close(self):
        if self.h:
            self.L.hostsim_close(self.h)
            self.h = None

    def bwt_rank(self, c, i, inclusive=1):
        return self.L.hostsim_bwt_rank(self.h, "ACGT".index(c), i, inclusive)

    def bwt_access(self, i):
        return "ACGT"[self.L.hostsim_bwt_access(self.h, i)]

    def locate(self, row):
        return self.L.hostsim_locate(self.h, row)

    def pair_check(self, n_ranges=20000, seed=1):
        """disagreements between the pair layout's two-step extend and two literal BackwardExtend calls"""
        return int(self.L.hostsim_pair_check(self.h, n_ranges, seed))

    def reduce_taxids(self, tax_ids, k):
        arr = (C.c_uint64 * len(tax_ids))(*tax_ids)
        out = (C.c_uint64 * (len(tax_ids) + max(k, 1) + 1))()
        n = self.L.hostsim_reduce_taxids(self.h, arr, len(tax_ids), k, out)
        return [int(out[i]) for i in range(n)]

    def expand_taxids(self, tax_ids, k):
        """(promoted ids, child lists) of the scoring stage's tax_reduce + tax_expand"""
        ids = self.reduce_taxids(tax_ids, k)
        arr = (C.c_uint64 * len(tax_ids))(*tax_ids)
        child = (C.c_uint64 * (len(tax_ids) + 1))()
        cnt = (C.c_uint32 * (max(k, 1) + 1))()
        total = self.L.hostsim_expand_taxids(self.h, arr, len(tax_ids), k, child, cnt)
        assert total >= 0
        lists, at = [], 0
        for i in range(len(ids)):
            lists.append([int(x) for x in child[at:at + cnt[i]]])
            at += cnt[i]
        assert at == total
        return ids, lists

    def classify_expanded(self, reads1, reads2=None, arena_rows=0):
        """classify() plus, per read, the list of child-id lists (one per reported id)"""
        b, keep = make_batch(reads1, reads2)
        n, k = len(reads1), self.stride
        res = np.zeros(n, dtype=RESULT_DTYPE)
        ids = np.zeros(max(1, n * k), dtype=np.uint64)
        exp_cnt = np.zeros(max(1, n * k), dtype=np.uint32)
        exp_off = np.zeros(max(1, n), dtype=np.uint64)
        cap = 64 * n + 1024
        exp_ids = np.zeros(cap, dtype=np.uint64)
        exp_n = C.c_uint64(0)
        st = self.L.hostsim_classify_expanded(self.h, self.p.dust, arena_rows, C.byref(b), res.ctypes.data,
                                              ids.ctypes.data, None, exp_cnt.ctypes.data, exp_off.ctypes.data,
                                              exp_ids.ctypes.data, cap, C.byref(exp_n))
        if st != 0:
            raise RuntimeError("hostsim_classify_expanded status %d" % st)
        return res, ids.reshape(-1, k) if n else ids, expansion_lists(res, exp_cnt, exp_off, exp_ids, k)

    def classify(self, reads1, reads2=None, arena_rows=0):
        b, keep = make_batch(reads1, reads2)
        n = len(reads1)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        ids = np.zeros(max(1, n * self.stride), dtype=np.uint64)
        cnt = Counters()
        st = self.L.hostsim_classify(self.h, self.p.dust, arena_rows, C.byref(b), res.ctypes.data,
                                     ids.ctypes.data, C.byref(cnt))
        if st != 0:
            raise RuntimeError("hostsim_classify status %d" % st)
        return res, ids.reshape(-1, self.stride) if n else ids, \
            {k: getattr(cnt, k) for k, _ in Counters._fields_}


def result_tuples(res, ids, k):
    """comparable tuples (same shape as Oracle.result_tuple)"""
    out = []
    for i in range(len(res)):
        n = int(res["n_assign"][i])
        m = min(n, k)
        out.append((int(res["score"][i]), int(res["secondary_score"][i]), int(res["hit_length"][i]),
                    int(res["query_length"][i]), n, int(res["by_rank"][i]),
                    tuple(int(x) for x in ids[i][:m])))
    return out


def expansion_lists(res, exp_cnt, exp_off, exp_ids, k):
    """per read: one list of compact tax ids per reported id (cfr_fetch_expanded layout)"""
    out = []
    for i in range(len(res)):
        at = int(exp_off[i])
        lists = []
        for j in range(min(int(res["n_assign"][i]), k)):
            c = int(exp_cnt[i * k + j])
            lists.append([int(x) for x in exp_ids[at:at + c]])
            at += c
        out.append(lists)
    return out
