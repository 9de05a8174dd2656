"""centrifuger_b200 -- B200-native drop-in for the centrifuger classification path.

Host-side mirror of the reference's `Classifier` interface (Classifier.hpp:902
`Init`, :950 `Query`) over the C ABI of include/centrifuger_b200.h.  The work is
done by hand-written sm_100a kernels in libcfrb200.so; there is NO CPU fallback:
importing works anywhere (so the symbol table can be checked), but opening an
index without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcfrb200.so")

LAYOUT_AUTO, LAYOUT_RUNBLOCK, LAYOUT_OCCLINE = 0, 1, 2

TSV_HEADER = "readID\tseqID\ttaxID\tscore\t2ndBestScore\thitLength\tqueryLength\tnumMatches\n"


class CfrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cfr status %d: %s" % (code, msg))
        self.code = code


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


class PackedBatch(C.Structure):
    """cfr_packed_batch: 2-bit codes + N bits per base, offsets = positions in the batch buffer"""
    _fields_ = [("n_reads", C.c_uint64), ("codes", C.c_void_p), ("nmask", C.c_void_p), ("n_words", C.c_uint64),
                ("off1", C.c_void_p), ("off2", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in
                ("n_rank", "n_access", "n_search", "n_locate", "n_lf", "n_extend",
                 "n_bases", "n_reads", "n_launches")]


class StageTimes(C.Structure):
    _fields_ = [("ms", C.c_double * 6), ("launches", C.c_uint64 * 6)]


STAGES = ("dust", "search", "select", "locate", "score", "other")

RESULT_DTYPE = np.dtype([("score", "<u8"), ("secondary_score", "<u8"), ("hit_length", "<i4"),
                         ("query_length", "<i4"), ("n_assign", "<i4"), ("by_rank", "<i4")])

# every symbol include/centrifuger_b200.h declares
ABI_SYMBOLS = [
    "cfr_default_params", "cfr_open", "cfr_close", "cfr_last_error", "cfr_classify_batch",
    "cfr_submit_batch", "cfr_submit_batch_masked", "cfr_wait_batch", "cfr_packed_words", "cfr_pack_reads", "cfr_submit_packed", "cfr_batch_upload", "cfr_classify_resident", "cfr_batch_fetch", "cfr_batch_free",
    "cfr_fetch_expanded", "cfr_batch_fetch_expanded",
    "cfr_host_alloc", "cfr_host_free", "cfr_index_info", "cfr_seq_name", "cfr_rank_name", "cfr_orig_taxid", "cfr_seq_taxid",
    "cfr_format_tsv", "cfr_taxon_counts_device", "cfr_taxon_counts_read", "cfr_taxon_counts_reset",
    "cfr_counts_allreduce", "cfr_counts_allreduce_local",
    "cfr_quant_enable", "cfr_quant_reset", "cfr_quant_stats", "cfr_quant_report",
    "cfr_get_counters", "cfr_reset_counters", "cfr_set_profiling", "cfr_get_stage_times",
    "cfr_get_stage_counters", "cfr_debug_bwt_rank", "cfr_debug_bwt_access",
    "cfr_debug_locate", "cfr_debug_dust",
]
# include/centrifuger_b200_build.h (the index builder, bound in centrifuger_b200/builder.py)
BUILD_ABI_SYMBOLS = [
    "cfr_build_default_params", "cfr_build_last_error", "cfr_build_fm_index", "cfr_build_synthetic_fm_index",
    "cfr_synth_bases", "cfr_synth_fragments",
]

_lib = None


def load_library():
    """dlopen libcfrb200.so (built in-tree by centrifuger_b200.build); fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m centrifuger_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.cfr_default_params.argtypes = [C.POINTER(Params)]
    L.cfr_default_params.restype = None
    L.cfr_open.argtypes = [C.c_char_p, C.POINTER(Params), i32, C.POINTER(vp)]
    L.cfr_close.argtypes = [vp]
    L.cfr_close.restype = None
    L.cfr_last_error.restype = C.c_char_p
    L.cfr_classify_batch.argtypes = [vp, C.POINTER(ReadBatch), vp, vp, vp]
    L.cfr_submit_batch.argtypes = [vp, C.POINTER(ReadBatch), vp, vp, vp, C.POINTER(C.c_int)]
    L.cfr_submit_batch_masked.argtypes = [vp, C.POINTER(ReadBatch), vp, vp, vp, vp, vp, C.POINTER(C.c_int)]
    L.cfr_wait_batch.argtypes = [vp, C.c_int]
    L.cfr_packed_words.argtypes = [C.POINTER(ReadBatch)]
    L.cfr_packed_words.restype = u64
    L.cfr_pack_reads.argtypes = [C.POINTER(ReadBatch), vp, vp, vp, vp, i32, C.POINTER(PackedBatch)]
    L.cfr_submit_packed.argtypes = [vp, C.POINTER(PackedBatch), vp, vp, vp, C.POINTER(C.c_int)]
    L.cfr_batch_upload.argtypes = [vp, C.POINTER(ReadBatch), vp, C.POINTER(vp)]
    L.cfr_classify_resident.argtypes = [vp, vp, vp]
    L.cfr_batch_fetch.argtypes = [vp, vp, vp, vp, vp]
    L.cfr_fetch_expanded.argtypes = [vp, C.c_int, vp, vp, vp, u64, C.POINTER(u64)]
    L.cfr_batch_fetch_expanded.argtypes = [vp, vp, vp, vp, vp, u64, C.POINTER(u64), vp]
    L.cfr_batch_free.argtypes = [vp, vp]
    L.cfr_batch_free.restype = None
    L.cfr_host_alloc.argtypes = [C.c_size_t]
    L.cfr_host_alloc.restype = vp
    L.cfr_host_free.argtypes = [vp]
    L.cfr_host_free.restype = None
    L.cfr_index_info.argtypes = [vp, i32]
    L.cfr_index_info.restype = u64
    L.cfr_seq_name.argtypes = [vp, u64]
    L.cfr_seq_name.restype = C.c_char_p
    L.cfr_rank_name.argtypes = [vp, u64]
    L.cfr_rank_name.restype = C.c_char_p
    L.cfr_orig_taxid.argtypes = [vp, u64]
    L.cfr_orig_taxid.restype = u64
    L.cfr_seq_taxid.argtypes = [vp, u64]
    L.cfr_seq_taxid.restype = u64
    L.cfr_format_tsv.argtypes = [vp, C.c_char_p, vp, vp, C.c_char_p, C.c_size_t]
    L.cfr_taxon_counts_device.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.cfr_taxon_counts_read.argtypes = [vp, vp, u64, vp]
    L.cfr_taxon_counts_reset.argtypes = [vp, vp]
    L.cfr_get_counters.argtypes = [vp, C.POINTER(Counters), vp]
    L.cfr_reset_counters.argtypes = [vp, vp]
    L.cfr_set_profiling.argtypes = [vp, i32]
    L.cfr_get_stage_times.argtypes = [vp, C.POINTER(StageTimes), i32]
    L.cfr_get_stage_counters.argtypes = [vp, i32, C.POINTER(Counters), vp]
    L.cfr_debug_bwt_rank.argtypes = [vp, vp, vp, vp, u64, vp]
    L.cfr_debug_bwt_access.argtypes = [vp, vp, u64, vp]
    L.cfr_debug_locate.argtypes = [vp, vp, u64, vp]
    L.cfr_debug_dust.argtypes = [vp, C.POINTER(ReadBatch), vp, vp]
    _lib = L
    return L


def pack_reads(reads):
    """list of bytes -> (uint8 buffer, uint64 offsets[n+1])"""
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    if len(reads):
        off[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    buf = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if len(reads) else np.zeros(0, np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)
    return buf, off


def pack_batch(seq1, off1, seq2=None, off2=None, threads=8, pinned=False):
    """cfr_pack_reads: host buffers -> (PackedBatch, keep): the bases as 2-bit codes + N bits (what the producer of a batch
    hands to cfr_submit_packed; 2.25 bits per base over the host link).  `keep` holds the arrays the struct points into."""
    L = load_library()
    n = len(off1) - 1
    b = make_batch(seq1, off1, seq2, off2, n)
    nw = int(L.cfr_packed_words(C.byref(b)))
    if pinned:
        import torch
        alloc = lambda k, dt: torch.empty(k, dtype=dt).pin_memory()
        codes, nmask = alloc(nw, torch.int64), alloc(nw, torch.int32)
        o1, o2 = alloc(n + 1, torch.int64), (alloc(n + 1, torch.int64) if seq2 is not None else None)
    else:
        codes, nmask = np.zeros(nw, dtype=np.uint64), np.zeros(nw, dtype=np.uint32)
        o1, o2 = np.zeros(n + 1, dtype=np.uint64), (np.zeros(n + 1, dtype=np.uint64) if seq2 is not None else None)
    pk = PackedBatch()
    st = L.cfr_pack_reads(C.byref(b), _ptr(codes), _ptr(nmask), _ptr(o1), _ptr(o2), threads, C.byref(pk))
    if st != 0:
        raise CfrError(st, L.cfr_last_error().decode())
    return pk, (codes, nmask, o1, o2)


def _ptr(a):
    """host pointer of a numpy array or a (pinned) torch CPU tensor"""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


def make_batch(seq1, off1, seq2=None, off2=None, n=None):
    b = ReadBatch()
    b.n_reads = int(n if n is not None else (len(off1) - 1))
    b.seq1, b.off1 = _ptr(seq1), _ptr(off1)
    b.seq2, b.off2 = _ptr(seq2), _ptr(off2)
    return b


class DeviceBatch:
    """A read batch resident in HBM (cfr_batch_upload)."""

    def __init__(self, clf, handle, n):
        self.clf, self.h, self.n = clf, handle, n

    def free(self):
        if self.h:
            self.clf.L.cfr_batch_free(self.clf.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Classifier:
    """Mirror of `Classifier<Sequence_RunBlock>` (Classifier.hpp:87-967).

    Classifier(idx_prefix, k=1, ...)   ~  Init(idxPrefix, _classifierParam)
    query(r1, r2=None)                 ~  Query(r1, r2, result)  (+ DUST, as ClassifyReads_Thread does)
    classify(reads1, reads2=None)      ~  one batch through ClassifyReads_Thread
    """

    def __init__(self, idx_prefix, k=1, min_hit_len=0, hitk_factor=40, dust=True,
                 secondary_len=2000, secondary_factor=0.995, layout=LAYOUT_AUTO,
                 device=0, max_batch_reads=0, arena_rows=0, expand_taxid=False, unlimited_cap=0):
        self.L = load_library()
        p = Params()
        self.L.cfr_default_params(C.byref(p))
        p.max_result, p.min_hit_len, p.max_result_per_hit_factor = k, min_hit_len, hitk_factor
        p.dust = 1 if dust else 0
        p.consider_secondary_hit_len = secondary_len
        p.consider_secondary_score_factor = secondary_factor
        p.layout, p.max_batch_reads, p.arena_rows = layout, max_batch_reads, arena_rows
        p.expand_taxid = 1 if expand_taxid else 0
        p.unlimited_cap = unlimited_cap  # id slots per read when k <= 0 (0 = the library's default, 64)
        self.params = p
        self.max_result = k
        self.k = k if k > 0 else 64  # id slots per read (the stride of the ids arrays); -k <= 0 = unlimited, see info(24)
        self.h = C.c_void_p()
        st = self.L.cfr_open(idx_prefix.encode(), C.byref(p), device, C.byref(self.h))
        if st != 0:
            self.h = None
            raise CfrError(st, self.L.cfr_last_error().decode())
        self.k = int(self.L.cfr_index_info(self.h, 24))

    # -- lifetime ----------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.cfr_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != 0:
            raise CfrError(st, self.L.cfr_last_error().decode())

    # -- facts -------------------------------------------------------------
    def info(self, which):
        return int(self.L.cfr_index_info(self.h, which))

    @property
    def n(self):
        return self.info(0)

    @property
    def min_hit_len(self):
        return self.info(4)

    @property
    def node_cnt(self):
        return self.info(5)

    @property
    def layout(self):
        return self.info(8)

    @property
    def hbm_bytes(self):
        return self.info(9)

    # -- classification ----------------------------------------------------
    def classify_packed(self, seq1, off1, seq2=None, off2=None, stream=None, out=None):
        """Host buffers (numpy or pinned torch CPU tensors) in, (results, ids) out."""
        n = len(off1) - 1
        b = make_batch(seq1, off1, seq2, off2, n)
        if out is None:
            res = np.zeros(n, dtype=RESULT_DTYPE)
            ids = np.zeros(max(1, n * self.k), dtype=np.uint64)
        else:
            res, ids = out
        self._check(self.L.cfr_classify_batch(self.h, C.byref(b), _ptr(res), _ptr(ids), stream))
        return res, ids

    def classify(self, reads1, reads2=None):
        s1, o1 = pack_reads(reads1)
        s2 = o2 = None
        if reads2 is not None:
            s2, o2 = pack_reads(reads2)
        res, ids = self.classify_packed(s1, o1, s2, o2)
        return res, ids.reshape(-1, self.k) if len(reads1) else ids

    def query(self, r1, r2=None):
        res, ids = self.classify([r1], None if r2 is None else [r2])
        return res[0], ids[0]

    def submit(self, seq1, off1, seq2=None, off2=None, stream=None, out=None):
        """Streaming form: enqueue one batch (pinned host buffers) and return (ticket, results, ids, keep).
        Results are valid after wait(ticket); keep the returned objects alive until then."""
        n = len(off1) - 1
        b = make_batch(seq1, off1, seq2, off2, n)
        if out is None:
            res = np.zeros(n, dtype=RESULT_DTYPE)
            ids = np.zeros(max(1, n * self.k), dtype=np.uint64)
        else:
            res, ids = out
        t = C.c_int(-1)
        self._check(self.L.cfr_submit_batch(self.h, C.byref(b), _ptr(res), _ptr(ids), stream, C.byref(t)))
        return t.value, res, ids, (b, seq1, off1, seq2, off2)

    def submit_packed(self, packed, stream=None, out=None):
        """cfr_submit_packed: `packed` = the PackedBatch of pack_batch (keep its arrays alive until wait)"""
        n = int(packed.n_reads)
        if out is None:
            res = np.zeros(n, dtype=RESULT_DTYPE)
            ids = np.zeros(max(1, n * self.k), dtype=np.uint64)
        else:
            res, ids = out
        t = C.c_int(-1)
        self._check(self.L.cfr_submit_packed(self.h, C.byref(packed), _ptr(res), _ptr(ids), stream, C.byref(t)))
        return t.value, res, ids

    def wait(self, ticket):
        self._check(self.L.cfr_wait_batch(self.h, ticket))

    def classify_masked(self, reads1, reads2=None):
        """One batch through the streaming form that also returns the reads as they were classified
        (DUST intervals as N): (results, ids, masked1, masked2)."""
        s1, o1 = pack_reads(reads1)
        s2, o2 = pack_reads(reads2) if reads2 is not None else (None, None)
        n = len(o1) - 1
        b = make_batch(s1, o1, s2, o2, n)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        ids = np.zeros(max(1, n * self.k), dtype=np.uint64)
        m1 = np.zeros(max(1, len(s1)), dtype=np.uint8)
        m2 = np.zeros(max(1, len(s2)), dtype=np.uint8) if s2 is not None else None
        t = C.c_int(-1)
        self._check(self.L.cfr_submit_batch_masked(self.h, C.byref(b), _ptr(res), _ptr(ids), _ptr(m1),
                                                   _ptr(m2) if m2 is not None else None, None, C.byref(t)))
        self.wait(t.value)
        split = lambda m, o: [bytes(m[int(o[i]):int(o[i + 1])]) for i in range(n)]
        return res, ids.reshape(n, self.k), split(m1, o1), (split(m2, o2) if m2 is not None else None)

    def _expansion_lists(self, res, fetch):
        """fetch(exp_cnt, exp_off, exp_ids, cap, &n) -> per read, one list of compact tax ids per reported id"""
        n = len(res)
        cnt = np.zeros(max(1, n * self.k), dtype=np.uint32)
        off = np.zeros(max(1, n), dtype=np.uint64)
        need = C.c_uint64(0)
        ids = np.zeros(1024, dtype=np.uint64)
        st = fetch(_ptr(cnt), _ptr(off), _ptr(ids), len(ids), C.byref(need))
        if st != 0 and need.value > len(ids):  # sized on the second try
            ids = np.zeros(need.value, dtype=np.uint64)
            st = fetch(_ptr(cnt), _ptr(off), _ptr(ids), len(ids), C.byref(need))
        self._check(st)
        out = []
        for i in range(n):
            at = int(off[i])
            lists = []
            for j in range(min(int(res["n_assign"][i]), self.k)):
                c = int(cnt[i * self.k + j])
                lists.append([int(x) for x in ids[at:at + c]])
                at += c
            out.append(lists)
        return out

    def classify_expanded(self, reads1, reads2=None):
        """One batch with --expand-taxid (handle opened with expand_taxid=True): (results, ids, lists) where
        lists[i][j] are the compact tax ids promoted into assignment j of read i (Classifier.hpp:807-838)."""
        s1, o1 = pack_reads(reads1)
        s2, o2 = pack_reads(reads2) if reads2 is not None else (None, None)
        t, res, ids, keep = self.submit(s1, o1, s2, o2)
        self.wait(t)
        lists = self._expansion_lists(
            res, lambda c, o, i, cap, n: self.L.cfr_fetch_expanded(self.h, t, c, o, i, cap, n))
        return res, ids.reshape(-1, self.k) if len(reads1) else ids, lists

    def fetch_expanded(self, batch, res, stream=None):
        """the lists of a resident batch, after fetch()"""
        return self._expansion_lists(
            res, lambda c, o, i, cap, n: self.L.cfr_batch_fetch_expanded(self.h, batch.h, c, o, i, cap, n, stream))

    def upload(self, seq1, off1, seq2=None, off2=None, stream=None):
        n = len(off1) - 1
        b = make_batch(seq1, off1, seq2, off2, n)
        h = C.c_void_p()
        self._check(self.L.cfr_batch_upload(self.h, C.byref(b), stream, C.byref(h)))
        return DeviceBatch(self, h, n)

    def classify_resident(self, batch, stream=None):
        self._check(self.L.cfr_classify_resident(self.h, batch.h, stream))

    def fetch(self, batch, stream=None, out=None):
        if out is None:
            res = np.zeros(batch.n, dtype=RESULT_DTYPE)
            ids = np.zeros(max(1, batch.n * self.k), dtype=np.uint64)
        else:
            res, ids = out
        self._check(self.L.cfr_batch_fetch(self.h, batch.h, _ptr(res), _ptr(ids), stream))
        return res, ids

    # -- output ------------------------------------------------------------
    def format_tsv(self, read_id, res_row, ids_row):
        buf = C.create_string_buffer(1 << 16)
        r = np.array([res_row], dtype=RESULT_DTYPE)
        i = np.ascontiguousarray(ids_row, dtype=np.uint64)
        w = self.L.cfr_format_tsv(self.h, read_id.encode(), r.ctypes.data, i.ctypes.data, buf, len(buf))
        if w < 0:
            raise CfrError(w, "format_tsv")
        return buf.raw[:w].decode()

    def classify_tsv(self, read_ids, reads1, reads2=None, header=True):
        res, ids = self.classify(reads1, reads2)
        out = [TSV_HEADER] if header else []
        for i, rid in enumerate(read_ids):
            out.append(self.format_tsv(rid, res[i], ids[i]))
        return "".join(out)

    def seq_name(self, seq_id):
        return self.L.cfr_seq_name(self.h, seq_id).decode()

    def rank_name(self, ctid):
        return self.L.cfr_rank_name(self.h, ctid).decode()

    def orig_taxid(self, ctid):
        return int(self.L.cfr_orig_taxid(self.h, ctid))

    def seq_taxid(self, seq_id):
        return int(self.L.cfr_seq_taxid(self.h, seq_id))

    # -- counters ----------------------------------------------------------
    def counters(self, stream=None):
        c = Counters()
        self._check(self.L.cfr_get_counters(self.h, C.byref(c), stream))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}

    def reset_counters(self, stream=None):
        self._check(self.L.cfr_reset_counters(self.h, stream))

    def set_profiling(self, on=True):
        self._check(self.L.cfr_set_profiling(self.h, 1 if on else 0))

    def stage_times(self, reset=False):
        """{stage: (milliseconds, launches)} measured with CUDA events on the launch stream"""
        t = StageTimes()
        self._check(self.L.cfr_get_stage_times(self.h, C.byref(t), 1 if reset else 0))
        return {STAGES[i]: (float(t.ms[i]), int(t.launches[i])) for i in range(6)}

    def stage_counters(self, stage, stream=None):
        c = Counters()
        self._check(self.L.cfr_get_stage_counters(self.h, STAGES.index(stage), C.byref(c), stream))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}

    def taxon_counts(self, stream=None):
        n = self.node_cnt + 3
        out = np.zeros(n, dtype=np.uint64)
        self._check(self.L.cfr_taxon_counts_read(self.h, out.ctypes.data, n, stream))
        return out

    def taxon_counts_reset(self, stream=None):
        self._check(self.L.cfr_taxon_counts_reset(self.h, stream))

    # -- quantification (replaces centrifuger-quant) --------------------------
    def quant_enable(self, min_score=0, min_hit_length=0):
        """from now on every finished batch is coalesced on the device (Quantifier.hpp:490-622)"""
        self.L.cfr_quant_enable.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        self._check(self.L.cfr_quant_enable(self.h, min_score, min_hit_length))

    def quant_reset(self):
        self.L.cfr_quant_reset.argtypes = [C.c_void_p]
        self._check(self.L.cfr_quant_reset(self.h))

    def quant_stats(self):
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.L.cfr_quant_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        self._check(self.L.cfr_quant_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"distinct_records": a.value, "batches": b.value, "records_moved": c.value}

    def quant_report(self, idx_prefix, path, fmt=0, others=()):
        """abundance report of everything classified since quant_enable / quant_reset (with `others`: more
        Classifier objects, one per GPU, whose records are merged in)"""
        hs = [self.h] + [o.h for o in others]
        arr = (C.c_void_p * len(hs))(*hs)
        self.L.cfr_quant_report.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p]
        self._check(self.L.cfr_quant_report(arr, len(hs), idx_prefix.encode(), fmt, path.encode()))

    def taxon_counts_device(self):
        """(device pointer, entries) of the uint64 per-taxon counters (for NCCL all-reduce)."""
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self.L.cfr_taxon_counts_device(self.h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    # -- diagnostics ---------------------------------------------------------
    def debug_rank(self, codes, pos, inclusive):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        inclusive = np.ascontiguousarray(inclusive, dtype=np.int32)
        out = np.zeros(len(pos), dtype=np.uint64)
        self._check(self.L.cfr_debug_bwt_rank(self.h, codes.ctypes.data, pos.ctypes.data,
                                              inclusive.ctypes.data, len(pos), out.ctypes.data))
        return out

    def debug_access(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        out = np.zeros(len(pos), dtype=np.uint8)
        self._check(self.L.cfr_debug_bwt_access(self.h, pos.ctypes.data, len(pos), out.ctypes.data))
        return out

    def debug_locate(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.zeros(len(rows), dtype=np.uint64)
        self._check(self.L.cfr_debug_locate(self.h, rows.ctypes.data, len(rows), out.ctypes.data))
        return out

    def debug_dust(self, reads1, reads2=None):
        s1, o1 = pack_reads(reads1)
        s2 = o2 = None
        m2 = None
        if reads2 is not None:
            s2, o2 = pack_reads(reads2)
            m2 = np.zeros_like(s2)
        m1 = np.zeros_like(s1)
        b = make_batch(s1, o1, s2, o2, len(reads1))
        self._check(self.L.cfr_debug_dust(self.h, C.byref(b), m1.ctypes.data,
                                          None if m2 is None else m2.ctypes.data))
        r1 = [m1[int(o1[i]):int(o1[i + 1])].tobytes() for i in range(len(reads1))]
        if reads2 is None:
            return r1
        return r1, [m2[int(o2[i]):int(o2[i + 1])].tobytes() for i in range(len(reads2))]
