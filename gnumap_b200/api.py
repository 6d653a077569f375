"""Host-side mirror of the reference interface over the gmx C ABI (include/gmx.h).

Method names follow the reference classes they stand in for, so parity tests read like the
reference's own call sites:

    GenomeBwt::get_sa_int / get_sa_coord / GetString        -> Mapper.get_sa_int / get_sa_coord / GetString
    bin_seq::get_align_score / _w_traceback / pairHMM        -> Mapper.get_align_score / ..._w_traceback / pairHMM
    set_top_matches + create_match_output (Driver.cpp)       -> Mapper.process_batch
    MPI reduce + PrintFinal inputs (Driver.cpp:1615-1823)    -> Mapper.finish / accumulator tensors

The CUDA library is mandatory: importing this module where `libgmx.so` has not been built, or
creating a Mapper without a CUDA device, raises -- there is no CPU fallback (and nothing here
touches oracle/).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (GmxIndex, GmxParams, GmxReads, GmxStageStats, HIT_DTYPE, READ_RESULT_DTYPE, IndexHandle, ReadBatch, ptr)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMX_LIB") or os.path.join(HERE, "libgmx.so")      # GMX_LIB: an alternative build of the same library (kernel experiments)

_lib = None

EXPORTS = [
    "gmx_default_params", "gmx_create", "gmx_destroy", "gmx_strerror", "gmx_last_error", "gmx_abi_version",
    "gmx_set_stream", "gmx_synchronize", "gmx_fm_search", "gmx_sa_locate", "gmx_get_windows", "gmx_self_score",
    "gmx_nw_score", "gmx_nw_traceback", "gmx_pair_hmm", "gmx_map_batch", "gmx_score_batch", "gmx_process_batch",
    "gmx_get_hits", "gmx_get_best_alignments", "gmx_accumulators_device", "gmx_reset_accumulators", "gmx_finish",
    "gmx_get_stage_stats", "gmx_set_option", "gmx_fastq_scan_host", "gmx_fastq_scan", "gmx_process_fastq", "gmx_format_sam", "gmx_format_sgr", "gmx_format_gmp", "gmx_snp_call",
    "gmx_comm_create", "gmx_comm_reduce", "gmx_comm_stats", "gmx_comm_destroy", "gmx_measure_alu_peak", "gmx_chunk_stats", "gmx_format_g", "gmx_index_sizes", "gmx_index_build",
]

OPT_COLLECT_HITS, OPT_CHUNK_READS, OPT_VOTE_FILTER, OPT_FILTER_SHIFT, OPT_CIGAR_STRIDE, OPT_VOTE_SLOTS, OPT_VOTE_COMPACT, OPT_SAM_DEVICE, OPT_FASTQ_PIECE, OPT_OPTIMISTIC, OPT_STAGE_TIMING = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11
COMM_AUTO, COMM_PEER, COMM_NCCL = 0, 1, 2


class GmxError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: {code} ({detail})")


def load_library():
    """dlopen gnumap_b200/libgmx.so; fails loudly when the CUDA extension is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(gnumap_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.gmx_strerror.restype = C.c_char_p
        L.gmx_last_error.restype = C.c_char_p
        L.gmx_last_error.argtypes = [C.c_void_p]
        L.gmx_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_int]
        L.gmx_destroy.argtypes = [C.c_void_p]
        L.gmx_destroy.restype = None
        L.gmx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.gmx_synchronize.argtypes = [C.c_void_p]
        L.gmx_fm_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
        L.gmx_sa_locate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.gmx_get_windows.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
        L.gmx_self_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gmx_nw_score.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.gmx_nw_traceback.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.gmx_pair_hmm.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        for f in ("gmx_map_batch", "gmx_process_batch"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gmx_score_batch.argtypes = [C.c_void_p, C.c_void_p]
        L.gmx_get_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.gmx_get_best_alignments.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]
        L.gmx_accumulators_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_uint64)]
        L.gmx_reset_accumulators.argtypes = [C.c_void_p]
        L.gmx_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gmx_get_stage_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.gmx_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        L.gmx_fastq_scan_host.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.gmx_fastq_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.gmx_format_sam.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.gmx_format_sgr.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.gmx_snp_call.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_char_p, C.c_int]
        L.gmx_format_gmp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.gmx_process_fastq.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]
        L.gmx_index_sizes.argtypes = [C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gmx_index_build.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_int32), C.c_char_p, C.c_int]
        L.gmx_measure_alu_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.gmx_chunk_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gmx_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int]
        L.gmx_comm_reduce.argtypes = [C.c_void_p, C.c_int]
        L.gmx_comm_stats.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.gmx_comm_destroy.argtypes = [C.c_void_p]
        L.gmx_comm_destroy.restype = None
        _lib = L
    return _lib


def default_params() -> GmxParams:
    p = GmxParams()
    load_library().gmx_default_params(C.byref(p))
    return p


def fastq_scan_host(text: bytes, illumina: int = 0) -> np.ndarray:
    """SeqReader::get_more_fastq record index of `text` (reference semantics incl. recovery); needs no GPU."""
    L = load_library()
    buf = np.frombuffer(text, dtype=np.uint8)
    n = C.c_int64(0)
    cap = max(text.count(b"\n") // 4 + 2, 4)
    recs = np.zeros(cap, dtype=_abi.FASTQ_REC_DTYPE)
    rc = L.gmx_fastq_scan_host(buf.ctypes.data if len(buf) else recs.ctypes.data, len(buf), illumina, recs.ctypes.data, cap, C.byref(n))
    if rc != 0:
        raise GmxError(rc, "gmx_fastq_scan_host", L.gmx_strerror(rc).decode())
    return recs[: n.value]


def snp_call(counts, genome_base: int, monoploid: bool = False, snp_pval: float = 0.001):
    """GenomeBwt::is_snp + the PrintSNPCall column for one position's read counts (A,C,G,T,N); needs no GPU.
    Returns (first, second, diploid, pval, text)."""
    L = load_library()
    c = np.ascontiguousarray(counts, dtype=np.float32)
    first, second, dip, pv = C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0.0)
    text = C.create_string_buffer(96)
    rc = L.gmx_snp_call(c.ctypes.data, genome_base, int(monoploid), snp_pval, C.byref(first), C.byref(second), C.byref(dip), C.byref(pv), text, 96)
    if rc != 0:
        raise GmxError(rc, "gmx_snp_call", L.gmx_strerror(rc).decode())
    return first.value, second.value, bool(dip.value), pv.value, text.value


def batch_from_fastq(text: bytes, recs: np.ndarray):
    """(names, ReadBatch) of a record index: Read::name / seq / fq with the quality cut to the sequence length."""
    names = [text[int(r["name_off"]): int(r["name_off"]) + int(r["name_len"])].decode() for r in recs]
    seqs = [text[int(r["seq_off"]): int(r["seq_off"]) + int(r["seq_len"])] for r in recs]
    quals = [text[int(r["qual_off"]): int(r["qual_off"]) + int(r["seq_len"])] for r in recs]
    return names, ReadBatch(seqs, quals)


def index_build(codes: np.ndarray, device: int = 0):
    """bwa_index on the GPU (gmx_index_build): codes uint8[n] in 0..3 -> dict(bwt uint32[], primary, L2 uint64[5], sa uint64[],
    pac uint8[], rounds)."""
    L = load_library()
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n = len(codes)
    w, ns, pb = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    rc = L.gmx_index_sizes(n, C.byref(w), C.byref(ns), C.byref(pb))
    if rc != 0:
        raise GmxError(rc, "gmx_index_sizes", L.gmx_strerror(rc).decode())
    bwt = np.zeros(w.value, dtype=np.uint32); sa = np.zeros(ns.value, dtype=np.uint64); pac = np.zeros(pb.value, dtype=np.uint8)
    L2 = np.zeros(5, dtype=np.uint64)
    primary = C.c_uint64(0); rounds = C.c_int32(0)
    err = C.create_string_buffer(512)
    rc = L.gmx_index_build(ptr(codes), n, device, ptr(bwt), C.byref(primary), ptr(L2), ptr(sa), ptr(pac), C.byref(rounds), err, 512)
    if rc != 0:
        raise GmxError(rc, "gmx_index_build", err.value.decode() or L.gmx_strerror(rc).decode())
    return dict(bwt=bwt, primary=int(primary.value), L2=L2, sa=sa, pac=pac, rounds=int(rounds.value))


class Mapper:
    """One GPU's worth of the GNUMAP hot path (one process per GPU)."""

    def __init__(self, index, params: GmxParams | None = None, device: int = 0):
        self.L = load_library()
        self.index = index
        self.params = params if params is not None else default_params()
        self._ih = IndexHandle(index)
        self._ctx = C.c_void_p()
        rc = self.L.gmx_create(C.byref(self._ctx), C.addressof(self._ih.struct), C.addressof(self.params), device)
        if rc != 0:
            detail = self.L.gmx_last_error(self._ctx).decode() if self._ctx else ""
            if self._ctx:
                self.L.gmx_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise GmxError(rc, "gmx_create", detail or self.L.gmx_strerror(rc).decode())

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc: int, where: str):
        if rc != 0:
            raise GmxError(rc, where, self.L.gmx_last_error(self._ctx).decode() or self.L.gmx_strerror(rc).decode())

    def close(self):
        if self._ctx:
            self.L.gmx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._ck(self.L.gmx_set_stream(self._ctx, C.c_void_p(cuda_stream)), "gmx_set_stream")

    def set_option(self, option: int, value: int):
        self._ck(self.L.gmx_set_option(self._ctx, option, value), "gmx_set_option")

    def synchronize(self):
        self._ck(self.L.gmx_synchronize(self._ctx), "gmx_synchronize")

    # -- GenomeBwt ----------------------------------------------------------------------------
    def get_sa_int(self, kmers):
        """kmers: list of equal-length bytes, or uint8[n, len] -> (start uint64[n], end uint64[n])."""
        arr = np.frombuffer(b"".join(kmers), dtype=np.uint8).reshape(len(kmers), -1).copy() if isinstance(kmers, (list, tuple)) else np.ascontiguousarray(kmers, dtype=np.uint8)
        n, ln = arr.shape
        k = np.zeros(n, dtype=np.uint64); l = np.zeros(n, dtype=np.uint64)
        self._ck(self.L.gmx_fm_search(self._ctx, ptr(arr), ln, n, ptr(k), ptr(l)), "gmx_fm_search")
        return k, l

    def get_sa_coord(self, ranks, sampled: bool = False):
        r = np.ascontiguousarray(ranks, dtype=np.uint64)
        out = np.zeros(len(r), dtype=np.uint64)
        self._ck(self.L.gmx_sa_locate(self._ctx, ptr(r), len(r), 1 if sampled else 0, ptr(out)), "gmx_sa_locate")
        return out

    def GetString(self, begins, size: int):
        b = np.ascontiguousarray(begins, dtype=np.uint64)
        chars = np.zeros((len(b), size), dtype=np.uint8); lens = np.zeros(len(b), dtype=np.int32)
        self._ck(self.L.gmx_get_windows(self._ctx, ptr(b), len(b), size, ptr(chars), ptr(lens)), "gmx_get_windows")
        return [chars[i, : lens[i]].tobytes() for i in range(len(b))]

    # -- bin_seq ------------------------------------------------------------------------------
    def self_score(self, batch: ReadBatch):
        out = np.zeros(batch.n_reads, dtype=np.float32)
        self._ck(self.L.gmx_self_score(self._ctx, C.addressof(batch.struct), ptr(out)), "gmx_self_score")
        return out

    @staticmethod
    def _pack_windows(windows, stride=None):
        stride = stride or max(len(w) for w in windows)
        buf = np.zeros((len(windows), stride), dtype=np.uint8)
        for i, w in enumerate(windows):
            buf[i, : len(w)] = np.frombuffer(w, dtype=np.uint8)
        return buf, stride

    def get_align_score(self, batch: ReadBatch, read_idx, strands, windows):
        ri = np.ascontiguousarray(read_idx, dtype=np.int32); st = np.ascontiguousarray(strands, dtype=np.uint8)
        wb, stride = self._pack_windows(windows)
        out = np.zeros(len(ri), dtype=np.float32)
        self._ck(self.L.gmx_nw_score(self._ctx, C.addressof(batch.struct), len(ri), ptr(ri), ptr(st), ptr(wb), stride, ptr(out)), "gmx_nw_score")
        return out

    def get_align_score_w_traceback(self, batch: ReadBatch, read_idx, strands, windows, consensus=None):
        ri = np.ascontiguousarray(read_idx, dtype=np.int32); st = np.ascontiguousarray(strands, dtype=np.uint8)
        wb, stride = self._pack_windows(windows)
        cb = None
        if consensus is not None:
            cb, _ = self._pack_windows(consensus, stride)
        a_stride, c_stride = 2 * stride + 16, 128
        aligned = np.zeros((len(ri), a_stride), dtype=np.uint8); alen = np.zeros(len(ri), dtype=np.int32)
        cigar = np.zeros((len(ri), c_stride), dtype=np.uint8)
        self._ck(self.L.gmx_nw_traceback(self._ctx, C.addressof(batch.struct), len(ri), ptr(ri), ptr(st), ptr(wb), stride, ptr(cb),
                                         ptr(aligned), a_stride, ptr(alen), ptr(cigar), c_stride), "gmx_nw_traceback")
        return [(aligned[i, : alen[i]].tobytes(), bytes(cigar[i]).split(b"\0")[0].decode()) for i in range(len(ri))]

    def pairHMM(self, batch: ReadBatch, read_idx, strands, windows):
        ri = np.ascontiguousarray(read_idx, dtype=np.int32); st = np.ascontiguousarray(strands, dtype=np.uint8)
        wb, stride = self._pack_windows(windows)
        out = np.zeros((len(ri), stride, 5), dtype=np.float32)
        self._ck(self.L.gmx_pair_hmm(self._ctx, C.addressof(batch.struct), len(ri), ptr(ri), ptr(st), ptr(wb), stride, ptr(out)), "gmx_pair_hmm")
        return out

    # -- batch pipeline -----------------------------------------------------------------------
    def _fetch(self, batch, out, scored: bool):
        n = C.c_int64(0)
        self._ck(self.L.gmx_get_hits(self._ctx, None, 0, C.byref(n)), "gmx_get_hits")
        hits = np.zeros(max(n.value, 1), dtype=HIT_DTYPE)
        self._ck(self.L.gmx_get_hits(self._ctx, ptr(hits), len(hits), C.byref(n)), "gmx_get_hits")
        out["hits"] = hits[: n.value]
        if scored:
            cs, as_ = 64, 512
            cig = np.zeros((batch.n_reads, cs), dtype=np.uint8); al = np.zeros((batch.n_reads, as_), dtype=np.uint8)
            self._ck(self.L.gmx_get_best_alignments(self._ctx, ptr(cig), cs, ptr(al), as_), "gmx_get_best_alignments")
            out["cigars"] = [bytes(r).split(b"\0")[0].decode() for r in cig]
            out["aligned"] = al
        return out

    def process_batch(self, batch, score: bool = True, fetch: bool = True, results: np.ndarray | None = None):
        """PHASE A (+ PHASE B).  Returns dict(results, hits, cigars, aligned).  `batch` is a host ReadBatch or any
        object with a `.struct` gmx_reads (e.g. DeviceReadBatch); `results` may be a caller-owned (pinned) array."""
        if results is None:
            results = np.zeros(batch.n_reads, dtype=READ_RESULT_DTYPE)
        fn = self.L.gmx_process_batch if score else self.L.gmx_map_batch
        self._ck(fn(self._ctx, C.addressof(batch.struct), C.c_void_p(results.ctypes.data)), "gmx_process_batch" if score else "gmx_map_batch")
        out = dict(results=results)
        return self._fetch(batch, out, score) if fetch else out

    def score_batch(self, batch, results: np.ndarray | None = None, fetch: bool = True):
        """PHASE B for the batch last passed to process_batch(score=False) (gmx_map_batch)."""
        if results is None:
            results = np.zeros(batch.n_reads, dtype=READ_RESULT_DTYPE)
        self._ck(self.L.gmx_score_batch(self._ctx, C.c_void_p(results.ctypes.data)), "gmx_score_batch")
        out = dict(results=results)
        return self._fetch(batch, out, True) if fetch else out

    def fastq_scan(self, text: bytes) -> np.ndarray:
        """Device FASTQ indexer (well-formed text only: GmxError(GMX_ERR_FORMAT) otherwise)."""
        buf = np.frombuffer(text, dtype=np.uint8)
        cap = max(text.count(b"\n") // 4 + 2, 4)
        recs = np.zeros(cap, dtype=_abi.FASTQ_REC_DTYPE)
        n = C.c_int64(0)
        self._ck(self.L.gmx_fastq_scan(self._ctx, buf.ctypes.data, len(buf), 0, recs.ctypes.data, cap, C.byref(n)), "gmx_fastq_scan")
        return recs[: n.value]

    def process_fastq(self, text: bytes, fetch: bool = True):
        """FASTQ text -> PHASE A + B with the reads used in place on the device; falls back to the host scan (the
        reference's recovery path) when the text is not well-formed.  Returns (names, dict as process_batch)."""
        buf = np.frombuffer(text, dtype=np.uint8)
        cap = max(text.count(b"\n") // 4 + 2, 4)
        results = np.zeros(cap, dtype=READ_RESULT_DTYPE); recs = np.zeros(cap, dtype=_abi.FASTQ_REC_DTYPE)
        n = C.c_int64(0)
        rc = self.L.gmx_process_fastq(self._ctx, buf.ctypes.data if len(buf) else recs.ctypes.data, len(buf), 0, results.ctypes.data, cap, C.byref(n), recs.ctypes.data)
        if rc == _abi.GMX_ERR_FORMAT and n.value > 0:
            # a pipelined text whose later part needs the reference's recovery path: the first n reads are done (mapped and
            # scored); the rest goes through the host scan
            done = n.value
            cut = int(recs[done - 1]["qual_off"]) + int(recs[done - 1]["qual_len"]) + 1
            rest = text[cut:]
            recs2 = fastq_scan_host(rest, self.params.illumina)
            names2, batch2 = batch_from_fastq(rest, recs2)
            out2 = self.process_batch(batch2, fetch=False)
            names1, _ = batch_from_fastq(text, recs[:done])
            return names1 + names2, dict(results=np.concatenate([results[:done], out2["results"]]))
        if rc == _abi.GMX_ERR_FORMAT:
            recs = fastq_scan_host(text, self.params.illumina)
            names, batch = batch_from_fastq(text, recs)
            return names, self.process_batch(batch, fetch=fetch)
        self._ck(rc, "gmx_process_fastq")
        recs = recs[: n.value]
        names, batch = batch_from_fastq(text, recs)
        out = dict(results=results[: n.value])
        return names, (self._fetch(batch, out, True) if fetch else out)

    def format_sam(self, text: bytes, recs: np.ndarray, results: np.ndarray) -> bytes:
        """SAM body of the batch last scored (ScoredSeq::get_SAM + the reference's writer), as bytes."""
        buf = np.frombuffer(text, dtype=np.uint8)
        recs = np.ascontiguousarray(recs); results = np.ascontiguousarray(results)
        names = (C.c_char_p * len(self.index.names))(*[nm.encode() for nm in self.index.names])
        n = C.c_int64(0)
        cap = int(len(results)) * 64 + int(recs["seq_len"].sum() + recs["qual_len"].sum() + recs["name_len"].sum()) * 2 + 4096
        while True:
            out = np.zeros(cap, dtype=np.uint8)
            rc = self.L.gmx_format_sam(self._ctx, buf.ctypes.data, recs.ctypes.data, results.ctypes.data, len(results), names, out.ctypes.data, cap, C.byref(n))
            if rc == _abi.GMX_ERR_OVERFLOW:
                cap = n.value + 16
                continue
            self._ck(rc, "gmx_format_sam")
            return out[: n.value].tobytes()

    def format_sgr(self, min_print: float = 0.001) -> bytes:
        """The .sgr text of the accumulators as they stand on the device (GenomeBwt::PrintFinalSGR)."""
        names = (C.c_char_p * len(self.index.names))(*[nm.encode() for nm in self.index.names])
        n = C.c_int64(0)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=np.uint8)
            rc = self.L.gmx_format_sgr(self._ctx, names, min_print, out.ctypes.data, cap, C.byref(n))
            if rc == _abi.GMX_ERR_OVERFLOW:
                cap = n.value + 16
                continue
            self._ck(rc, "gmx_format_sgr")
            return out[: n.value].tobytes()

    def format_gmp(self, target_base: int = -1, min_print: float = 0.001, snp_pval: float = 0.001, monoploid: bool = False) -> bytes:
        """The .gmp text of the accumulators as they stand on the device: GenomeBwt::PrintFinalSNP with the
        likelihood-ratio SNP call in SNP mode, GenomeBwt::PrintFinalBisulfite (rows at genome base `target_base`,
        0..3 = a,c,g,t) in bisulfite / A->G mode."""
        names = (C.c_char_p * len(self.index.names))(*[nm.encode() for nm in self.index.names])
        n = C.c_int64(0)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=np.uint8)
            rc = self.L.gmx_format_gmp(self._ctx, names, target_base, min_print, snp_pval, int(monoploid), out.ctypes.data, cap, C.byref(n))
            if rc == _abi.GMX_ERR_OVERFLOW:
                cap = n.value + 16
                continue
            self._ck(rc, "gmx_format_gmp")
            return out[: n.value].tobytes()

    def best_cigars(self, n_reads: int, stride: int = 64) -> np.ndarray:
        cig = np.zeros((n_reads, stride), dtype=np.uint8)
        self._ck(self.L.gmx_get_best_alignments(self._ctx, ptr(cig), stride, None, 0), "gmx_get_best_alignments")
        return cig

    def accumulators_device(self):
        amount = C.c_void_p(); n_amount = C.c_uint64(); planes = (C.c_void_p * 5)(); n_plane = C.c_uint64()
        self._ck(self.L.gmx_accumulators_device(self._ctx, C.byref(amount), C.byref(n_amount), planes, C.byref(n_plane)), "gmx_accumulators_device")
        return amount.value, n_amount.value, [planes[b] for b in range(5)], n_plane.value

    def reset_accumulators(self):
        self._ck(self.L.gmx_reset_accumulators(self._ctx), "gmx_reset_accumulators")

    def finish(self):
        """Download the accumulators: (amount float32[n_amount], planes float32[5, l_pac] | None)."""
        _, n_amount, _, n_plane = self.accumulators_device()
        amount = np.zeros(n_amount, dtype=np.float32)
        planes = np.zeros((5, n_plane), dtype=np.float32) if n_plane else None
        pl = (C.c_void_p * 5)()
        for b in range(5):
            pl[b] = planes[b].ctypes.data if planes is not None else None
        self._ck(self.L.gmx_finish(self._ctx, ptr(amount), pl), "gmx_finish")
        return amount, planes

    def alu_peak(self, kind: int = 0) -> float:
        """Measured tera lane-operations per second: 0 = FP32 mul + add (no FMA), 1 = FP32 add + max, 2 = FP64 mul + add."""
        v = C.c_double(0.0)
        self._ck(self.L.gmx_measure_alu_peak(self._ctx, kind, C.byref(v)), "gmx_measure_alu_peak")
        return v.value

    def chunk_stats(self):
        """(chunks issued without a host wait, of those run again) since the context was created -- GMX_OPT_OPTIMISTIC."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self.L.gmx_chunk_stats(self._ctx, C.byref(a), C.byref(b)), "gmx_chunk_stats")
        return a.value, b.value

    def stage_stats(self):
        s = GmxStageStats()
        self._ck(self.L.gmx_get_stage_stats(self._ctx, C.byref(s)), "gmx_get_stage_stats")
        return {s.name[i].decode(): dict(ms=s.ms[i], units=s.units[i], bytes=s.bytes[i], launches=s.launches[i]) for i in range(s.n_stages)}


class Comm:
    """Several Mappers (one per GPU, or several on one GPU) of ONE process whose accumulators are terms of one sum --
    the in-process counterpart of the reference's MPI reduce (reference src/Driver.cpp:1615-1811).  mappers[0] is the
    root: its finish() reduces first."""

    def __init__(self, mappers, backend: int = COMM_AUTO):
        self.L = load_library()
        self.mappers = list(mappers)
        arr = (C.c_void_p * len(self.mappers))(*[m._ctx for m in self.mappers])
        self._comm = C.c_void_p()
        rc = self.L.gmx_comm_create(C.byref(self._comm), arr, len(self.mappers), backend)
        if rc != 0:
            raise GmxError(rc, "gmx_comm_create", self.L.gmx_last_error(self.mappers[0]._ctx).decode() or self.L.gmx_strerror(rc).decode())

    def reduce(self, all: bool = False):
        rc = self.L.gmx_comm_reduce(self._comm, int(all))
        if rc != 0:
            raise GmxError(rc, "gmx_comm_reduce", self.L.gmx_last_error(self.mappers[0]._ctx).decode())

    def stats(self):
        ms, nb, be = C.c_float(0), C.c_uint64(0), C.c_int(0)
        self.L.gmx_comm_stats(self._comm, C.byref(ms), C.byref(nb), C.byref(be))
        return dict(ms=ms.value, bytes=nb.value, backend={COMM_PEER: "peer", COMM_NCCL: "nccl"}.get(be.value, str(be.value)))

    def close(self):
        if self._comm:
            self.L.gmx_comm_destroy(self._comm)
            self._comm = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
