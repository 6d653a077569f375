"""ctypes wrapper around oracle/liboracle.so (and oracle/_ref/libref_probe.so when present).

TEST INFRASTRUCTURE ONLY -- see oracle/gnumap_oracle.h.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by
gnumap_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from gnumap_b200 import _abi
from gnumap_b200._abi import GmxIndex, GmxParams, GmxReads, HIT_DTYPE, READ_RESULT_DTYPE, ptr

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_DIR = os.environ.get("GMX_REF_DIR") or os.path.join(HERE, "_ref")     # GMX_REF_DIR: another build of the reference (e.g. against real GSL)
REF_BIN = os.path.join(REF_DIR, "gnumap")
REF_PROBE = os.path.join(REF_DIR, "libref_probe.so")

_lib = None


def build(force: bool = False) -> None:
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "gnumap_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_self_score.restype = C.c_float
        L.orc_nw_score.restype = C.c_float
        L.orc_nw_score.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int]
        L.orc_nw_traceback.restype = C.c_int
        L.orc_nw_traceback.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_bwt_occ.restype = C.c_uint64
        L.orc_bwt_occ.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.orc_bwt_sa.restype = C.c_uint64
        L.orc_bwt_sa.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_get_string.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        L.orc_get_sa_int.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_max_char.restype = C.c_char
        L.orc_process_batch.restype = C.c_int
        L.orc_process_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def default_params() -> GmxParams:
    p = GmxParams()
    lib().orc_default_params(C.byref(p))
    return p


def table_np(t) -> np.ndarray:
    return np.ctypeslib.as_array(t).reshape(256, 4).copy()


def fastq_pwm(seq: bytes, qual: bytes, illumina: int = 0) -> np.ndarray:
    n = len(seq)
    out = np.zeros((n, 4), dtype=np.float32)
    s = np.frombuffer(seq, dtype=np.uint8).copy(); q = np.frombuffer(qual, dtype=np.uint8).copy()
    lib().orc_fastq_pwm(ptr(s), ptr(q), n, illumina, ptr(out))
    return out


def onehot_pwm(seq: bytes) -> np.ndarray:
    """PWM of a FASTA read as the reference's tests build it (one-hot rows)."""
    out = np.zeros((len(seq), 4), dtype=np.float32)
    for i, c in enumerate(seq.lower()):
        k = b"acgt".find(bytes([c]))
        if k >= 0:
            out[i, k] = 1.0
    return out


def revcomp_pwm(pwm: np.ndarray) -> np.ndarray:
    out = np.zeros_like(pwm)
    lib().orc_revcomp_pwm(ptr(np.ascontiguousarray(pwm)), len(pwm), ptr(out))
    return out


def max_char_consensus(pwm: np.ndarray) -> bytes:
    pwm = np.ascontiguousarray(pwm, dtype=np.float32)
    return b"".join(lib().orc_max_char(C.c_void_p(pwm.ctypes.data + 16 * i)) for i in range(len(pwm)))


def self_score(pwm: np.ndarray, consensus: bytes, params: GmxParams) -> float:
    pwm = np.ascontiguousarray(pwm, dtype=np.float32)
    c = np.frombuffer(consensus, dtype=np.uint8).copy()
    return float(lib().orc_self_score(ptr(pwm), ptr(c), len(pwm), C.byref(params.align_scores)))


def nw_score(pwm: np.ndarray, gen: bytes, params: GmxParams) -> float:
    pwm = np.ascontiguousarray(pwm, dtype=np.float32)
    g = np.frombuffer(gen, dtype=np.uint8).copy()
    return float(lib().orc_nw_score(ptr(pwm), len(pwm), ptr(g), C.addressof(params.align_scores),
                                    params.gap, params.max_gap))


def nw_traceback(pwm: np.ndarray, consensus: bytes, gen: bytes, params: GmxParams):
    pwm = np.ascontiguousarray(pwm, dtype=np.float32)
    c = np.frombuffer(consensus + b"\0", dtype=np.uint8).copy()
    g = np.frombuffer(gen, dtype=np.uint8).copy()
    cap = len(pwm) + len(gen) + 16
    aligned = np.zeros(cap, dtype=np.uint8); cigar = np.zeros(1024, dtype=np.uint8)
    n = lib().orc_nw_traceback(ptr(pwm), len(pwm), ptr(c), ptr(g), len(gen), C.addressof(params.align_scores),
                               params.gap, params.max_gap, ptr(aligned), cap, ptr(cigar), 1024)
    return aligned[:n].tobytes(), cigar.tobytes().split(b"\0")[0].decode()


def pair_hmm(pwm: np.ndarray, consensus: bytes, gen: bytes, params: GmxParams) -> np.ndarray:
    pwm = np.ascontiguousarray(pwm, dtype=np.float32)
    c = np.frombuffer(consensus, dtype=np.uint8).copy(); g = np.frombuffer(gen, dtype=np.uint8).copy()
    out = np.zeros((len(gen), 5), dtype=np.float32)
    lib().orc_pair_hmm(ptr(pwm), len(pwm), ptr(c), ptr(g), len(gen), C.byref(params.phmm_scores), ptr(out))
    return out


class OracleIndex:
    def __init__(self, ix):
        self.ix = ix
        self.h = _abi.IndexHandle(ix)
        self.p = C.addressof(self.h.struct)

    def get_sa_int(self, kmer: bytes):
        k = np.frombuffer(kmer, dtype=np.uint8).copy()
        s = C.c_uint64(); e = C.c_uint64()
        lib().orc_get_sa_int(self.p, ptr(k), len(kmer), C.byref(s), C.byref(e))
        return s.value, e.value

    def bwt_sa(self, k: int) -> int:
        return int(lib().orc_bwt_sa(self.p, k))

    def bwt_occ(self, k: int, c: int) -> int:
        return int(lib().orc_bwt_occ(self.p, C.c_uint64(k & 0xFFFFFFFFFFFFFFFF), c))

    def get_string(self, begin: int, size: int) -> bytes:
        out = np.zeros(size, dtype=np.uint8)
        n = lib().orc_get_string(self.p, begin, size, ptr(out))
        return out[:n].tobytes()


def n_accum(ix, params) -> int:
    return (int(ix.l_pac) + params.gen_size - 1) // params.gen_size


def process_batch(oix: OracleIndex, params: GmxParams, batch: _abi.ReadBatch, do_score: bool = True,
                  amount: np.ndarray | None = None, planes: np.ndarray | None = None, hits_cap: int | None = None):
    """Returns dict(results, hits, cigars, aligned, amount, planes)."""
    n = batch.n_reads
    results = np.zeros(n, dtype=READ_RESULT_DTYPE)
    cap = hits_cap or max(1024, 64 * n)
    hits = np.zeros(cap, dtype=HIT_DTYPE)
    n_hits = C.c_int64(0)
    cig_stride, al_stride = 256, 512
    cigars = np.zeros((n, cig_stride), dtype=np.uint8)
    aligned = np.zeros((n, al_stride), dtype=np.uint8)
    if amount is None:
        amount = np.zeros(n_accum(oix.ix, params), dtype=np.float32)
    if planes is None and params.mode != _abi.MODE_NORMAL:
        planes = np.zeros((5, n_accum(oix.ix, params)), dtype=np.float32)
    pl = (C.c_void_p * 5)()
    for b in range(5):
        pl[b] = None if planes is None else planes[b].ctypes.data
    rc = lib().orc_process_batch(oix.p, C.addressof(params), C.addressof(batch.struct), int(do_score),
                                 ptr(results), ptr(hits), cap, C.byref(n_hits),
                                 ptr(cigars), cig_stride, ptr(aligned), al_stride, ptr(amount), pl)
    if rc == _abi.GMX_ERR_OVERFLOW:
        return process_batch(oix, params, batch, do_score, None, None, hits_cap=int(n_hits.value) + 16)
    if rc != 0:
        raise RuntimeError(f"orc_process_batch failed: {rc}")
    return dict(results=results, hits=hits[: n_hits.value].copy(),
                cigars=[bytes(r).split(b"\0")[0].decode() for r in cigars],
                aligned=aligned, amount=amount, planes=planes)


# ----------------------------------------------------------------------------------------------
# the unmodified reference, compiled into oracle/_ref/ by oracle/Makefile
# ----------------------------------------------------------------------------------------------
def have_ref_binary() -> bool:
    return os.path.exists(REF_BIN)


def run_reference(genome_fa: str, reads_fq: str, out_prefix: str, threads: int = 1, extra=(), timeout=None, mmap_threshold: int = 65536):
    """Run oracle/_ref/gnumap with the hygiene SURVEY.md §8(c) requires: amount_genome is malloc'd and never
    zeroed (reference src/GenomeBwt.cpp:323); a malloc threshold at or below the array size makes glibc serve it
    from fresh zero pages.  Small test genomes need a small `mmap_threshold`."""
    env = dict(os.environ, MALLOC_MMAP_THRESHOLD_=str(mmap_threshold))
    cmd = [REF_BIN, "-g", genome_fa, "-o", out_prefix, "-a", ".9", "-c", str(threads), *extra, reads_fq]
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference run failed ({p.returncode}):\n{p.stdout[-2000:]}")
    return p.stdout


def align_score_range(pwm: np.ndarray, gen: bytes, begin: int, end: int, params: GmxParams) -> float:
    pwm = np.ascontiguousarray(pwm, dtype=np.float32)
    g = np.frombuffer(gen, dtype=np.uint8).copy()
    f = lib().orc_align_score_range
    f.restype = C.c_float
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.c_float, C.c_int]
    return float(f(ptr(pwm), len(pwm), ptr(g), begin, end, C.addressof(params.align_scores), params.gap, params.max_gap))


class RefProbe:
    """The unmodified reference behind a C shim (oracle/ref_probe.cpp).  Only available where
    oracle/_ref/libref_probe.so has been built (this container; it also travels to the GPU box)."""

    def __init__(self):
        L = C.CDLL(REF_PROBE)
        L.refp_init()
        L.refp_self_score.restype = C.c_float
        L.refp_nw_score.restype = C.c_float
        L.refp_align_score_range.restype = C.c_float
        L.refp_get_sa_coord.restype = C.c_uint64
        L.refp_get_sa_coord.argtypes = [C.c_uint64]
        L.refp_get_string.argtypes = [C.c_uint64, C.c_uint, C.c_void_p]
        self.L = L

    def tables(self):
        a = np.zeros((256, 4), np.float32); p = np.zeros((256, 4), np.float32); s = np.zeros(5, np.float32)
        self.L.refp_get_tables(ptr(a), ptr(p), ptr(s))
        return a, p, s

    def set_mode(self, mode: int):
        self.L.refp_set_mode(mode)

    def self_score(self, pwm, consensus: bytes) -> float:
        pwm = np.ascontiguousarray(pwm, dtype=np.float32)
        return float(self.L.refp_self_score(ptr(pwm), len(pwm), consensus))

    def nw_score(self, pwm, gen: bytes) -> float:
        pwm = np.ascontiguousarray(pwm, dtype=np.float32)
        return float(self.L.refp_nw_score(ptr(pwm), len(pwm), gen, len(gen)))

    def align_score_range(self, pwm, gen: bytes, begin: int, end: int) -> float:
        pwm = np.ascontiguousarray(pwm, dtype=np.float32)
        return float(self.L.refp_align_score_range(ptr(pwm), len(pwm), gen, len(gen), begin, end))

    def nw_traceback(self, pwm, consensus: bytes, gen: bytes):
        pwm = np.ascontiguousarray(pwm, dtype=np.float32)
        cap = len(pwm) + len(gen) + 16
        aligned = np.zeros(cap, np.uint8); cigar = np.zeros(1024, np.uint8)
        n = self.L.refp_nw_traceback(ptr(pwm), len(pwm), consensus, gen, len(gen), ptr(aligned), cap, ptr(cigar), 1024)
        return aligned[:n].tobytes(), cigar.tobytes().split(b"\0")[0].decode()

    def pair_hmm(self, pwm, consensus: bytes, gen: bytes) -> np.ndarray:
        pwm = np.ascontiguousarray(pwm, dtype=np.float32)
        out = np.zeros((len(pwm), 5), np.float32)
        self.L.refp_pair_hmm(ptr(pwm), len(pwm), consensus, gen, len(gen), ptr(out))
        return out

    def load_genome(self, fasta: str):
        self.L.refp_load_genome(fasta.encode())

    def get_sa_int(self, kmer: bytes):
        s = C.c_uint64(); e = C.c_uint64()
        self.L.refp_get_sa_int(kmer, len(kmer), C.byref(s), C.byref(e))
        return s.value, e.value

    def get_sa_coord(self, k: int) -> int:
        return int(self.L.refp_get_sa_coord(k))

    def get_string(self, begin: int, size: int) -> bytes:
        out = np.zeros(size + 1, np.uint8)
        n = self.L.refp_get_string(begin, size, ptr(out))
        return out[:n].tobytes()

    def snp_call(self, count: int, counts, monop: bool = False, pval: float = 0.001) -> bytes:
        """Call column GenomeBwt::PrintSNPCall prints for read counts (A,C,G,T,N) at genome position `count`
        (set_mode(2) before load_genome)."""
        c = np.ascontiguousarray(counts, dtype=np.float32)
        out = C.create_string_buffer(128)
        f = self.L.refp_snp_call
        f.argtypes = [C.c_uint64, C.c_void_p, C.c_int, C.c_float, C.c_char_p, C.c_int]
        n = f(count, ptr(c), int(monop), pval, out, 128)
        if n < 0:
            raise RuntimeError("refp_snp_call: no SNP-mode genome loaded")
        return out.value

    def score_once(self, kind, pwm, gen_string: bytes, align_score: float, positions, denom: float, cap: int):
        pwm = np.ascontiguousarray(pwm, dtype=np.float32)
        pos = np.asarray([p for p, _ in positions], dtype=np.uint64)
        st = np.asarray([s for _, s in positions], dtype=np.int32)
        amount = np.zeros(cap, np.float32)
        planes = np.zeros((5, cap), np.float32) if kind != 0 else None
        f = self.L.refp_score_once
        f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                      C.c_double, C.c_void_p, C.c_uint64, C.c_void_p]
        f(kind, ptr(pwm), len(pwm), gen_string, align_score, ptr(pos), ptr(st), len(pos), denom, ptr(amount), cap, ptr(planes))
        return amount, planes
