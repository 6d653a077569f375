"""ctypes mirror of include/gmx.h (structs and constants only -- no library is loaded here)."""
from __future__ import annotations

import ctypes as C

import numpy as np

GMX_OK = 0
GMX_ERR_INVALID, GMX_ERR_CUDA, GMX_ERR_NOMEM, GMX_ERR_UNSUPPORTED = -1, -2, -3, -4
GMX_ERR_OVERFLOW, GMX_ERR_NO_DEVICE, GMX_ERR_STATE, GMX_ERR_FORMAT = -5, -6, -7, -8

READ_MAPPED, READ_UNMATCHED, READ_TOO_SHORT, READ_TOO_POOR, READ_TOO_MANY = 0, 1, 2, 3, 4
POS_STRAND, NEG_STRAND = 0, 1
MODE_NORMAL, MODE_BS, MODE_SNP = 0, 1, 2
GMX_N_STAGES = 12


class GmxIndex(C.Structure):
    _fields_ = [
        ("bwt", C.c_void_p), ("bwt_words", C.c_uint64), ("primary", C.c_uint64),
        ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64),
        ("sa", C.c_void_p), ("n_sa", C.c_uint64), ("sa_intv", C.c_int32), ("n_seqs", C.c_int32),
        ("pac", C.c_void_p), ("l_pac", C.c_int64),
        ("seq_offset", C.c_void_p), ("seq_len_arr", C.c_void_p),
    ]


class GmxParams(C.Structure):
    _fields_ = [
        ("align_scores", (C.c_float * 4) * 256), ("phmm_scores", (C.c_float * 4) * 256),
        ("gap", C.c_float), ("max_gap", C.c_int32), ("mer", C.c_int32), ("jump", C.c_int32),
        ("min_seed_hits", C.c_int32), ("max_kmer_hits", C.c_uint32), ("max_matches", C.c_uint32),
        ("gen_size", C.c_uint32), ("align_score", C.c_float), ("perc", C.c_int32),
        ("cutoff", C.c_float), ("match_pos", C.c_int32), ("match_neg", C.c_int32),
        ("unique_only", C.c_int32), ("fast", C.c_int32), ("use_nw", C.c_int32),
        ("mode", C.c_int32), ("illumina", C.c_int32), ("adjust", C.c_float),
    ]


class GmxReads(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("offsets", C.c_void_p), ("seq", C.c_void_p),
                ("qual", C.c_void_p), ("pwm", C.c_void_p), ("on_device", C.c_int32), ("max_len", C.c_int32),
                ("qual_offsets", C.c_void_p), ("lens", C.c_void_p)]


class GmxStageStats(C.Structure):
    _fields_ = [("name", C.c_char_p * GMX_N_STAGES), ("ms", C.c_float * GMX_N_STAGES),
                ("units", C.c_uint64 * GMX_N_STAGES), ("bytes", C.c_uint64 * GMX_N_STAGES),
                ("launches", C.c_int32 * GMX_N_STAGES), ("n_stages", C.c_int32)]


# numpy views of gmx_read_result / gmx_hit (natural C alignment, checked against sizeof in tests)
READ_RESULT_DTYPE = np.dtype([
    ("top_score", "<f8"), ("denominator", "<f8"), ("max_align_score", "<f4"), ("status", "<i4"),
    ("n_groups", "<i4"), ("n_candidates", "<i4"), ("best_score", "<f4"), ("best_posterior", "<f4"),
    ("best_n_positions", "<i4"), ("best_first_strand", "<i4"), ("best_first_pos", "<u8"),
    ("hit_begin", "<i4"), ("hit_end", "<i4"), ("best_group", "<i4"), ("best_aligned_len", "<i4"),
], align=True)

FASTQ_REC_DTYPE = np.dtype([("name_off", "<i8"), ("seq_off", "<i8"), ("qual_off", "<i8"), ("name_len", "<i4"),
                            ("seq_len", "<i4"), ("qual_len", "<i4"), ("pad", "<i4")], align=True)

HIT_DTYPE = np.dtype([
    ("pos", "<u8"), ("score", "<f4"), ("read", "<i4"), ("group", "<i2"), ("strand", "u1"),
    ("first_strand", "u1"),
], align=True)


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class IndexHandle:
    """Keeps the numpy arrays alive next to the ctypes struct that borrows them."""

    def __init__(self, ix):
        self.arrays = dict(
            bwt=np.ascontiguousarray(ix.bwt, dtype=np.uint32),
            sa=np.ascontiguousarray(ix.sa, dtype=np.uint64),
            pac=np.ascontiguousarray(ix.pac, dtype=np.uint8),
            off=np.ascontiguousarray(ix.seq_offset, dtype=np.int64),
            lens=np.ascontiguousarray(ix.seq_len_arr, dtype=np.int32),
        )
        s = GmxIndex()
        s.bwt = ptr(self.arrays["bwt"]); s.bwt_words = len(self.arrays["bwt"])
        s.primary = int(ix.primary)
        for i in range(5):
            s.L2[i] = int(ix.L2[i])
        s.seq_len = int(ix.seq_len)
        s.sa = ptr(self.arrays["sa"]); s.n_sa = len(self.arrays["sa"]); s.sa_intv = int(ix.sa_intv)
        s.n_seqs = len(self.arrays["off"])
        s.pac = ptr(self.arrays["pac"]); s.l_pac = int(ix.l_pac)
        s.seq_offset = ptr(self.arrays["off"]); s.seq_len_arr = ptr(self.arrays["lens"])
        self.struct = s


class ReadBatch:
    """Host-side batch in the layout of gmx_reads (concatenated ASCII seq / qual + offsets)."""

    def __init__(self, seqs, quals=None, pwm=None):
        lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
        self.offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(lens, out=self.offsets[1:])
        self.seq = np.frombuffer(b"".join(bytes(s) for s in seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
        self.qual = None
        if quals is not None:
            self.qual = np.frombuffer(b"".join(bytes(q) for q in quals), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
            assert len(self.qual) == len(self.seq)
        self.pwm = None if pwm is None else np.ascontiguousarray(pwm, dtype=np.float32)
        self._mk()

    @classmethod
    def from_arrays(cls, bases_codes: np.ndarray, quals_phred: np.ndarray):
        """bases_codes uint8[n, L] in 0..4, quals_phred uint8[n, L] -> ASCII batch (fixed length)."""
        self = cls.__new__(cls)
        n, L = bases_codes.shape
        self.offsets = (np.arange(n + 1, dtype=np.int64) * L)
        self.seq = np.frombuffer(b"ACGTN", dtype=np.uint8)[bases_codes].reshape(-1).copy()
        self.qual = (quals_phred.astype(np.uint8) + 33).reshape(-1).copy()
        self.pwm = None
        self._mk()
        return self

    def _mk(self):
        s = GmxReads()
        s.n_reads = len(self.offsets) - 1
        s.offsets = ptr(self.offsets); s.seq = ptr(self.seq); s.qual = ptr(self.qual); s.pwm = ptr(self.pwm)
        self.struct = s

    @property
    def n_reads(self) -> int:
        return len(self.offsets) - 1

    def slice(self, lo: int, hi: int) -> "ReadBatch":
        out = ReadBatch.__new__(ReadBatch)
        a, b = int(self.offsets[lo]), int(self.offsets[hi])
        out.offsets = (self.offsets[lo:hi + 1] - a).copy()
        out.seq = self.seq[a:b].copy()
        out.qual = None if self.qual is None else self.qual[a:b].copy()
        out.pwm = None if self.pwm is None else self.pwm[a:b].copy()
        out._mk()
        return out
