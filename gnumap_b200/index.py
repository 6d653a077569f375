"""GNUMAP / BWA-style FM index: build, load and save in the reference's on-disk format.

The product's C ABI consumes exactly what the reference keeps in memory after
`GenomeBwt::LoadGenome()` (reference src/GenomeBwt.cpp:59-140,282-328): the occ-interleaved BWT
(`bwt_t::bwt`, reference src/bwtindex.c:128-150), the 1/32-sampled suffix array
(`bwt_cal_sa`, reference src/bwt.c:62-84), the forward-only 2-bit `pac`
(reference src/bntseq.c:224-225,273-314) and the per-sequence annotations.

`load_index` / `save_index` read and write `<fasta>.gnumap.{bwt,sa,pac,ann,amb}` byte-for-byte
as the reference does (reference src/bwt.c:389-447, src/bntseq.c:66-96,304-314), so an index
built by either side is usable by the other.

`build_index` constructs the same arrays from scratch.  The BWT of a text is unique, so any
correct suffix sorter reproduces the reference's files bit-for-bit (tests compare against
indexes written by the compiled reference).  On a GPU (`device="cuda"`) it is the library's
own builder, `gmx_index_build` (CUDA, csrc/index_build.cuh: prefix doubling over radix sorts,
then streaming kernels for the BWT, the occ interleave, the SA sample and the pac); on the CPU
the same construction is expressed with torch sort / cumsum / numpy, for the small genomes of
the CPU tests and fixtures.  SURVEY.md §8f rank 4.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
import torch

OCC_INTERVAL = 128
SA_INTV = 32


@dataclass
class FMIndex:
    bwt: np.ndarray            # uint32[bwt_words], occ-interleaved
    primary: int
    L2: np.ndarray             # uint64[5]
    seq_len: int
    sa: np.ndarray             # uint64[n_sa], sa[0] == 2**64-1
    sa_intv: int
    pac: np.ndarray            # uint8[ceil(l_pac/4)]
    l_pac: int
    names: list = field(default_factory=list)
    seq_offset: np.ndarray = None   # int64[n_seqs]
    seq_len_arr: np.ndarray = None  # int32[n_seqs]

    @property
    def n_seqs(self) -> int:
        return len(self.names)

    def codes(self) -> np.ndarray:
        """Unpack the 2-bit genome to uint8 codes 0..3."""
        b = self.pac
        out = np.empty(len(b) * 4, dtype=np.uint8)
        out[0::4] = b >> 6
        out[1::4] = (b >> 4) & 3
        out[2::4] = (b >> 2) & 3
        out[3::4] = b & 3
        return out[: self.l_pac]

    def pos2chr(self, pos: int):
        """GenomeBwt::GetPosPair (reference src/GenomeBwt.cpp:629-635)."""
        rid = int(np.searchsorted(self.seq_offset, pos, side="right") - 1)
        return self.names[rid], pos - int(self.seq_offset[rid])


# --------------------------------------------------------------------------------------------
# suffix array
# --------------------------------------------------------------------------------------------
def suffix_array(codes: np.ndarray, device: str | torch.device = "cpu") -> np.ndarray:
    """Suffix array (int64[n]) of `codes` (values 0..3) under '$'-terminated order
    (a suffix that is a proper prefix of another sorts first)."""
    n = int(codes.shape[0])
    dev = torch.device(device)
    t = torch.from_numpy(np.ascontiguousarray(codes)).to(dev).to(torch.int64) + 1   # 1..4, 0 = past the end

    def shifted(x: torch.Tensor, h: int) -> torch.Tensor:
        out = torch.zeros_like(x)
        if h < n:
            out[: n - h] = x[h:]
        return out

    # pack the first 16 symbols, 3 bits each, by doubling
    key = t
    width = 3
    h = 1
    while h < 16:
        key = (key << (width * h)) | shifted(key, h)
        h *= 2
    del t

    def ranks_from(key: torch.Tensor):
        skey, order = torch.sort(key)
        newgrp = torch.ones(n, dtype=torch.int64, device=dev)
        newgrp[1:] = (skey[1:] != skey[:-1]).to(torch.int64)
        del skey
        r_sorted = torch.cumsum(newgrp, 0)
        del newgrp
        nranks = int(r_sorted[-1].item())
        rank = torch.empty(n, dtype=torch.int64, device=dev)
        rank[order] = r_sorted
        return rank, order, nranks

    rank, order, nranks = ranks_from(key)
    del key
    while nranks < n:
        key = rank * (n + 1) + shifted(rank, h)
        del rank, order
        rank, order, nranks = ranks_from(key)
        del key
        h *= 2
    return order.cpu().numpy()


# --------------------------------------------------------------------------------------------
# build
# --------------------------------------------------------------------------------------------
def pack_pac(codes: np.ndarray) -> np.ndarray:
    n = len(codes)
    padded = np.zeros(((n + 3) // 4) * 4, dtype=np.uint8)
    padded[:n] = codes
    q = padded.reshape(-1, 4)
    return ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)


def _interleave_occ(bwt_codes: np.ndarray) -> np.ndarray:
    """bwt_bwtupdate_core (reference src/bwtindex.c:128-150): every 128 symbols, 4 x uint64
    running counts followed by 8 x uint32 of 16 two-bit symbols (MSB first); a final count block."""
    n = len(bwt_codes)
    n_blocks = (n + OCC_INTERVAL - 1) // OCC_INTERVAL
    padded = np.zeros(n_blocks * OCC_INTERVAL, dtype=np.uint8)
    padded[:n] = bwt_codes
    # 16 symbols -> one uint32, MSB first
    sym = padded.reshape(-1, 16).astype(np.uint32)
    shifts = (np.uint32(30) - np.arange(16, dtype=np.uint32) * np.uint32(2))
    words = np.bitwise_or.reduce(sym << shifts[None, :], axis=1).astype(np.uint32)
    n_words = (n + 15) // 16                       # bwt_pac2bwt: bwt_size = (seq_len + 15) >> 4
    # running counts at block starts
    counts = np.zeros((n_blocks + 1, 4), dtype=np.uint64)
    blk = np.minimum(np.arange(n) // OCC_INTERVAL, n_blocks - 1) if n else np.zeros(0, dtype=np.int64)
    for c in range(4):
        per_block = np.bincount(blk[bwt_codes == c], minlength=n_blocks).astype(np.uint64) if n else np.zeros(0, np.uint64)
        counts[1:, c] = np.cumsum(per_block)
    out = np.zeros(n_words + (n_blocks + 1) * 8, dtype=np.uint32)
    full = np.zeros((n_blocks, 16), dtype=np.uint32)
    full[:, :8] = counts[:n_blocks].view(np.uint32).reshape(n_blocks, 8)
    full[:, 8:] = words.reshape(n_blocks, 8)
    flat = full.reshape(-1)
    # the last block may be short: only ceil(rem/16) symbol words are present in the file layout
    body_words = n_blocks * 8 + n_words
    if n_blocks:
        last_syms = n - (n_blocks - 1) * OCC_INTERVAL
        last_words = (last_syms + 15) // 16
        keep = (n_blocks - 1) * 16 + 8 + last_words
        out[:keep] = flat[:keep]
        assert keep == body_words
    out[body_words:body_words + 8] = counts[n_blocks].view(np.uint32)
    return out


def build_index(contigs, device: str | torch.device = "cpu") -> FMIndex:
    """contigs: [(name, uint8 codes 0..3)] -> FMIndex identical to what bwa_index writes
    (reference src/bwtindex.c:187-293) for a FASTA holding those sequences (no N)."""
    codes = np.concatenate([c for _, c in contigs]).astype(np.uint8)
    n = len(codes)
    dev = torch.device(device)
    if dev.type == "cuda":
        # the product's builder: gmx_index_build (csrc/index_build.cuh), CUDA behind the C ABI
        from . import api
        r = api.index_build(codes, dev.index or 0)
        offs, lens, names = [], [], []
        o = 0
        for name, c in contigs:
            names.append(name); offs.append(o); lens.append(len(c)); o += len(c)
        return FMIndex(bwt=r["bwt"], primary=r["primary"], L2=r["L2"], seq_len=n, sa=r["sa"], sa_intv=SA_INTV, pac=r["pac"], l_pac=n,
                       names=names, seq_offset=np.asarray(offs, dtype=np.int64), seq_len_arr=np.asarray(lens, dtype=np.int32))
    # CPU: the same construction with torch / numpy (tests without a GPU, fixtures)
    sa = suffix_array(codes, device)                       # int64[n], ranks 1..n of the $-augmented SA
    # full SA including '$' suffix at rank 0
    prev = sa - 1                                           # position of the preceding symbol
    primary = int(np.nonzero(sa == 0)[0][0]) + 1           # rank of suffix 0 in the augmented SA
    bwt_codes = codes[np.where(prev >= 0, prev, 0)]
    # augmented rank r = i + 1 for i in 0..n-1; '$' suffix (rank 0) contributes T[n-1]
    aug = np.empty(n + 1, dtype=np.uint8)
    aug[0] = codes[n - 1]
    aug[1:] = bwt_codes
    stored = np.delete(aug, primary)                        # BWT string without '$' (n symbols)
    L2 = np.zeros(5, dtype=np.uint64)
    L2[1:] = np.cumsum(np.bincount(codes, minlength=4)[:4]).astype(np.uint64)
    bwt = _interleave_occ(stored)
    # sampled SA: bwt_cal_sa (reference src/bwt.c:62-84)
    n_sa = (n + SA_INTV) // SA_INTV
    sa_s = np.zeros(n_sa, dtype=np.uint64)
    full = np.empty(n + 1, dtype=np.int64)
    full[0] = n
    full[1:] = sa
    sa_s[:] = full[::SA_INTV][:n_sa].astype(np.uint64)
    sa_s[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    offs, lens, names = [], [], []
    o = 0
    for name, c in contigs:
        names.append(name); offs.append(o); lens.append(len(c)); o += len(c)
    return FMIndex(bwt=bwt, primary=primary, L2=L2, seq_len=n, sa=sa_s, sa_intv=SA_INTV,
                   pac=pack_pac(codes), l_pac=n, names=names,
                   seq_offset=np.asarray(offs, dtype=np.int64), seq_len_arr=np.asarray(lens, dtype=np.int32))


def full_suffix_array(contigs, device="cpu") -> np.ndarray:
    """Augmented suffix array uint64[n+1] (rank 0 = '$' suffix = n): what bwt_sa returns for every rank
    except rank 0, whose reference value is seq_len via the -1 sentinel arithmetic."""
    codes = np.concatenate([c for _, c in contigs]).astype(np.uint8)
    sa = suffix_array(codes, device)
    full = np.empty(len(codes) + 1, dtype=np.uint64)
    full[0] = len(codes)
    full[1:] = sa.astype(np.uint64)
    return full


# --------------------------------------------------------------------------------------------
# on-disk format
# --------------------------------------------------------------------------------------------
def save_index(ix: FMIndex, prefix: str) -> None:
    with open(prefix + ".gnumap.bwt", "wb") as f:           # bwt_dump_bwt, reference src/bwt.c:389-399
        f.write(np.uint64(ix.primary).tobytes())
        f.write(ix.L2[1:5].astype(np.uint64).tobytes())
        f.write(ix.bwt.astype(np.uint32).tobytes())
    with open(prefix + ".gnumap.sa", "wb") as f:            # bwt_dump_sa, reference src/bwt.c:401-413
        f.write(np.uint64(ix.primary).tobytes())
        f.write(ix.L2[1:5].astype(np.uint64).tobytes())
        f.write(np.uint64(ix.sa_intv).tobytes())
        f.write(np.uint64(ix.seq_len).tobytes())
        f.write(ix.sa[1:].astype(np.uint64).tobytes())
    with open(prefix + ".gnumap.pac", "wb") as f:           # reference src/bntseq.c:304-314
        f.write(ix.pac[: (ix.l_pac >> 2) + (0 if ix.l_pac & 3 == 0 else 1)].tobytes())
        if ix.l_pac % 4 == 0:
            f.write(b"\x00")
        f.write(bytes([ix.l_pac % 4]))
    with open(prefix + ".gnumap.ann", "w") as f:            # bns_dump, reference src/bntseq.c:66-82
        f.write(f"{ix.l_pac} {ix.n_seqs} 11\n")
        for name, off, ln in zip(ix.names, ix.seq_offset, ix.seq_len_arr):
            f.write(f"0 {name} (null)\n{int(off)} {int(ln)} 0\n")
    with open(prefix + ".gnumap.amb", "w") as f:            # reference src/bntseq.c:83-95
        f.write(f"{ix.l_pac} {ix.n_seqs} 0\n")


def load_index(prefix: str) -> FMIndex:
    """bwa_idx_load_from_disk (reference src/GenomeBwt.cpp:112-140): bwt_restore_bwt
    (src/bwt.c:443-462), bwt_restore_sa (:421-441), bns_restore (src/bntseq.c:98-196), pac fread."""
    raw = np.fromfile(prefix + ".gnumap.bwt", dtype=np.uint8)
    head = raw[:40].view(np.uint64)
    primary = int(head[0])
    L2 = np.zeros(5, dtype=np.uint64)
    L2[1:] = head[1:5]
    bwt = raw[40:].view(np.uint32).copy()
    seq_len = int(L2[4])
    sraw = np.fromfile(prefix + ".gnumap.sa", dtype=np.uint64)
    assert int(sraw[0]) == primary, "SA-BWT inconsistency: primary is not the same."
    sa_intv = int(sraw[5])
    assert int(sraw[6]) == seq_len, "SA-BWT inconsistency: seq_len is not the same."
    n_sa = (seq_len + sa_intv) // sa_intv
    sa = np.empty(n_sa, dtype=np.uint64)
    sa[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    sa[1:] = sraw[7:7 + n_sa - 1]
    names, offs, lens = [], [], []
    with open(prefix + ".gnumap.ann") as f:
        l_pac, n_seqs, _seed = f.readline().split()
        l_pac, n_seqs = int(l_pac), int(n_seqs)
        for _ in range(n_seqs):
            parts = f.readline().rstrip("\n").split(" ", 2)
            names.append(parts[1])
            o, ln, _ = f.readline().split()
            offs.append(int(o)); lens.append(int(ln))
    pac = np.fromfile(prefix + ".gnumap.pac", dtype=np.uint8)[: (l_pac + 3) // 4].copy()
    return FMIndex(bwt=bwt, primary=primary, L2=L2, seq_len=seq_len, sa=sa, sa_intv=sa_intv, pac=pac,
                   l_pac=l_pac, names=names, seq_offset=np.asarray(offs, dtype=np.int64),
                   seq_len_arr=np.asarray(lens, dtype=np.int32))


def index_files_exist(prefix: str) -> bool:
    return all(os.path.exists(prefix + ext) for ext in (".gnumap.bwt", ".gnumap.sa", ".gnumap.pac", ".gnumap.ann", ".gnumap.amb"))
