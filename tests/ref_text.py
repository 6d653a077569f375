"""SAM and SGR emission from batch results -- the reference's writers, unchanged in format.

Mirrors, for the fields the hot path produces:
  * `ScoredSeq::get_SAM` (reference inc/ScoredSeq.h:293-404): MAPQ, CIGAR, one record per
    (pos, strand) of the best group;
  * the SAM writer `single_write_cond_wait` (reference src/Driver.cpp:2146-2217);
  * `GenomeBwt::PrintFinalSGR` (reference src/GenomeBwt.cpp:1212-1273).

TEST INFRASTRUCTURE (the product formatters are gmx_format_sam / _sgr / _gmp): they exist so that whole-program parity
(SAM body, .sgr) can be checked against the compiled reference.
"""
from __future__ import annotations

import math

import numpy as np

from gnumap_b200 import _abi

_RC = bytes.maketrans(b"acgtACGT-", b"tgcaTGCA-")


def reverse_comp(s: bytes) -> bytes:
    """reference inc/SequenceOperations.h:56-96: anything that is not acgtACGT- becomes 'n'."""
    out = bytearray()
    for c in reversed(s):
        ch = bytes([c])
        out += ch.translate(_RC) if ch in b"acgtACGT-" else b"n"
    return bytes(out)


def reverse_cigar(c: str) -> str:
    """reference inc/SequenceOperations.h:109-123 (note: digits test is 48..58)."""
    out, num = "", ""
    for ch in c:
        if 48 <= ord(ch) <= 58:
            num += ch
        else:
            out = num + ch + out
            num = ""
    return out


def cfmt(x: float) -> str:
    """C++ ostream default float formatting (== printf %g)."""
    return "%g" % x


def mapq(total_score: float) -> int:
    """reference inc/ScoredSeq.h:302-309"""
    if total_score == 1:
        q = 30
    else:
        v = 1 - total_score
        q = 30 if v <= 0 else int(_c_round(-10 * math.log(v) / math.log(10)))
    return min(q, 30)


def _c_round(x: float) -> float:
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def sam_records(index, names, batch: _abi.ReadBatch, results: np.ndarray, hits: np.ndarray, cigars, adjust: float):
    """Yield the SAM body lines of one batch, in read order (unmapped reads print nothing:
    reference src/Driver.cpp:620-629)."""
    for r in range(len(results)):
        res = results[r]
        if res["status"] != _abi.READ_MAPPED or res["best_group"] < 0:
            continue
        if not float(res["best_score"]) > float(res["top_score"]) - 0.00001:     # SAME_DIFF, reference src/Driver.cpp:695
            continue
        a, b = int(batch.offsets[r]), int(batch.offsets[r + 1])
        seq = batch.seq[a:b].tobytes()
        qual = batch.qual[a:b].tobytes() if batch.qual is not None else b"*"
        post = np.float32(res["best_posterior"])
        # POST_PROB is a float in TopReadOutput; MAPQ is computed from the double before the cast
        total = math.exp(float(res["best_score"])) / float(res["denominator"])
        q = mapq(total)
        xa = float(np.float32(res["best_score"])) * (1.0 / adjust)
        cigar = cigars[r]
        hs = hits[int(res["hit_begin"]):int(res["hit_end"])]
        hs = hs[hs["group"] == res["best_group"]]
        for h in hs:
            chrom, cpos = index.pos2chr(int(h["pos"]))
            neg = int(h["strand"]) == _abi.NEG_STRAND
            fields = [
                names[r], "16" if neg else "0", chrom, str(cpos + 1), str(q),
                reverse_cigar(cigar) if neg else cigar, "*", "0", "0",
                (reverse_comp(seq) if neg else seq).decode(), (qual[::-1] if neg else qual).decode(),
                "XA:f:" + cfmt(xa), "XP:f:" + cfmt(float(post)), "X0:i:%d" % int(res["best_n_positions"]),
            ]
            yield "\t".join(fields)


def sgr_lines(index, amount: np.ndarray, gen_size: int, min_print: float = 0.001):
    """GenomeBwt::PrintFinalSGR.  `count` runs on across sequence boundaries exactly as the
    reference's shared loop counter does."""
    bounds = list(index.seq_offset[1:]) + [index.l_pac]
    count = 0
    for i, end in enumerate(bounds):
        name = index.names[i]
        start = int(index.seq_offset[i])
        idx = np.arange(count, int(end), gen_size, dtype=np.int64)
        if len(idx):
            bins = idx // gen_size
            ok = bins < len(amount)
            vals = np.zeros(len(idx), dtype=np.float32)
            vals[ok] = amount[bins[ok]]
            sel = np.nonzero(vals.astype(np.float64) > min_print)[0]      # float against the double literal MIN_PRINT
            for k in sel:
                yield "%s\t%d\t%.5f" % (name, int(idx[k]) - start + 1, float(vals[k]))
            count = int(idx[-1]) + gen_size


def gmp_rows(index, amount: np.ndarray, planes: np.ndarray, mode: int, min_print: float = 0.001):
    """The numeric columns of the .gmp file: (chrom, 1-based pos, amount, A, C, G, T, N).

    SNP mode: GenomeBwt::PrintFinalSNP (reference src/GenomeBwt.cpp:930-1003), every position with
    amount > MIN_PRINT.  BS mode (-b, + strand): PrintFinalBisulfite (:1092-1205), positions whose
    genome base is 'c' and amount > 0.  gen_size is 1 in both modes."""
    codes = index.codes()
    if mode == _abi.MODE_SNP:
        sel = np.nonzero(amount[: index.l_pac].astype(np.float64) > min_print)[0]
    else:
        sel = np.nonzero((amount[: index.l_pac] > 0) & (codes == 1))[0]
    rid = np.searchsorted(index.seq_offset, sel, side="right") - 1
    for p, r in zip(sel, rid):
        yield (index.names[r], int(p) - int(index.seq_offset[r]) + 1, float(amount[p]),
               *[float(planes[b][p]) for b in range(5)])


def parse_gmp(path: str):
    rows = []
    with open(path) as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            rows.append((t[0], int(t[1]), float(t[2]), *[float(x) for x in t[3:8]]))
    return rows
