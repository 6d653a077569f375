"""Shared builders and comparators for the parity tests."""
from __future__ import annotations

import numpy as np

from gnumap_b200 import _abi, index, synth

COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def world_plain(seed=1, length=300_000, n_reads=1500, read_len=100):
    contigs = synth.make_genome(length, seed, n_contigs=2)
    codes = np.concatenate([c for _, c in contigs])
    reads = synth.simulate_reads(codes, n_reads, read_len, seed + 1, indel_rate=0.15, n_rate=0.003)
    return contigs, _abi.ReadBatch.from_arrays(reads["bases"], reads["quals"]), reads


def world_repeats(seed=5, n_reads=600, read_len=100):
    """Small genome with exact and reverse-complemented repeats and a homopolymer run: multi-position
    groups, cross-strand groups, huge SA intervals (global vote table), READ_TOO_MANY."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=60_000, dtype=np.uint8)
    unit = base[1000:3000].copy()
    parts = [base, unit, rng.integers(0, 4, size=5000, dtype=np.uint8), unit, COMP[unit[::-1]],
             np.zeros(6000, dtype=np.uint8), rng.integers(0, 4, size=4000, dtype=np.uint8), unit[:700]]
    # near-identical repeat: one substitution every 97 bases
    near = unit.copy(); near[::97] = (near[::97] + 1) & 3
    parts += [near, rng.integers(0, 4, size=3000, dtype=np.uint8)]
    codes = np.concatenate(parts).astype(np.uint8)
    contigs = [("rep1", codes[:70_000]), ("rep2", codes[70_000:])]
    reads = synth.simulate_reads(codes, n_reads, read_len, seed + 1, sub_rate=0.005, indel_rate=0.05)
    # force a good share of the reads into the repeat units and the homopolymer
    k = n_reads // 2
    starts = rng.integers(1000, 3000 - read_len, size=k)
    fwd = codes[starts[:, None] + np.arange(read_len)[None, :]]
    reads["bases"][:k] = fwd
    reads["bases"][k:k + 10] = 0          # poly-A reads
    reads["bases"][k + 10:k + 14] = 3     # poly-T reads
    return contigs, _abi.ReadBatch.from_arrays(reads["bases"], reads["quals"]), reads


def world_genome_start(seed=21, n_reads=240, read_len=100):
    """Reads whose k-mers hit the first bases of the genome, so that `sa <= offset` and the reference clamps the
    diagonal to 0 (inc/align_seq2_raw.cpp:270): reads starting at 0..6, reads with 1..12 extra leading bases in front
    of genome[0:], a short tandem repeat at the genome start (several clamped hits per k-mer), plus ordinary reads."""
    rng = np.random.default_rng(seed)
    unit = rng.integers(0, 4, size=23, dtype=np.uint8)
    head = np.concatenate([unit, unit, unit, rng.integers(0, 4, size=40, dtype=np.uint8)])
    body = rng.integers(0, 4, size=80_000, dtype=np.uint8)
    codes = np.concatenate([head, body]).astype(np.uint8)
    contigs = [("s1", codes[:50_000]), ("s2", codes[50_000:])]
    reads = synth.simulate_reads(codes, n_reads, read_len, seed + 1, sub_rate=0.005, indel_rate=0.05)
    k = 0
    for start in range(0, 7):
        reads["bases"][k] = codes[start:start + read_len]; k += 1
    for lead in range(1, 13):
        for rep in range(3):
            pre = rng.integers(0, 4, size=lead, dtype=np.uint8)
            r = np.concatenate([pre, codes[: read_len - lead]])
            if rep == 2:
                r = COMP[r[::-1]]
            reads["bases"][k] = r; k += 1
    for start in (0, 23, 46):                       # inside the tandem repeat
        reads["bases"][k] = codes[start:start + read_len]; k += 1
    return contigs, _abi.ReadBatch.from_arrays(reads["bases"], reads["quals"]), reads


def world_ragged(seed=9):
    """Variable read lengths in one batch incl. too-short, all-N and lowest-quality reads."""
    contigs = synth.make_genome(150_000, seed, n_contigs=3)
    codes = np.concatenate([c for _, c in contigs])
    rng = np.random.default_rng(seed + 1)
    seqs, quals = [], []
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    for k in range(400):
        L = int(rng.integers(8, 151))
        r = synth.simulate_reads(codes, 1, L, seed * 1000 + k, n_rate=0.01 if k % 7 == 0 else 0.0, qlo=2 if k % 11 == 0 else 15)
        s = lut[r["bases"][0]].tobytes()
        if k % 13 == 0:
            s = s.lower()
        seqs.append(s); quals.append((r["quals"][0] + 33).astype(np.uint8).tobytes())
    seqs.append(b"N" * 60); quals.append(b"I" * 60)
    seqs.append(b"ACGTACGTAC"); quals.append(b"!!!!!!!!!!")
    return contigs, _abi.ReadBatch(seqs, quals), None


def set_mode(params, mode):
    params.mode = mode
    if mode != _abi.MODE_NORMAL:
        params.gen_size = 1
    if mode == _abi.MODE_BS:
        params.align_scores[ord("c")][3] = params.align_scores[ord("a")][0]    # Driver.cpp:1266
    return params


def canon_hits(hits):
    h = np.sort(hits, order=["read", "pos", "strand"])
    return h


def partition(hits):
    """{read: frozenset of frozensets of (pos, strand)}"""
    out = {}
    for h in hits:
        out.setdefault(int(h["read"]), {}).setdefault(int(h["group"]), set()).add((int(h["pos"]), int(h["strand"])))
    return {r: frozenset(frozenset(g) for g in gs.values()) for r, gs in out.items()}


def compare_batches(got, want, check_score_fields=True, rtol_post=1e-6):
    """Raise AssertionError with a readable message on the first parity violation."""
    g, w = got["results"], want["results"]
    assert len(g) == len(w)
    for f in ("status", "n_groups", "max_align_score", "top_score", "best_score", "best_first_pos",
              "best_n_positions", "best_first_strand"):
        bad = np.nonzero(g[f] != w[f])[0]
        assert len(bad) == 0, f"{f}: {len(bad)} reads differ, first {bad[:5]}: got {g[f][bad[:5]]} want {w[f][bad[:5]]}"
    ok = w["status"] != _abi.READ_TOO_MANY
    bad = np.nonzero((g["n_candidates"] != w["n_candidates"]) & ok)[0]
    assert len(bad) == 0, f"n_candidates differ at {bad[:5]}: {g['n_candidates'][bad[:5]]} vs {w['n_candidates'][bad[:5]]}"
    assert np.allclose(g["denominator"], w["denominator"], rtol=1e-12, atol=0), "denominator"
    assert np.allclose(g["best_posterior"], w["best_posterior"], rtol=rtol_post, atol=1e-12), "posterior"
    gh, wh = canon_hits(got["hits"]), canon_hits(want["hits"])
    assert len(gh) == len(wh), f"hit count {len(gh)} vs {len(wh)}"
    for f in ("read", "pos", "strand", "score", "first_strand"):
        assert np.array_equal(gh[f], wh[f]), f"hits.{f} differ"
    assert partition(got["hits"]) == partition(want["hits"]), "group partition differs"
    if check_score_fields:
        assert np.array_equal(g["best_aligned_len"], w["best_aligned_len"]), "best_aligned_len"
        assert got["cigars"] == want["cigars"], "cigars"


def accum_close(got, want, hits, offsets, gen_size, l_pac, what="amount"):
    """Accumulator parity.  FP32 sums into one bin depend on the order of the adds (the reference's
    own `-c N` runs differ among themselves), so the bound per bin is
        1e-6 + 1e-5 * |want|  +  adds_into_bin * 2^-24 * |want|
    where the last term is the worst-case rounding of the reference's own sequential accumulation
    (one add per aligned base, reference src/NormalScoredSeq.cpp:69-73).  On ordinary genomes a bin
    sees ~10 adds and the 1e-5 term dominates; it only matters for kilo-fold repeats."""
    lens = (offsets[1:] - offsets[:-1])[hits["read"]] + 8
    cov = np.zeros(len(want) + 1, dtype=np.float64)
    lo = np.minimum(hits["pos"].astype(np.int64) // gen_size, len(want))
    hi = np.minimum((hits["pos"].astype(np.int64) + lens) // gen_size + 1, len(want))
    np.add.at(cov, lo, gen_size * 1.0); np.add.at(cov, hi, -gen_size * 1.0)
    cov = np.cumsum(cov)[: len(want)] + gen_size
    tol = 1e-6 + 1e-5 * np.abs(want) + cov * 2.0 ** -24 * np.abs(want)
    bad = np.nonzero(np.abs(got.astype(np.float64) - want) > tol)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} bins out of tolerance, first {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]} tol {tol[bad[:5]]}"
    assert abs(got.sum(dtype=np.float64) - want.sum(dtype=np.float64)) <= 1e-5 * abs(want.sum(dtype=np.float64)) + 1e-6
