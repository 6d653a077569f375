"""Synthetic genomes and simulated reads of the shapes named in BASELINE.json / SURVEY.md §8(d).

Genomes are i.i.d. uniform ACGT (no N); reads are uniform-start, 50/50 strand, per-base
substitution rate `sub_rate`, Phred qualities ~ U{qlo..qhi} (+33).  Everything is a pure function
of the numpy seed so that tests, the oracle, the reference binary and the CUDA path all see the
same bytes.
"""
from __future__ import annotations

import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


def make_genome(length: int, seed: int, n_contigs: int = 1) -> list[tuple[str, np.ndarray]]:
    """Return [(name, codes uint8[len] in 0..3)], contig lengths splitting `length` evenly."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=length, dtype=np.uint8)
    if n_contigs == 1:
        return [("chrS", codes)]
    bounds = np.linspace(0, length, n_contigs + 1).astype(np.int64)
    return [(f"chrS{i + 1}", codes[bounds[i]:bounds[i + 1]]) for i in range(n_contigs)]


def write_fasta(path: str, contigs: list[tuple[str, np.ndarray]], width: int = 70) -> None:
    with open(path, "wb") as f:
        for name, codes in contigs:
            f.write(b">" + name.encode() + b"\n")
            seq = _BASES[codes]
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((n // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


def simulate_reads(genome_codes: np.ndarray, n_reads: int, read_len: int, seed: int,
                   sub_rate: float = 0.01, qlo: int = 15, qhi: int = 40,
                   n_rate: float = 0.0, bisulfite: float = 0.0, indel_rate: float = 0.0):
    """Simulate reads from a concatenated genome (codes 0..3).

    Returns dict(bases uint8[n_reads, read_len] codes 0..4 (4 = N),
                 quals uint8[n_reads, read_len] Phred 0..93,
                 pos int64[n_reads] 0-based forward start, strand uint8[n_reads]).
    `bisulfite` = fraction of C's converted to T on + strand reads (G->A on - strand reads).
    `indel_rate` = per-read probability of one 1-base deletion (read keeps length by extending).
    """
    rng = np.random.default_rng(seed)
    L = len(genome_codes)
    pos = rng.integers(0, L - read_len - 2, size=n_reads, dtype=np.int64)
    strand = rng.integers(0, 2, size=n_reads, dtype=np.uint8)
    idx = pos[:, None] + np.arange(read_len, dtype=np.int64)[None, :]
    if indel_rate > 0:
        has = rng.random(n_reads) < indel_rate
        at = rng.integers(read_len // 4, 3 * read_len // 4, size=n_reads)
        shift = (np.arange(read_len)[None, :] >= at[:, None]) & has[:, None]
        idx = idx + shift
    fwd = genome_codes[idx]
    if bisulfite > 0:
        conv = rng.random(fwd.shape) < bisulfite
        plus = (strand == 0)[:, None]
        fwd = np.where(conv & plus & (fwd == 1), 3, fwd)      # C->T on + reads
        fwd = np.where(conv & ~plus & (fwd == 2), 0, fwd)     # G->A seen on - reads
        fwd = fwd.astype(np.uint8)
    sub = rng.random(fwd.shape) < sub_rate
    delta = rng.integers(1, 4, size=fwd.shape, dtype=np.uint8)
    fwd = np.where(sub, (fwd + delta) & 3, fwd).astype(np.uint8)
    rc = _COMP[fwd[:, ::-1]]
    bases = np.where((strand == 1)[:, None], rc, fwd).astype(np.uint8)
    if n_rate > 0:
        isn = rng.random(bases.shape) < n_rate
        bases = np.where(isn, 4, bases).astype(np.uint8)
    quals = rng.integers(qlo, qhi + 1, size=bases.shape, dtype=np.uint8)
    return {"bases": bases, "quals": quals, "pos": pos, "strand": strand}


def write_fastq(path: str, reads: dict, prefix: str = "r") -> None:
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    bases, quals = reads["bases"], reads["quals"]
    with open(path, "wb") as f:
        for k in range(bases.shape[0]):
            name = f"@{prefix}{k}_{int(reads['pos'][k]) + 1}_{'+-'[int(reads['strand'][k])]}\n".encode()
            f.write(name + lut[bases[k]].tobytes() + b"\n+\n" + (quals[k] + 33).astype(np.uint8).tobytes() + b"\n")


def read_fastq(path: str):
    """Minimal FASTQ reader -> (names, list of (seq bytes, qual bytes))."""
    names, recs = [], []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    while i + 3 < len(lines) + 1 and i < len(lines):
        if not lines[i]:
            i += 1
            continue
        names.append(lines[i][1:].decode())
        recs.append((lines[i + 1], lines[i + 3]))
        i += 4
    return names, recs


def simulate_reads_torch(codes_t, n_reads: int, read_len: int, seed: int, sub_rate: float = 0.01, qlo: int = 15, qhi: int = 40,
                         bisulfite: float = 0.0, block_reads: int = 131072, first_block: int = 0):
    """The same read model on the GPU, for read sets too large for the numpy generator (10 M x 150 bp): reads are made in
    blocks of `block_reads`, block b from torch seed `seed + b`, so any rank can make any part of the set on its own.
    `codes_t`: uint8 tensor (codes 0..3) of the concatenated genome on the target device.  Returns dict(seq uint8[n, L] ASCII
    ACGT, qual uint8[n, L] ASCII Phred+33, pos int64[n], strand uint8[n]) as device tensors."""
    import torch
    dev = codes_t.device
    L = codes_t.numel()
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    comp = torch.tensor([3, 2, 1, 0], dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev, dtype=torch.int64)
    outs = {"seq": [], "qual": [], "pos": [], "strand": []}
    b = first_block
    done = 0
    while done < n_reads:
        n = min(block_reads, n_reads - done)
        g = torch.Generator(device=dev)
        g.manual_seed(seed + b)
        pos = torch.randint(0, L - read_len - 2, (n,), generator=g, device=dev, dtype=torch.int64)
        strand = torch.randint(0, 2, (n,), generator=g, device=dev, dtype=torch.uint8)
        fwd = codes_t[pos[:, None] + ar[None, :]]
        if bisulfite > 0:
            conv = torch.rand(fwd.shape, generator=g, device=dev) < bisulfite
            plus = (strand == 0)[:, None]
            fwd = torch.where(conv & plus & (fwd == 1), torch.full_like(fwd, 3), fwd)
            fwd = torch.where(conv & ~plus & (fwd == 2), torch.zeros_like(fwd), fwd)
        sub = torch.rand(fwd.shape, generator=g, device=dev) < sub_rate
        delta = torch.randint(1, 4, fwd.shape, generator=g, device=dev, dtype=torch.uint8)
        fwd = torch.where(sub, (fwd + delta) & 3, fwd)
        rc = comp[fwd.flip(1).long()]
        bases = torch.where((strand == 1)[:, None], rc, fwd)
        quals = torch.randint(qlo, qhi + 1, fwd.shape, generator=g, device=dev, dtype=torch.uint8) + 33
        outs["seq"].append(lut[bases.long()]); outs["qual"].append(quals); outs["pos"].append(pos); outs["strand"].append(strand)
        done += n
        b += 1
    return {k: torch.cat(v) for k, v in outs.items()}
