#!/usr/bin/env python
"""BASELINE configs[0] fixture: the reference's own example reads on the surrogate genome of SURVEY.md §8(d).

`examples/Cel_gen.fa` is missing from the reference tree (.MISSING_LARGE_BLOBS); `examples/Cel_gen.reads.aln` (ART's
truth file) holds, for each of the 24 869 reads of `examples/Cel_gen.reads.fq`, the 50 reference bases it was drawn
from and where.  The surrogate is one contig `chrI_third` of the stated length (4 973 850) carrying those true
50-mers at their offsets ('-' records: forward offset = L - pos - len, stored line = reverse complement) and seeded
random ACGT elsewhere.

Run in the build container (needs /root/reference and oracle/_ref/gnumap):

    python tests/golden/make_cfg0.py

Writes tests/golden/cfg0_surrogate.npz (planted segments, reads, and the UNMODIFIED reference binary's SAM for
`gnumap -a .9 -c 1` on that genome, sequence / quality columns stripped: they are the read's own).  The GPU box has no
/root/reference and only reads this file.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gnumap_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

EX = "/root/reference/examples"
SEED = 4973850
CODE = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
OUT = os.path.join(HERE, "cfg0_surrogate.npz")


def parse_aln(path):
    """-> (L, [(id, forward offset, forward codes uint8[])])"""
    L = None
    recs = []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith(b"@SQ"):
            L = int(ln.split(b"\t")[2])
        if ln.startswith(b">"):
            _, name, pos, strand = ln[1:].split(b"\t")
            ref = lines[i + 1].replace(b"-", b"")            # gaps of the one indel record are not genome bases
            codes = np.array([CODE[c] for c in ref.upper()], dtype=np.uint8)
            pos = int(pos)
            if strand.strip() == b"-":
                codes = (3 - codes[::-1]).astype(np.uint8)
                pos = L - pos - len(codes)
            recs.append((int(name.rsplit(b"-", 1)[1]), pos, codes))
            i += 3
            continue
        i += 1
    return L, recs


def surrogate_genome(L, seed, off, lens, flat):
    """The surrogate: seeded random ACGT with the planted segments written in file order (they agree where they overlap)."""
    g = np.random.default_rng(seed).integers(0, 4, size=L, dtype=np.uint8)
    at = 0
    for o, n in zip(off, lens):
        g[o:o + n] = flat[at:at + n]
        at += n
    return g


def load(path=OUT):
    z = np.load(path, allow_pickle=False)
    L = int(z["length"]); off = z["planted_off"]; lens = z["planted_len"]
    flat = np.unpackbits(z["planted_2bit"]).reshape(-1, 2)
    flat = (flat[:, 0] * 2 + flat[:, 1]).astype(np.uint8)[: int(lens.sum())]
    genome = surrogate_genome(L, int(z["seed"]), off, lens, flat)
    n = len(z["read_ids"]); rl = int(z["read_len"])
    seqs = z["read_seq"].reshape(n, rl); quals = z["read_qual"].reshape(n, rl)
    names = [f"chrI_third-{int(i)}" for i in z["read_ids"]]
    sam = bytes(z["ref_sam_stripped"]).decode().split("\n")
    return {"genome": genome, "names": names, "seq": seqs, "qual": quals, "sam_stripped": [s for s in sam if s],
            "matched": int(z["ref_matched"]), "not_matched": int(z["ref_not_matched"]), "total_nw": int(z["ref_total_nw"])}


def strip(line: str) -> str:
    f = line.split("\t")
    return "\t".join(f[:9] + f[11:])


def main():
    O.build()
    if not O.have_ref_binary():
        raise SystemExit("oracle/_ref/gnumap is missing")
    L, recs = parse_aln(os.path.join(EX, "Cel_gen.reads.aln"))
    off = np.array([r[1] for r in recs], dtype=np.int64)
    lens = np.array([len(r[2]) for r in recs], dtype=np.int32)
    flat = np.concatenate([r[2] for r in recs])
    genome = surrogate_genome(L, SEED, off, lens, flat)
    # overlapping true segments must agree (SURVEY: 73 854 overlapping bases agree, 0 disagree)
    at = 0; bad = 0
    for o, n in zip(off, lens):
        bad += int((genome[o:o + n] != flat[at:at + n]).sum()); at += n
    print(f"surrogate: L={L}, {len(recs)} planted segments, {bad} conflicting bases")
    assert bad == 0
    names, fq = synth.read_fastq(os.path.join(EX, "Cel_gen.reads.fq"))
    rl = len(fq[0][0])
    assert all(len(s) == rl and len(q) == rl for s, q in fq)
    ids = np.array([int(nm.rsplit("-", 1)[1]) for nm in names], dtype=np.int32)
    tmp = tempfile.mkdtemp(prefix="gmx_cfg0_")
    try:
        fa = os.path.join(tmp, "chrI_third.fa")
        synth.write_fasta(fa, [("chrI_third", genome)])
        empty = os.path.join(tmp, "empty.fq"); open(empty, "w").close()
        O.run_reference(fa, empty, os.path.join(tmp, "warm"), threads=1, mmap_threshold=1024)        # builds the index
        log = O.run_reference(fa, os.path.join(EX, "Cel_gen.reads.fq"), os.path.join(tmp, "out"), threads=1, mmap_threshold=1024)
        sam = sorted(ln.rstrip("\n") for ln in open(os.path.join(tmp, "out.sam")) if not ln.startswith("@"))
        matched = int([ln for ln in log.splitlines() if "Sequences matched" in ln][0].split(":")[1])
        not_matched = int([ln for ln in log.splitlines() if "Sequences not matched" in ln][0].split(":")[1])
        nw = [ln for ln in log.splitlines() if "Total NW" in ln]
        total_nw = int(nw[0].split("Total NW:")[1].split(",")[0]) if nw else -1
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    bits = np.zeros((len(flat) + 3) // 4 * 4, dtype=np.uint8); bits[: len(flat)] = flat
    packed = np.packbits(np.stack([bits >> 1, bits & 1], axis=1).reshape(-1))
    np.savez_compressed(OUT, length=L, seed=SEED, planted_off=off, planted_len=lens, planted_2bit=packed,
                        read_ids=ids, read_len=rl,
                        read_seq=np.frombuffer(b"".join(s for s, _ in fq), dtype=np.uint8),
                        read_qual=np.frombuffer(b"".join(q for _, q in fq), dtype=np.uint8),
                        ref_sam_stripped=np.frombuffer("\n".join(strip(s) for s in sam).encode(), dtype=np.uint8),
                        ref_matched=matched, ref_not_matched=not_matched, ref_total_nw=total_nw)
    print(f"{OUT}: {len(fq)} reads, reference: {matched} matched / {not_matched} not matched, {len(sam)} SAM records, Total NW {total_nw}, "
          f"{os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
