#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/, built by `make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (committed; the GPU box has no /root/reference and only reads these files):

  ref_functions.npz     function-level vectors produced by oracle/_ref/libref_probe.so, i.e. by the reference's own
                        bin_seq / GenomeBwt objects: self score, banded NW score (exact float bits), traceback strings
                        and CIGARs, pair-HMM posteriors, FM-index intervals, SA coordinates, GetString windows,
                        ScoredSeq::score() accumulator footprints for the three modes;
  ref_index.npz         the index files the reference's bwa_index wrote for a small two-contig FASTA (bytes), next
                        to the FASTA's bases -- pins gnumap_b200/index.py byte for byte;
  ref_fastq.json        FASTQ texts (well-formed and malformed) with what the reference's SeqReader returns for them
                        (name, sequence, quality string per Read) -- pins gmx_fastq_scan_host incl. its recovery paths;
  ref_program_<mode>.json.gz   whole-program runs of oracle/_ref/gnumap (`-c 1`, zero-initialised accumulators):
                        the reads, the sorted SAM body and the .sgr / .gmp rows.

Everything is seeded; re-running reproduces the files bit for bit.
"""
from __future__ import annotations

import gzip
import json
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

BASES = b"acgt"


def random_cases(rng, n_cases):
    """(seq, qual, strand, window) tuples: windows are the read's source with substitutions / indels / an 'n'."""
    cases = []
    for k in range(n_cases):
        n = int(rng.integers(20, 151))
        src = rng.integers(0, 4, size=n + 8)
        seq = bytearray(b"ACGT"[c] for c in src[:n])
        if k % 5 == 0:
            seq[int(rng.integers(0, n))] = ord("N")
        if k % 9 == 0:
            seq = bytearray(bytes(seq).lower())
        qual = bytes(int(q) + 33 for q in rng.integers(2, 41, size=n))
        win = list(src[:n])
        for _ in range(int(rng.integers(0, 4))):
            win[int(rng.integers(0, n))] = int(rng.integers(0, 4))
        kind = k % 4
        if kind == 1:                                  # deletion in the window (read has an extra base)
            p = int(rng.integers(5, n - 5)); del win[p]; win.append(int(src[n]))
        elif kind == 2:                                # insertion in the window
            p = int(rng.integers(5, n - 5)); win.insert(p, int(rng.integers(0, 4))); win = win[:n]
        elif kind == 3 and n > 40:                     # two-base gap
            p = int(rng.integers(5, n - 8)); del win[p:p + 2]; win += [int(src[n]), int(src[n + 1])]
        w = bytearray(BASES[c] for c in win)
        if k % 7 == 0:
            w[int(rng.integers(0, n))] = ord("n")
        strand = int(rng.integers(0, 2))
        cases.append((bytes(seq), qual, strand, bytes(w)))
    return cases


def make_functions(tmp):
    rp = O.RefProbe()
    rng = np.random.default_rng(20260101)
    out = {}
    cases = random_cases(rng, 48)
    seqs, quals, strands, wins = zip(*cases)
    out["seq"] = np.array(seqs, dtype=object); out["qual"] = np.array(quals, dtype=object)
    out["strand"] = np.array(strands, dtype=np.uint8); out["window"] = np.array(wins, dtype=object)
    for mode, tag in ((0, "normal"), (1, "bs")):
        rp.set_mode(mode)
        a, p, sc = rp.tables()
        out[f"align_scores_{tag}"] = a; out[f"phmm_scores_{tag}"] = p; out[f"scalars_{tag}"] = sc
        selfs, scores, aligned, cigars, hmms = [], [], [], [], []
        for seq, qual, strand, win in cases:
            pwm = O.fastq_pwm(seq, qual)
            selfs.append(rp.self_score(pwm, seq))
            if strand:
                pwm = O.revcomp_pwm(pwm)
            cons = O.max_char_consensus(pwm)
            scores.append(rp.nw_score(pwm, win))
            al, cg = rp.nw_traceback(pwm, cons, win)
            aligned.append(al); cigars.append(cg)
            if mode == 0:
                hmms.append(rp.pair_hmm(pwm, cons, win))
        out[f"self_{tag}"] = np.array(selfs, dtype=np.float32); out[f"score_{tag}"] = np.array(scores, dtype=np.float32)
        out[f"aligned_{tag}"] = np.array(aligned, dtype=object); out[f"cigar_{tag}"] = np.array(cigars, dtype=object)
        if mode == 0:
            out["phmm_flat"] = np.concatenate([h.reshape(-1) for h in hmms]).astype(np.float32)
    rp.set_mode(0)
    # PWM rows of FASTQ (base, quality) pairs as the reference's reader builds them are covered by the whole-program runs.

    # genome-backed probes on a small two-contig genome
    contigs = synth.make_genome(6000, 11, n_contigs=2)
    fa = os.path.join(tmp, "probe.fa")
    synth.write_fasta(fa, contigs)
    rp.load_genome(fa)
    codes = np.concatenate([c for _, c in contigs])
    kmers = []
    for _ in range(200):
        p = int(rng.integers(0, len(codes) - 12)); ln = int(rng.integers(4, 13))
        k = bytearray(BASES[c] for c in codes[p:p + ln])
        if rng.random() < 0.3:
            k[int(rng.integers(0, ln))] = b"acgtn"[int(rng.integers(0, 5))]
        if rng.random() < 0.2:
            k = bytearray(bytes(k).upper())
        kmers.append(bytes(k))
    kmers += [b"nnnnnnnnnn", b"a", b"acgtacgtacgtacgtacgtacgt"]
    out["kmer"] = np.array(kmers, dtype=object)
    out["sa_int"] = np.array([rp.get_sa_int(k) for k in kmers], dtype=np.uint64)
    out["sa_coord"] = np.array([rp.get_sa_coord(k) for k in range(1, len(codes) + 1)], dtype=np.uint64)
    begins = [0, 1, 2999 - 50, 3000 - 40, 3000 - 39, 3000, 5999 - 40, 6000 - 40, 6000 - 39, 5999]
    out["string_begin"] = np.array(begins, dtype=np.uint64)
    out["string_40"] = np.array([rp.get_string(b, 40) for b in begins], dtype=object)

    # ScoredSeq::score footprints: one group with two positions on opposite strands, per mode
    seq, qual, _, _ = cases[1]
    seq = seq[:50]; qual = qual[:50]
    pwm = O.fastq_pwm(seq, qual)
    gen_string = rp.get_string(1000, 50)
    for kind, tag in ((0, "normal"), (1, "bs"), (2, "snp")):
        rp.set_mode(kind)
        rp.load_genome(fa)                               # re-allocates the accumulators for the mode's bin size
        gs = 1 if kind else 8
        amount, planes = rp.score_once(kind, pwm, gen_string, 30.5, [(1000, 0), (2100, 1)], 3.0 * np.exp(30.5), 6000 // gs)
        out[f"score_once_amount_{tag}"] = amount
        if planes is not None:
            out[f"score_once_planes_{tag}"] = planes
    out["score_once_seq"] = np.array([seq, qual, gen_string], dtype=object)
    rp.set_mode(0)
    np.savez_compressed(os.path.join(HERE, "ref_functions.npz"), **out)
    print("ref_functions.npz:", len(cases), "alignment cases,", len(kmers), "k-mers")


def make_index(tmp):
    contigs = synth.make_genome(5003, 21, n_contigs=3)         # odd length: exercises the pac tail byte
    fa = os.path.join(tmp, "ix.fa")
    synth.write_fasta(fa, contigs)
    fq = os.path.join(tmp, "none.fq")
    open(fq, "w").close()
    O.run_reference(fa, fq, os.path.join(tmp, "ixout"), threads=1, mmap_threshold=1024)
    out = {"lens": np.array([len(c) for _, c in contigs], dtype=np.int64), "codes": np.concatenate([c for _, c in contigs])}
    for ext in ("bwt", "sa", "pac", "ann", "amb"):
        out[ext] = np.fromfile(fa + ".gnumap." + ext, dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_index.npz"), **out)
    print("ref_index.npz written")


def make_program(tmp):
    for mode, extra, seed in (("normal", [], 31), ("snp", ["--snp"], 32), ("bs", ["-b"], 33)):
        d = os.path.join(tmp, mode)
        os.makedirs(d)
        contigs = synth.make_genome(24000, seed, n_contigs=2)
        codes = np.concatenate([c for _, c in contigs])
        # plant a repeat so that multi-position groups and X0 > 1 occur
        codes[15000:15400] = codes[3000:3400]
        contigs = [("chrS1", codes[:12000]), ("chrS2", codes[12000:])]
        fa = os.path.join(d, "g.fa"); fq = os.path.join(d, "r.fq")
        synth.write_fasta(fa, contigs)
        reads = synth.simulate_reads(codes, 400, 62, seed + 100, indel_rate=0.15, n_rate=0.003, bisulfite=0.6 if mode == "bs" else 0.0)
        reads["pos"][:40] = np.arange(3000, 3400 - 62, 8)[:40]            # reads inside the repeat
        fwd = codes[reads["pos"][:40, None] + np.arange(62)[None, :]]
        reads["bases"][:40] = np.where((reads["strand"][:40] == 1)[:, None], (3 - fwd[:, ::-1]), fwd)
        synth.write_fastq(fq, reads)
        # first run only builds the index: a process that has run bwa_index hands the accumulators recycled
        # (non-zero) heap memory whatever the malloc threshold (reference src/GenomeBwt.cpp:323 never zeroes them)
        empty = os.path.join(d, "empty.fq")
        open(empty, "w").close()
        O.run_reference(fa, empty, os.path.join(d, "warm"), threads=1, extra=extra, mmap_threshold=1024)
        log = O.run_reference(fa, fq, os.path.join(d, "out"), threads=1, extra=extra, mmap_threshold=1024)
        sam = [ln.rstrip("\n") for ln in open(os.path.join(d, "out.sam")) if not ln.startswith("@")]
        rec = {"mode": mode, "extra": extra, "genome_seed": seed,
               "contigs": [[n, "".join("ACGT"[c] for c in cs)] for n, cs in contigs],
               "reads": [[nm, s.decode(), q.decode()] for nm, (s, q) in zip(*synth.read_fastq(fq))],
               "sam": sam,
               "matched": int([ln for ln in log.splitlines() if "Sequences matched" in ln][0].split(":")[1])}
        if mode == "normal":
            rec["sgr"] = [ln.rstrip("\n") for ln in open(os.path.join(d, "out.sgr"))]
        else:
            rec["gmp"] = [ln.rstrip("\n") for ln in open(os.path.join(d, "out.gmp"))]
        with gzip.GzipFile(os.path.join(HERE, f"ref_program_{mode}.json.gz"), "wb", mtime=0) as f:
            f.write(json.dumps(rec).encode())
        print(f"ref_program_{mode}.json.gz: {len(sam)} SAM records, matched {rec['matched']}")


def make_program_snpdepth(tmp):
    """Whole-program `--snp` run at 15x coverage over a two-haplotype sample with planted homozygous and heterozygous
    SNPs, so that the .gmp carries confident mono- and diploid calls (the 1.7x fixtures above carry none)."""
    d = os.path.join(tmp, "snpdepth")
    os.makedirs(d)
    contigs = synth.make_genome(6000, 77, n_contigs=2)
    codes = np.concatenate([c for _, c in contigs])
    hap_a = codes.copy(); hap_b = codes.copy()
    hom = np.arange(150, 5900, 300); het = np.arange(300, 5900, 300)
    hap_a[hom] = (hap_a[hom] + 1) & 3; hap_b[hom] = hap_a[hom]
    hap_b[het] = (hap_b[het] + 2) & 3
    ra = synth.simulate_reads(hap_a, 750, 62, 78, sub_rate=0.01)
    rb = synth.simulate_reads(hap_b, 750, 62, 79, sub_rate=0.01)
    reads = {k: np.concatenate([ra[k], rb[k]]) for k in ra}
    fa = os.path.join(d, "g.fa"); fq = os.path.join(d, "r.fq")
    synth.write_fasta(fa, contigs); synth.write_fastq(fq, reads)
    empty = os.path.join(d, "empty.fq")
    open(empty, "w").close()
    O.run_reference(fa, empty, os.path.join(d, "warm"), threads=1, extra=["--snp"], mmap_threshold=1024)
    log = O.run_reference(fa, fq, os.path.join(d, "out"), threads=1, extra=["--snp"], mmap_threshold=1024)
    rec = {"mode": "snp", "extra": ["--snp"], "genome_seed": 77,
           "contigs": [[n, "".join("ACGT"[c] for c in cs)] for n, cs in contigs],
           "reads": [[nm, s.decode(), q.decode()] for nm, (s, q) in zip(*synth.read_fastq(fq))],
           "sam": [ln.rstrip("\n") for ln in open(os.path.join(d, "out.sam")) if not ln.startswith("@")],
           "matched": int([ln for ln in log.splitlines() if "Sequences matched" in ln][0].split(":")[1]),
           "gmp": [ln.rstrip("\n") for ln in open(os.path.join(d, "out.gmp"))]}
    with gzip.GzipFile(os.path.join(HERE, "ref_program_snpdepth.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(rec).encode())
    calls = [ln.split("\t")[-1][:2] for ln in rec["gmp"]]
    print("ref_program_snpdepth.json.gz:", len(rec["gmp"]), "rows,", sum(c == "Y:" for c in calls), "Y calls,", sum("/" in ln.split("\t")[-1] for ln in rec["gmp"]), "diploid")


def make_snp_calls(tmp):
    """Call columns of GenomeBwt::PrintSNPCall (reference src/GenomeBwt.cpp:1011-1092 over is_snp / LRT / dipLRT
    :739-898) for random read-count vectors, through the unmodified reference objects LINKED WITH THE REFERENCE'S OWN
    GSL 1.9 (built once from /root/reference/lib/gsl-1.9.tar.gz; `make -C oracle ref OUT=<dir> GSL=<prefix>` and
    GMX_REF_DIR=<dir> for this script).  Count vectors for which GSL's error handler fires -- the unmodified program
    aborts there, e.g. gsl_cdf_chisq_P(inf, 1) after a likelihood ratio underflowed -- are left out."""
    import ctypes as C
    R = O.RefProbe()
    gsl_errors = C.CDLL(O.REF_PROBE).refp_gsl_errors
    real_gsl = gsl_errors() >= 0
    if not real_gsl:
        raise SystemExit("ref_snp_calls.json.gz is generated against the real GSL: build it and set GMX_REF_DIR (see the docstring)")
    R.set_mode(2)
    contigs = synth.make_genome(2000, 5, n_contigs=1)
    fa = os.path.join(tmp, "lrt.fa")
    synth.write_fasta(fa, contigs)
    R.load_genome(fa)
    codes = contigs[0][1]
    rng = np.random.default_rng(2024)
    cases = []
    for t in range(3000):
        cov = float(rng.choice([1, 2, 3, 5, 8, 15, 30, 60, 200, 1000]))
        c = (rng.dirichlet(np.ones(5) * rng.choice([0.05, 0.3, 1.0])) * cov).astype(np.float32)
        if t % 5 == 1:
            c = np.round(c).astype(np.float32)
        if t % 5 == 2:
            c = np.zeros(5, np.float32); c[rng.integers(0, 5)] = cov; c[rng.integers(0, 5)] += np.float32(cov * rng.choice([0.0, 0.3, 0.34, 0.5, 1.0]))
        if t % 150 == 3:
            c = np.zeros(5, np.float32)
        pos = int(rng.integers(0, 2000))
        monop = bool(t % 3 == 0)
        pval = float(rng.choice([0.001, 0.05]))
        call = R.snp_call(pos, c, monop, pval).decode()
        if gsl_errors() > 0:
            continue                                   # the reference aborts on this input
        cases.append({"counts": [float(x) for x in c], "base": int(codes[pos]), "monop": monop, "pval": pval, "call": call})
    R.set_mode(0)
    with gzip.GzipFile(os.path.join(HERE, "ref_snp_calls.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(cases).encode())
    kinds = {}
    for c in cases:
        kinds[c["call"][:3]] = kinds.get(c["call"][:3], 0) + 1
    print("ref_snp_calls.json.gz:", kinds)


def fastq_cases():
    """FASTQ texts that exercise SeqReader::get_more_fastq incl. its recovery paths."""
    rng = np.random.default_rng(77)
    def rec(name, n, qn=None, plus=b"+"):
        seq = bytes(b"ACGTN"[int(c)] for c in rng.integers(0, 5, size=n))
        q = bytes(int(x) + 33 for x in rng.integers(0, 41, size=n if qn is None else qn))
        return b"@" + name + b"\n" + seq + b"\n" + plus + b"\n" + q + b"\n"
    good = b"".join(rec(b"g%d" % i, int(rng.integers(1, 120))) for i in range(40))
    cases = {
        "well_formed": good,
        "no_final_newline": good[:-1],
        "plus_repeats_name": b"".join(rec(b"p%d" % i, 30, plus=b"+p%d" % i) for i in range(5)),
        "blank_lines_between": rec(b"a", 10) + b"\n\n" + rec(b"b", 12) + b"\n" + rec(b"c", 9),
        "quality_longer": rec(b"a", 10, qn=15) + rec(b"b", 8),
        "quality_shorter_then_recover": rec(b"a", 10, qn=6) + rec(b"b", 8) + rec(b"c", 7),
        "garbage_line": rec(b"a", 10) + b"GARBAGE\n" + rec(b"b", 8) + rec(b"c", 5),
        "missing_plus": b"@a\nACGT\nIIII\n" + rec(b"b", 8) + rec(b"c", 6),
        "truncated_last_record": rec(b"a", 10) + b"@b\nACGT\n+\n",
        "crlf": rec(b"a", 10).replace(b"\n", b"\r\n") + rec(b"b", 6).replace(b"\n", b"\r\n"),
        "empty_sequence": b"@e\n\n+\n\n" + rec(b"b", 6),
        "lowercase_and_other_letters": b"@x\nacgtRYKM\n+\nIIIIIIII\n",
        "bad_quality_char": b"@x\nACGT\n+\nII I\n",
        "empty_file": b"",
        # a getline on a stream that already hit end-of-file leaves its string untouched: the reference then emits
        # the truncated last record with the previous record's lines
        "truncated_after_plus_no_newline": rec(b"a", 30) + rec(b"b", 12)[: -14],
        "stale_lines_at_eof": b"@r2\nnntAc\n+\n@extra\n@#-2+?/B\n",
    }
    return cases


def make_fastq(tmp):
    import ctypes as C
    L = C.CDLL(O.REF_PROBE)
    out = {}
    for name, text in fastq_cases().items():
        fn = os.path.join(tmp, name + ".fq")
        with open(fn, "wb") as f:
            f.write(text)
        buf = C.create_string_buffer(1 << 20)
        n = L.refp_read_fastq(fn.encode(), buf, 1 << 20)
        reads = [ln.split("\t") for ln in buf.value.decode("latin-1").split("\n")[:-1]] if n > 0 else []
        out[name] = {"text": text.decode("latin-1"), "n": n, "reads": reads}
    with open(os.path.join(HERE, "ref_fastq.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("ref_fastq.json:", {k: v["n"] for k, v in out.items()})


def main():
    O.build()
    if not (O.have_ref_binary() and os.path.exists(O.REF_PROBE)):
        raise SystemExit("oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists")
    tmp = tempfile.mkdtemp(prefix="gmx_golden_")
    try:
        make_fastq(tmp)
        make_functions(tmp)
        make_index(tmp)
        make_program(tmp)
        make_program_snpdepth(tmp)
        make_snp_calls(tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
