"""Whole-program drop-in: the reference program ITSELF -- its option parser, FASTQ reader, worker threads, SAM writer and
.sgr / .gmp printers -- with integration/driver_gmx.patch applied to its Driver.cpp, linked with the reference-side binding
and libgmx.so (oracle/_ref/gnumap_gmx, built where /root/reference is present; it travels to the GPU box).  Run with real
option strings, it must write what the UNMODIFIED binary (oracle/_ref/gnumap) writes for the same command line."""
import os
import subprocess

import numpy as np
import pytest

from gnumap_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "gnumap")
GMX = os.path.join(ROOT, "oracle", "_ref", "gnumap_gmx")

pytestmark = pytest.mark.gpu


def _world(tmp_path, seed=61, length=300_000, n_reads=6000, read_len=100, bisulfite=0.0, qlo=15, qoff=33):
    contigs = synth.make_genome(length, seed, n_contigs=3)
    codes = np.concatenate([c for _, c in contigs])
    # a repeat (multi-position groups, X0 > 1) and its reverse complement (one group on both strands)
    codes[200_000:200_600] = codes[50_000:50_600]
    codes[260_000:260_400] = 3 - codes[90_000:90_400][::-1]
    b = np.cumsum([0] + [len(c) for _, c in contigs])
    contigs = [(n, codes[b[i]:b[i + 1]]) for i, (n, _) in enumerate(contigs)]
    reads = synth.simulate_reads(codes, n_reads, read_len, seed + 1, indel_rate=0.2, n_rate=0.002, bisulfite=bisulfite, qlo=qlo)
    k = n_reads // 10
    reads["pos"][:k] = np.random.default_rng(seed).integers(50_000, 50_600 - read_len, size=k)
    fwd = codes[reads["pos"][:k, None] + np.arange(read_len)[None, :]]
    reads["bases"][:k] = np.where((reads["strand"][:k] == 1)[:, None], 3 - fwd[:, ::-1], fwd)
    fa = str(tmp_path / "g.fa"); fq = str(tmp_path / "r.fq")
    synth.write_fasta(fa, contigs)
    if qoff != 33:
        reads = dict(reads); reads["quals"] = (reads["quals"].astype(np.int16) + (qoff - 33)).astype(np.uint8)
    synth.write_fastq(fq, reads)
    return fa, fq


def _run(binary, fa, fq, out, opts, threads, env_extra=None):
    env = dict(os.environ, MALLOC_MMAP_THRESHOLD_="1024")
    env.update(env_extra or {})
    p = subprocess.run([binary, "-g", fa, "-o", out, "-a", ".9", "-c", str(threads), *opts, fq], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=1200)
    assert p.returncode == 0, p.stdout[-3000:]
    return p.stdout


def _stat(log, key):
    return int([ln for ln in log.splitlines() if key in ln][0].split(":")[1].split(",")[0])


def _rows(path, ncol):
    t = {}
    for ln in open(path):
        f = ln.rstrip("\n").split("\t")
        t[(f[0], int(f[1]))] = ([float(x) for x in f[2:2 + ncol]], f[2 + ncol:])
    return t


CASES = {
    # name: (reference options, worker threads of the patched run, world overrides)
    "normal_c8": ([], 8, {}),
    "snp_c4": (["--snp"], 4, {}),
    "snp_monop": (["--snp", "--snp_monop", "--snp_pval=0.01"], 2, {}),
    "bs": (["-b"], 2, {"bisulfite": 0.6}),
    "b2_down_strand": (["--b2"], 2, {"bisulfite": 0.6}),
    "a_to_g": (["-d"], 2, {}),
    "max_gap_5": (["-M", "5"], 3, {}),
    "max_gap_1": (["--max_gap=1"], 2, {}),
    "gap_penalty": (["-G", "-0.5"], 2, {}),
    "bin_size_4": (["--bin_size=4"], 2, {}),
    "bin_size_1_unique": (["--bin_size=1", "-u", "1"], 2, {}),          # the reference's -u consumes one token (src/Driver.cpp:2767)
    "illumina": (["--illumina"], 2, {"qoff": 64, "qlo": 5}),
    "mer12_jump4_seeds3": (["-m", "12", "-j", "4", "-k", "3"], 2, {}),
    "max_kmer_and_matches": (["-h", "40", "-T", "3"], 2, {}),
    "up_strand_raw": (["--up_strand", "-r", "-a", "60"], 2, {}),
    "read_quality_fast": (["-q", "70", "--fast", "-k", "1"], 2, {}),      # --fast stops at the first k-mer that hits: one seed must do
    "subst_file": (["-S", "@SUBST@"], 2, {}),                             # readPWM, reference src/Driver.cpp:768-859: own score table, gADJUST = 1
}

SUBST_TABLE = "\tA\tC\tG\tT\nA\t2.5\t-1.75\t-0.5\t-1.75\nC\t-1.75\t2.5\t-1.75\t-0.5\nG\t-0.5\t-1.75\t2.5\t-1.75\nT\t-1.75\t-0.5\t-1.75\t2.5\nN\t0.25\t0.25\t0.25\t0.25\n"


@pytest.mark.parametrize("case", list(CASES))
def test_patched_reference_binary_matches_unmodified(tmp_path, case):
    if not (os.path.exists(GMX) and os.path.exists(REF)):
        pytest.skip("oracle/_ref/gnumap_gmx has not been built (needs /root/reference at build time)")
    opts, threads, world = CASES[case]
    fa, fq = _world(tmp_path, **world)
    if "@SUBST@" in opts:
        open(str(tmp_path / "subst.txt"), "w").write(SUBST_TABLE)
        opts = [str(tmp_path / "subst.txt") if o == "@SUBST@" else o for o in opts]
    empty = str(tmp_path / "empty.fq"); open(empty, "w").close()
    _run(REF, fa, empty, str(tmp_path / "warm"), opts, 1)               # builds the index; later runs start from zero pages
    ref_log = _run(REF, fa, fq, str(tmp_path / "ref"), opts, 1)
    gmx_log = _run(GMX, fa, fq, str(tmp_path / "gmx"), opts, threads, {"GMX_NATIVE_PRINT": "1", "GMX_SLICE_READS": "1500"})
    assert "GPU context(s)" in gmx_log
    for key in ("Sequences matched", "Sequences not matched"):
        assert _stat(gmx_log, key) == _stat(ref_log, key), key
    body = lambda p: sorted(ln for ln in open(p) if not ln.startswith("@PG"))
    assert body(str(tmp_path / "gmx.sam")) == body(str(tmp_path / "ref.sam")), "SAM of the patched binary differs"
    assert _stat(ref_log, "Sequences matched") > 500
    gmp = any(o in opts for o in ("--snp", "-b", "--b2", "-d"))
    ext, ncol = (".gmp", 6) if gmp else (".sgr", 1)
    got, want = _rows(str(tmp_path / "gmx") + ext, ncol), _rows(str(tmp_path / "ref") + ext, ncol)
    same_call = 0
    for k in set(got) | set(want):
        if k not in got or k not in want:
            assert (got.get(k) or want.get(k))[0][0] < 0.00103, (k, got.get(k), want.get(k))      # print-threshold edge
            continue
        assert np.allclose(got[k][0], want[k][0], rtol=1e-5, atol=1.1e-5), (k, got[k], want[k])
        same_call += got[k][1] == want[k][1]
    assert len(want) > 100 and same_call >= 0.995 * len(set(got) & set(want))
    # the library's own printers (rows selected on the device) against the reference's PrintFinal on the same accumulators
    assert open(str(tmp_path / "gmx.native") + ext, "rb").read() == open(str(tmp_path / "gmx") + ext, "rb").read()


def test_patched_binary_counts_nw_like_the_reference(tmp_path):
    """DEBUG_NW "Total NW" (reference src/Driver.cpp:1596) of a single-threaded run."""
    if not (os.path.exists(GMX) and os.path.exists(REF)):
        pytest.skip("oracle/_ref/gnumap_gmx has not been built")
    fa, fq = _world(tmp_path, n_reads=3000)
    ref_log = _run(REF, fa, fq, str(tmp_path / "ref"), [], 1)
    gmx_log = _run(GMX, fa, fq, str(tmp_path / "gmx"), [], 1)
    assert _stat(gmx_log, "Total NW") == _stat(ref_log, "Total NW") > 3000


@pytest.mark.parametrize("backend", ["peer", "nccl"])
def test_patched_binary_on_several_gpus(tmp_path, backend):
    """`gnumap -c N` drives N GPUs: worker thread t -> GPU t % N, accumulators summed inside gmx_finish."""
    import torch
    if not (os.path.exists(GMX) and os.path.exists(REF)):
        pytest.skip("oracle/_ref/gnumap_gmx has not been built")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two or more GPUs")
    n = min(torch.cuda.device_count(), 8)
    fa, fq = _world(tmp_path, n_reads=12000)
    ref_log = _run(REF, fa, fq, str(tmp_path / "ref"), ["--snp"], 1)
    gmx_log = _run(GMX, fa, fq, str(tmp_path / "gmx"), ["--snp"], n, {"GMX_COMM": backend, "GMX_SLICE_READS": "700"})
    assert f"{n} GPU context(s)" in gmx_log
    assert _stat(gmx_log, "Sequences matched") == _stat(ref_log, "Sequences matched")
    body = lambda p: sorted(ln for ln in open(p) if not ln.startswith("@PG"))
    assert body(str(tmp_path / "gmx.sam")) == body(str(tmp_path / "ref.sam"))
    got, want = _rows(str(tmp_path / "gmx.gmp"), 6), _rows(str(tmp_path / "ref.gmp"), 6)
    for k in set(got) & set(want):
        assert np.allclose(got[k][0], want[k][0], rtol=1e-5, atol=1.1e-5), (k, got[k], want[k])
    assert len(set(got) ^ set(want)) <= 0.001 * len(want) + 2
