"""Parity at the BASELINE.json configs themselves, against the UNMODIFIED reference binary.

configs[0]  the reference's own example reads (examples/Cel_gen.reads.fq, 24 869 x 50 bp, real ART qualities) on the
            surrogate genome of SURVEY.md 8(d): the reference binary's SAM is a committed fixture
            (tests/golden/cfg0_surrogate.npz, made by tests/golden/make_cfg0.py) -- CPU test: oracle == fixture on a
            slice; GPU test: all reads through CUDA == fixture, byte for byte, and the NW count == the reference's
            DEBUG_NW "Total NW" line (reference src/Driver.cpp:1596).
configs[1..4]  the synthetic shapes at their FULL genome sizes (100 Mb / 156 Mb; ~95 / ~149 suffix-array hits per k-mer,
            the regime the benchmark runs in).  8 192 reads of each workload are mapped by `oracle/_ref/gnumap -c 1`
            (the compiled reference travels to the GPU box) and by the CUDA path; SAM bodies must be identical and
            the .sgr / .gmp numeric columns agree to rtol 1e-5 (+ the files' print precision).
"""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from gnumap_b200 import _abi, index, synth
from tests import common
from tests import ref_text
from tests.golden import make_cfg0

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "gnumap")
CACHE = os.environ.get("GMX_BENCH_CACHE", "/tmp/gnumap_b200_bench")


# ---------------------------------------------------------------------------------------------------------------
# configs[0]
# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cfg0():
    fx = make_cfg0.load()
    fx["index"] = index.build_index([("chrI_third", fx["genome"])])
    return fx


def _full_sam(fx, stripped):
    """Re-attach the sequence / quality columns (the read's own, reverse-complemented on the - strand)."""
    by_name = {nm: i for i, nm in enumerate(fx["names"])}
    out = []
    for ln in stripped:
        f = ln.split("\t")
        i = by_name[f[0]]
        seq = fx["seq"][i].tobytes(); qual = fx["qual"][i].tobytes()
        if f[1] == "16":
            seq = ref_text.reverse_comp(seq); qual = qual[::-1]
        out.append("\t".join(f[:9] + [seq.decode(), qual.decode()] + f[9:]))
    return out


def test_cfg0_oracle_matches_reference_binary(cfg0):
    """CPU: the oracle on all 24 869 example reads against the reference binary's SAM, matched count and NW count."""
    from oracle import oracle as O
    n = len(cfg0["names"])
    batch = _abi.ReadBatch([cfg0["seq"][i].tobytes() for i in range(n)], [cfg0["qual"][i].tobytes() for i in range(n)])
    p = O.default_params()
    got = O.process_batch(O.OracleIndex(cfg0["index"]), p, batch)
    r = got["results"]
    sam = sorted(ref_text.sam_records(cfg0["index"], cfg0["names"], batch, r, got["hits"], got["cigars"], p.adjust))
    assert sam == sorted(_full_sam(cfg0, cfg0["sam_stripped"]))
    assert int((r["status"] == _abi.READ_MAPPED).sum()) == cfg0["matched"]
    assert int(r["n_candidates"].sum()) == cfg0["total_nw"]                     # DEBUG_NW "Total NW", src/Driver.cpp:1596
    # the example reads hold two mapped reads whose best group scores a few ulps under the read's top NW score (one genome
    # string met on both strands): the reference prints no SAM record for them (SAME_DIFF, src/Driver.cpp:695)
    silent = (r["status"] == _abi.READ_MAPPED) & ~(r["best_score"].astype(np.float64) > r["top_score"] - 0.00001)
    assert int(silent.sum()) == 2


@pytest.mark.gpu
def test_cfg0_cuda_matches_reference_binary(cfg0):
    """GPU: all 24 869 example reads; SAM body, matched counts and the NW count of the reference binary."""
    from gnumap_b200 import api
    n = len(cfg0["names"])
    text = b"".join(b"@" + cfg0["names"][i].encode() + b"\n" + cfg0["seq"][i].tobytes() + b"\n+\n" + cfg0["qual"][i].tobytes() + b"\n" for i in range(n))
    m = api.Mapper(cfg0["index"])
    for collect in (1, 0):
        m.reset_accumulators()
        m.set_option(api.OPT_COLLECT_HITS, collect)
        names, out = m.process_fastq(text, fetch=False)
        res = out["results"]
        assert names == cfg0["names"]
        assert int((res["status"] == _abi.READ_MAPPED).sum()) == cfg0["matched"]
        assert int((res["status"] != _abi.READ_MAPPED).sum()) == cfg0["not_matched"]
        recs = api.fastq_scan_host(text)
        sam = m.format_sam(text, recs, res).decode().split("\n")
        assert sorted(s for s in sam if s) == sorted(_full_sam(cfg0, cfg0["sam_stripped"])), "SAM differs from the reference binary's"
        # DEBUG_NW counts every bin_seq alignment of PHASE A; a read that ends as READ_TOO_MANY stops early in the reference
        many = res["status"] == _abi.READ_TOO_MANY
        assert int(res["n_candidates"][~many].sum()) <= cfg0["total_nw"] <= int(res["n_candidates"].sum())
    m.close()


# ---------------------------------------------------------------------------------------------------------------
# configs[1..4] at full genome size, against the compiled reference run on the spot
# ---------------------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (genome length, genome seed, read length, reads seed, mode, reference flags)
    "cfg1_normal_100Mb_100bp": (100_000_000, 100, 100, 101, "normal", []),
    "cfg2_normal_156Mb_150bp": (156_000_000, 156, 150, 157, "normal", []),
    "cfg3_snp_156Mb_150bp": (156_000_000, 156, 150, 157, "snp", ["--snp"]),
    "cfg4_bs_100Mb_100bp": (100_000_000, 500, 100, 501, "bs", ["-b"]),
    # the same regime (95 SA hits per k-mer) through the vote paths the defaults do not take: three agreeing k-mers (the
    # byte-counter filter classes), a cap on the hits of a k-mer (half of the k-mers are skipped), a 12-mer with jump 3
    "cfg1_min_seed_hits_3": (100_000_000, 100, 100, 101, "normal", ["-k", "3"], {"min_seed_hits": 3}),
    "cfg1_max_kmer_hits_95": (100_000_000, 100, 100, 101, "normal", ["-h", "95"], {"max_kmer_hits": 95}),
    "cfg1_mer12_jump3": (100_000_000, 100, 100, 101, "normal", ["-m", "12", "-j", "3"], {"mer": 12, "jump": 3}),
}
N_SAMPLE = 8192


def _index_for(length, seed):
    """Index in the reference's on-disk format, shared with bench.py's cache."""
    os.makedirs(CACHE, exist_ok=True)
    prefix = os.path.join(CACHE, f"g{length}_s{seed}.fa")
    if index.index_files_exist(prefix):
        return index.load_index(prefix), prefix
    ix = index.build_index(synth.make_genome(length, seed), device="cuda")
    tmp = prefix + f".tmp{os.getpid()}"
    index.save_index(ix, tmp)
    for ext in (".gnumap.bwt", ".gnumap.sa", ".gnumap.pac", ".gnumap.ann", ".gnumap.amb"):
        os.replace(tmp + ext, prefix + ext)
    if not os.path.exists(prefix):
        with open(prefix, "w") as f:
            f.write(">chrS\n")
    return ix, prefix


def _run_reference(prefix, fqs, outs, flags):
    env = dict(os.environ, MALLOC_MMAP_THRESHOLD_="65536")
    ps = [subprocess.Popen([REF_BIN, "-g", prefix, "-o", o, "-a", ".9", "-c", "1", *flags, fq], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True) for fq, o in zip(fqs, outs)]
    logs = []
    for p in ps:
        out, _ = p.communicate(timeout=1500)
        assert p.returncode == 0, out[-2000:]
        logs.append(out)
    return logs


def _close_rows(got_lines, want_lines, ncol, what):
    """Numeric columns of .sgr / .gmp rows: rtol 1e-5 + the print precision (5 decimals); rows whose amount sits at the
    print threshold may come and go."""
    def table(lines):
        t = {}
        for ln in lines:
            f = ln.split("\t")
            t[(f[0], int(f[1]))] = [float(x) for x in f[2:2 + ncol]]
        return t
    g, w = table(got_lines), table(want_lines)
    for k in set(g) | set(w):
        a, b = g.get(k), w.get(k)
        if a is None or b is None:
            v = (a or b)[0]
            assert v < 0.00103, f"{what}: row {k} only on one side with amount {v}"
            continue
        assert np.allclose(a, b, rtol=1e-5, atol=1.1e-5), f"{what}: row {k}: {a} vs {b}"
    return len(w)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_full_size_config_matches_reference_binary(cfg):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gnumap has not been built (needs /root/reference at build time)")
    from gnumap_b200 import api
    length, gseed, L, rseed, mode, flags = CONFIGS[cfg][:6]
    edits = CONFIGS[cfg][6] if len(CONFIGS[cfg]) > 6 else {}
    ix, prefix = _index_for(length, gseed)
    codes = ix.codes()
    reads = synth.simulate_reads(codes, N_SAMPLE, L, rseed, sub_rate=0.01, qlo=15, qhi=40, bisulfite=0.95 if mode == "bs" else 0.0)
    # one `gnumap -c 1` process per sub-sample; each holds the index (~2.5 B / base) and its own accumulators
    # (4 B / bin in Normal mode, 6 x 4 B / base in SNP and bisulfite mode): stay inside the box's free memory
    import psutil
    per_proc = length * (2.5 + (0.5 if mode == "normal" else 24.0)) + 0.5e9
    procs = max(1, min(os.cpu_count() or 1, 16, int(0.6 * psutil.virtual_memory().available / per_proc)))
    per = N_SAMPLE // procs
    tmp = tempfile.mkdtemp(prefix="gmx_cfgtest_", dir=CACHE)
    try:
        fqs, outs = [], []
        for p in range(procs):
            fq = os.path.join(tmp, f"s{p}.fq")
            synth.write_fastq(fq, {k: v[p * per:(p + 1) * per] for k, v in reads.items()}, prefix=f"r{p}_")
            fqs.append(fq); outs.append(os.path.join(tmp, f"o{p}"))
        logs = _run_reference(prefix, fqs, outs, flags)
        params = common.set_mode(api.default_params(), {"normal": _abi.MODE_NORMAL, "bs": _abi.MODE_BS, "snp": _abi.MODE_SNP}[mode])
        for k, v in edits.items():
            assert hasattr(params, k), k
            setattr(params, k, v)
        m = api.Mapper(ix, params)
        total_nw_ref = total_nw = 0
        n_sam = n_rows = 0
        for p in range(procs):
            text = open(fqs[p], "rb").read()
            m.reset_accumulators()
            for collect in ((1, 0) if p == 0 else (p & 1,)):          # both download paths
                m.reset_accumulators()
                m.set_option(api.OPT_COLLECT_HITS, collect)
                names, out = m.process_fastq(text, fetch=False)
                res = out["results"]
                recs = api.fastq_scan_host(text)
                sam = sorted(s for s in m.format_sam(text, recs, res).decode().split("\n") if s)
                want = sorted(ln.rstrip("\n") for ln in open(outs[p] + ".sam") if not ln.startswith("@"))
                assert sam == want, f"{cfg}: SAM of sub-sample {p} differs ({len(sam)} vs {len(want)} records)"
            n_sam += len(want)
            matched = int([ln for ln in logs[p].splitlines() if "Sequences matched" in ln][0].split(":")[1])
            assert int((res["status"] == _abi.READ_MAPPED).sum()) == matched
            nw = [ln for ln in logs[p].splitlines() if "Total NW" in ln]
            if nw and not (res["status"] == _abi.READ_TOO_MANY).any():
                total_nw_ref += int(nw[0].split("Total NW:")[1].split(",")[0]); total_nw += int(res["n_candidates"].sum())
            if mode == "normal":
                n_rows += _close_rows([s for s in m.format_sgr().decode().split("\n") if s],
                                      [ln.rstrip("\n") for ln in open(outs[p] + ".sgr")], 1, f"{cfg} sgr {p}")
            else:
                target = 1 if mode == "bs" else -1
                n_rows += _close_rows([s for s in m.format_gmp(target_base=target).decode().split("\n") if s],
                                      [ln.rstrip("\n") for ln in open(outs[p] + ".gmp")], 6, f"{cfg} gmp {p}")
        assert total_nw == total_nw_ref, f"{cfg}: NW alignments {total_nw} vs the reference's DEBUG_NW total {total_nw_ref}"
        assert n_sam > (0.05 if mode == "bs" else 0.9) * N_SAMPLE and n_rows > 0
        m.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
