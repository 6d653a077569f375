"""Whole-program drop-in: the reference's OWN host code (index loader, FASTQ reader, .sgr / .gmp printers incl. the SNP
LRT calls) linked with the reference-side binding of INTEGRATION.md and libgmx.so (oracle/_ref/gnumap_gmx_demo, built
where /root/reference is present) must reproduce what the unmodified reference binary wrote for the same inputs."""
import os
import subprocess

import numpy as np
import pytest

from tests import test_oracle_golden as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "gnumap_gmx_demo")

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["normal", "snp", "bs"])
def test_reference_host_code_over_libgmx(tmp_path, mode):
    if not os.path.exists(DEMO):
        pytest.skip("oracle/_ref/gnumap_gmx_demo has not been built (needs /root/reference at build time)")
    rec = G.load_program(mode)
    fa = tmp_path / "g.fa"
    with open(fa, "w") as f:
        for name, seq in rec["contigs"]:
            f.write(f">{name}\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i:i + 70] + "\n")
    fq = tmp_path / "r.fq"
    with open(fq, "w") as f:
        for nm, s, q in rec["reads"]:
            f.write(f"@{nm}\n{s}\n+\n{q}\n")
    out = tmp_path / "out"
    env = dict(os.environ, MALLOC_MMAP_THRESHOLD_="1024")
    # first run builds the index with the reference's own bwa_index; the second is the measured one
    for _ in range(2):
        p = subprocess.run([DEMO, str(fa), str(fq), str(out), mode], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:]
    assert f"Sequences matched: {rec['matched']}" in p.stdout
    sam = [ln.rstrip("\n") for ln in open(str(out) + ".sam")]
    assert sorted(sam) == sorted(rec["sam"]), "SAM written by the reference's host code over libgmx differs"
    if mode == "normal":
        def table(lines):
            return {(c, int(pos)): float(v) for c, pos, v in (ln.split("\t") for ln in lines)}
        got = table(ln.rstrip("\n") for ln in open(str(out) + ".sgr")); want = table(rec["sgr"])
        for k in set(got) | set(want):
            a, b = got.get(k, 0.0), want.get(k, 0.0)
            assert abs(a - b) <= 1e-5 * abs(b) + 1.1e-5 or max(a, b) < 0.00102, (k, a, b)
    else:
        got = [ln.rstrip("\n").split("\t") for ln in open(str(out) + ".gmp")]
        want = [ln.split("\t") for ln in rec["gmp"]]
        gk = {(g[0], g[1]): g for g in got}; wk = {(w[0], w[1]): w for w in want}
        common = set(gk) & set(wk)
        assert len(common) >= 0.999 * len(wk) and len(gk) <= 1.001 * len(wk) + 2      # threshold-edge rows may come and go
        same_call = 0
        for k in common:
            g, w = gk[k], wk[k]
            assert np.allclose([float(x) for x in g[2:8]], [float(x) for x in w[2:8]], rtol=1e-5, atol=1.1e-5), (g, w)
            same_call += g[8:] == w[8:]
        # the SNP / methylation call column comes from the reference's own LRT code on the GPU's accumulators
        assert same_call >= 0.995 * len(common), (same_call, len(common))
    # the library's printers (gmx_format_sgr / gmx_format_gmp, rows selected on the device, LRT call
    # and text on the host) against the reference's own PrintFinal on the very same accumulators: byte for byte
    ext = ".sgr" if mode == "normal" else ".gmp"
    assert open(str(out) + ".native" + ext, "rb").read() == open(str(out) + ext, "rb").read()


@pytest.mark.parametrize("mode", ["snp", "snp_monop"])
def test_native_gmp_calls_equal_reference_printer_at_depth(tmp_path, mode):
    """15x coverage over a two-haplotype sample with planted homozygous and heterozygous SNPs: the rows, the
    likelihood-ratio calls (mono- and diploid, Y and N) and the p-values gmx_format_gmp prints must be the bytes
    GenomeBwt::PrintFinalSNP / PrintSNPCall (reference src/GenomeBwt.cpp:930-1092) print for the same accumulators."""
    if not os.path.exists(DEMO):
        pytest.skip("oracle/_ref/gnumap_gmx_demo has not been built (needs /root/reference at build time)")
    from gnumap_b200 import synth
    contigs = synth.make_genome(6000, 77, n_contigs=2)
    codes = np.concatenate([c for _, c in contigs])
    hap_a = codes.copy(); hap_b = codes.copy()
    hom = np.arange(150, 5900, 300); het = np.arange(300, 5900, 300)
    hap_a[hom] = (hap_a[hom] + 1) & 3; hap_b[hom] = hap_a[hom]
    hap_b[het] = (hap_b[het] + 2) & 3
    ra = synth.simulate_reads(hap_a, 750, 62, 78, sub_rate=0.01)
    rb = synth.simulate_reads(hap_b, 750, 62, 79, sub_rate=0.01)
    reads = {k: np.concatenate([ra[k], rb[k]]) for k in ra}
    fa = tmp_path / "g.fa"; fq = tmp_path / "r.fq"; out = tmp_path / "out"
    synth.write_fasta(str(fa), contigs); synth.write_fastq(str(fq), reads)
    env = dict(os.environ, MALLOC_MMAP_THRESHOLD_="1024")
    for _ in range(2):
        p = subprocess.run([DEMO, str(fa), str(fq), str(out), mode], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:]
    want = open(str(out) + ".gmp", "rb").read()
    got = open(str(out) + ".native.gmp", "rb").read()
    calls = [ln.split(b"\t")[-1] for ln in want.split(b"\n") if ln]
    assert sum(c.startswith(b"Y:") for c in calls) >= 15, "the sample must produce confident SNP calls"
    if mode == "snp":
        assert sum(b"/" in c for c in calls) >= 5, "the sample must produce diploid calls"
    if got != want:
        gl, wl = got.split(b"\n"), want.split(b"\n")
        diff = [(a, b) for a, b in zip(gl, wl) if a != b][:5]
        raise AssertionError(f"{len(gl)} vs {len(wl)} rows; first differences: {diff}")
