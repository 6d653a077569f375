"""CPU suite, part 4: host-side logic of the package (index builder, batch packing, formatters)."""
import numpy as np

from gnumap_b200 import _abi, index, synth
from tests import ref_text as output
from oracle import oracle as O


def naive_sa(codes):
    s = bytes(codes.tolist())
    return np.array(sorted(range(len(s)), key=lambda i: s[i:]), dtype=np.int64)


def test_suffix_array_matches_naive_sort():
    rng = np.random.default_rng(0)
    for n in (1, 2, 17, 300, 2500):
        codes = rng.integers(0, 4, size=n, dtype=np.uint8)
        assert np.array_equal(index.suffix_array(codes), naive_sa(codes))
    rep = np.tile(np.array([0, 1, 0, 1, 2], dtype=np.uint8), 200)          # long repeats: many doubling rounds
    assert np.array_equal(index.suffix_array(rep), naive_sa(rep))
    assert np.array_equal(index.suffix_array(np.zeros(500, np.uint8)), np.arange(499, -1, -1))


def test_occ_and_backward_search_against_brute_force():
    contigs = synth.make_genome(3000, 5, n_contigs=2)
    ix = index.build_index(contigs)
    oix = O.OracleIndex(ix)
    codes = ix.codes()
    text = bytes(codes.tolist())
    rng = np.random.default_rng(1)
    for _ in range(100):
        ln = int(rng.integers(1, 9))
        p = int(rng.integers(0, len(codes) - ln))
        kmer = codes[p:p + ln]
        k, l = oix.get_sa_int(bytes(b"acgt"[c] for c in kmer))
        want = [i for i in range(len(codes) - ln + 1) if text[i:i + ln] == bytes(kmer.tolist())]
        got = sorted(oix.bwt_sa(r) for r in range(k, l + 1))
        assert got == want
    assert oix.get_sa_int(b"acgn") == (0, 0)


def test_pack_and_codes_round_trip():
    rng = np.random.default_rng(2)
    for n in (1, 4, 5, 1023):
        codes = rng.integers(0, 4, size=n, dtype=np.uint8)
        ix = index.build_index([("x", codes)])
        assert np.array_equal(ix.codes(), codes)
        assert ix.l_pac == n and ix.seq_len == n and int(ix.L2[4]) == n


def test_read_batch_layout_and_slice():
    b = _abi.ReadBatch([b"ACGT", b"", b"GGNNA"], [b"IIII", b"", b"#####"])
    assert b.offsets.tolist() == [0, 4, 4, 9] and b.n_reads == 3
    s = b.slice(1, 3)
    assert s.offsets.tolist() == [0, 0, 5] and s.seq.tobytes() == b"GGNNA"
    fa = _abi.ReadBatch.from_arrays(np.array([[0, 1, 2, 3, 4]], np.uint8), np.array([[0, 10, 20, 30, 40]], np.uint8))
    assert fa.seq.tobytes() == b"ACGTN" and fa.qual.tobytes() == bytes([33, 43, 53, 63, 73])


def test_sam_helpers_follow_the_reference():
    assert output.reverse_comp(b"acgtNx-") == b"-nnacgt"                      # inc/SequenceOperations.h:56-96
    assert output.reverse_cigar("28M1I21M2D") == "2D21M1I28M"                  # :109-123
    assert output.mapq(1.0) == 30 and output.mapq(0.5) == 3 and output.mapq(0.999999) == 30 and output.mapq(0.9) == 10
    assert output.cfmt(163.69800000001) == "163.698"


def test_simulated_reads_are_reproducible_and_true():
    contigs = synth.make_genome(50_000, 9)
    codes = contigs[0][1]
    a = synth.simulate_reads(codes, 100, 80, 3, sub_rate=0.0)
    b = synth.simulate_reads(codes, 100, 80, 3, sub_rate=0.0)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    comp = np.array([3, 2, 1, 0], np.uint8)
    for r in range(100):
        src = codes[a["pos"][r]:a["pos"][r] + 80]
        want = comp[src[::-1]] if a["strand"][r] else src
        assert np.array_equal(a["bases"][r], want)


def test_snp_call_matches_reference_printer():
    """the library's likelihood-ratio SNP call (gmx_snp_call, host arithmetic, what gmx_format_gmp
    runs per row) against the call column the UNMODIFIED reference's GenomeBwt::PrintSNPCall printed for 3000 read-count
    vectors (tests/golden/ref_snp_calls.json.gz, made by tests/golden/make_golden.py:make_snp_calls): byte for byte,
    p-values included."""
    import gzip, json, os
    from gnumap_b200 import api
    cases = json.loads(gzip.open(os.path.join(os.path.dirname(__file__), "golden", "ref_snp_calls.json.gz")).read())
    kinds = set()
    for c in cases:
        first, second, dip, pval, text = api.snp_call(c["counts"], c["base"], c["monop"], c["pval"])
        assert text.decode() == c["call"], (c, text)
        kinds.add(c["call"][:3] + ("/" if "/" in c["call"] else ""))
    assert {"\tN", "\tN:", "\tY:", "\tN:/", "\tY:/"} <= kinds
