"""-m gpu: the CUDA path (through the C ABI) against the oracle on the same seeded inputs."""
import numpy as np
import pytest

from gnumap_b200 import _abi, index
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from gnumap_b200 import api as a
    return a


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def plain():
    contigs, batch, reads = common.world_plain()
    return index.build_index(contigs), batch, reads


@pytest.fixture(scope="module")
def repeats():
    contigs, batch, reads = common.world_repeats()
    return index.build_index(contigs), batch, reads


def test_fm_search_and_locate(api, O, plain):
    ix, batch, _ = plain
    m = api.Mapper(ix)
    oix = O.OracleIndex(ix)
    rng = np.random.default_rng(3)
    codes = ix.codes()
    kmers = []
    for _ in range(3000):
        p = int(rng.integers(0, ix.l_pac - 10))
        k = bytearray(b"acgt"[c] for c in codes[p:p + 10])
        if rng.random() < 0.3:
            k[int(rng.integers(0, 10))] = b"acgtn"[int(rng.integers(0, 5))]
        kmers.append(bytes(k))
    kmers += [b"nnnnnnnnnn", b"ACGTACGTAC", b"aaaaaaaaaa"]
    k, l = m.get_sa_int(kmers)
    want = np.array([oix.get_sa_int(x) for x in kmers], dtype=np.uint64)
    assert np.array_equal(k, want[:, 0]) and np.array_equal(l, want[:, 1])
    ranks = rng.integers(1, ix.seq_len + 1, size=5000).astype(np.uint64)
    want_pos = np.array([oix.bwt_sa(int(r)) for r in ranks], dtype=np.uint64)
    assert np.array_equal(m.get_sa_coord(ranks), want_pos)               # de-sampled SA
    assert np.array_equal(m.get_sa_coord(ranks[:500], sampled=True), want_pos[:500])   # LF walk, as bwt_sa
    assert m.get_sa_coord(np.array([0], dtype=np.uint64))[0] == np.uint64(0xFFFFFFFFFFFFFFFF)
    m.close()


def test_get_string(api, O, plain):
    ix, _, _ = plain
    m = api.Mapper(ix)
    oix = O.OracleIndex(ix)
    b0 = int(ix.seq_offset[1])
    begins = [0, 5, b0 - 100, b0 - 99, b0 - 50, b0, ix.l_pac - 100, ix.l_pac - 99, ix.l_pac - 1]
    got = m.GetString(begins, 100)
    assert got == [oix.get_string(b, 100) for b in begins]
    m.close()


def test_nw_kernels_explicit_windows(api, O, plain):
    ix, batch, reads = plain
    params = O.default_params()
    m = api.Mapper(ix)
    oix = O.OracleIndex(ix)
    rng = np.random.default_rng(4)
    n = 300
    ridx = rng.integers(0, batch.n_reads, size=n).astype(np.int32)
    strands = rng.integers(0, 2, size=n).astype(np.uint8)
    wins, want_s, want_t, want_h = [], [], [], []
    for t in range(n):
        r = int(ridx[t]); a, b = int(batch.offsets[r]), int(batch.offsets[r + 1])
        pwm = O.fastq_pwm(batch.seq[a:b].tobytes(), batch.qual[a:b].tobytes())
        if strands[t]:
            pwm = O.revcomp_pwm(pwm)
        p = int(reads["pos"][r]) + int(rng.integers(-3, 4))
        w = oix.get_string(max(p, 0), b - a)
        if len(w) != b - a:
            w = oix.get_string(1000, b - a)
        if t % 17 == 0:
            w = w[:40] + b"n" + w[41:]
        wins.append(w)
        want_s.append(O.nw_score(pwm, w, params))
        cons = O.max_char_consensus(pwm)
        want_t.append(O.nw_traceback(pwm, cons, w, params))
        if t < 60:
            want_h.append(O.pair_hmm(pwm, cons, w, params))
    got_s = m.get_align_score(batch, ridx, strands, wins)
    assert np.array_equal(got_s, np.array(want_s, dtype=np.float32)), "K2a score is not bit-exact"
    got_t = m.get_align_score_w_traceback(batch, ridx, strands, wins)
    assert got_t == want_t
    got_h = m.pairHMM(batch, ridx[:60], strands[:60], wins[:60])
    want_h = np.stack(want_h)
    assert np.allclose(got_h[:, : want_h.shape[1]], want_h, rtol=1e-5, atol=1e-7)   # K2c tolerance (float column sums)
    assert np.array_equal(m.self_score(batch), np.array(
        [O.self_score(O.fastq_pwm(batch.seq[batch.offsets[r]:batch.offsets[r + 1]].tobytes(), batch.qual[batch.offsets[r]:batch.offsets[r + 1]].tobytes()),
                      batch.seq[batch.offsets[r]:batch.offsets[r + 1]].tobytes(), params) for r in range(batch.n_reads)], dtype=np.float32))
    m.close()


def test_reference_kats(api, O):
    """The reference's own known-answer tests (reference src/bin_seq.cpp:1046-1215) through the CUDA path."""
    import json, os
    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bin_seq_kat.json")))
    contigs, _, _ = common.world_plain(length=50_000, n_reads=4)
    m = api.Mapper(index.build_index(contigs))
    cons = kat["consensus"].encode()
    batch = _abi.ReadBatch([cons], None, pwm=O.onehot_pwm(cons))
    tb = m.get_align_score_w_traceback(batch, [0, 0], [0, 0], [kat["traceback"][0]["genome"].encode(), kat["traceback"][1]["genome"].encode()],
                                       consensus=[cons, cons])
    for got, want in zip(tb, kat["traceback"]):
        assert got[0].decode() == want["aligned"] and got[1] == want["cigar"]
    h = m.pairHMM(batch, [0], [0], [kat["phmm"]["genome"].encode()])[0]
    assert np.all(np.abs(h - np.array(kat["phmm"]["answer"], dtype=np.float32)) < 0.01)
    m.close()


@pytest.mark.parametrize("mode", [_abi.MODE_NORMAL, _abi.MODE_BS, _abi.MODE_SNP])
@pytest.mark.parametrize("world", ["plain", "repeats", "ragged", "genome_start"])
def test_pipeline_matches_oracle(api, O, world, mode):
    contigs, batch, _ = getattr(common, "world_" + world)()
    ix = index.build_index(contigs)
    pg = common.set_mode(api.default_params(), mode)
    po = common.set_mode(O.default_params(), mode)
    m = api.Mapper(ix, pg)
    got = m.process_batch(batch)
    amount, planes = m.finish()
    want = O.process_batch(O.OracleIndex(ix), po, batch)
    common.compare_batches(got, want)
    common.accum_close(amount, want["amount"], want["hits"], batch.offsets, pg.gen_size, ix.l_pac)
    if mode == _abi.MODE_NORMAL:
        # the exact hash-table vote kernels alone (the filter kernel's overflow path) must give the same answer
        m.reset_accumulators()
        m.set_option(api.OPT_VOTE_FILTER, 0)
        common.compare_batches(m.process_batch(batch), want)
        m.set_option(api.OPT_VOTE_FILTER, 1)
    if mode != _abi.MODE_NORMAL:
        for b in range(5):
            common.accum_close(planes[b], want["planes"][b], want["hits"], batch.offsets, pg.gen_size, ix.l_pac, what=f"plane {b}")
    assert (want["results"]["status"] == _abi.READ_MAPPED).sum() > 0
    m.close()


def test_too_many_and_unique(api, O, repeats):
    ix, batch, _ = repeats
    for kw in (dict(max_matches=1), dict(unique_only=1), dict(min_seed_hits=1), dict(min_seed_hits=3, jump=3), dict(max_kmer_hits=50),
               dict(match_neg=0), dict(match_pos=0), dict(fast=1, mer=14, jump=14), dict(perc=0, align_score=20.0), dict(cutoff=60.0)):
        pg, po = api.default_params(), O.default_params()
        for k, v in kw.items():
            setattr(pg, k, v); setattr(po, k, v)
        m = api.Mapper(ix, pg)
        got = m.process_batch(batch)
        amount, _ = m.finish()
        want = O.process_batch(O.OracleIndex(ix), po, batch)
        common.compare_batches(got, want)
        common.accum_close(amount, want["amount"], want["hits"], batch.offsets, pg.gen_size, ix.l_pac, what=str(kw))
        m.close()


def test_accumulate_across_batches_and_empty(api, O, plain):
    ix, batch, _ = plain
    m = api.Mapper(ix)
    empty = _abi.ReadBatch([], [])
    assert len(m.process_batch(empty)["results"]) == 0
    half = batch.n_reads // 2
    m.process_batch(batch.slice(0, half)); m.process_batch(batch.slice(half, batch.n_reads))
    amount, _ = m.finish()
    want = O.process_batch(O.OracleIndex(ix), O.default_params(), batch)
    assert np.allclose(amount, want["amount"], rtol=1e-5, atol=1e-6)
    m.reset_accumulators()
    assert not m.finish()[0].any()
    m.close()


@pytest.mark.parametrize("mode", ["normal", "snp", "bs"])
def test_whole_program_fixture_through_cuda(api, mode):
    """The CUDA path against the UNMODIFIED reference binary's SAM / SGR / GMP (tests/golden/ref_program_*.json.gz)."""
    from tests import ref_text as output
    from tests import test_oracle_golden as G
    rec = G.load_program(mode)
    lut = {c: i for i, c in enumerate("ACGT")}
    contigs = [(n, np.array([lut[c] for c in s], dtype=np.uint8)) for n, s in rec["contigs"]]
    ix = index.build_index(contigs)
    names = [r[0] for r in rec["reads"]]
    batch = _abi.ReadBatch([r[1].encode() for r in rec["reads"]], [r[2].encode() for r in rec["reads"]])
    p = common.set_mode(api.default_params(), {"normal": _abi.MODE_NORMAL, "bs": _abi.MODE_BS, "snp": _abi.MODE_SNP}[mode])
    m = api.Mapper(ix, p)
    got = m.process_batch(batch)
    amount, planes = m.finish()
    assert int((got["results"]["status"] == _abi.READ_MAPPED).sum()) == rec["matched"]
    sam = list(output.sam_records(ix, names, batch, got["results"], got["hits"], got["cigars"], p.adjust))
    assert sorted(sam) == sorted(rec["sam"]), "SAM body differs from the reference binary's"
    if mode == "normal":
        want = {}
        for ln in rec["sgr"]:
            c, pos, v = ln.split("\t"); want[(c, int(pos))] = float(v)
        mine = {}
        for ln in output.sgr_lines(ix, amount, p.gen_size, min_print=0.0005):
            c, pos, v = ln.split("\t"); mine[(c, int(pos))] = float(v)
        for k, v in want.items():
            assert k in mine and abs(mine[k] - v) <= 1e-5 * abs(v) + 1.1e-5, (k, v, mine.get(k))
    else:
        want = [ln.split("\t") for ln in rec["gmp"]]
        rows = {(r[0], r[1]): r for r in output.gmp_rows(ix, amount, planes, p.mode, min_print=0.0005)}
        for w in want:
            g = rows[(w[0], int(w[1]))]
            assert np.allclose(np.array(g[2:8], dtype=np.float64), np.array([float(x) for x in w[2:8]]), rtol=1e-5, atol=1.1e-5), (g, w)
    m.close()


def test_fast_path_split_phases_chunking_and_device_input(api, O, plain):
    """Every way of driving the batch pipeline gives the per-read results of the default one."""
    import torch
    ix, batch, _ = plain
    want = O.process_batch(O.OracleIndex(ix), O.default_params(), batch)
    fields = ("status", "n_groups", "top_score", "best_score", "best_first_pos", "best_n_positions",
              "best_first_strand", "best_aligned_len", "n_candidates")

    def same(res, what):
        for f in fields:
            assert np.array_equal(res[f], want["results"][f]), f"{what}: {f}"
        # exp() of the device and of libm may differ in the last place
        assert np.allclose(res["denominator"], want["results"]["denominator"], rtol=1e-12, atol=0), what
        assert np.allclose(res["best_posterior"], want["results"]["best_posterior"], rtol=1e-6, atol=1e-12), what

    m = api.Mapper(ix)
    # (1) fast download path: no hit list, CIGARs still available
    m.set_option(api_mod(api).OPT_COLLECT_HITS, 0)
    out = m.process_batch(batch, fetch=False)
    same(out["results"], "fast path")
    cig = [bytes(r).split(b"\0")[0].decode() for r in m.best_cigars(batch.n_reads)]
    assert cig == want["cigars"]
    amount_fast, _ = m.finish()
    # (2) small chunks
    m.reset_accumulators()
    m.set_option(api_mod(api).OPT_CHUNK_READS, 257)
    out = m.process_batch(batch, fetch=False)
    same(out["results"], "chunked")
    assert [bytes(r).split(b"\0")[0].decode() for r in m.best_cigars(batch.n_reads)] == want["cigars"]
    amount_chunked, _ = m.finish()
    assert np.allclose(amount_chunked, amount_fast, rtol=1e-5, atol=1e-6)
    m.set_option(api_mod(api).OPT_CHUNK_READS, 1 << 19)
    m.set_option(api_mod(api).OPT_COLLECT_HITS, 1)
    # (3) gmx_map_batch + gmx_score_batch (resident single chunk) and (4) multi-chunk split (re-run from the kept copy)
    for chunk in (1 << 19, 300):
        m.reset_accumulators()
        m.set_option(api_mod(api).OPT_CHUNK_READS, chunk)
        a = m.process_batch(batch, score=False)
        assert np.array_equal(a["results"]["status"], want["results"]["status"])
        assert np.allclose(a["results"]["denominator"], want["results"]["denominator"], rtol=1e-12, atol=0)
        assert not m.finish()[0].any(), "PHASE A must not touch the accumulators"
        b = m.score_batch(batch)
        common.compare_batches(b, want)
        assert np.allclose(m.finish()[0], want["amount"], rtol=1e-5, atol=1e-6)
    m.set_option(api_mod(api).OPT_CHUNK_READS, 1 << 19)
    # (5) device-resident reads
    m.reset_accumulators()
    dev = torch.device("cuda", 0)
    off = torch.from_numpy(batch.offsets).to(dev); seq = torch.from_numpy(batch.seq).to(dev); qual = torch.from_numpy(batch.qual).to(dev)

    class DevBatch:
        pass
    db = DevBatch(); db.n_reads = batch.n_reads
    s = _abi.GmxReads(); s.n_reads = batch.n_reads; s.offsets = off.data_ptr(); s.seq = seq.data_ptr(); s.qual = qual.data_ptr()
    s.pwm = None; s.on_device = 1; s.max_len = int(np.diff(batch.offsets).max())
    db.struct = s
    out = m.process_batch(db, fetch=True)
    common.compare_batches(out, want)
    m.close()


def api_mod(api):
    return api


def test_full_size_properties(api):
    """BASELINE configs[1] at its full size (100 Mb genome, 1 M x 100 bp reads) through properties that do not need
    the oracle: truth recovery, conservation of posterior mass in the accumulators, run-to-run determinism of the
    per-read results, and equality of the FASTQ-text path with the packed path."""
    import torch
    from gnumap_b200 import synth
    contigs = synth.make_genome(100_000_000, 100)
    ix = index.build_index(contigs, device="cuda")
    codes = contigs[0][1]
    n, L = 1_000_000, 100
    reads = synth.simulate_reads(codes, n, L, 101, sub_rate=0.01, qlo=15, qhi=40)
    batch = _abi.ReadBatch.from_arrays(reads["bases"], reads["quals"])
    m = api.Mapper(ix)
    m.set_option(api_mod(api).OPT_COLLECT_HITS, 0)
    a = m.process_batch(batch, fetch=False)["results"].copy()
    amount, _ = m.finish()
    # (1) truth: a read without indels maps to its source position and strand
    mapped = a["status"] == _abi.READ_MAPPED
    assert mapped.mean() > 0.995
    ok = (a["best_first_pos"] == reads["pos"].astype(np.uint64)) & (a["best_first_strand"] == reads["strand"])
    assert ok[mapped].mean() > 0.999, ok[mapped].mean()
    # (2) every accepted (position, strand) adds posterior x aligned length; posteriors of one read sum to 1 over its
    # positions, so the accumulators hold one unit of mass per aligned base of every mapped read
    mass = float(amount.sum(dtype=np.float64))
    want = float(a["best_aligned_len"][mapped].sum())
    assert abs(mass - want) <= 2e-3 * want, (mass, want)
    # (3) a second pass gives bit-identical per-read records (atomics only reorder the accumulator sums)
    m.reset_accumulators()
    b = m.process_batch(batch, fetch=False)["results"]
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f]), f
    # (4) the same reads as FASTQ text, used in place on the device
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    rec_len = 1 + 7 + 1 + L + 3 + L + 1
    txt = np.empty((n, rec_len), dtype=np.uint8)
    txt[:, 0] = ord("@")
    ids = np.arange(n)
    for d in range(7):
        txt[:, 7 - d] = ord("0") + (ids // 10 ** d) % 10
    txt[:, 8] = 10
    txt[:, 9:9 + L] = lut[reads["bases"]]
    txt[:, 9 + L:12 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    txt[:, 12 + L:12 + 2 * L] = reads["quals"].astype(np.uint8) + 33
    txt[:, -1] = 10
    m.reset_accumulators()
    _, c = m.process_fastq(txt.tobytes(), fetch=False)
    for f in ("status", "best_first_pos", "best_first_strand", "best_score", "n_groups", "best_n_positions", "best_aligned_len"):
        assert np.array_equal(a[f], c["results"][f]), f
    m.close()


def test_raw_pwm_reads_match_fastq_reads(api, O, plain):
    """PRB / INT inputs reach the path as raw float PWMs (reference src/SeqReader.cpp:541-571,901-978).  A batch whose PWM
    rows are the FASTQ rows must give the same per-read results through the raw-PWM code paths of every kernel."""
    ix, batch, _ = plain
    sub = batch.slice(0, 400)
    pwm = np.concatenate([O.fastq_pwm(sub.seq[sub.offsets[r]:sub.offsets[r + 1]].tobytes(), sub.qual[sub.offsets[r]:sub.offsets[r + 1]].tobytes())
                          for r in range(sub.n_reads)])
    seqs = [sub.seq[sub.offsets[r]:sub.offsets[r + 1]].tobytes() for r in range(sub.n_reads)]
    raw = _abi.ReadBatch(seqs, None, pwm=pwm)
    for mode in (_abi.MODE_NORMAL, _abi.MODE_SNP):
        pg = common.set_mode(api.default_params(), mode); po = common.set_mode(O.default_params(), mode)
        m = api.Mapper(ix, pg)
        got = m.process_batch(raw)
        amount, planes = m.finish()
        want = O.process_batch(O.OracleIndex(ix), po, sub)
        common.compare_batches(got, want)
        assert np.allclose(amount, want["amount"], rtol=1e-5, atol=1e-6)
        if planes is not None:
            assert np.allclose(planes, want["planes"], rtol=1e-5, atol=1e-6)
        m.close()


def test_long_reads_take_the_exact_vote_path(api, O):
    """Reads beyond the filter kernel's span (k-mer offset + mer > 448) fall back to the exact hash-table kernels; the
    generic band / traceback / pair-HMM paths handle lengths up to GMX_MAX_READ_LEN."""
    from gnumap_b200 import synth
    contigs = synth.make_genome(400_000, 41, n_contigs=2)
    codes = np.concatenate([c for _, c in contigs])
    ix = index.build_index(contigs)
    seqs, quals = [], []
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    for k, L in enumerate((300, 449, 470, 600, 900, 120, 200, 256)):     # exp(score) stays finite up to ~940 bp
        r = synth.simulate_reads(codes, 12, L, 900 + k, indel_rate=0.3)
        for i in range(12):
            seqs.append(lut[r["bases"][i]].tobytes()); quals.append((r["quals"][i] + 33).astype(np.uint8).tobytes())
    batch = _abi.ReadBatch(seqs, quals)
    m = api.Mapper(ix)
    got = m.process_batch(batch)
    amount, _ = m.finish()
    want = O.process_batch(O.OracleIndex(ix), O.default_params(), batch)
    common.compare_batches(got, want)
    assert np.allclose(amount, want["amount"], rtol=1e-5, atol=1e-6)
    assert (want["results"]["status"] == _abi.READ_MAPPED).sum() > 80
    m.close()
    # SNP mode: the pair-HMM's generic path -- register strips up to 256 bp, local-memory strips (32 columns per lane) up
    # to GMX_MAX_READ_LEN; bin_seq::pairHMM itself has no length limit (reference src/bin_seq.cpp:60-244)
    ps = common.set_mode(api.default_params(), _abi.MODE_SNP)
    po = common.set_mode(O.default_params(), _abi.MODE_SNP)
    m = api.Mapper(ix, ps)
    for sub in (_abi.ReadBatch(seqs[60:], quals[60:]),            # 120, 200 and 256 bp
                batch):                                           # ... and 300 to 900 bp
        m.reset_accumulators()
        got = m.process_batch(sub)
        amount, planes = m.finish()
        want = O.process_batch(O.OracleIndex(ix), po, sub)
        common.compare_batches(got, want)
        # the reference's unscaled FP64 forward pass underflows for the 900-bp reads (fE = 0, posteriors 0 / 0): the same
        # NaNs must come out of the kernel
        assert np.allclose(planes, want["planes"], rtol=1e-5, atol=1e-6, equal_nan=True)
        assert np.nansum(np.abs(want["planes"])) > 0
    assert np.isnan(want["planes"]).any() and not np.isnan(want["planes"]).all()
    m.close()


def test_boundary_misuse_is_safe(api, O, plain):
    """Call sequences and inputs gmx.h allows or must refuse: map(results) -> score(NULL), an option toggled between the
    two calls, offsets that run backwards, a device-resident read longer than the declared max_len, a CIGAR slot that
    is too small (GMX_ERR_OVERFLOW instead of a cut string)."""
    import ctypes as C
    import torch
    ix, batch, _ = plain
    want = O.process_batch(O.OracleIndex(ix), O.default_params(), batch)
    m = api.Mapper(ix)
    # (1) PHASE A into a caller buffer, PHASE B into the library's own storage, in both download modes
    for collect_a, collect_b in ((0, 0), (0, 1), (1, 0)):
        m.reset_accumulators()
        m.set_option(api.OPT_COLLECT_HITS, collect_a)
        res = np.zeros(batch.n_reads, dtype=_abi.READ_RESULT_DTYPE)
        m._ck(m.L.gmx_map_batch(m._ctx, C.addressof(batch.struct), C.c_void_p(res.ctypes.data)), "gmx_map_batch")
        assert np.array_equal(res["status"], want["results"]["status"])
        m.set_option(api.OPT_COLLECT_HITS, collect_b)
        m._ck(m.L.gmx_score_batch(m._ctx, None), "gmx_score_batch(NULL)")
        cig = [bytes(r).split(b"\0")[0].decode() for r in m.best_cigars(batch.n_reads)]
        assert cig == want["cigars"]
        assert np.allclose(m.finish()[0], want["amount"], rtol=1e-5, atol=1e-6)
    m.set_option(api.OPT_COLLECT_HITS, 1)
    # (2) offsets must not run backwards
    bad = _abi.ReadBatch([batch.seq[:100].tobytes(), batch.seq[100:200].tobytes()], [batch.qual[:100].tobytes(), batch.qual[100:200].tobytes()])
    bad.offsets[1] = 250
    with pytest.raises(api.GmxError) as e:
        m.process_batch(bad)
    assert e.value.code == _abi.GMX_ERR_INVALID
    # (3) device-resident reads: one read longer than the declared max_len is refused, nothing is overrun
    dev = torch.device("cuda", 0)
    off = torch.from_numpy(batch.offsets).to(dev); seq = torch.from_numpy(batch.seq).to(dev); qual = torch.from_numpy(batch.qual).to(dev)

    class DevBatch:
        pass
    db = DevBatch(); db.n_reads = batch.n_reads
    s = _abi.GmxReads(); s.n_reads = batch.n_reads; s.offsets = off.data_ptr(); s.seq = seq.data_ptr(); s.qual = qual.data_ptr()
    s.pwm = None; s.on_device = 1; s.max_len = int(np.diff(batch.offsets).max()) - 1
    db.struct = s
    with pytest.raises(api.GmxError) as e:
        m.process_batch(db, fetch=False)
    assert e.value.code == _abi.GMX_ERR_INVALID
    s.max_len += 1
    common.compare_batches(m.process_batch(db), want)                 # the context is usable after the error
    # (4) a CIGAR that does not fit its slot is an error, never a cut string: reads with three separate deletions
    codes = ix.codes()
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = []
    for p0 in (5000, 60000, 123456, 250000):
        keep = np.ones(103, dtype=bool); keep[[20, 51, 82]] = False
        seqs.append(lut[codes[p0:p0 + 103][keep]].tobytes())
    gapped = _abi.ReadBatch(seqs, [b"I" * 100] * len(seqs))
    m.close()
    pg, po = api.default_params(), O.default_params()
    pg.align_score = po.align_score = 0.8                       # three gaps and the shifted tail stay above the threshold
    want_g = O.process_batch(O.OracleIndex(ix), po, gapped)
    assert (want_g["results"]["status"] == _abi.READ_MAPPED).all() and min(len(c) for c in want_g["cigars"]) > 15, want_g["cigars"]
    m = api.Mapper(ix, pg)
    common.compare_batches(m.process_batch(gapped), want_g)
    m.set_option(api.OPT_CIGAR_STRIDE, 16)
    with pytest.raises(api.GmxError) as e:
        m.process_batch(gapped)
    assert e.value.code == _abi.GMX_ERR_OVERFLOW
    m.set_option(api.OPT_CIGAR_STRIDE, 128)
    m.reset_accumulators()
    common.compare_batches(m.process_batch(gapped), want_g)
    m.close()


@pytest.mark.parametrize("mode", [_abi.MODE_NORMAL, _abi.MODE_BS, _abi.MODE_SNP])
def test_optimistic_chunks_equal_synchronous_chunks(api, O, repeats, mode):
    """GMX_OPT_OPTIMISTIC: chunks issued without a host wait (bounds predicted from the previous chunk), chunks whose bounds
    did not hold and were run again (value 2 forces that for every chunk), and chunks that wait for their counts give the
    same per-read results, CIGARs, multi-position lists and accumulators -- over calls that keep the prediction, too."""
    ix, batch, _ = repeats
    po = common.set_mode(O.default_params(), mode)
    want = O.process_batch(O.OracleIndex(ix), po, batch)
    m = api.Mapper(ix, common.set_mode(api.default_params(), mode))
    m.set_option(api.OPT_COLLECT_HITS, 0)
    m.set_option(api.OPT_CHUNK_READS, 193)
    runs = {}
    for opt in (0, 1, 2):
        m.set_option(api.OPT_OPTIMISTIC, opt)
        m.reset_accumulators()
        before = m.chunk_stats()
        out = m.process_batch(batch, fetch=False)
        half = batch.n_reads // 2
        m.process_batch(batch.slice(0, half), fetch=False)                # a second and third call: single-chunk and multi-chunk
        m.process_batch(batch.slice(half, batch.n_reads), fetch=False)
        after = m.chunk_stats()
        issued, rerun = after[0] - before[0], after[1] - before[1]
        if opt == 0:
            assert issued == 0
        elif opt == 1:
            assert issued > 0 and rerun <= issued
        else:
            assert issued > 0 and rerun >= issued // 2          # a chunk with half the candidates of its predecessor still fits
        runs[opt] = (out["results"].copy(), m.finish())
    for f in ("status", "n_groups", "top_score", "best_score", "best_first_pos", "best_n_positions", "best_first_strand",
              "best_aligned_len", "n_candidates", "denominator", "best_posterior"):
        assert f in ("denominator", "best_posterior") or np.array_equal(runs[0][0][f], want["results"][f]), f
        for opt in (1, 2):
            assert np.array_equal(runs[opt][0][f], runs[0][0][f]), (opt, f)
    for opt in (1, 2):
        assert np.allclose(runs[opt][1][0], runs[0][1][0], rtol=1e-5, atol=1e-6)
        if runs[0][1][1] is not None:
            assert np.allclose(runs[opt][1][1], runs[0][1][1], rtol=1e-5, atol=1e-6)
    assert runs[0][1][0].sum() > 0
    m.close()
