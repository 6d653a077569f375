"""SURVEY.md §8(f) rank 1 -- FASTQ text -> reads.

CPU: gmx_fastq_scan_host (C++ restatement of SeqReader::get_more_fastq incl. its recovery from malformed records)
against what the reference's own SeqReader returned for the same texts (tests/golden/ref_fastq.json).
GPU: the device indexer equals the host scan on well-formed text and refuses malformed text; gmx_process_fastq
(reads used in place inside the text) equals gmx_process_batch on the packed reads."""
import json
import os

import numpy as np
import pytest

from gnumap_b200 import _abi, api, index, synth
from tests import common

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_fastq.json")


def fields(text, recs):
    return [[text[int(r["name_off"]): int(r["name_off"]) + int(r["name_len"])].decode("latin-1"),
             text[int(r["seq_off"]): int(r["seq_off"]) + int(r["seq_len"])].decode("latin-1"),
             text[int(r["qual_off"]): int(r["qual_off"]) + int(r["qual_len"])].decode("latin-1")] for r in recs]


def test_host_scan_matches_the_reference_reader():
    cases = json.load(open(GOLD))
    assert len(cases) >= 14
    for name, c in cases.items():
        text = c["text"].encode("latin-1")
        if c["n"] < 0:                                    # the reference threw "Invalid Fastq Character"
            with pytest.raises(api.GmxError) as e:
                api.fastq_scan_host(text)
            assert e.value.code == _abi.GMX_ERR_FORMAT, name
            continue
        got = fields(text, api.fastq_scan_host(text))
        assert got == c["reads"], f"{name}: {got[:3]} vs {c['reads'][:3]}"


def test_host_scan_fuzz_against_the_reference_reader(tmp_path):
    """Differential fuzz where the compiled reference is present (oracle/_ref/libref_probe.so, built by
    __graft_entry__.build() next to /root/reference): random FASTQ texts with deleted, blank, garbage, shortened,
    lengthened and duplicated lines through the reference's own SeqReader and through gmx_fastq_scan_host."""
    import ctypes as C
    from oracle import oracle as O
    if not os.path.exists(O.REF_PROBE):
        pytest.skip("oracle/_ref/libref_probe.so has not been built (needs /root/reference at build time)")
    L = C.CDLL(O.REF_PROBE)
    rng = np.random.default_rng(2718)
    fn = str(tmp_path / "f.fq")
    buf = C.create_string_buffer(1 << 18)

    def rec(i):
        n = int(rng.integers(1, 40))
        return [b"@r%d" % i, bytes(b"ACGTNacgtn"[int(c)] for c in rng.integers(0, 10, size=n)), b"+",
                bytes(int(x) + 33 for x in rng.integers(0, 41, size=n))]

    checked = recovered = 0
    while checked < 400:
        lines = []
        for i in range(int(rng.integers(1, 8))):
            lines += rec(i)
        for _ in range(int(rng.integers(0, 4))):
            if not lines:
                break
            op, k = int(rng.integers(0, 8)), int(rng.integers(0, len(lines)))
            if op == 0: del lines[k]
            elif op == 1: lines.insert(k, b"")
            elif op == 2: lines.insert(k, b"GARBAGE")
            elif op == 3: lines[k] = lines[k][: max(0, len(lines[k]) - int(rng.integers(1, 5)))]
            elif op == 4: lines[k] = lines[k] + b"XX"
            elif op == 5: lines.insert(k, b"@extra")
            elif op == 6: lines.insert(k, b"+")
            else: lines[k] = b"@" + lines[k]
        text = b"\n".join(lines) + (b"\n" if rng.random() < 0.8 else b"")
        if not text.startswith(b"@"):
            continue                                      # the reference sniffs the format from the first character
        with open(fn, "wb") as f:
            f.write(text)
        n = L.refp_read_fastq(fn.encode(), buf, 1 << 18)
        want = None if n < 0 else ([ln.split("\t") for ln in buf.value.decode("latin-1").split("\n")[:-1]] if n > 0 else [])
        try:
            got = fields(text, api.fastq_scan_host(text))
        except api.GmxError as e:
            assert e.code == _abi.GMX_ERR_FORMAT
            got = None
        assert got == want, (text, got, want)
        checked += 1
        recovered += want is not None and len(want) != text.count(b"@r")
    assert recovered > 50


def test_illumina_offset_falls_back_like_the_reference():
    # quality chars below 64 with --illumina: the reference turns the flag off and re-reads at offset 33 (SeqReader.cpp:1180-1188)
    text = b"@a\nACGT\n+\n5555\n"
    assert len(api.fastq_scan_host(text, illumina=1)) == 1


def make_fastq_text(reads):
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    out = []
    for k in range(reads["bases"].shape[0]):
        out.append(b"@r%d\n" % k + lut[reads["bases"][k]].tobytes() + b"\n+\n" + (reads["quals"][k] + 33).astype(np.uint8).tobytes() + b"\n")
    return b"".join(out)


@pytest.mark.gpu
def test_device_indexer_equals_host_scan():
    cases = json.load(open(GOLD))
    contigs = synth.make_genome(5000, 3)
    m = api.Mapper(index.build_index(contigs))
    well = ("well_formed", "no_final_newline", "plus_repeats_name", "quality_longer", "empty_sequence", "lowercase_and_other_letters", "empty_file")
    for name, c in cases.items():
        text = c["text"].encode("latin-1")
        if name in well:
            assert fields(text, m.fastq_scan(text)) == c["reads"], name
        else:                                             # needs the recovery path (or is invalid): the device refuses
            with pytest.raises(api.GmxError) as e:
                m.fastq_scan(text)
            assert e.value.code == _abi.GMX_ERR_FORMAT, name
    m.close()


@pytest.mark.gpu
def test_process_fastq_equals_process_batch():
    contigs, batch, reads = common.world_plain(seed=31, length=200_000, n_reads=3000, read_len=80)
    ix = index.build_index(contigs)
    text = make_fastq_text(reads)
    m = api.Mapper(ix)
    want = m.process_batch(batch)
    amount_want, _ = m.finish()
    m.reset_accumulators()
    names, got = m.process_fastq(text)
    assert names == [f"r{k}" for k in range(batch.n_reads)]
    common.compare_batches(got, want)
    assert np.allclose(m.finish()[0], amount_want, rtol=1e-5, atol=1e-6)
    # malformed text: same answer through the host-scan fallback (one garbage line and one blank line inserted)
    lines = text.split(b"\n")
    messy = b"\n".join(lines[:400] + [b"GARBAGE"] + lines[400:800] + [b""] + lines[800:])
    m.reset_accumulators()
    names2, got2 = m.process_fastq(messy)
    assert names2 == names
    common.compare_batches(got2, want)
    # chunked + fast download path
    m.set_option(api.OPT_CHUNK_READS, 700); m.set_option(api.OPT_COLLECT_HITS, 0)
    m.reset_accumulators()
    _, got3 = m.process_fastq(text, fetch=False)
    for f in ("status", "best_first_pos", "best_score", "n_groups"):
        assert np.array_equal(got3["results"][f], want["results"][f]), f
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("collect", [1, 0])
def test_native_sam_equals_the_reference_binary(collect):
    """gmx_format_sam against the SAM body the unmodified reference binary wrote (whole-program
    fixture with a planted repeat: multi-position best groups, both strands, indels)."""
    from tests import test_oracle_golden as G
    rec = G.load_program("normal")
    lut = {c: i for i, c in enumerate("ACGT")}
    contigs = [(n, np.array([lut[c] for c in s], dtype=np.uint8)) for n, s in rec["contigs"]]
    ix = index.build_index(contigs)
    text = "".join(f"@{nm}\n{s}\n+\n{q}\n" for nm, s, q in rec["reads"]).encode()
    m = api.Mapper(ix)
    m.set_option(api.OPT_COLLECT_HITS, collect)
    names, got = m.process_fastq(text, fetch=False)
    recs = api.fastq_scan_host(text)
    sam = m.format_sam(text, recs, got["results"]).decode().split("\n")[:-1]
    assert sorted(sam) == sorted(rec["sam"])
    assert max(int(ln.split("X0:i:")[1]) for ln in sam) > 1, "fixture must contain multi-position groups"
    m.close()


@pytest.mark.gpu
def test_native_sgr_equals_the_reference_binary():
    """SURVEY.md §8(f) rank 3 (Normal-mode part): gmx_format_sgr against the .sgr the unmodified reference binary wrote.
    FP32 atomics reorder the sums, so values are compared numerically at the file's five decimals; the set of printed
    bins may differ only where the value sits on the MIN_PRINT threshold."""
    from tests import test_oracle_golden as G
    from tests import ref_text as output
    rec = G.load_program("normal")
    lut = {c: i for i, c in enumerate("ACGT")}
    contigs = [(n, np.array([lut[c] for c in s], dtype=np.uint8)) for n, s in rec["contigs"]]
    ix = index.build_index(contigs)
    text = "".join(f"@{nm}\n{s}\n+\n{q}\n" for nm, s, q in rec["reads"]).encode()
    m = api.Mapper(ix)
    m.process_fastq(text, fetch=False)
    got = m.format_sgr().decode().split("\n")[:-1]
    amount, _ = m.finish()
    assert got == list(output.sgr_lines(ix, amount, 8)), "native formatter differs from the restated printer on the same accumulators"
    def table(lines):
        return {(c, int(p)): float(v) for c, p, v in (ln.split("\t") for ln in lines)}
    g, w = table(got), table(rec["sgr"])
    for k in set(g) | set(w):
        a, b = g.get(k, 0.0), w.get(k, 0.0)
        assert abs(a - b) <= 1e-5 * abs(b) + 1.1e-5 or max(a, b) < 0.00102, (k, a, b)
    assert len(w) > 1000
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["snp", "bs", "snpdepth"])
def test_native_gmp_rows_and_calls(mode):
    """SURVEY.md §8(f) rank 3 (SNP / bisulfite part): gmx_format_gmp -- rows selected and gathered on the device --
    against the restated row printer on the same accumulators (text equality; the call column through gmx_snp_call,
    itself pinned to the reference's PrintSNPCall by tests/test_host_logic.py), and against the .gmp the unmodified
    reference binary wrote (numbers at the file's decimals: FP32 atomics reorder the sums)."""
    from tests import test_oracle_golden as G, common
    from tests import ref_text as output
    rec = G.load_program(mode)
    lut = {c: i for i, c in enumerate("ACGT")}
    contigs = [(n, np.array([lut[c] for c in s], dtype=np.uint8)) for n, s in rec["contigs"]]
    ix = index.build_index(contigs)
    p = common.set_mode(api.default_params(), _abi.MODE_BS if mode == "bs" else _abi.MODE_SNP)
    text = "".join(f"@{nm}\n{s}\n+\n{q}\n" for nm, s, q in rec["reads"]).encode()
    m = api.Mapper(ix, p)
    m.process_fastq(text, fetch=False)
    got = m.format_gmp(target_base=1 if mode == "bs" else -1).decode().split("\n")
    assert got[-1] == "" and len(got) > 1000
    got = got[:-1]
    amount, planes = m.finish()
    codes = ix.codes()
    want = []
    for r in output.gmp_rows(ix, amount, planes, p.mode):
        line = ("%s\t%d\t%f" if mode == "bs" else "%s\t%d\t%.5f") % r[:3] + "".join("\t%.5f" % x for x in r[3:8])
        if mode != "bs":
            pos = int(ix.seq_offset[ix.names.index(r[0])]) + r[1] - 1
            line += api.snp_call(np.array(r[3:8], dtype=np.float32), int(codes[pos]))[4].decode()
        want.append(line)
    assert got == want
    ref = {tuple(ln.split("\t")[:2]): ln.split("\t") for ln in rec["gmp"]}
    mine = {tuple(ln.split("\t")[:2]): ln.split("\t") for ln in got}
    both = set(ref) & set(mine)
    assert len(both) >= 0.999 * len(ref) and len(mine) <= 1.001 * len(ref) + 2
    same = 0
    for k in both:
        assert np.allclose([float(x) for x in mine[k][2:8]], [float(x) for x in ref[k][2:8]], rtol=1e-5, atol=1.1e-5), (mine[k], ref[k])
        same += mine[k][8:] == ref[k][8:]
    assert same >= 0.995 * len(both)
    if mode == "snpdepth":                                 # 15x two-haplotype sample: confident mono- and diploid calls
        assert sum(mine[k][8].startswith("Y:") for k in both) >= 30 and sum("/" in mine[k][8] for k in both) >= 10
    m.close()
    m0 = api.Mapper(ix)
    with pytest.raises(api.GmxError):
        m0.format_gmp()                      # Normal mode has no .gmp
    m0.close()


@pytest.mark.gpu
def test_device_sam_formatter_equals_host_formatter():
    """gmx_format_sam after gmx_process_fastq formats on the GPU (csrc/sam_out.cuh): the bytes, in read order, must be the
    host formatter's -- both strands, indels, three contigs, multi-position best groups (written by the host into the
    places the device reserved), reads that print nothing."""
    from gnumap_b200 import synth
    contigs = synth.make_genome(300_000, 61, n_contigs=3)
    codes = np.concatenate([c for _, c in contigs])
    codes[200_000:200_600] = codes[50_000:50_600]
    codes[260_000:260_400] = 3 - codes[90_000:90_400][::-1]
    b = np.cumsum([0] + [len(c) for _, c in contigs])
    contigs = [(n, codes[b[i]:b[i + 1]]) for i, (n, _) in enumerate(contigs)]
    n_reads, L = 20000, 100
    reads = synth.simulate_reads(codes, n_reads, L, 62, indel_rate=0.2, n_rate=0.002, qlo=5)
    k = n_reads // 10
    reads["pos"][:k] = np.random.default_rng(61).integers(50_000, 50_600 - L, size=k)
    fwd = codes[reads["pos"][:k, None] + np.arange(L)[None, :]]
    reads["bases"][:k] = np.where((reads["strand"][:k] == 1)[:, None], 3 - fwd[:, ::-1], fwd)
    reads["bases"][k:k + 50] = 4                                             # all-N reads: unmapped
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    text = b"".join(b"@read_%d/%d\n" % (i, i % 7) + lut[reads["bases"][i]].tobytes() + b"\n+\n" + (reads["quals"][i] + 33).astype(np.uint8).tobytes() + b"\n"
                    for i in range(n_reads))
    ix = index.build_index(contigs)
    m = api.Mapper(ix)
    m.set_option(api.OPT_COLLECT_HITS, 0)
    m.set_option(api.OPT_CHUNK_READS, 6000)
    _, got = m.process_fastq(text, fetch=False)
    recs = api.fastq_scan_host(text)
    dev = m.format_sam(text, recs, got["results"])
    m.set_option(api.OPT_SAM_DEVICE, 0)
    host = m.format_sam(text, recs, got["results"])
    assert dev == host
    lines = host.decode().split("\n")[:-1]
    assert len(lines) > 0.9 * n_reads and any(ln.split("\t")[1] == "16" for ln in lines)
    assert max(int(ln.split("X0:i:")[1]) for ln in lines) > 1 and any("I" in ln.split("\t")[5] or "D" in ln.split("\t")[5] for ln in lines)
    m.close()


def test_g_format_of_the_device_sam_writer_is_printf_g():
    """The "%g" writer the device SAM formatter uses (exact 128-bit decimal conversion) against printf, through the
    host-callable copy exported for this test."""
    import ctypes as C
    L = api.load_library()
    L.gmx_format_g.argtypes = [C.c_double, C.c_char_p, C.c_int]
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.random(20000).astype(np.float32).astype(np.float64), (rng.random(20000) * 300).astype(np.float32).astype(np.float64) * 4,
                           10.0 ** (-rng.random(5000) * 15), np.array([0.0, 1.0, 0.5, 999999.5, 999999.49, 1e6, 1e-5, 9.9999995e-5, 123456.5, 1234565.0, 131.322, -2.5])])
    buf = C.create_string_buffer(64)
    for v in vals:
        n = L.gmx_format_g(float(v), buf, 64)
        assert n > 0 and buf.raw[:n].decode() == "%g" % v, (v, buf.raw[:n])


@pytest.mark.gpu
def test_pipelined_fastq_text_equals_whole_text_path():
    """A host FASTQ text of several pieces (GMX_OPT_FASTQ_PIECE) is cut at record boundaries and piece p + 1 is uploaded and
    indexed while piece p is mapped: per-read results, record index, accumulators and the device-formatted SAM must be
    those of the whole-text path; a text whose LATER part is malformed reports the reads done so far and the wrapper
    finishes it through the host scan."""
    from gnumap_b200 import synth
    contigs = synth.make_genome(250_000, 91, n_contigs=2)
    codes = np.concatenate([c for _, c in contigs])
    n_reads = 6000
    rng = np.random.default_rng(92)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    recs_txt = []
    for i in range(n_reads):
        L = int(rng.integers(40, 121))
        r = synth.simulate_reads(codes, 1, L, 1000 + i, indel_rate=0.2, qlo=0)      # Q0 -> '!' ... quality lines may start with '@' (Q31)
        q = (r["quals"][0] + 33).astype(np.uint8)
        if i % 5 == 0:
            q[0] = ord("@")                                                          # a quality line that looks like a name line
        recs_txt.append(b"@r%d\n" % i + lut[r["bases"][0]].tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    text = b"".join(recs_txt)
    ix = index.build_index(contigs)
    m = api.Mapper(ix)
    m.set_option(api.OPT_COLLECT_HITS, 0)
    m.set_option(api.OPT_FASTQ_PIECE, 0)
    names0, whole = m.process_fastq(text, fetch=False)
    recs0 = api.fastq_scan_host(text)
    sam0 = m.format_sam(text, recs0, whole["results"])
    amount0, _ = m.finish()
    # chunks issued without a host wait (default), every such chunk voided and run again (2), every chunk waiting (0)
    for piece, chunk, optimistic in ((100_000, 1 << 19, 1), (150_000, 700, 1), (150_000, 700, 2), (120_000, 900, 0)):
        m.reset_accumulators()
        m.set_option(api.OPT_FASTQ_PIECE, piece); m.set_option(api.OPT_CHUNK_READS, chunk); m.set_option(api.OPT_OPTIMISTIC, optimistic)
        before = m.chunk_stats()
        names1, piped = m.process_fastq(text, fetch=False)
        issued, rerun = (x - y for x, y in zip(m.chunk_stats(), before))
        assert (issued == 0) if optimistic == 0 else (issued > 0 and (optimistic == 1 or rerun >= issued // 2))
        assert names1 == names0
        for f in whole["results"].dtype.names:
            assert np.array_equal(piped["results"][f], whole["results"][f]), f
        assert m.format_sam(text, recs0, piped["results"]) == sam0
        assert np.allclose(m.finish()[0], amount0, rtol=1e-5, atol=1e-6)
    # malformed late in the text: a record with a quality line shorter than its sequence
    k = 4000
    bad = recs_txt[k].split(b"\n")
    bad[3] = bad[3][:-3]
    messy = b"".join(recs_txt[:k]) + b"\n".join(bad) + b"".join(recs_txt[k + 1:])
    m.set_option(api.OPT_FASTQ_PIECE, 0); m.set_option(api.OPT_CHUNK_READS, 1 << 19); m.set_option(api.OPT_OPTIMISTIC, 1)
    m.reset_accumulators()
    names_a, res_a = m.process_fastq(messy, fetch=False)              # whole text: device indexer refuses, host scan does it all
    amount_a, _ = m.finish()
    m.set_option(api.OPT_FASTQ_PIECE, 100_000)
    m.reset_accumulators()
    names_b, res_b = m.process_fastq(messy, fetch=False)              # pipelined: the first pieces on the device, the rest through the host scan
    assert names_b == names_a and len(names_a) in (n_reads - 1, n_reads)
    for f in ("status", "best_first_pos", "best_score", "n_groups", "best_first_strand"):
        assert np.array_equal(res_b["results"][f], res_a["results"][f]), f
    assert np.allclose(m.finish()[0], amount_a, rtol=1e-5, atol=1e-6)
    m.close()
