"""-m gpu: the library's GPU index builder (gmx_index_build) against the files the reference's own
bwa_index wrote (tests/golden/ref_index.npz; and, where the compiled reference is present, bwa_index run on the spot) and
against the torch / numpy construction of gnumap_b200/index.py, on genomes around the occ-block and SA-sample boundaries,
a homopolymer, tandem repeats (many doubling rounds) and a 3 Mb random genome."""
import os

import numpy as np
import pytest

from gnumap_b200 import index, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _same(a, b, what):
    assert a.primary == b.primary and a.seq_len == b.seq_len and a.l_pac == b.l_pac, what
    assert np.array_equal(a.L2, b.L2), what
    assert np.array_equal(a.bwt, b.bwt), f"{what}: bwt"
    assert np.array_equal(a.sa, b.sa), f"{what}: sa"
    assert np.array_equal(a.pac, b.pac), f"{what}: pac"


def test_native_builder_reproduces_bwa_index_files(tmp_path):
    g = np.load(os.path.join(GOLD, "ref_index.npz"))
    lens = g["lens"]; codes = g["codes"]
    b = np.cumsum([0] + list(lens))
    contigs = [(f"chrS{i + 1}", codes[b[i]:b[i + 1]]) for i in range(len(lens))]
    ix = index.build_index(contigs, device="cuda")
    prefix = str(tmp_path / "ix.fa")
    index.save_index(ix, prefix)
    for ext in ("bwt", "sa", "pac"):
        assert np.array_equal(np.fromfile(prefix + ".gnumap." + ext, dtype=np.uint8), g[ext]), f".gnumap.{ext} differs from bwa_index's"


def test_native_builder_equals_the_torch_construction():
    from gnumap_b200 import api
    rng = np.random.default_rng(99)
    shapes = [[1], [2], [15, 1], [127], [128, 129], [31, 32, 33], [1000, 1, 64], [4097], [255, 256, 257, 4, 12], [70000]]
    for k, lens in enumerate(shapes):
        contigs = []
        for j, n in enumerate(lens):
            c = rng.integers(0, 4, size=n, dtype=np.uint8)
            if k == 6 and j == 0:
                c[100:900] = 0                                     # homopolymer: one round per doubling of the run
            if k == 7:
                c[1000:3400] = np.tile(c[1000:1012], 200)           # tandem repeat
            if k == 9:
                c[20000:40000] = c[:20000]                          # a 20 kb exact repeat
            contigs.append((f"c{k}_{j}", c))
        _same(index.build_index(contigs, device="cuda"), index.build_index(contigs, device="cpu"), f"genome {k} {lens}")
    r = api.index_build(np.zeros(5000, dtype=np.uint8))            # all-A: the worst case for prefix doubling
    assert r["rounds"] >= 9 and r["primary"] == 5000


def test_native_builder_at_scale_and_live_bwa_index(tmp_path):
    from oracle import oracle as O
    contigs = synth.make_genome(3_000_000, 17, n_contigs=3)
    ix = index.build_index(contigs, device="cuda")
    _same(ix, index.build_index(contigs, device="cpu"), "3 Mb")
    if not O.have_ref_binary():
        return
    small = [(n, c[:40_000]) for n, c in contigs]
    fa = str(tmp_path / "g.fa")
    synth.write_fasta(fa, small)
    empty = str(tmp_path / "e.fq"); open(empty, "w").close()
    O.run_reference(fa, empty, str(tmp_path / "out"), threads=1, mmap_threshold=1024)
    mine = str(tmp_path / "mine.fa")
    index.save_index(index.build_index(small, device="cuda"), mine)
    for ext in ("bwt", "sa", "pac", "ann", "amb"):
        assert np.array_equal(np.fromfile(mine + ".gnumap." + ext, dtype=np.uint8), np.fromfile(fa + ".gnumap." + ext, dtype=np.uint8)), ext
