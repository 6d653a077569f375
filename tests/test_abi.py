"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/gmx.h declares, the ctypes /
numpy mirrors agree with the C struct layouts, and the product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from gnumap_b200 import _abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gmx.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmx_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = api.load_library()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/gmx.h but not exported by libgmx.so"
    assert set(api.EXPORTS) == set(names), "gnumap_b200.api.EXPORTS is out of date with include/gmx.h"
    assert lib.gmx_abi_version() == 1


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof as gcc sees the header vs the ctypes and numpy mirrors."""
    prog = tmp_path / "layout.c"
    prog.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "gmx.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(gmx_index), sizeof(gmx_params), sizeof(gmx_reads), sizeof(gmx_read_result), sizeof(gmx_hit), sizeof(gmx_stage_stats));
  printf("%zu %zu %zu %zu %zu\\n", offsetof(gmx_read_result, best_first_pos), offsetof(gmx_read_result, hit_begin), offsetof(gmx_read_result, best_aligned_len), offsetof(gmx_hit, group), offsetof(gmx_params, gap));
  printf("%zu %zu\\n", offsetof(gmx_reads, on_device), offsetof(gmx_index, seq_offset));
  printf("%zu %zu %zu\\n", sizeof(gmx_fastq_rec), offsetof(gmx_fastq_rec, seq_len), offsetof(gmx_reads, lens));
  return 0; }
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = [int(x) for x in out]
    assert sizes[0] == C.sizeof(_abi.GmxIndex)
    assert sizes[1] == C.sizeof(_abi.GmxParams)
    assert sizes[2] == C.sizeof(_abi.GmxReads)
    assert sizes[3] == _abi.READ_RESULT_DTYPE.itemsize
    assert sizes[4] == _abi.HIT_DTYPE.itemsize
    assert sizes[5] == C.sizeof(_abi.GmxStageStats)
    assert sizes[6] == _abi.READ_RESULT_DTYPE.fields["best_first_pos"][1]
    assert sizes[7] == _abi.READ_RESULT_DTYPE.fields["hit_begin"][1]
    assert sizes[8] == _abi.READ_RESULT_DTYPE.fields["best_aligned_len"][1]
    assert sizes[9] == _abi.HIT_DTYPE.fields["group"][1]
    assert sizes[10] == _abi.GmxParams.gap.offset
    assert sizes[11] == _abi.GmxReads.on_device.offset
    assert sizes[12] == _abi.GmxIndex.seq_offset.offset
    assert sizes[13] == _abi.FASTQ_REC_DTYPE.itemsize
    assert sizes[14] == _abi.FASTQ_REC_DTYPE.fields["seq_len"][1]
    assert sizes[15] == _abi.GmxReads.lens.offset


def test_default_params_are_the_reference_defaults():
    """inc/const_define.h:46-107 and setup_alignment_matrices() (inc/a_matrices.c:25-126)."""
    p = api.default_params()
    S = np.ctypeslib.as_array(p.align_scores).reshape(256, 4)
    assert S[ord("a")].tolist() == [0.75, -0.75, -0.5, -0.75]          # match 3, transversion -3, transition -2, x 0.25
    assert S[ord("A")].tolist() == S[ord("a")].tolist() and S[ord("n")].tolist() == [-0.75] * 4
    assert (p.gap, p.max_gap, p.mer, p.jump, p.min_seed_hits) == (-1.0, 3, 10, 5, 2)
    assert (p.max_matches, p.gen_size, p.perc) == (1000, 8, 1) and abs(p.align_score - 0.9) < 1e-7
    P = np.ctypeslib.as_array(p.phmm_scores).reshape(256, 4)
    assert np.allclose(P[ord("c")], [0.005, 0.98, 0.005, 0.01])
    from oracle import oracle as O
    q = O.default_params()
    assert bytes(p) == bytes(q), "library and oracle defaults differ"


def test_no_cpu_fallback(monkeypatch):
    """Without a CUDA device the library refuses to create a context (GMX_ERR_NO_DEVICE); it never computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gnumap_b200 import index, synth
    ix = index.build_index(synth.make_genome(2000, 1))
    with pytest.raises(api.GmxError) as e:
        api.Mapper(ix)
    assert e.value.code == _abi.GMX_ERR_NO_DEVICE
    assert b"no CPU fallback" in api.load_library().gmx_strerror(_abi.GMX_ERR_NO_DEVICE)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "LIB_PATH", str(tmp_path / "libgmx.so"))
    with pytest.raises(ImportError):
        api.load_library()


def test_product_never_touches_the_oracle():
    """Nothing under gnumap_b200/ or include/ may import, link or execute oracle/."""
    for base in ("gnumap_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "liboracle" not in txt and "gnumap_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, \
                        f"{base}/{f} references the oracle"


def test_reference_side_binding_compiles(tmp_path):
    """integration/gnumap_gmx_bridge.cpp (the stub shown in INTEGRATION.md) must compile against the reference's own
    headers: the binding uses the reference's real types (Read, TopReadOutput, bwaidx_t, GenomeBwt)."""
    ref_inc = "/root/reference/inc"
    if not os.path.isdir(ref_inc):
        pytest.skip("the reference checkout is not present on this machine")
    src = os.path.join(ROOT, "integration", "gnumap_gmx_bridge.cpp")
    obj = tmp_path / "bridge.o"
    subprocess.check_call(["g++", "-std=c++0x", "-w", "-I", ref_inc, "-I", os.path.join(ROOT, "oracle", "gsl_stub"),
                           "-I", os.path.join(ROOT, "include"), "-DGMX_BRIDGE_TEST_ACCESS", "-c", src, "-o", str(obj)])
    syms = subprocess.check_output(["nm", "-C", str(obj)], text=True)
    for fn in ("gmx_attach(GenomeBwt&, unsigned int)", "gmx_run_slice(GenomeBwt&", "gmx_collect(GenomeBwt&, char const*)", "gmx_slice_reads(unsigned int)"):
        assert fn in syms
    for dep in ("gmx_create", "gmx_process_batch", "gmx_get_hits", "gmx_get_best_alignments", "gmx_finish", "gmx_destroy", "gmx_comm_create"):
        assert f"U {dep}" in syms, f"the binding should call {dep} through the C ABI"
    # INTEGRATION.md shows this file and the Driver.cpp patch, verbatim
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert open(src).read() in doc
    assert open(os.path.join(ROOT, "integration", "driver_gmx.patch")).read() in doc


def test_driver_patch_applies_to_the_reference(tmp_path):
    """integration/driver_gmx.patch applies cleanly to the reference's src/Driver.cpp and the result compiles."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("the reference checkout is not present on this machine")
    work = tmp_path / "Driver.cpp"
    work.write_bytes(open(os.path.join(ref, "src", "Driver.cpp"), "rb").read())
    subprocess.check_call(["patch", "-s", "-p2", str(work), os.path.join(ROOT, "integration", "driver_gmx.patch")])
    txt = work.read_text()
    for call in ("gmx_attach(gGen, gNUM_THREADS)", "gmx_run_slice(", "gmx_collect(gGen, output_file)", "gmx_slice_reads(READS_PER_PROC)"):
        assert call in txt
    subprocess.check_call(["g++", "-std=c++0x", "-w", "-DDEBUG_NW", "-DDEBUG_TIME", "-I", os.path.join(ref, "inc"),
                           "-I", os.path.join(ROOT, "oracle", "gsl_stub"), "-fsyntax-only", str(work)])
