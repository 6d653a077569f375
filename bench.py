#!/usr/bin/env python
"""bench.py -- reads/sec of the GNUMAP hot path (seed -> probabilistic NW -> posterior scatter).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extra] [--no-cpu] [--no-fastq]

One "step" = one pass of the hot path over one batch of synthetic reads.  The headline line is BASELINE.json
configs[1] (synthetic 100 Mb genome, 1 M simulated 100-bp reads with Phred qualities, Normal mode): with N ranks every
rank maps its own 1 M-read shard against its own replica of the index (weak scaling) and the job ends with the one
collective of the path, the NCCL sum-reduce of the accumulators (timed).  The same JSON line carries, under
"configs", one row per remaining BASELINE config, each measured the way the config is stated:
  cfg2  156 Mb genome, 10 M x 150 bp reads, Normal mode, STRONG scaling: the one read set is dealt to the ranks in
        2048-read slices round robin (reference inc/SeqManager.h:329-345)
  cfg3  the same genome in --snp mode (2 M of those reads), the 6 x 624 MB all-reduce INSIDE every timed step
  cfg4  100 Mb genome, 5 M x 100 bp C->T converted reads, -b mode, strong scaling, reduce inside every step

Own arm (JSON keys, see DESIGN.md "Measurement"):
  value        reads/s, inputs resident in HBM when the timed region starts (device-resident gmx_reads), CUDA events
  e2e          reads/s through the C ABI with HOST (pinned) buffers: H2D of the reads and D2H of the per-read
               results + best CIGARs inside the timed region; WALL clock, max over ranks
  roofline     dominant kernel: algorithmic bytes / CUDA-event time of that stage vs MEASURED_PEAKS.json
  cpu_baseline the unmodified reference (oracle/_ref/gnumap) on a bounded sample, all host cores; parity_sample: the
               GPU's SAM for that very sample against the reference's
Reference arm (--impl reference): the unmodified reference on bounded samples of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CACHE = os.environ.get("GMX_BENCH_CACHE", "/tmp/gnumap_b200_bench")
MODES = {"normal": 0, "bs": 1, "snp": 2}
SLICE = 2048                      # READS_PER_PROC, reference inc/const_include.h:64


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_REAL_STDOUT = None


def claim_stdout():
    """Libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line on stdout.  Point fd 1
    at stderr for the run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------------
# workloads: i.i.d. genomes, 1 % substitutions, Q15..40
# ------------------------------------------------------------------------------------------------
def workload_name(genome, reads, read_len, mode, tag):
    return f"synthetic {genome // 1_000_000} Mb genome, {reads} x {read_len} bp reads, {mode} mode ({tag})"


def main_tag(a):
    shape = {("normal", 100_000_000, 100): "BASELINE configs[1]", ("normal", 156_000_000, 150): "BASELINE configs[2] shape, 1 M reads per GPU",
             ("snp", 156_000_000, 150): "BASELINE configs[3] shape, weak", ("bs", 100_000_000, 100): "BASELINE configs[4] shape, weak"}
    return shape.get((a.mode, a.genome, a.read_len), "custom shape")


def get_index(genome, genome_seed, device):
    """Build (GPU suffix sort) or load the cached index, in the reference's on-disk format."""
    from gnumap_b200 import index, synth
    os.makedirs(CACHE, exist_ok=True)
    prefix = os.path.join(CACHE, f"g{genome}_s{genome_seed}.fa")
    if index.index_files_exist(prefix):
        t = time.time()
        ix = index.load_index(prefix)
        log(f"[bench] index loaded from {prefix} in {time.time() - t:.1f}s")
        return ix, prefix
    t = time.time()
    contigs = synth.make_genome(genome, genome_seed)
    ix = index.build_index(contigs, device=device)
    log(f"[bench] index of {genome} bp built in {time.time() - t:.1f}s")
    tmp = prefix + f".tmp{os.getpid()}"
    index.save_index(ix, tmp)
    for ext in (".gnumap.bwt", ".gnumap.sa", ".gnumap.pac", ".gnumap.ann", ".gnumap.amb"):
        os.replace(tmp + ext, prefix + ext)
    if not os.path.exists(prefix):
        with open(prefix, "w") as f:          # the reference never opens the FASTA once the index files exist
            f.write(">chrS\n")
    return ix, prefix


def get_reads_numpy(a, ix, shard: int):
    from gnumap_b200 import synth
    t = time.time()
    reads = synth.simulate_reads(ix.codes(), a.reads, a.read_len, a.reads_seed + 1000 * shard, sub_rate=0.01, qlo=15, qhi=40,
                                 bisulfite=0.95 if a.mode == "bs" else 0.0)
    log(f"[bench] {a.reads} reads simulated in {time.time() - t:.1f}s")
    return reads


# ------------------------------------------------------------------------------------------------
# the reference on a bounded sample: P single-threaded processes of the unmodified binary
# ------------------------------------------------------------------------------------------------
class ReferenceRunner:
    def __init__(self, mode, prefix, reads):
        from oracle import oracle as O
        if not O.have_ref_binary():
            raise RuntimeError("oracle/_ref/gnumap is missing (built by __graft_entry__.build() where /root/reference exists)")
        self.bin = O.REF_BIN
        self.prefix = prefix
        self.reads = reads
        self.mode = mode
        self.cores = os.cpu_count() or 1
        self.dir = tempfile.mkdtemp(prefix="gmx_ref_", dir=CACHE)
        self.cursor = 0
        self.load_s = None
        self.extra = {"normal": [], "bs": ["-b"], "snp": ["--snp"]}[mode]
        self.env = dict(os.environ, MALLOC_MMAP_THRESHOLD_="65536")

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)

    def _fastq(self, path, lo, hi):
        from gnumap_b200 import synth
        sub = {k: v[lo:hi] for k, v in self.reads.items()}
        synth.write_fastq(path, sub, prefix=f"r{lo}_")

    def _startup(self, threads=1):
        empty = os.path.join(self.dir, "empty.fq")
        open(empty, "w").close()
        best = None
        for _ in range(2):                             # second run: page cache warm
            t = time.time()
            subprocess.run([self.bin, "-g", self.prefix, "-o", os.path.join(self.dir, "oe"), "-a", ".9", "-c", str(threads), *self.extra, empty],
                           env=self.env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            best = time.time() - t
        return best

    def run(self, n_reads: int, procs: int | None = None, keep_sam: bool = False):
        """Map `n_reads` reads of the workload with `procs` concurrent `gnumap -c 1` processes (the reference deals
        work to its own threads only in 2048-read slices, so small samples would idle a `-c N` run).  Returns
        (seconds of mapping, reads mapped by the sample, reads in the sample, processes, [(lo, hi, sam path)])."""
        procs = procs or self.cores
        procs = max(1, min(procs, n_reads))
        total = len(self.reads["pos"])
        per = n_reads // procs
        jobs = []
        for p in range(procs):
            lo = self.cursor % max(total - per, 1)
            self.cursor += per
            fq = os.path.join(self.dir, f"s{p}.fq")
            self._fastq(fq, lo, lo + per)
            jobs.append((fq, os.path.join(self.dir, f"o{p}"), lo, lo + per))
        if self.load_s is None:                       # index load + start-up, measured once on an empty read file
            self.load_s = self._startup()
        t0 = time.time()
        ps = [subprocess.Popen([self.bin, "-g", self.prefix, "-o", out, "-a", ".9", "-c", "1", "--no_gmp", *self.extra, fq], env=self.env,
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for fq, out, _, _ in jobs]
        mapped = 0
        for p in ps:
            outp, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError(f"reference run failed ({p.returncode}): {outp[-500:]}")
        wall = time.time() - t0
        sams = []
        for _, out, lo, hi in jobs:                   # matched reads = distinct read names among the SAM records
            sam = out + ".sam"
            if os.path.exists(sam):
                names = set()
                with open(sam) as f:
                    for ln in f:
                        if ln and ln[0] != "@":
                            names.add(ln.split("\t", 1)[0])
                mapped += len(names)
                if keep_sam:
                    sams.append((lo, hi, sam))
        return max(wall - self.load_s, 1e-6), mapped, per * procs, procs, sams

    def run_threads(self, threads: int, slices_per_thread: int = 1):
        """ONE process, `-c threads`, on 2048 x threads x slices_per_thread reads (every thread busy): the reference's own
        multithreaded mode."""
        n = SLICE * threads * slices_per_thread
        n = min(n, len(self.reads["pos"]))
        fq = os.path.join(self.dir, "cn.fq")
        self._fastq(fq, 0, n)
        load = self._startup(threads)
        t0 = time.time()
        p = subprocess.run([self.bin, "-g", self.prefix, "-o", os.path.join(self.dir, "ocn"), "-a", ".9", "-c", str(threads), "--no_gmp", *self.extra, fq],
                           env=self.env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if p.returncode != 0:
            raise RuntimeError(f"reference -c {threads} run failed ({p.returncode})")
        return n / max(time.time() - t0 - load, 1e-6), n


    def run_patched(self, threads: int, n_reads: int, gpus: int = 1):
        """The reference program itself with integration/driver_gmx.patch applied (oracle/_ref/gnumap_gmx): its own option parser,
        FASTQ reader, worker threads and SAM writer around gmx_process_batch.  Returns None when that binary was not built."""
        exe = os.path.join(os.path.dirname(self.bin), "gnumap_gmx")
        if not os.path.exists(exe):
            return None
        n = min(n_reads, len(self.reads["pos"]))
        fq = os.path.join(self.dir, "dropin.fq")
        self._fastq(fq, 0, n)
        env = dict(self.env, GMX_GPUS=str(gpus))
        cmd = [exe, "-g", self.prefix, "-a", ".9", "-c", str(threads), "--no_gmp", *self.extra]
        empty = os.path.join(self.dir, "empty.fq")
        open(empty, "w").close()
        load = None
        for _ in range(2):                             # start-up: index load, context creation, index upload
            t = time.time()
            subprocess.run(cmd + ["-o", os.path.join(self.dir, "de"), empty], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            load = time.time() - t
        t0 = time.time()
        p = subprocess.run(cmd + ["-o", os.path.join(self.dir, "do"), fq], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        wall = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError(f"patched binary failed ({p.returncode}): {p.stdout[-400:]}")
        return {"value": n / max(wall - load, 1e-6), "unit": "reads/s", "reads": n, "threads": threads, "gpus": gpus, "seconds": wall, "startup_seconds": load,
                "what": f"`gnumap_gmx -c {threads}` (the reference binary with integration/driver_gmx.patch: its FASTQ reader, worker threads and SAM "
                        "writer around gmx_process_batch in 65 536-read slices), wall clock minus the start-up measured on an empty read file"}


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ix, prefix = get_index(a.genome, a.genome_seed, dev)
    reads = get_reads_numpy(a, ix, 0)
    rr = ReferenceRunner(a.mode, prefix, reads)
    try:
        budget_s = float(os.environ.get("GMX_REF_BUDGET_S", "150"))
        t, _, n, procs, _ = rr.run(max(rr.cores * 8, 64))            # calibration (always untimed)
        rate = n / t
        per_step = max(int(rate * budget_s / max(a.steps + a.warmup, 1)), rr.cores)
        per_step = min(per_step, a.reads)
        log(f"[bench] reference calibration: {rate:.1f} reads/s on {procs} processes; {per_step} reads per step")
        for _ in range(a.warmup):
            rr.run(per_step)
        times, done, mapped = [], 0, 0
        for _ in range(a.steps):
            t, m, n, procs, _ = rr.run(per_step)
            times.append(t); done += n; mapped += m
        total = sum(times)
        value = done / total
        sample = f"{done // a.steps} reads per step ({procs} concurrent `gnumap -c 1` processes, index pre-built, start-up {rr.load_s:.2f}s subtracted)"
        line = {
            "metric": "reads/sec (probabilistic-NW mapping)", "value": value, "unit": "reads/s", "impl": "reference",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.genome, a.reads, a.read_len, a.mode, main_tag(a)), "sample": sample, "mapped_fraction": mapped / max(done, 1)},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": procs, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
    finally:
        rr.close()


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="gmx_clocks_", suffix=".csv")
        self.f = open(self.path, "w")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.close()
        sm, mx, reasons = [], 0.0, set()
        try:
            for ln in open(self.path):
                parts = [x.strip() for x in ln.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx = max(mx, float(parts[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        busy = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


class Job:
    """One rank's context, batches and timing loops for one workload."""

    def __init__(self, ix, mode, local, world, seq_h, qual_h, L, name):
        import torch
        from gnumap_b200 import _abi, api, sharding
        self.torch = torch; self.api = api; self._abi = _abi
        self.world = world; self.mode = mode; self.L = L; self.name = name
        self.dev = torch.device("cuda", local)
        params = api.default_params()
        params.mode = MODES[mode]
        if mode != "normal":
            params.gen_size = 1
        if mode == "bs":
            params.align_scores[ord("c")][3] = params.align_scores[ord("a")][0]          # reference src/Driver.cpp:1266
        t = time.time()
        self.m = m = api.Mapper(ix, params, device=local)
        m.synchronize()
        log(f"[bench] {name}: context created (index upload + SA de-sampling) in {time.time() - t:.1f}s")
        m.set_option(api.OPT_COLLECT_HITS, 0)
        if os.environ.get("GMX_CHUNK_READS"):
            m.set_option(api.OPT_CHUNK_READS, int(os.environ["GMX_CHUNK_READS"]))
        if os.environ.get("GMX_FILTER_SHIFT"):
            m.set_option(api.OPT_FILTER_SHIFT, int(os.environ["GMX_FILTER_SHIFT"]))
        if os.environ.get("GMX_FASTQ_PIECE"):
            m.set_option(api.OPT_FASTQ_PIECE, int(os.environ["GMX_FASTQ_PIECE"]))
        if os.environ.get("GMX_OPTIMISTIC"):
            m.set_option(api.OPT_OPTIMISTIC, int(os.environ["GMX_OPTIMISTIC"]))
        if os.environ.get("GMX_VOTE_COMPACT"):
            m.set_option(api.OPT_VOTE_COMPACT, int(os.environ["GMX_VOTE_COMPACT"]))
        if os.environ.get("GMX_VOTE_SLOTS"):
            m.set_option(api.OPT_VOTE_SLOTS, int(os.environ["GMX_VOTE_SLOTS"]))
        self.stream = torch.cuda.Stream(device=self.dev)
        m.set_stream(self.stream.cuda_stream)
        n = self.n = seq_h.numel() // L
        self.seq_h, self.qual_h = seq_h, qual_h                                           # pinned host
        self.off_h = torch.from_numpy(np.arange(n + 1, dtype=np.int64) * L).pin_memory()
        self.res_h = torch.zeros(max(n, 1) * _abi.READ_RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        self.res_np = self.res_h.numpy().view(_abi.READ_RESULT_DTYPE)[:n]
        self.host_batch = self._batch(self.off_h, seq_h, qual_h, 0)
        self.dev_batch = self._batch(self.off_h.to(self.dev), seq_h.to(self.dev), qual_h.to(self.dev), 1)
        self.acc = sharding.device_accumulators(m, self.dev)
        self.acc_bytes = int(sum(t.numel() for t in self.acc) * 4)
        self.reduce_ms = 0.0

    def _batch(self, off, seq, qual, on_device):
        class B:
            pass
        b = B()
        s = self._abi.GmxReads()
        s.n_reads = self.n; s.offsets = off.data_ptr(); s.seq = seq.data_ptr(); s.qual = qual.data_ptr(); s.pwm = None
        s.on_device = on_device; s.max_len = self.L
        b.struct = s; b.n_reads = self.n; b.keep = (off, seq, qual)
        return b

    def step(self, batch):
        self.m.process_batch(batch, fetch=False, results=self.res_np)

    def small_calls(self, reads_per_call, calls):
        """The reference's natural call size: its worker threads hand over 2048 reads each per iteration (src/Driver.cpp:2339),
        2048 x threads reads per round.  `calls` gmx_process_batch calls of `reads_per_call` reads from pinned host memory,
        wall clock."""
        torch = self.torch
        n = min(reads_per_call, self.n)
        calls = max(1, min(calls, self.n // n))
        batches = []
        for c in range(calls):
            class B:
                pass
            b = B()
            s = self._abi.GmxReads()
            s.n_reads = n; s.offsets = self.off_h.data_ptr() + 8 * c * n; s.seq = self.seq_h.data_ptr(); s.qual = self.qual_h.data_ptr(); s.pwm = None
            s.on_device = 0; s.max_len = self.L
            b.struct = s; b.n_reads = n
            batches.append((b, self.res_np[c * n:(c + 1) * n]))
        def run():
            for b, r in batches[:2]:
                self.m.process_batch(b, fetch=False, results=r)
            torch.cuda.synchronize()
            t0 = time.time()
            for b, r in batches:
                self.m.process_batch(b, fetch=False, results=r)
            torch.cuda.synchronize()
            return time.time() - t0
        dt_timed = run()
        self.m.set_option(self.api.OPT_STAGE_TIMING, 0)         # what a caller that does not read gmx_get_stage_stats sets
        dt = run()
        self.m.set_option(self.api.OPT_STAGE_TIMING, 1)
        return {"reads_per_call": n, "calls": calls, "value": n * calls / dt, "unit": "reads/s", "ms_per_call": 1e3 * dt / calls,
                "with_stage_timing": {"value": n * calls / dt_timed, "ms_per_call": 1e3 * dt_timed / calls},
                "what": "gmx_process_batch from pinned host buffers at the reference's call granularity (2048 reads x 16 worker threads per round), "
                        "wall clock, GMX_OPT_STAGE_TIMING 0 (the per-stage CUDA events off, as the reference-side binding sets it)"}

    def reduce(self, timed=True):
        """The path's one collective (MPI Allreduce / Reduce of the accumulators, reference src/Driver.cpp:1615-1811)."""
        if self.world <= 1:
            return
        from gnumap_b200 import sharding
        torch = self.torch
        with torch.cuda.stream(self.stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            sharding.all_reduce_accumulators(self.acc)
            e1.record(self.stream)
        if timed:
            self._reduce_events.append((e0, e1))

    def timed(self, batch, steps, reduce_every_step):
        import torch.distributed as dist
        torch = self.torch
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        self._reduce_events = []
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        st = {"ms": {}, "units": {}, "bytes": {}, "launches": {}}
        e0.record(self.stream)
        t0 = time.time()
        for _ in range(steps):
            self.step(batch)
            for k, v in self.m.stage_stats().items():
                for f in st:
                    st[f][k] = st[f].get(k, 0) + v[f]
            if reduce_every_step:
                self.reduce()
        if not reduce_every_step:
            self.reduce()
        e1.record(self.stream)
        torch.cuda.synchronize()
        wall_ms = (time.time() - t0) * 1e3
        if self.world > 1:
            dist.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        red = sum(a.elapsed_time(b) for a, b in self._reduce_events)
        self.last_ranks = None
        if self.world > 1:
            # what every rank measured on its own: the line's figures are the maxima, this is where they come from
            mine = torch.tensor([ms, red, sum(st["ms"].values())], device=self.dev, dtype=torch.float64)
            every = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine)
            self.last_ranks = [{"rank": r, "ms_per_step": float(t[0]) / steps, "of_which_inside_the_collective_ms_per_step": float(t[1]) / steps,
                                "stage_ms_per_step": float(t[2]) / steps} for r, t in enumerate(every)]
            tt = torch.tensor([ms, wall_ms, red], device=self.dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms, wall_ms, red = (float(x) for x in tt.tolist())
        return ms, wall_ms, st, red

    def reduce_check(self):
        """Sum of the all-reduced accumulators == sum over ranks of the local sums == aligned-base mass of the mapped reads."""
        import torch.distributed as dist
        torch = self.torch
        self.m.reset_accumulators()
        self.step(self.dev_batch)
        torch.cuda.synchronize()
        local = torch.stack([t.sum(dtype=torch.float64) for t in self.acc])
        mapped = self.res_np["status"] == 0
        mass = torch.tensor([float(self.res_np["best_aligned_len"][mapped].sum())], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(local); dist.all_reduce(mass)
        self.reduce(timed=False)
        torch.cuda.synchronize()
        after = torch.stack([t.sum(dtype=torch.float64) for t in self.acc])
        rel = float(((after - local).abs() / local.abs().clamp_min(1e-30)).max())
        amount_mass = float(after[0]); want = float(mass[0])
        return {"sum_after_reduce": [float(x) for x in after.tolist()], "sum_of_local_sums": [float(x) for x in local.tolist()],
                "max_rel_diff": rel, "ok": bool(rel < 1e-5 and abs(amount_mass - want) <= 3e-3 * want),
                "amount_mass": amount_mass, "aligned_bases_of_mapped_reads": want}

    def close(self):
        self.m.close()


def roofline_of(st, steps, peaks, alu):
    """Dominant stage by CUDA-event time + the per-kernel table."""
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kstages = [k for k in st["ms"] if k not in ("upload", "download") and st["launches"].get(k, 0) > 0]
    top = max(kstages, key=lambda k: st["ms"][k])
    kernels = {}
    for k in ("seed_walk", "locate_vote", "scatter", "prep_reads"):
        if st["ms"].get(k, 0) > 0 and st["bytes"].get(k, 0) > 0:
            gbs = st["bytes"][k] / (st["ms"][k] * 1e-3) / 1e9
            kernels[k] = {"bound": "hbm", "ms_per_step": st["ms"][k] / steps, "algorithmic_GB_per_step": st["bytes"][k] / steps / 1e9,
                          "achieved_GBs": gbs, "frac_of_hbm_peak": gbs / peak}
    for k in ("nw_score", "nw_traceback", "pair_hmm"):
        if st["ms"].get(k, 0) > 0 and st["units"].get(k, 0) > 0:
            gc = st["units"][k] / (st["ms"][k] * 1e-3) / 1e9
            kernels[k] = {"bound": "fp64 pipe" if k == "pair_hmm" else "fp32 alu (no fma)", "ms_per_step": st["ms"][k] / steps, "GCUPS": gc}
            if k == "pair_hmm" and alu.get("fp64_mul_add"):
                # 20 FP64 operations per cell: 8 forward, 12 backward + posterior (pair_hmm.cuh)
                kernels[k]["tera_fp64_ops"] = gc * 20 / 1e3
                kernels[k]["frac_of_measured_fp64_peak"] = gc * 20 / 1e3 / alu["fp64_mul_add"]
            elif k != "pair_hmm" and alu.get("fp32_mul_add_nofma"):
                # 12 FP32 operations per NW cell (4 mul + 3 add for the substitution value, 3 adds, 2 max)
                kernels[k]["frac_of_measured_fp32_nofma_peak"] = gc * 12 / 1e3 / alu["fp32_mul_add_nofma"]
    if top in ("nw_score", "nw_traceback", "pair_hmm"):
        unit = "TFLOP/s"
        achieved = kernels[top].get("tera_fp64_ops") if top == "pair_hmm" else kernels[top]["GCUPS"] * 12 / 1e3
        pk = alu.get("fp64_mul_add") if top == "pair_hmm" else alu.get("fp32_mul_add_nofma")
        roof = {"kernel": top, "bound": "fp64 pipe (non-tensor)" if top == "pair_hmm" else "fp32 alu (non-tensor)", "achieved": achieved, "peak": pk, "unit": unit,
                "frac": (achieved / pk) if (achieved and pk) else None, "traffic": None,
                "peak_source": "measured here by gmx_measure_alu_peak (mul + add without FMA)"}
    else:
        achieved = st["bytes"][top] / (st["ms"][top] * 1e-3) / 1e9 if st["ms"][top] > 0 else 0.0
        roof = {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)"}
    roof.update({"ms_per_step": st["ms"][top] / steps, "launches_per_step": st["launches"][top] / steps,
                 "algorithmic_bytes_per_step": st["bytes"][top] / steps,
                 "stages_ms_per_step": {k: round(v / steps, 3) for k, v in st["ms"].items()},
                 "stages_units_per_step": {k: v // steps for k, v in st["units"].items()}, "kernels": kernels})
    return roof


def strong_reads(ix_codes_dev, total, L, seed, mode, rank, world, dev):
    """This rank's share of ONE read set of `total` reads, dealt in 2048-read slices round robin; generated on the GPU in
    blocks that any rank reproduces on its own.  Returns pinned host tensors (seq, qual) of the share."""
    import torch
    from gnumap_b200 import synth
    block = 131072
    seqs, quals = [], []
    for b0 in range(0, total, block):
        n = min(block, total - b0)
        r = synth.simulate_reads_torch(ix_codes_dev, n, L, seed, bisulfite=0.95 if mode == "bs" else 0.0, block_reads=block, first_block=b0 // block)
        idx = torch.arange(b0, b0 + n, device=dev)
        keep = ((idx // SLICE) % world) == rank
        seqs.append(r["seq"][keep].cpu()); quals.append(r["qual"][keep].cpu())
    seq = torch.cat(seqs).reshape(-1).contiguous().pin_memory(); qual = torch.cat(quals).reshape(-1).contiguous().pin_memory()
    return seq, qual


EXTRA = [
    # name, genome, genome seed, read length, total reads, reads seed, mode, reduce inside every step
    ("cfg2_normal_156Mb_10Mx150bp_strong", 156_000_000, 156, 150, 10_000_000, 157, "normal", False),
    ("cfg3_snp_156Mb_2Mx150bp_strong", 156_000_000, 156, 150, 2_000_000, 157, "snp", True),
    ("cfg4_bs_100Mb_5Mx100bp_strong", 100_000_000, 500, 100, 5_000_000, 501, "bs", True),
]


def extra_row(spec, a, rank, world, local, dev, peaks, alu):
    import torch
    import torch.distributed as dist
    from gnumap_b200 import _abi
    name, genome, gseed, L, total, rseed, mode, reduce_each = spec
    total = int(total * a.extra_scale) // SLICE * SLICE
    if rank == 0:
        ix, _ = get_index(genome, gseed, dev)
    if world > 1:
        dist.barrier()
    if rank != 0:
        ix, _ = get_index(genome, gseed, dev)
    t = time.time()
    codes_dev = torch.from_numpy(ix.codes()).to(dev)
    seq_h, qual_h = strong_reads(codes_dev, total, L, rseed, mode, rank, world, dev)
    del codes_dev
    torch.cuda.empty_cache()
    log(f"[bench] {name}: {seq_h.numel() // L} of {total} reads on rank {rank} in {time.time() - t:.1f}s")
    job = Job(ix, mode, local, world, seq_h, qual_h, L, name)
    steps, warm = max(2, min(a.steps, 3)), 2
    for _ in range(warm):
        job.step(job.dev_batch)
    job.step(job.host_batch)
    job.reduce(timed=False)
    job.m.reset_accumulators()
    ms_dev, _, st, red_dev = job.timed(job.dev_batch, steps, reduce_each)
    ms_e2e, wall_e2e, _, _ = job.timed(job.host_batch, steps, reduce_each)
    mapped = torch.tensor([int((job.res_np["status"] == 0).sum())], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(mapped)
    check = job.reduce_check()
    n_red = steps if reduce_each else 1
    row = None
    if rank == 0:
        roof = roofline_of(st, steps, peaks, alu)
        red_ms = red_dev / n_red if world > 1 else 0.0
        row = {"workload": workload_name(genome, total, L, mode, "BASELINE " + name.split("_")[0].replace("cfg", "configs[") + "]"),
               "scaling": "strong", "reads_per_step_all_gpus": total, "n_gpus": world, "steps": steps, "warmup": warm + 1,
               "value": total * steps / (ms_dev * 1e-3), "unit": "reads/s", "ms_per_step": ms_dev / steps,
               "e2e": {"value": total * steps / (wall_e2e * 1e-3), "unit": "reads/s", "ms_per_step": wall_e2e / steps, "clock": "wall, max over ranks",
                       "h2d_bytes_per_step": int(2 * total * L + 8 * (total + world)), "d2h_bytes_per_step": int(total * (_abi.READ_RESULT_DTYPE.itemsize + 64))},
               "dealing": f"{SLICE}-read slices round robin over the ranks (reference inc/SeqManager.h:329-345)",
               "l2": f"inputs exceed L2: reads {2 * total * L // world // 1_000_000} MB per rank per step, suffix array {4 * genome // 1_000_000} MB",
               "mapped_fraction": int(mapped.item()) / total,
               "collective": (f"ncclAllReduce(sum, f32) of {job.acc_bytes / 1e6:.0f} MB per GPU, " + ("inside every timed step" if reduce_each else "once after the last step, inside the timed region")) if world > 1 else "none (1 GPU)",
               "reduce_ms": red_ms, "reduce_busbw_GBs": (2 * (world - 1) / world * job.acc_bytes / (red_ms * 1e-3) / 1e9) if red_ms > 0 else None,
               "reduce_check": check, "gpu_launches": int(sum(st["launches"].values())), "roofline": roof}
        stages = roof["stages_ms_per_step"]
        lim = max((k for k in stages if k not in ("upload", "download")), key=lambda k: stages[k])
        row["limiter"] = f"{lim} ({stages[lim]:.1f} of {ms_dev / steps:.1f} ms per step)" + (f"; reduce {red_ms:.1f} ms" if world > 1 and reduce_each else "")
    job.close()
    del job
    torch.cuda.empty_cache()
    return row


def own_arm(a):
    import torch
    import torch.distributed as dist
    from gnumap_b200 import _abi, api

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gnumap_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # index: rank 0 builds and caches it, the others load the cache
    if rank == 0:
        ix, prefix = get_index(a.genome, a.genome_seed, dev)
    if world > 1:
        dist.barrier()
    if rank != 0:
        ix, prefix = get_index(a.genome, a.genome_seed, dev)
    reads = get_reads_numpy(a, ix, rank)
    n, L = a.reads, a.read_len
    seq_h = torch.from_numpy(np.frombuffer(b"ACGTN", dtype=np.uint8)[reads["bases"]].reshape(-1).copy()).pin_memory()
    qual_h = torch.from_numpy((reads["quals"].astype(np.uint8) + 33).reshape(-1).copy()).pin_memory()
    job = Job(ix, a.mode, local, world, seq_h, qual_h, L, "main")
    m, res_np, res_h = job.m, job.res_np, job.res_h
    alu = {}
    if rank == 0:
        alu = {"fp32_mul_add_nofma": m.alu_peak(0), "fp32_add_max": m.alu_peak(1), "fp64_mul_add": m.alu_peak(2), "unit": "tera lane-operations / s",
               "how": "gmx_measure_alu_peak: 8 independent chains per thread, 16 CTAs x 256 threads per SM, best of 5"}

    # nvidia-smi needs a few hundred ms to deliver its first sample: start it before the warm-up (same kernels, same
    # load) and keep it running through the timed region
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(a.warmup):
        job.step(job.dev_batch)
    for _ in range(max(a.warmup // 2, 1)):
        job.step(job.host_batch)
    job.reduce(timed=False)             # warm NCCL up too
    m.reset_accumulators()

    ms_dev, wall_dev, st, red_ms = job.timed(job.dev_batch, a.steps, False)
    ranks_dev = job.last_ranks
    clocks = sampler.stop() if sampler else None
    ms_e2e, wall_e2e, _, _ = job.timed(job.host_batch, a.steps, False)
    mapped = int((res_np["status"] == 0).sum())
    small = job.small_calls(SLICE * 16, 24) if rank == 0 else None

    # next row: the same step fed with FASTQ TEXT (pinned host buffer): H2D of the raw text, device record
    # indexer, reads used in place, D2H of the results and of the record index
    fastq = None
    txt = None
    if not a.no_fastq:
        name_w = 9
        rec_len = 1 + name_w + 1 + L + 3 + L + 1
        txt = np.empty((n, rec_len), dtype=np.uint8)
        txt[:, 0] = ord("@")
        ids = np.arange(n, dtype=np.int64)
        for d in range(name_w):
            txt[:, name_w - d] = ord("0") + (ids // 10 ** d) % 10
        txt[:, 1 + name_w] = 10
        txt[:, 2 + name_w: 2 + name_w + L] = np.frombuffer(b"ACGTN", dtype=np.uint8)[reads["bases"]]
        txt[:, 2 + name_w + L: 5 + name_w + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
        txt[:, 5 + name_w + L: 5 + name_w + 2 * L] = reads["quals"].astype(np.uint8) + 33
        txt[:, -1] = 10
        text_h = torch.from_numpy(txt.reshape(-1)).pin_memory()
        recs_h = torch.zeros(n * _abi.FASTQ_REC_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        n_out = C.c_int64(0)

        def fq_step():
            rc = m.L.gmx_process_fastq(m._ctx, text_h.data_ptr(), text_h.numel(), 0, res_h.data_ptr(), n, C.byref(n_out), recs_h.data_ptr())
            if rc != 0:
                raise RuntimeError(f"gmx_process_fastq: {rc} {m.L.gmx_last_error(m._ctx).decode()}")

        fq_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(a.steps):
            fq_step()
        torch.cuda.synchronize()
        fq_ms = (time.time() - t0) * 1e3
        fq_stages = {k: round(v["ms"], 3) for k, v in m.stage_stats().items()}          # of the last call
        if world > 1:
            tt = torch.tensor([fq_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            fq_ms = float(tt.item())
        assert n_out.value == n and int((res_np["status"] == 0).sum()) == mapped
        # SAM row: native formatting of the batch just scored
        sam_cap = n * (2 * L + 96)
        sam_buf = torch.empty(sam_cap, dtype=torch.uint8).pin_memory()
        names_c = (C.c_char_p * len(ix.names))(*[nm.encode() for nm in ix.names])
        sam_len = C.c_int64(0)
        sam_s = None
        for _ in range(3):                      # first call allocates the formatter's device buffers
            t0 = time.time()
            rc = m.L.gmx_format_sam(m._ctx, text_h.data_ptr(), recs_h.data_ptr(), res_h.data_ptr(), n, names_c, sam_buf.data_ptr(), sam_cap, C.byref(sam_len))
            dt = time.time() - t0
            sam_s = dt if sam_s is None else min(sam_s, dt)
        if rc != 0:
            raise RuntimeError(f"gmx_format_sam: {rc} {m.L.gmx_last_error(m._ctx).decode()}")
        # accumulator output row: .sgr text of the accumulators as they stand (device row selection + host text)
        sgr_cap = 64 << 20
        sgr_buf = np.empty(sgr_cap, dtype=np.uint8)
        sgr_len = C.c_int64(0)
        sgr_rows = None; sgr_s = None
        if a.mode == "normal":
            t0 = time.time()
            rc = m.L.gmx_format_sgr(m._ctx, names_c, 0.001, sgr_buf.ctypes.data, sgr_cap, C.byref(sgr_len))
            if rc == _abi.GMX_ERR_OVERFLOW:
                sgr_cap = int(sgr_len.value) + 16; sgr_buf = np.empty(sgr_cap, dtype=np.uint8)
                t0 = time.time()
                rc = m.L.gmx_format_sgr(m._ctx, names_c, 0.001, sgr_buf.ctypes.data, sgr_cap, C.byref(sgr_len))
            sgr_s = time.time() - t0
            if rc != 0:
                raise RuntimeError(f"gmx_format_sgr: {rc} {m.L.gmx_last_error(m._ctx).decode()}")
            sgr_rows = int((sgr_buf[: sgr_len.value] == 10).sum())
        fastq = {"value": n * world * a.steps / (fq_ms * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": int(text_h.numel()),
                 "d2h_bytes_per_step": int(n * (_abi.READ_RESULT_DTYPE.itemsize + 64 + _abi.FASTQ_REC_DTYPE.itemsize)),
                 "what": "gmx_process_fastq: FASTQ text in pinned host memory -> device record indexer -> reads used in place -> results (wall clock)",
                 "ms_per_step": fq_ms / a.steps, "stages_ms_last_step": fq_stages,
                 "sam": {"value": n / sam_s, "unit": "reads/s", "bytes": int(sam_len.value),
                         "what": "gmx_format_sam: SAM body of the batch (ScoredSeq::get_SAM + writer) formatted on the GPU from the resident text / results / CIGARs, "
                                 "finished text copied to pinned host memory (wall clock, best of 3)"},
                 "sgr": {"seconds": sgr_s, "rows": sgr_rows, "bytes": int(sgr_len.value), "bins_scanned": int(job.acc[0].numel()),
                         "what": "gmx_format_sgr: GenomeBwt::PrintFinalSGR of the accumulators (device scan + select, host text)"}}

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        roofline = roofline_of(st, a.steps, peaks, alu)
        try:   # DRAM bytes per SA hit of the vote kernel from the committed ncu --set full capture, scaled to this launch
            tj = json.load(open(os.path.join(ROOT, "profiles", "vote_kernel_traffic.json")))
            if roofline["kernel"] == "locate_vote":
                chunks = max(st["launches"].get("classify", 0), 1)          # one k_classify per chunk
                roofline["traffic"] = tj["dram_bytes_per_sa_hit"] * st["units"]["locate_vote"] / chunks
                roofline["algorithmic_bytes_per_loaded_launch"] = st["bytes"]["locate_vote"] / chunks
                roofline["note"] = ("stage = one launch per non-empty task class and chunk (6 filter + 6 exact classes exist; a uniform workload fills one); "
                                    "achieved / traffic are per chunk")
        except (OSError, KeyError, ValueError):
            pass
        nw_cells = st["units"].get("nw_score", 0)
        gcups = nw_cells / (st["ms"]["nw_score"] * 1e-3) / 1e9 if st["ms"].get("nw_score", 0) > 0 else None
        cpu = None
        parity = None
        if not a.no_cpu:
            try:
                rr = ReferenceRunner(a.mode, prefix, reads)
                try:
                    t, _, nn, procs, _ = rr.run(max(rr.cores * 4, 32))
                    rate = nn / t
                    nn2 = min(max(int(rate * float(os.environ.get("GMX_CPU_BASELINE_S", "20"))), rr.cores), a.reads)
                    t, mp, nn2, procs, sams = rr.run(nn2, keep_sam=True)
                    cpu = {"value": nn2 / t, "unit": "reads/s", "cores": procs, "kind": "reference",
                           "sample": f"{nn2} reads of the same workload ({procs} concurrent `gnumap -c 1` processes of the unmodified reference, "
                                     f"index pre-built, start-up {rr.load_s:.2f}s subtracted), mapped fraction {mp / max(nn2, 1):.3f}"}
                    # parity on that very sample: the GPU's SAM records for the same reads against the reference's
                    same, n_rec, n_reads_cmp, first_diff = True, 0, 0, None
                    for lo, hi, sam_path in sams:
                        fq = os.path.join(rr.dir, "parity.fq")
                        rr._fastq(fq, lo, hi)
                        text = open(fq, "rb").read()
                        _, out = m.process_fastq(text, fetch=False)
                        got = sorted(s for s in m.format_sam(text, api.fastq_scan_host(text), out["results"]).decode().split("\n") if s)
                        want = sorted(ln.rstrip("\n") for ln in open(sam_path) if not ln.startswith("@"))
                        n_rec += len(want); n_reads_cmp += hi - lo
                        if got != want:
                            same = False
                            first_diff = first_diff or next(((g, w) for g, w in zip(got, want) if g != w), (len(got), len(want)))
                    parity = {"reads": n_reads_cmp, "sam_records": n_rec, "sam_identical": same,
                              "what": "SAM body (flag, position, MAPQ, CIGAR, XA, XP, X0 of every record) of the cpu_baseline sample: gmx_process_fastq + gmx_format_sam vs the unmodified reference binary"}
                    if first_diff:
                        parity["first_difference"] = str(first_diff)[:600]
                    try:       # the reference's own multithreaded mode, one process
                        v, nr = rr.run_threads(rr.cores)
                        cpu["single_process_c_n"] = {"value": v, "unit": "reads/s", "threads": rr.cores, "reads": nr,
                                                     "what": f"one `gnumap -c {rr.cores}` process on 2048 x {rr.cores} reads (every thread gets one slice)"}
                    except Exception as e:
                        cpu["single_process_c_n"] = {"value": None, "what": f"unavailable: {e}"}
                    try:       # the same program with its hot path on the GPU (the drop-in): bounded by the reference's own reader / writer
                        cpu["e2e_dropin"] = rr.run_patched(rr.cores, min(a.reads, 1 << 19)) or {"value": None, "what": "oracle/_ref/gnumap_gmx was not built"}
                    except Exception as e:
                        cpu["e2e_dropin"] = {"value": None, "what": f"unavailable: {e}"}
                finally:
                    rr.close()
            except Exception as e:  # the baseline is reported, never required
                cpu = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        if fastq is not None and not a.no_cpu:
            try:      # the reference's own FASTQ reader (SeqReader, one thread) on a sample of the same text, in its own process
                ns = min(n, 200_000)
                fn = os.path.join(CACHE, f"sample_{os.getpid()}.fq")
                with open(fn, "wb") as f:
                    f.write(txt[:ns].tobytes())
                code = ("import ctypes as C, sys, time\nfrom oracle import oracle as O\nL = C.CDLL(O.REF_PROBE)\n"
                        f"buf = C.create_string_buffer({ns * (2 * L + 32)})\nt0 = time.time()\ngot = L.refp_read_fastq({fn!r}.encode(), buf, len(buf))\n"
                        "print(got, time.time() - t0)\n")
                outp = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True, timeout=300)
                os.unlink(fn)
                got, dt = outp.split()[-2:]
                fastq["cpu_reader"] = {"value": int(got) / float(dt), "unit": "reads/s", "cores": 1, "kind": "reference",
                                       "sample": f"{ns} reads of the same text through the reference's SeqReader::get_more_fastq (PWM construction included), separate process"}
            except Exception as e:
                fastq["cpu_reader"] = {"value": None, "sample": f"unavailable: {e}"}
        total_reads = n * world * a.steps
        line = {
            "metric": "reads/sec (probabilistic-NW mapping)", "value": total_reads / (ms_dev * 1e-3), "unit": "reads/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_dev / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.genome, n, L, a.mode, main_tag(a)), "reads_per_gpu_per_step": n,
                       "l2": f"inputs exceed L2: reads {2 * n * L // 1_000_000} MB + suffix array {4 * a.genome // 1_000_000} MB streamed per step (L2 126 MB)",
                       "collective": (f"one ncclAllReduce(sum, f32) of {job.acc_bytes / 1e6:.0f} MB per GPU after the last step, inside the timed region ({red_ms:.2f} ms)") if world > 1 else "none (1 GPU)",
                       "mapped_fraction": mapped / n, "nw_gcups": gcups},
            "clocks": clocks,
            "e2e": {"value": total_reads / (wall_e2e * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": int(2 * n * L + 8 * (n + 1)),
                    "d2h_bytes_per_step": int(n * (_abi.READ_RESULT_DTYPE.itemsize + 64)), "ms_per_step": wall_e2e / a.steps,
                    "clock": "wall, max over ranks", "cuda_event_ms_per_step": ms_e2e / a.steps},
            "gpu_launches": int(sum(st["launches"].values())),
            "roofline": roofline,
            "ranks": ranks_dev,
            "small_calls": small,
            "chunks": dict(zip(("issued_without_host_wait", "run_again"), job.m.chunk_stats())),
            "alu_peaks": alu,
            "cpu_baseline": cpu,
            "parity_sample": parity,
            "fastq_row": fastq,
        }
    job.close()
    del job, seq_h, qual_h
    torch.cuda.empty_cache()
    if not a.no_extra:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        rows = {}
        for spec in EXTRA:
            try:
                row = extra_row(spec, a, rank, world, local, dev, peaks, alu)
            except Exception as e:          # a row is reported, never allowed to take the headline down
                if world > 1:
                    raise
                row = {"error": f"{type(e).__name__}: {e}"}
            if rank == 0:
                rows[spec[0]] = row
        if rank == 0:
            line["configs"] = rows
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--mode", default="normal", choices=list(MODES))
    ap.add_argument("--genome-seed", type=int, default=100)
    ap.add_argument("--reads-seed", type=int, default=101)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity_sample leg")
    ap.add_argument("--no-fastq", action="store_true", help="skip the FASTQ-text leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the rows of BASELINE configs[2..4]")
    ap.add_argument("--extra-scale", type=float, default=1.0, help="scale the read counts of the configs[2..4] rows")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not (a.impl == "ours" and world == 1 and a.gpus > 1):
        claim_stdout()
    if a.impl == "reference":
        return reference_arm(a)
    if world == 1 and a.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29517", os.path.abspath(__file__), *sys.argv[1:]]
        raise SystemExit(subprocess.call(cmd))
    import __graft_entry__ as g
    g.build()                       # serialised by a file lock: a no-op on every rank but the first when something is stale
    own_arm(a)


if __name__ == "__main__":
    main()
