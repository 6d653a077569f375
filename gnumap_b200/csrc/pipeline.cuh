// pipeline.cuh -- kernels of the batched mapping pipeline (PHASE A + PHASE B of the reference's
// per-read loops, reference src/Driver.cpp:2344-2373) for one chunk of reads.
//
//   k_prep_reads      set_top_matches prologue           reference src/Driver.cpp:446-497
//   k_seed_walk       align_sequence k-mer walk + K1     reference inc/align_seq2_raw.cpp:192-243
//   k_classify        tasks by SA-hit count -> filter / exact class lists
//   k_vote_filter     K1b locate + K1c diagonal vote     reference inc/align_seq2_raw.cpp:262-274,28-35
//                     (Bloom / counting filter + exact verification against the packed genome: the fast path)
//   k_vote_smem/_gmem same, exact hash tables (kmin == 1, very long reads, > 16 k hits per task)
//   k_cand_score      GetString + K2a                    reference inc/align_seq2_raw.cpp:43-64
//   k_cand_ranges     candidate range of every read in the sorted key list
//   k_finalize_reads  acceptance, grouping, denominator  reference inc/align_seq2_raw.cpp:95-165,
//                     best group                          reference src/Driver.cpp:593-611,640-680
//   k_traceback       K2b per group                      reference src/NormalScoredSeq.cpp:40-62
//   k_scatter         K3                                 reference src/NormalScoredSeq.cpp:68-75 etc.
//   k_gather_best / k_gather_multi   fixed-size per-read records, positions of multi-position best groups
#pragma once

#include "fm_index.cuh"
#include "nw.cuh"

struct DevParams {
    float gap, align_score, cutoff;
    int max_gap, mer, jump, kmin, perc, match_pos, match_neg, unique_only, fast, mode;
    uint32_t max_kmer_hits, max_matches, gen_size;
};

// per-read scratch produced by k_prep_reads
struct ReadPrep {
    double min_align;     // min_align_score
    float  max_align;     // self score
    int32_t status;       // GMX_READ_* (MAPPED here means "eligible", decided later)
};

__device__ __forceinline__ ReadView gmx_read_view(const DevReads &R, int r, int neg)
{
    ReadView v;
    int64_t off = R.offsets[r];
    v.n = gmx_read_len(R, r);
    v.seq = R.seq + off;
    v.qual = R.qual ? R.qual + (R.qoffsets ? R.qoffsets[r] : off) : nullptr;
    v.pwm = R.pwm ? R.pwm + 4 * off : nullptr;
    v.neg = neg;
    return v;
}

// ---- a3 + status ------------------------------------------------------------------------------
// One thread per read; the score is a left-to-right FP32 sum over the bases, so the read is walked sequentially.
// Byte loads with a stride of one read length between lanes would cost 32 L1 wavefronts each, so the block first
// stages its (contiguous) reads in shared memory with coalesced word loads; the per-base term
// get_val(pwm(base, Q), seq[i]) is a pure function of the two characters and comes from a 96 KB table.
#define GMX_PREP_THREADS 128
#define GMX_PREP_STAGE_BYTES (GMX_PREP_THREADS * 176)        // reads of up to 176 bases on average per block
__global__ void __launch_bounds__(GMX_PREP_THREADS) k_prep_reads(DevReads R, DevTables T, DevParams P, ReadPrep *prep, unsigned long long *bad_len)
{
    __shared__ __align__(16) uint8_t s_buf[2 * GMX_PREP_STAGE_BYTES + 16];
    uint8_t *s_seq = s_buf, *s_qual = s_buf + GMX_PREP_STAGE_BYTES + 8;
    const int r0 = blockIdx.x * blockDim.x;
    const int r = r0 + threadIdx.x;
    const int r1 = min(r0 + (int)blockDim.x, R.n_reads);
    // staged = 1, contiguous layout (no raw PWM, no in-place FASTQ offsets): the block's reads are one byte range of seq and
    // one of qual.  staged = 2, reads used in place in a FASTQ text (seq == qual == the text, per-read offsets of the two
    // lines): the block's records are one byte range of the text, staged whole (names and '+' lines included)
    int staged = 0;
    int64_t lead = 0, t0 = 0, t1 = 0;
    if (!R.pwm && !R.qoffsets && !R.lens && R.qual) {
        const int64_t b0 = R.offsets[r0], b1 = R.offsets[r1];
        const uintptr_t a_seq = (uintptr_t)(R.seq + b0), a_qual = (uintptr_t)(R.qual + b0);
        lead = (int64_t)(a_seq & 3u);
        if ((a_qual & 3u) == (uintptr_t)lead && b1 - b0 + lead <= GMX_PREP_STAGE_BYTES) {
            staged = 1;
            const uint32_t *g_seq = reinterpret_cast<const uint32_t *>(a_seq - lead), *g_qual = reinterpret_cast<const uint32_t *>(a_qual - lead);
            const int words = (int)((b1 - b0 + lead + 3) >> 2);
            for (int w = threadIdx.x; w < words; w += blockDim.x) {
                reinterpret_cast<uint32_t *>(s_seq)[w] = g_seq[w];
                reinterpret_cast<uint32_t *>(s_qual)[w] = g_qual[w];
            }
        }
        lead -= b0;                                        // shared index of genome-order byte x is x + lead
    } else if (!R.pwm && R.qoffsets && R.lens && R.qual == R.seq && r1 > r0) {
        t0 = R.offsets[r0];
        t1 = R.qoffsets[r1 - 1] + (int64_t)max(R.lens[r1 - 1], 0);
        const uintptr_t a = (uintptr_t)(R.seq + t0);
        lead = (int64_t)(a & 3u);
        if (t1 > t0 && t1 - t0 + lead <= 2 * GMX_PREP_STAGE_BYTES + 8) {
            staged = 2;
            const uint32_t *g = reinterpret_cast<const uint32_t *>(a - lead);
            const int words = (int)((t1 - t0 + lead + 3) >> 2);
            for (int w = threadIdx.x; w < words; w += blockDim.x) reinterpret_cast<uint32_t *>(s_buf)[w] = g[w];
        }
        lead -= t0;
    }
    __syncthreads();
    if (r >= R.n_reads) return;
    ReadView rd = gmx_read_view(R, r, 0);
    ReadPrep out; out.min_align = 0; out.max_align = 0; out.status = GMX_READ_MAPPED;
    // device-resident input: a read longer than the caller's max_len (or a negative length) would overrun every buffer
    // sized from it -- it takes no part in the batch and the host turns the count into GMX_ERR_INVALID
    if (R.max_len > 0 && (rd.n < 0 || rd.n > R.max_len)) {
        if (bad_len) atomicAdd(bad_len, 1ull);
        out.status = GMX_READ_TOO_SHORT; prep[r] = out; return;
    }
    if ((unsigned)rd.n < (unsigned)P.mer) { out.status = GMX_READ_TOO_SHORT; prep[r] = out; return; }
    // get_align_score(read, consensus, 0, n-1) == get_align_score_mid  (reference src/bin_seq.cpp:860-893)
    float score = 0.f;
    if (staged == 2) {
        // a caller's own record index need not be in text order: a read outside the staged range is walked in place
        const int64_t so = R.offsets[r], qo = R.qoffsets[r];
        if (so < t0 || so + rd.n > t1 || qo < t0 || qo + rd.n > t1) staged = 0;
        else {
            const uint8_t *sq = s_buf + (so + lead), *ql = s_buf + (qo + lead);
            for (int i = 0; i < rd.n; ++i) score = __fadd_rn(score, __ldg(T.self + (int)sq[i] * GMX_NQ + gmx_qidx(ql[i], 0)));
        }
    }
    if (staged == 1) {
        const int64_t at = R.offsets[r] + lead;
        for (int i = 0; i < rd.n; ++i)
            score = __fadd_rn(score, __ldg(T.self + (int)s_seq[at + i] * GMX_NQ + gmx_qidx(s_qual[at + i], 0)));
    } else if (staged == 0) {
        for (int i = 0; i < rd.n; ++i) {
            float4 p = rd.pwm_row(T, i);
            uint8_t ch = rd.seq[i];                  // GetConsensus(): read.seq (reference src/Driver.cpp:352-356)
            const float *s = T.S + 4 * (int)ch;
            float t = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, s[0]), __fmul_rn(p.y, s[1])), __fmul_rn(p.z, s[2])), __fmul_rn(p.w, s[3]));
            score = __fadd_rn(score, t);
        }
    }
    out.max_align = score;
    double max_align = (double)score;
    if (max_align < (double)P.cutoff) { out.status = GMX_READ_TOO_POOR; prep[r] = out; return; }
    out.min_align = P.perc ? __dmul_rn((double)P.align_score, max_align) : (double)P.align_score;
    prep[r] = out;
}

// ---- k-mer walk ---------------------------------------------------------------------------------
// One thread per task = (read, strand).  The walk is sequentially data dependent (the next offset
// depends on whether the previous k-mer hit), so the parallelism is across the 2 x n_reads tasks.
struct SeedStore {
    uint4 *rec;        // [n_tasks][max_seeds] {first SA rank of the interval, interval size, the k-mer itself (low, high word):
                       //   2 bits per base, first base most significant} -- one 16-byte store / load per k-mer
    uint16_t *offset;  // [n_tasks][max_seeds] k-mer offset i in the oriented read
    uint8_t  *n_seeds; // [n_tasks]
    uint32_t *hits;    // [n_tasks] total SA hits of the task
    int max_seeds;
    int64_t n_tasks;
    // one task's seeds are contiguous: the vote kernels read them with the lanes of one warp
    __host__ __device__ __forceinline__ int64_t at(int64_t task, int s) const { return task * max_seeds + s; }
};

// stats[0] += k-mer lookups, stats[1] += backward-search steps (one bwt_2occ each), stats[2] += SA hits
__global__ void k_seed_walk(DevIndex ix, DevReads R, DevParams P, const ReadPrep *prep, SeedStore S, unsigned long long *stats)
{
    int64_t task = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = task < S.n_tasks;
    int r = in_range ? (int)(task >> 1) : 0, neg = (int)(task & 1);
    int ns = 0; uint32_t total = 0, n_lookups = 0, n_steps = 0;
    bool active = in_range && prep[r].status == GMX_READ_MAPPED && (neg ? P.match_neg : P.match_pos);
    if (active && !R.seq) active = false;
    if (active) {
        int64_t off = R.offsets[r];
        int n = gmx_read_len(R, r);
        const uint8_t *seq = R.seq + off;
        // oriented consensus symbol: POS = nt4(seq[x]); NEG = complement of seq[n-1-x]
        // (reverse_comp maps every non-acgt character to 'n', which never matches)
        // bytes come through an 8-byte register buffer: a byte load per base costs a full L1 wavefront per lane
        // (the lanes of a warp sit one read apart), an aligned 8-byte load serves 8 bases of the walk
        unsigned long long buf = 0; uintptr_t buf_at = ~(uintptr_t)0;
        auto sym_at = [&](int x) -> uint32_t {
            const uintptr_t a = (uintptr_t)(seq + (neg ? n - 1 - x : x));
            if ((a >> 3) != buf_at) { buf_at = a >> 3; buf = __ldg(reinterpret_cast<const unsigned long long *>(buf_at << 3)); }
            int c = gmx_nt4((uint8_t)(buf >> ((a & 7u) * 8u)));
            return (uint32_t)((neg && c < 4) ? 3 - c : c);
        };
        // rolling window: the k-mer at `wbase`, 2 bits per base (first base most significant), and its non-ACGT mask.
        // The walk moves by one base after a miss and by `jump` after a hit, so a lookup usually adds <= jump bases.
        const int mer = P.mer;
        const unsigned long long wmask = mer < 32 ? ((1ull << (2 * mer)) - 1ull) : ~0ull, bmask = (1ull << mer) - 1ull;
        unsigned long long wcode = 0, wbad = 0;
        int wbase = -(1 << 20);                                   // nothing loaded
        auto window_at = [&](int base) {
            int d = base - wbase;
            if (d < 0 || d >= mer) { d = mer; wcode = 0; wbad = 0; }
            for (int t = mer - d; t < mer; ++t) {
                const uint32_t c = sym_at(base + t);
                wcode = (wcode << 2) | (unsigned long long)(c & 3u);
                wbad = (wbad << 1) | (unsigned long long)(c > 3u);
            }
            wcode &= wmask; wbad &= bmask;
            wbase = base;
        };
        unsigned last = (unsigned)n - (unsigned)P.mer;
        for (unsigned i = 0; i < last; i += (unsigned)P.jump) {
            unsigned j;
            bool found = false;
            uint64_t k = 0, l = 0;
            for (j = 0; j + i < last; j++) {
                window_at((int)(i + j));
                bool hit = gmx_match_exact_packed(ix, mer, wcode, wbad, k, l, n_steps);
                n_lookups++;
                if (!hit) continue;
                if (P.max_kmer_hits > 0 && l - k + 1 > (uint64_t)P.max_kmer_hits) continue;
                found = true;
                break;
            }
            i += j;
            if (!found) break;
            if (ns < S.max_seeds) {
                S.rec[S.at(task, ns)] = make_uint4((uint32_t)k, (uint32_t)(l - k + 1), (uint32_t)wcode, (uint32_t)(wcode >> 32));   // the window is at i
                S.offset[S.at(task, ns)] = (uint16_t)i;
                total += (uint32_t)(l - k + 1);
                ns++;
            }
            if (P.fast) break;
        }
    }
    if (in_range) { S.n_seeds[task] = (uint8_t)ns; S.hits[task] = total; }
    if (stats) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            n_lookups += __shfl_xor_sync(0xffffffffu, n_lookups, o);
            n_steps += __shfl_xor_sync(0xffffffffu, n_steps, o);
            total += __shfl_xor_sync(0xffffffffu, total, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (n_lookups) atomicAdd(&stats[0], (unsigned long long)n_lookups);
            if (n_steps) atomicAdd(&stats[1], (unsigned long long)n_steps);
            if (total) atomicAdd(&stats[2], (unsigned long long)total);
        }
    }
}

// ---- task classes for the vote -----------------------------------------------------------------
// Two families of lists: `filter` classes (counting-filter kernel, the fast path, sized by SA hits) and
// `exact` classes (exact hash-table kernels: 5 shared-memory table sizes + 1 global-memory class), which
// take the tasks the filter kernel cannot handle (vote-queue overflow on repeat-rich reads, very long reads,
// kmin == 1, more hits than the largest filter).
#define GMX_N_CLASSES 6
#define GMX_FILTER_LOG2_MIN 12   // filter class c uses a (1 << (12 + c))-byte counting filter, hits <= 512 << c

struct ClassLists {
    uint32_t *list;     // [GMX_N_CLASSES][n_tasks]
    uint32_t *count;    // [GMX_N_CLASSES]
    uint32_t *cursor;   // [GMX_N_CLASSES] work-stealing cursors of the persistent vote kernels
    int64_t n_tasks;
};

__host__ __device__ __forceinline__ uint32_t gmx_class_max_hits(int cls) { return (5u << (10 + cls)) >> 3; }   // load <= 5/8
__host__ __device__ __forceinline__ uint32_t gmx_filter_max_hits(int cls) { return 512u << cls; }

__device__ __forceinline__ int gmx_exact_class(uint32_t h)
{
    int cls = 5;
#pragma unroll
    for (int c = 4; c >= 0; --c) if (h <= gmx_class_max_hits(c)) cls = c;
    return cls;
}

__device__ __forceinline__ void gmx_class_append(const ClassLists &C, int cls, uint32_t task)
{
    uint32_t at = atomicAdd(&C.count[cls], 1u);
    C.list[(int64_t)cls * C.n_tasks + at] = task;
}

__global__ void k_classify(const uint32_t *hits, ClassLists F, ClassLists E, int use_filter)
{
    int64_t task = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t h = task < E.n_tasks ? hits[task] : 0u;
    int cls = -1;                                   // 0..5 filter classes, 6..11 exact classes
    if (h) {
        if (use_filter) {
#pragma unroll
            for (int c = GMX_N_CLASSES - 1; c >= 0; --c) if (h <= gmx_filter_max_hits(c)) cls = c;
        }
        if (cls < 0) cls = GMX_N_CLASSES + gmx_exact_class(h);
    }
    // one atomic per (warp, class) instead of one per task
    const uint32_t peers = __match_any_sync(0xffffffffu, cls);
    if (cls < 0) return;
    const ClassLists &C = cls < GMX_N_CLASSES ? F : E;
    const int c = cls < GMX_N_CLASSES ? cls : cls - GMX_N_CLASSES;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&C.count[c], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    C.list[(int64_t)c * C.n_tasks + base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = (uint32_t)task;
}

// ---- K1b + K1c: locate + diagonal vote ---------------------------------------------------------
// One warp per task.  Rounds (= k-mers of the walk) are processed in order; within a round the 32
// lanes stride over the SA interval, read the de-sampled suffix array (contiguous ranks: coalesced)
// and insert diag = max(0, sa - i) into a warp-private open-addressing table in shared memory
// (uint32 keys + packed 8-bit counters).  The insert that raises a counter to kmin emits the
// candidate, tagged with the round so that the reference's processing order (round, then ascending
// position; reference inc/align_seq2_raw.cpp:28-35,292) can be restored by one radix sort.
struct CandSink {
    unsigned long long *keys;  // (task << 40) | (round << 32) | diag
    uint32_t *count;
    uint32_t *overflow;
    uint32_t cap;
};

__device__ __forceinline__ uint32_t gmx_hash32(uint32_t x) { return (x * 0x9E3779B1u) ^ (x >> 15); }

// exact vote of one task with atomics on a table in global memory (k_vote_gmem: tasks beyond the shared-memory classes)
__device__ __forceinline__ void gmx_vote_task(const DevIndex &ix, const SeedStore &S, uint32_t task, int kmin,
                                              uint32_t *keys, uint32_t *cnts, uint32_t mask, CandSink sink, int lane)
{
    int ns = S.n_seeds[task];
    for (int s = 0; s < ns; ++s) {
        const uint4 rec = S.rec[S.at(task, s)];
        uint32_t rank0 = rec.x, cnt = rec.y;
        uint32_t off = S.offset[S.at(task, s)];
        for (uint32_t t0 = 0; t0 < cnt; t0 += 32) {
            uint32_t t = t0 + lane;
            bool valid = t < cnt;
            bool emit = false;
            uint32_t diag = 0;
            if (valid) {
                uint32_t sa = __ldg(ix.sa_full + rank0 + t);
                diag = (sa <= off) ? 0u : sa - off;
                uint32_t h = gmx_hash32(diag) & mask;
                while (true) {
                    uint32_t prev = atomicCAS(&keys[h], GMX_EMPTY_KEY, diag);
                    if (prev == GMX_EMPTY_KEY || prev == diag) break;
                    h = (h + 1) & mask;
                }
                uint32_t sh = (h & 3u) << 3;
                uint32_t cur = (reinterpret_cast<volatile uint32_t *>(cnts)[h >> 2] >> sh) & 0xffu;
                if ((int)cur < kmin) {
                    uint32_t old = (atomicAdd(&cnts[h >> 2], 1u << sh) >> sh) & 0xffu;
                    emit = ((int)old + 1 == kmin);
                }
            }
            uint32_t em = __ballot_sync(0xffffffffu, emit);
            if (em) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(sink.count, (uint32_t)__popc(em));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (emit) {
                    uint32_t at = base + (uint32_t)__popc(em & ((1u << lane) - 1u));
                    if (at < sink.cap) sink.keys[at] = ((unsigned long long)task << 40) | ((unsigned long long)s << 32) | diag;
                    else *sink.overflow = 1u;
                }
            }
        }
        __syncwarp();
    }
}

// Shared-memory variant, no atomics.  The table is private to one warp, and within one warp step all
// 32 lanes hold hits of the SAME k-mer: their suffix-array values are distinct, hence so are their
// diagonals (after the lanes clamped to diagonal 0 have been folded into one).  Distinct keys can only
// race for an empty slot; that race is settled by "store, __syncwarp, re-read": the lane that reads
// its own key back owns the slot, every other lane moves on.  Counters are one byte per slot and
// are written by the slot's owner of the step only, so plain LDS/STS replace the ATOMS.CAS +
// ATOMS.ADD pair of the generic path (2 cycles per lane each on this part: profiles/r01_*).
#define GMX_VOTE_UNROLL 4
#define GMX_SA_INVALID 0xFFFFFFFFu

template <int SLOTS_LOG2>
__device__ __forceinline__ void gmx_vote_step(uint32_t sa, uint32_t off, int kmin, volatile uint32_t *keys, volatile uint8_t *cnt8,
                                              uint32_t task, int round, CandSink sink, int lane)
{
    bool valid = sa != GMX_SA_INVALID;
    const bool clamp = valid && sa <= off;                    // diag = max(0, sa - off) (reference inc/align_seq2_raw.cpp:270)
    uint32_t diag = clamp ? 0u : sa - off;
    uint32_t inc = 1;
    const uint32_t cm = __ballot_sync(0xffffffffu, clamp);
    if (cm && clamp) {                                        // several hits of this k-mer on diagonal 0: one lane votes for all
        if (lane != __ffs(cm) - 1) valid = false; else inc = (uint32_t)__popc(cm);
    }
    uint32_t h = (diag * 0x9E3779B1u) >> (32 - SLOTS_LOG2);
    bool pending = valid, claimed = false;
    while (__any_sync(0xffffffffu, pending)) {
        if (pending) {
            uint32_t k = keys[h];
            if (k == diag) pending = false;
            else if (k == GMX_EMPTY_KEY) { keys[h] = diag; claimed = true; }
            else { h = (h + 1) & ((1u << SLOTS_LOG2) - 1u); claimed = false; }
        }
        __syncwarp();
    }
    bool emit = false;
    if (valid) {
        uint32_t old = claimed ? 0u : (uint32_t)cnt8[h];
        if ((int)old < kmin) {
            uint32_t nw = old + inc; if (nw > 255u) nw = 255u;
            cnt8[h] = (uint8_t)nw;
            emit = (int)nw >= kmin;
        }
    }
    const uint32_t em = __ballot_sync(0xffffffffu, emit);
    if (em) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(sink.count, (uint32_t)__popc(em));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (emit) {
            uint32_t at = base + (uint32_t)__popc(em & ((1u << lane) - 1u));
            if (at < sink.cap) sink.keys[at] = ((unsigned long long)task << 40) | ((unsigned long long)round << 32) | diag;
            else *sink.overflow = 1u;
        }
    }
    __syncwarp();
}

template <int SLOTS_LOG2, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_vote_smem(DevIndex ix, SeedStore S, ClassLists C, int cls, int kmin, CandSink sink)
{
    constexpr uint32_t SLOTS = 1u << SLOTS_LOG2;
    extern __shared__ __align__(16) uint32_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *keys = smem + (size_t)warp * (SLOTS + SLOTS / 4);
    uint32_t *cnts = keys + SLOTS;
    const uint32_t n_list = C.count[cls];
    const uint32_t *list = C.list + (int64_t)cls * C.n_tasks;
    while (true) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(&C.cursor[cls], 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_list) break;
        const uint32_t task = list[w];
        const int ns = S.n_seeds[task];
        // pull every suffix-array line this task will read into L2 while the table is being cleared
        for (int s = 0; s < ns; ++s) {
            const uint4 rec = S.rec[S.at(task, s)];
            const uint32_t rank0 = rec.x, cnt = rec.y;
            for (uint32_t t = (uint32_t)lane * 32u; t < cnt; t += 1024u)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.sa_full + rank0 + t));
        }
        uint4 *k4 = reinterpret_cast<uint4 *>(keys);
        for (uint32_t x = lane; x < SLOTS / 4; x += 32) k4[x] = make_uint4(GMX_EMPTY_KEY, GMX_EMPTY_KEY, GMX_EMPTY_KEY, GMX_EMPTY_KEY);
        uint4 *c4 = reinterpret_cast<uint4 *>(cnts);
        for (uint32_t x = lane; x < SLOTS / 16; x += 32) c4[x] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        for (int s = 0; s < ns; ++s) {
            const uint4 rec = S.rec[S.at(task, s)];
            const uint32_t rank0 = rec.x, cnt = rec.y;
            const uint32_t off = S.offset[S.at(task, s)];
            for (uint32_t t0 = 0; t0 < cnt; t0 += 32u * GMX_VOTE_UNROLL) {
                uint32_t sa[GMX_VOTE_UNROLL];
#pragma unroll
                for (int u = 0; u < GMX_VOTE_UNROLL; ++u) {
                    const uint32_t t = t0 + 32u * u + lane;
                    sa[u] = t < cnt ? __ldg(ix.sa_full + rank0 + t) : GMX_SA_INVALID;
                }
#pragma unroll
                for (int u = 0; u < GMX_VOTE_UNROLL; ++u)
                    if (t0 + 32u * u < cnt)
                        gmx_vote_step<SLOTS_LOG2>(sa[u], off, kmin, keys, reinterpret_cast<volatile uint8_t *>(cnts), task, s, sink, lane);
            }
        }
        __syncwarp();
    }
}

// ---- K1b + K1c, fast path: counting filter + exact verification against the packed genome ---------
// A diagonal becomes a candidate when `kmin` k-mers of the walk hit it.  Instead of counting every one of
// the ~L/4^mer hits per k-mer exactly, each warp keeps a counting filter (one byte per bucket, two probes per
// hit, plain LDS/STS: within a warp step the 32 lanes hold distinct diagonals of ONE k-mer, so a lost update
// between lanes can only under-count a bucket shared by two DIFFERENT diagonals, never a diagonal's own
// votes).  A hit whose two buckets already held kmin-1 votes is queued.  The queue is a superset of the
// reference's candidates; each queued diagonal d is then verified exactly: k-mer s of the walk hits d iff
// genome[d+off_s, d+off_s+mer) equals the k-mer (that is what membership of d+off_s in the k-mer's SA
// interval means), so one coalesced load of the 2-bit genome window and a few shuffles give the exact vote
// mask over all k-mers, hence the exact count and the exact round at which the reference's counter reaches
// kmin (inc/align_seq2_raw.cpp:28-35,262-274).  The entry queued by that very round's hit emits the candidate,
// which also de-duplicates without any exact table.
#define GMX_FQ_CAP 256           // vote-queue entries per task; drained whenever another step might not fit
#define GMX_FILTER_MAX_SEEDS 64
#define GMX_FILTER_MAX_SPAN 448  // max k-mer offset + mer: the window words must fit the 32 lanes ((15 + span) / 16 + 2 < 32)

// returning OR on a 32-bit shared-window address
__device__ __forceinline__ uint32_t gmx_atoms_or32(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }

template <int SEEDS>
struct FilterSmemT {             // per-warp layout behind the filter bytes; SEEDS = k-mers per task it holds (32 or 64)
    uint32_t queue[GMX_FQ_CAP];
    unsigned long long codes[SEEDS];
    unsigned long long outb[32];     // emitted keys, flushed to the candidate list 32 at a time
    uint32_t rank[SEEDS];
    uint32_t cnt[SEEDS];
    uint16_t offs[SEEDS];
};
typedef FilterSmemT<GMX_FILTER_MAX_SEEDS> FilterSmem;

__host__ __device__ constexpr size_t gmx_filter_warp_bytes(int f_log2) { return ((size_t)1 << f_log2) + sizeof(FilterSmem); }
// the compact variant: filter bytes need not be a power of two, at most 32 k-mers per task
// COMPACT 1: 7552 bytes, two bits per diagonal, six CTAs per SM; COMPACT 2: 5120 bytes, three bits, eight CTAs per SM
// (tasks of up to ~2 k hits: 0.8 false positives per task); COMPACT 3: 7552 bytes, three bits, six CTAs per SM (tasks of
// up to ~8 k hits: 150-bp reads on a 156 Mb genome have 4.2 k, which leaves 9 false positives per task where two bits
// in 8 KB leave 22 -- every one of them costs an exact verification against the genome)
__host__ __device__ constexpr uint32_t gmx_filter_compact_bytes(int compact) { return compact == 2 ? 5120u : 7552u; }
__host__ __device__ constexpr size_t gmx_filter_warp_bytes_compact(int compact) { return (size_t)gmx_filter_compact_bytes(compact) + sizeof(FilterSmemT<32>); }

// exact vote mask of diagonal d > 0 over the k-mers of the walk (bit s <=> k-mer s hits d), from its window words.
// Lane l holds k-mers l and l + 32 of the walk in registers: offset, code, and the last diagonal at which the
// k-mer still fits the genome (seq_len - off - mer).
template <bool SHORT_MER, int H>
__device__ __forceinline__ unsigned long long gmx_exact_mask(uint32_t w, uint32_t d, int ns, int mer, const uint32_t *off,
                                                             const unsigned long long *code, const uint32_t *limit, int lane)
{
    unsigned long long mask = 0;
#pragma unroll
    for (int h = 0; h < H; ++h) {
        if (h == 1 && ns <= 32) break;
        const bool active = lane + 32 * h < ns;
        const uint32_t bit = 2u * ((d & 15u) + off[h]);
        const int wi = (int)(bit >> 5); const uint32_t sh = bit & 31u;
        const uint32_t a = __shfl_sync(0xffffffffu, w, wi & 31), b = __shfl_sync(0xffffffffu, w, (wi + 1) & 31);
        bool hit;
        if (SHORT_MER) {                                             // mer <= 16: the k-mer fits one word
            const uint32_t top = __funnelshift_l(b, a, sh);
            hit = (top >> (32 - 2 * mer)) == (uint32_t)code[h];
        } else {
            const uint32_t c = __shfl_sync(0xffffffffu, w, (wi + 2) & 31);
            const unsigned long long top = ((unsigned long long)__funnelshift_l(b, a, sh) << 32) | __funnelshift_l(c, b, sh);
            hit = (top >> (64 - 2 * mer)) == code[h];
        }
        hit = hit && active && d <= limit[h];
        mask |= (unsigned long long)__ballot_sync(0xffffffffu, hit) << (32 * h);
    }
    return mask;
}

// round at which diagonal 0 reaches kmin votes, or -1.  Diagonal 0 collects every occurrence of k-mer s at a
// position p <= off_s (the reference clamps sa - i at 0, inc/align_seq2_raw.cpp:270), several per k-mer.
template <class FS>
__device__ int gmx_round_diag0(const DevIndex &ix, int ns, int mer, int kmin, const FS *fs, int lane)
{
    uint32_t carry = 0;
    for (int g = 0; g < ns; g += 32) {
        const int s = g + lane;
        uint32_t v = 0;
        if (s < ns) {
            const uint32_t off = fs->offs[s];
            const unsigned long long code = fs->codes[s];
            for (uint32_t p = 0; p <= off && (unsigned long long)p + (unsigned)mer <= ix.seq_len; ++p) {
                unsigned long long k = 0;
                for (int t = 0; t < mer; ++t) k = (k << 2) | (unsigned long long)gmx_pac_base(ix.pac, (int64_t)p + t);
                v += (k == code);
            }
        }
        uint32_t cum = v;                                            // inclusive scan over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += t; }
        cum += carry;
        const uint32_t reached = __ballot_sync(0xffffffffu, (int)cum >= kmin);
        if (reached) return g + __ffs(reached) - 1;
        carry = __shfl_sync(0xffffffffu, cum, 31);
    }
    return -1;
}

// BITS: kmin == 2 only needs "was this diagonal hit before": a blocked Bloom filter over BITS (one 32-bit word per
// diagonal, two bits inside it) -- half the filter bytes of the byte counters at a fifth of their false positives,
// and ONE shared-memory operation per hit: a returning atomic OR sets the two bits and reports whether both were
// already there.  Being atomic, lanes that share a word in one step cannot lose each other's bits.
// U: 32-hit slots per step.  A step handles the hits of ONE k-mer, so U is sized for the genome: a k-mer of a random
// genome has seq_len / 4^mer hits on average (95 at 100 Mb, 149 at 156 Mb for mer 10); with too few slots most k-mers
// need a second, mostly empty step.
// COMPACT (BITS only): the occupancy variant for tasks of at most 32 k-mers -- a 7552-byte filter (the word index is a
// multiply-high instead of a shift, so the size need not be a power of two), half the k-mer arrays in shared memory and
// in registers, and a register budget for six CTAs of four warps per SM: 24 warps instead of 20.
template <int F_LOG2, int WARPS, bool BITS, int U = GMX_VOTE_UNROLL, int COMPACT = 0>
__global__ void __launch_bounds__(WARPS * 32, COMPACT == 2 ? 8 : (COMPACT ? 6 : 1)) k_vote_filter(DevIndex ix, uint32_t pac_words, SeedStore S, ClassLists F, ClassLists E,
                                                            int cls, int kmin, int mer, CandSink sink)
{
    constexpr uint32_t FBYTES = COMPACT ? gmx_filter_compact_bytes(COMPACT) : (1u << F_LOG2);
    constexpr int SEEDS = COMPACT ? 32 : GMX_FILTER_MAX_SEEDS;
    constexpr int H = SEEDS / 32;                                   // k-mers per lane
    typedef FilterSmemT<SEEDS> FS;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *filt = smem_raw + (size_t)warp * (FBYTES + sizeof(FS));
    FS *fs = reinterpret_cast<FS *>(filt + FBYTES);
    const uint32_t n_list = F.count[cls];
    const uint32_t *list = F.list + (int64_t)cls * F.n_tasks;
    const uint32_t lt = (1u << lane) - 1u;
    const int need = kmin - 1;                                       // votes a bucket must already hold

    // seeds of a task in registers: seed `lane` and seed `lane + 32`; fetched one task ahead
    struct Meta { uint32_t task; int ns; uint32_t rank[H], cnt[H], off[H]; unsigned long long code[H]; };
    // Work items are taken GRAB at a time (one same-address atomic per task would serialise 2 M returning atomics
    // per step), and nothing on the way to a task's k-mers waits for a load it has just issued: the cursor of the
    // next grab is requested one grab ahead, its task ids (lanes 0..GRAB-1) one task later, and a task's k-mers --
    // whose addresses follow from the task id alone -- while the task before it is processed.
    constexpr uint32_t GRAB = 4;
    auto grab = [&]() -> uint32_t {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(&F.cursor[cls], GRAB);
        return __shfl_sync(0xffffffffu, w, 0);
    };
    auto load_ids = [&](uint32_t base) -> uint32_t { return ((uint32_t)lane < GRAB && base + lane < n_list) ? list[base + lane] : 0xffffffffu; };
    const int64_t last_seed = S.n_tasks * S.max_seeds - 1;
    const bool two_halves = S.max_seeds > 32;
    auto fetch = [&](uint32_t task, Meta &m) {                  // unmasked: lanes beyond the task's k-mers are cleared when used
        m.task = task;
        m.ns = S.n_seeds[task];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            m.rank[h] = 0u; m.cnt[h] = 0u; m.off[h] = 0u; m.code[h] = 0ull;
            if (h == 1 && !two_halves) break;
            const int64_t at = min(S.at(task, lane + 32 * h), last_seed);
            const uint4 rec = S.rec[at];
            m.rank[h] = rec.x; m.cnt[h] = rec.y; m.off[h] = S.offset[at]; m.code[h] = ((unsigned long long)rec.w << 32) | rec.z;
        }
    };

    uint32_t ne = 0;                                             // emitted keys waiting in fs->outb (kept across tasks)
    auto flush = [&]() {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(sink.count, ne);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        if ((uint32_t)lane < ne) {
            if (base + lane < sink.cap) sink.keys[base + lane] = fs->outb[lane];
            else *sink.overflow = 1u;
        }
        __syncwarp();
        ne = 0;
    };
    Meta cur, nxt;
    uint32_t g_base = grab();
    uint32_t g_ids = load_ids(g_base);
    uint32_t n_base = grab(), n_ids = 0xffffffffu;
    bool n_ids_pending = true;
    uint32_t slot = 0;
    bool has_cur = g_base < n_list;
    if (has_cur) fetch(__shfl_sync(0xffffffffu, g_ids, 0), cur);
    while (has_cur) {
        // step the cursor to the next work item and start loading its k-mers
        if (n_ids_pending && slot >= 1) { n_ids = load_ids(n_base); n_ids_pending = false; }
        if (++slot == GRAB) { g_base = n_base; g_ids = n_ids; slot = 0; n_base = grab(); n_ids_pending = true; }
        const bool has_next = g_base + slot < n_list;
        const uint32_t next_task = __shfl_sync(0xffffffffu, g_ids, (int)slot);
        if (has_next) fetch(next_task, nxt);

        const uint32_t task = cur.task;
        const int ns = cur.ns;
        bool unsupported = ns > SEEDS;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int s = lane + 32 * h;
            if (!(s < ns && s < SEEDS)) { cur.rank[h] = 0u; cur.cnt[h] = 0u; cur.off[h] = 0u; cur.code[h] = 0ull; }
            if (s < SEEDS) { fs->rank[s] = cur.rank[h]; fs->cnt[s] = cur.cnt[h]; fs->offs[s] = (uint16_t)cur.off[h]; fs->codes[s] = cur.code[h]; }
            if (s < ns && cur.off[h] + (uint32_t)mer > GMX_FILTER_MAX_SPAN) unsupported = true;
        }
        unsupported = __any_sync(0xffffffffu, unsupported);
        if (unsupported) {
            if (lane == 0) gmx_class_append(E, gmx_exact_class(S.hits[task]), task);
            cur = nxt; has_cur = has_next;
            continue;
        }
        // pull every suffix-array line this task will read into L2 while the filter is being cleared (asking one task
        // ahead instead measured 0.4 ms slower per step: the lines of two tasks per warp then compete for L2)
#pragma unroll
        for (int h = 0; h < H; ++h)
            for (uint32_t t = 0; t < cur.cnt[h]; t += 32u)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.sa_full + cur.rank[h] + t));
        uint4 *f4 = reinterpret_cast<uint4 *>(filt);
        for (uint32_t x = lane; x < FBYTES / 16; x += 32) f4[x] = make_uint4(0, 0, 0, 0);
        __syncwarp();

        // pass 1: count votes approximately, queue the hits that may complete kmin votes.  The hits of ONE k-mer
        // are distinct diagonals, so up to 32 * U of them are handled as one step: all loads, then all stores.
        // The suffix-array words of the next step are requested before the current one is processed.
        // queue -> candidates: distinct queued diagonals, exact votes from the genome, emission.  Runs whenever the queue
        // may not hold another step (repeat-rich reads) and once at the end; a diagonal drained twice is emitted twice
        // with the same key and dropped after the sort (k_cand_score).
        uint32_t qn = 0, d0_hits = 0;
        auto drain = [&]() {
            // distinct queued diagonals, compacted in place (a diagonal is queued once per k-mer that hits it)
            uint32_t n2 = 0;
            for (uint32_t q0 = 0; q0 < qn; q0 += 32) {
                const uint32_t q = q0 + lane;
                const bool in = q < qn;
                const uint32_t d = in ? fs->queue[q] : (GMX_EMPTY_KEY - (uint32_t)lane);      // padding lanes never match
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                bool first = in && (__ffs(peers) - 1 == lane);
                for (uint32_t j = 0; j < n2; ++j) first = first && fs->queue[j] != d;
                __syncwarp();
                const uint32_t fm = __ballot_sync(0xffffffffu, first);
                if (first) fs->queue[n2 + (uint32_t)__popc(fm & lt)] = d;   // n2 + rank <= q: never overtakes unread entries
                n2 += (uint32_t)__popc(fm);
                __syncwarp();
            }

            // pass 2: exact votes of each distinct diagonal from its genome window, four windows in flight
            uint32_t limit[H];
#pragma unroll
            for (int h = 0; h < H; ++h) limit[h] = (uint32_t)ix.seq_len - cur.off[h] - (uint32_t)mer;
            const uint32_t *pac32 = reinterpret_cast<const uint32_t *>(ix.pac);
            for (uint32_t q0 = 0; q0 < n2; q0 += 4) {
                uint32_t d[4], raw[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) d[i] = q0 + i < n2 ? fs->queue[q0 + i] : GMX_EMPTY_KEY;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t idx = min((d[i] >> 4) + (uint32_t)lane, pac_words - 1u);   // the pad words are zero
                    raw[i] = __ldg(pac32 + idx);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (d[i] == GMX_EMPTY_KEY) continue;
                    int round = -1;
                    if (d[i] != 0u) {
                        const uint32_t ww = __byte_perm(raw[i], 0, 0x0123);   // bases are packed most significant first
                        const unsigned long long m = mer <= 16 ? gmx_exact_mask<true, H>(ww, d[i], ns, mer, cur.off, cur.code, limit, lane)
                                                               : gmx_exact_mask<false, H>(ww, d[i], ns, mer, cur.off, cur.code, limit, lane);
                        if (__popcll(m) >= kmin) {
                            const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
                            const int pl = __popc(lo);
                            round = kmin <= pl ? (int)__fns(lo, 0, kmin) : 32 + (int)__fns(hi, 0, kmin - pl);
                        }
                    } else {
                        round = gmx_round_diag0(ix, ns, mer, kmin, fs, lane);
                    }
                    if (round >= 0) {
                        if (lane == 0) fs->outb[ne] = ((unsigned long long)task << 40) | ((unsigned long long)round << 32) | d[i];
                        ne++;
                        if (ne == 32) flush();
                    }
                }
            }
            qn = 0;
        };
        int s_cur = 0; uint32_t t_cur = 0;
        uint32_t sa_nxt[U];
        auto issue = [&](int s, uint32_t t0, uint32_t (&dst)[U]) {
            const uint32_t cnt = fs->cnt[s];
            const uint32_t *p = ix.sa_full + (fs->rank[s] + t0 + (uint32_t)lane);     // one address, U loads at +128 B
            const uint32_t t = t0 + (uint32_t)lane;
#pragma unroll
            for (int u = 0; u < U; ++u) dst[u] = t + 32u * u < cnt ? __ldg(p + 32 * u) : GMX_SA_INVALID;
        };
        if (ns > 0) issue(0, 0, sa_nxt);
        // one step: the hits in `sa` (requested one step ago) are voted, the next step's are requested into `nx`.  The loop
        // below runs two steps per iteration with the two register sets swapped, so no set is copied between steps.
        // (voting only the slots a step fills -- a k-mer's run seldom reaches the last one -- through a second instantiation
        // of the step body measured slower: 7.39 against 7.27 ms at U = 4, 219 against 184 ms at U = 6, registers and spills)
        auto step = [&](uint32_t (&sa)[U], uint32_t (&nx)[U]) {
            const uint32_t off = fs->offs[s_cur];
            // advance to the next step and request its words (every stored k-mer has at least one hit)
            int s_n = s_cur; uint32_t t_n = t_cur + 32u * U;
            if (t_n >= fs->cnt[s_cur]) { s_n = s_cur + 1; t_n = 0; }
            if (s_n < ns) issue(s_n, t_n, nx);

            uint32_t diag[U];
            bool valid[U];
            uint32_t lowest = sa[0];                                // invalid lanes hold 0xffffffff
#pragma unroll
            for (int u = 0; u < U; ++u) {
                valid[u] = sa[u] != GMX_SA_INVALID;
                diag[u] = sa[u] - off;
                if (u) lowest = min(lowest, sa[u]);
            }
            // hits clamped to diagonal 0 (sa <= off: only at the very start of the genome) bypass the filter: they are
            // counted per task and diagonal 0 is verified exactly from the genome once, at the task's last drain
            if (__any_sync(0xffffffffu, lowest <= off)) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool clamp = valid[u] && sa[u] <= off;
                    d0_hits += (uint32_t)__popc(__ballot_sync(0xffffffffu, clamp));
                    valid[u] = valid[u] && !clamp;
                }
            }
            bool flag[U];
            if (BITS) {
                // blocked Bloom filter: one word per diagonal, two bits inside it.  A 32-bit shared-window address
                // (multiply-add on the FMA pipe) keeps the integer pipe, the busier one here, for the bit masks.
                const uint32_t fbase = (uint32_t)__cvta_generic_to_shared(filt);
                uint32_t w[U], b[U], o[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    // one wide multiply: the word index from the top bits of the low half of the product; the two bit
                    // positions are the low five bits of the diagonal itself (hits are unrelated genome positions)
                    // and of the high half of the product -- shifts by a register wrap, so neither needs a mask
                    const unsigned long long pr = (unsigned long long)diag[u] * 0x9E3779B1ull;
                    const uint32_t idx = COMPACT ? __umulhi((uint32_t)pr, FBYTES / 4u) : (uint32_t)pr >> (32 - (F_LOG2 - 2));
                    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(w[u]) : "r"(idx), "r"(fbase));
                    b[u] = (1u << (diag[u] & 31u)) | (1u << ((uint32_t)(pr >> 32) & 31u));
                    if (COMPACT >= 2) b[u] |= 1u << ((uint32_t)(pr >> 37) & 31u);       // third bit: fewer false positives per filter byte
                }
                // one returning shared-memory atomic per hit: it sets the diagonal's two bits and tells whether both
                // were there already (measured on B200: 9.3 ms per step against 9.6 for load + reduction and 10.6 for
                // load + plain store + re-read + repair; lanes without a hit must skip it, the unit's cost is per lane)
#pragma unroll
                for (int u = 0; u < U; ++u) { o[u] = 0u; if (valid[u]) o[u] = gmx_atoms_or32(w[u], b[u]); }
#pragma unroll
                for (int u = 0; u < U; ++u) flag[u] = valid[u] && (o[u] & b[u]) == b[u];
            } else {
                uint32_t h1[U], h2[U], c1[U], c2[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    h1[u] = (diag[u] * 0x9E3779B1u) >> (32 - F_LOG2); h2[u] = (diag[u] * 0x85EBCA77u + 0x27D4EB2Fu) >> (32 - F_LOG2);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { c1[u] = valid[u] ? filt[h1[u]] : 0u; c2[u] = valid[u] ? filt[h2[u]] : 0u; }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    flag[u] = valid[u] && (int)(min(c1[u], c2[u]) + 1u) > need;
                    if (valid[u]) {
                        filt[h1[u]] = (uint8_t)min(c1[u] + 1u, 255u);
                        filt[h2[u]] = (uint8_t)min(c2[u] + 1u, 255u);
                    }
                }
            }
            bool any_flag = false;
#pragma unroll
            for (int u = 0; u < U; ++u) any_flag |= flag[u];
            // (appending the flagged hits lane by lane instead of slot by slot measured 0.5 ms slower per step)
            if (__any_sync(0xffffffffu, any_flag)) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t fm = __ballot_sync(0xffffffffu, flag[u]);
                    if (flag[u]) {
                        const uint32_t at = qn + (uint32_t)__popc(fm & lt);
                        if (at < GMX_FQ_CAP) fs->queue[at] = diag[u];
                    }
                    qn += (uint32_t)__popc(fm);
                }
            }
            __syncwarp();
            s_cur = s_n; t_cur = t_n;
        };
        uint32_t sa_alt[U];
        do {                                                           // pass-1 segments separated by queue drains (one drain site)
        while (s_cur < ns && qn < GMX_FQ_CAP - 32u * U) {         // room for one more step of flagged hits and diagonal 0
            // U == 4: two steps per iteration with the register sets swapped (7.40 -> 7.27 ms per step); with six slots
            // the second set of registers spills, so there the set is copied at the start of every step
            if (U != 4) {
#pragma unroll
                for (int u = 0; u < U; ++u) sa_alt[u] = sa_nxt[u];
                step(sa_alt, sa_nxt);
                continue;
            }
            step(sa_nxt, sa_alt);
            if (!(s_cur < ns && qn < GMX_FQ_CAP - 32u * U)) {
#pragma unroll
                for (int u = 0; u < U; ++u) sa_nxt[u] = sa_alt[u];
                break;
            }
            step(sa_alt, sa_nxt);
        }
        if (s_cur >= ns && d0_hits) {                                  // diagonal 0 joins the last drain
            if (lane == 0) fs->queue[qn] = 0u;
            qn++;
            __syncwarp();
        }
        drain();
        } while (s_cur < ns);
        __syncwarp();
        cur = nxt; has_cur = has_next;
    }
    if (ne) flush();
}

// Tasks whose hit count exceeds the largest shared-memory table (repeat-rich reads): same insert
// logic over a table carved out of a global scratch buffer by a bump allocator; one warp per task.
struct GlobalTableArena {
    uint32_t *words;
    unsigned long long *used;   // in words
    unsigned long long cap;     // in words
    uint32_t *overflow;
};

__global__ void __launch_bounds__(128) k_vote_gmem(DevIndex ix, SeedStore S, ClassLists C, int cls, int kmin, CandSink sink, GlobalTableArena A)
{
    int lane = threadIdx.x & 31;
    const uint32_t n_list = C.count[cls];
    const uint32_t *list = C.list + (int64_t)cls * C.n_tasks;
    while (true) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(&C.cursor[cls], 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_list) break;
        uint32_t task = list[w];
        uint32_t hits = S.hits[task];
        uint32_t slots = 1u << 15;
        while ((unsigned long long)slots * 5ull < (unsigned long long)hits * 8ull && slots < (1u << 31)) slots <<= 1;
        unsigned long long need = (unsigned long long)slots + slots / 4, base = 0;
        if (lane == 0) base = atomicAdd(A.used, need);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + need > A.cap) { if (lane == 0) *A.overflow = 1u; continue; }
        uint32_t *keys = A.words + base, *cnts = keys + slots;
        for (uint32_t x = lane; x < slots; x += 32) keys[x] = GMX_EMPTY_KEY;
        for (uint32_t x = lane; x < slots / 4; x += 32) cnts[x] = 0;
        __syncwarp();
        gmx_vote_task(ix, S, task, kmin, keys, cnts, slots - 1, sink, lane);
        __syncwarp();
    }
}

// ---- counts that live on the device ---------------------------------------------------------------
// A chunk issued without waiting for its candidate / leader counts (gmx.cu "optimistic chunk") launches its kernels over
// host-side BOUNDS and hands them the address of the real count: k_seal_candidates / k_seal_leaders store it there, or 0
// when the chunk has to be run again (a bound was exceeded, a task class was not launched), which turns every kernel behind
// them into a no-op.  A null address means the bound is the count.
__device__ __forceinline__ uint32_t gmx_live(const uint32_t *live, uint32_t bound)
{
    return live ? min(__ldg(live), bound) : bound;
}

struct SealIn {
    const uint32_t *n_cand, *cand_overflow, *arena_overflow;
    const uint32_t *cls_count, *fcls_count;     // [GMX_N_CLASSES] each
    uint32_t launched;                           // bit k: filter class k, bit GMX_N_CLASSES + k: exact class k
    uint32_t bound;                              // candidates the buffers and grids of this chunk cover
};

// pads the key list up to the sort's element count with keys that sort last, and decides whether the chunk stands
__global__ void __launch_bounds__(256) k_seal_candidates(SealIn in, unsigned long long *keys, uint32_t *live_cand, uint32_t *bad)
{
    const uint32_t n = *in.n_cand;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n && c < in.bound) keys[c] = ~0ull;
    if (c == 0) {
        uint32_t need = 0;
        for (int k = 0; k < GMX_N_CLASSES; ++k) {
            if (in.fcls_count[k]) need |= 1u << k;
            if (in.cls_count[k]) need |= 1u << (GMX_N_CLASSES + k);
        }
        const bool ok = n <= in.bound && !*in.cand_overflow && !*in.arena_overflow && !(need & ~in.launched);
        *live_cand = ok ? n : 0u;
        *bad = ok ? 0u : 1u;
    }
}

__global__ void k_seal_leaders(const uint32_t *n_leaders, uint32_t lead_cap, uint32_t *live_cand, uint32_t *live_lead, uint32_t *bad)
{
    const uint32_t n = *n_leaders;
    if (*bad || n > lead_cap) { *live_cand = 0; *live_lead = 0; *bad = 1; }
    else *live_lead = n;
}

// ---- candidate scoring -------------------------------------------------------------------------
__device__ __forceinline__ void gmx_decode_key(unsigned long long key, uint32_t &task, uint32_t &round, uint32_t &diag)
{
    task = (uint32_t)(key >> 40); round = (uint32_t)(key >> 32) & 0xffu; diag = (uint32_t)key;
}

// One thread per candidate (inter-task parallelism; the band lives in registers).
__global__ void __launch_bounds__(128) k_cand_score(DevIndex ix, DevReads R, DevTables T, DevParams P,
                                                    const unsigned long long *keys, uint32_t n_cand, float *score, const uint32_t *live)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= gmx_live(live, n_cand)) return;
    uint32_t task, round, diag;
    gmx_decode_key(keys[c], task, round, diag);
    ReadView rd = gmx_read_view(R, (int)(task >> 1), (int)(task & 1));
    float sc = __int_as_float(0x7fc00000);                       // NaN: window is "" (chromosome boundary)
    // a diagonal drained from the vote queue twice arrives twice with the same key: keep the first copy only
    const bool dup = c > 0 && keys[c - 1] == keys[c];
    if (!dup && gmx_window_valid(ix, diag, rd.n)) {
        WindowView win; win.pac = ix.pac; win.pos = diag; win.chars = nullptr;
        sc = gmx_nw_score_dispatch(rd, win, T, P.gap, P.max_gap);
    }
    score[c] = sc;
}

// ---- per-read finalisation ---------------------------------------------------------------------
// base j of the read-orientation key string of candidate (diag, neg): POS = window[j];
// NEG = complement(window[n-1-j])  (reference inc/align_seq2_raw.cpp:125-128)
__device__ __forceinline__ int gmx_key_base(const DevIndex &ix, uint32_t diag, int neg, int n, int j)
{
    int b = gmx_pac_base(ix.pac, (int64_t)diag + (neg ? n - 1 - j : j));
    return neg ? 3 - b : b;
}

__device__ uint64_t gmx_key_hash(const DevIndex &ix, uint32_t diag, int neg, int n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (int j = 0; j < n; ++j) { h ^= (uint64_t)gmx_key_base(ix, diag, neg, n, j) + 1; h *= 0x100000001b3ull; }
    return h;
}

// lexicographic compare of two key strings (<0, 0, >0)
__device__ int gmx_key_compare(const DevIndex &ix, uint32_t da, int na, uint32_t db, int nb, int n)
{
    for (int j = 0; j < n; ++j) {
        int a = gmx_key_base(ix, da, na, n, j), b = gmx_key_base(ix, db, nb, n, j);
        if (a != b) return a - b;
    }
    return 0;
}

// candidate range of every read in the sorted key list (replaces two binary searches per read): range[2r], range[2r+1];
// the array is zeroed first, so reads without candidates keep the empty range [0, 0)
__global__ void k_cand_ranges(const unsigned long long *keys, uint32_t n_cand, uint32_t *range, const uint32_t *live)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    n_cand = gmx_live(live, n_cand);
    if (c >= n_cand) return;
    const uint32_t r = (uint32_t)(keys[c] >> 41);                     // task >> 1
    if (c == 0 || (uint32_t)(keys[c - 1] >> 41) != r) range[2 * r] = c;
    if (c + 1 == n_cand || (uint32_t)(keys[c + 1] >> 41) != r) range[2 * r + 1] = c + 1;
}

__device__ __forceinline__ double gmx_shfl_f64(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double gmx_shfl_xor_f64(double v, int mask)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask); hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}

// leader rank within the read -> leader slot: slot = read_base[read] + rank (read_base = exclusive scan of the reads'
// group counts); also counts the leaders and the accepted candidates of the chunk
__global__ void __launch_bounds__(256) k_assign_slots(const unsigned long long *keys, const int32_t *leader, int32_t *slot, const uint32_t *read_base,
                                                      uint32_t *lead_cand, uint32_t n_cand, uint32_t *n_leaders, uint32_t *n_accepted,
                                                      const uint32_t *live, uint32_t lead_cap)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = c < gmx_live(live, n_cand);
    const bool acc = in && leader[c] >= 0;
    const bool lead = acc && slot[c] >= 0;
    if (lead) {
        const uint32_t r = (uint32_t)(keys[c] >> 41);
        const uint32_t s = read_base[r] + (uint32_t)slot[c];
        slot[c] = (int32_t)s;
        if (s < lead_cap) lead_cand[s] = c;              // beyond the bound: k_seal_leaders voids the chunk
    }
    const uint32_t am = __ballot_sync(0xffffffffu, acc), lm = __ballot_sync(0xffffffffu, lead);
    if ((threadIdx.x & 31) == 0) {
        if (am) atomicAdd(n_accepted, (uint32_t)__popc(am));
        if (lm) atomicAdd(n_leaders, (uint32_t)__popc(lm));
    }
}

struct FinalizeOut {
    gmx_read_result *results;   // [n_reads]
    int32_t *leader;            // [n_cand] candidate index of the group leader, or -1 (not accepted)
    int32_t *slot;              // [n_cand] leader-list slot for leaders, else -1
    uint32_t *lead_cand;        // [lead_cap] candidate index per leader slot
    uint32_t *n_leaders;
    uint32_t *n_accepted;
    uint32_t *groups_per_read;  // [n_reads] distinct accepted genome strings of the read (0 when not mapped)
    uint64_t *hashes;           // [n_cand] scratch: key hash of accepted candidates
    double   *expv;             // [n_cand] scratch: exp(score) of accepted candidates
    const uint32_t *range;      // [2 * n_reads] candidate range of every read (k_cand_ranges)
};

// One warp per read.  Candidates of the read are contiguous in the sorted list: POS strand then
// NEG strand, each in (round, position) order -- the order in which the reference meets them.
__global__ void __launch_bounds__(128) k_finalize_reads(DevIndex ix, DevReads R, DevParams P, const ReadPrep *prep,
                                                        const unsigned long long *keys, const float *score, uint32_t n_cand,
                                                        FinalizeOut O)
{
    int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= R.n_reads) return;
    ReadPrep pr = prep[r];
    gmx_read_result res;
    res.top_score = 0; res.denominator = 0; res.max_align_score = pr.max_align; res.status = pr.status;
    res.n_groups = 0; res.n_candidates = 0; res.best_score = 0; res.best_posterior = 0; res.best_n_positions = 0;
    res.best_first_strand = 0; res.best_first_pos = 0; res.hit_begin = 0; res.hit_end = 0; res.best_group = -1; res.best_aligned_len = 0;
    if (pr.status == GMX_READ_TOO_SHORT) { res.top_score = -2; if (lane == 0) O.results[r] = res; return; }
    if (pr.status == GMX_READ_TOO_POOR) { res.top_score = -3; if (lane == 0) O.results[r] = res; return; }
    const int n = gmx_read_len(R, r);
    const uint32_t lo = O.range[2 * r], hi = O.range[2 * r + 1];

    if (hi - lo <= 32u) {
        // ---- common case: at most one candidate per lane, everything stays in registers (no round trips through the
        // scratch arrays between the passes) -- same decisions, same order of the FP64 sums
        const uint32_t cnt = hi - lo;
        const uint32_t c = lo + (uint32_t)lane;
        const bool in = (uint32_t)lane < cnt;
        const float sc = in ? score[c] : 0.f;
        uint32_t task = 0, round = 0, diag = 0;
        if (in) gmx_decode_key(keys[c], task, round, diag);
        const bool valid = in && !isnan(sc);
        const bool acc = valid && ((double)sc >= pr.min_align);
        const uint32_t accm = __ballot_sync(0xffffffffu, acc);
        const int n_acc = __popc(accm);
        res.n_candidates = __popc(__ballot_sync(0xffffffffu, valid));
        double top = (valid && (double)sc > 0.0) ? (double)sc : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, top, o); if (t > top) top = t; }
        if (n_acc == 0) {
            res.status = GMX_READ_UNMATCHED;
            if (in) { O.leader[c] = -1; O.slot[c] = -1; }
            if (lane == 0) O.results[r] = res;
            return;
        }
        // group leaders: first accepted candidate (processing order) with the same key string
        int lead_lane = lane;
        if (n_acc > 1) {
            const uint64_t h = acc ? gmx_key_hash(ix, diag, (int)(task & 1), n) : 0ull;
            bool open = acc;
            for (uint32_t m = accm; m; m &= m - 1) {
                const int p = __ffs(m) - 1;
                const uint64_t hp = __shfl_sync(0xffffffffu, h, p);
                const uint32_t dp = __shfl_sync(0xffffffffu, diag, p), tp = __shfl_sync(0xffffffffu, task, p);
                if (open && p < lane && hp == h && gmx_key_compare(ix, diag, (int)(task & 1), dp, (int)(tp & 1), n) == 0) { lead_lane = p; open = false; }
            }
        }
        const bool lead = acc && lead_lane == lane;
        const uint32_t leadm = __ballot_sync(0xffffffffu, lead);
        const int n_groups = __popc(leadm), joined = n_acc - n_groups;
        if ((P.unique_only && joined > 0) || (uint32_t)n_groups > P.max_matches) {
            res.status = GMX_READ_TOO_MANY; res.top_score = 999999; res.denominator = 0; res.n_groups = 0;
            if (in) { O.leader[c] = -1; O.slot[c] = -1; }
            if (lane == 0) O.results[r] = res;
            return;
        }
        // denominator: exp(score) of every accepted candidate, added in processing order (one FP64 exp per lane: the
        // same value serves the denominator, the best-group comparison and the posterior)
        const double ev = acc ? exp((double)sc) : 0.0;
        double denom = 0.0;
        for (uint32_t m = accm; m; m &= m - 1) denom = __dadd_rn(denom, gmx_shfl_f64(ev, __ffs(m) - 1));
        // best group: largest exp(score), strict >, groups visited in key order, starting from exp(-1)
        const double e_floor = exp(-1.0);
        const bool cand = lead && ev > e_floor;
        double best_e = cand ? ev : 0.0;
        float best_sc = cand ? sc : -1.0f;
        int best_lane = cand ? lane : -1;
        uint32_t best_diag = diag, best_task = task;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double oe = gmx_shfl_xor_f64(best_e, o);
            const float osc = __shfl_xor_sync(0xffffffffu, best_sc, o);
            const int ol = __shfl_xor_sync(0xffffffffu, best_lane, o);
            const uint32_t od = __shfl_xor_sync(0xffffffffu, best_diag, o), ot = __shfl_xor_sync(0xffffffffu, best_task, o);
            bool better = false;
            if (ol >= 0) {
                if (best_lane < 0 || oe > best_e) better = true;
                else if (oe == best_e && ol != best_lane)
                    better = gmx_key_compare(ix, od, (int)(ot & 1), best_diag, (int)(best_task & 1), n) < 0;
            }
            if (better) { best_e = oe; best_sc = osc; best_lane = ol; best_diag = od; best_task = ot; }
        }
        // leader slots + per-candidate outputs for PHASE B
        // slot[] holds the leader's rank within its read here; k_assign_slots adds the read's base (an exclusive scan
        // of groups_per_read) -- no same-address atomic per read
        if (in) {
            O.leader[c] = acc ? (int32_t)(lo + (uint32_t)lead_lane) : -1;
            O.slot[c] = lead ? (int32_t)__popc(leadm & ((1u << lane) - 1u)) : -1;
        }
        if (lane == 0) O.groups_per_read[r] = (uint32_t)n_groups;
        res.status = GMX_READ_MAPPED;
        res.top_score = top; res.denominator = denom; res.n_groups = n_groups;
        res.hit_begin = (int32_t)lo; res.hit_end = (int32_t)hi;
        if (best_lane >= 0) {
            const bool mem = acc && lead_lane == best_lane;
            const int members = __popc(__ballot_sync(0xffffffffu, mem));
            uint64_t first = mem ? (((uint64_t)diag << 1) | (task & 1)) : ~0ull;
#pragma unroll
            for (int o = 16; o; o >>= 1) { uint64_t t = __shfl_xor_sync(0xffffffffu, first, o); if (t < first) first = t; }
            res.best_group = best_lane;
            res.best_score = best_sc;
            res.best_posterior = (float)(best_e / denom);
            res.best_n_positions = members;
            res.best_first_strand = (int)(best_task & 1);
            res.best_first_pos = first >> 1;
        }
        if (lane == 0) O.results[r] = res;
        return;
    }

    // pass 1: validity, top score, acceptance, exp(score), key hash
    int n_valid = 0, n_acc = 0;
    double top = 0.0;
    for (uint32_t c0 = lo; c0 < hi; c0 += 32) {
        uint32_t c = c0 + lane;
        bool in = c < hi;
        float sc = in ? score[c] : 0.f;
        bool valid = in && !isnan(sc);
        bool acc = valid && ((double)sc >= pr.min_align);
        if (valid && (double)sc > top) top = (double)sc;
        if (in) {
            O.leader[c] = acc ? (int32_t)c : -1;
            O.slot[c] = -1;
            if (acc) O.expv[c] = exp((double)sc);
        }
        n_valid += __popc(__ballot_sync(0xffffffffu, valid));
        n_acc += __popc(__ballot_sync(0xffffffffu, acc));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, top, o); if (t > top) top = t; }
    res.n_candidates = n_valid;
    __syncwarp();

    if (n_acc == 0) {                                    // reference src/Driver.cpp:593-602
        res.status = GMX_READ_UNMATCHED; res.top_score = 0; res.denominator = 0;
        if (lane == 0) O.results[r] = res;
        return;
    }

    // key-string hashes are only needed to group several accepted candidates
    if (n_acc > 1) {
        for (uint32_t c = lo + lane; c < hi; c += 32)
            if (O.leader[c] >= 0) {
                uint32_t task, round, diag; gmx_decode_key(keys[c], task, round, diag);
                O.hashes[c] = gmx_key_hash(ix, diag, (int)(task & 1), n);
            }
        __syncwarp();
    }

    // pass 2: group leaders = first accepted candidate (processing order) with the same key string
    int n_groups = 0, joined = 0;
    for (uint32_t c0 = lo; c0 < hi; c0 += 32) {
        uint32_t c = c0 + lane;
        bool acc = (c < hi) && O.leader[c] >= 0;
        bool is_leader = acc;
        if (acc) {
            uint32_t task, round, diag; gmx_decode_key(keys[c], task, round, diag);
            uint64_t h = O.hashes[c];
            for (uint32_t p = lo; p < c; ++p) {
                if (O.leader[p] < 0 || O.hashes[p] != h) continue;     // leader[p] >= 0 <=> accepted (pass 1 complete)
                uint32_t tp, rp, dp; gmx_decode_key(keys[p], tp, rp, dp);
                if (gmx_key_compare(ix, diag, (int)(task & 1), dp, (int)(tp & 1), n) == 0) { is_leader = false; O.slot[c] = -2 - (int32_t)(p - lo); break; }
            }
        }
        n_groups += __popc(__ballot_sync(0xffffffffu, is_leader));
        joined += __popc(__ballot_sync(0xffffffffu, acc && !is_leader));
    }
    __syncwarp();
    // resolve leader indices (slot[c] temporarily holds -2 - (first equal predecessor)); the first
    // equal predecessor of a non-leader is always a leader because equality is transitive
    for (uint32_t c0 = lo; c0 < hi; c0 += 32) {
        uint32_t c = c0 + lane;
        if (c < hi && O.leader[c] >= 0 && O.slot[c] <= -2) { O.leader[c] = (int32_t)(lo + (uint32_t)(-2 - O.slot[c])); O.slot[c] = -1; }
    }
    __syncwarp();

    // reference inc/align_seq2_raw.cpp:151-158 (gUNIQUE) and :299 (gMAX_MATCHES)
    if ((P.unique_only && joined > 0) || (uint32_t)n_groups > P.max_matches) {
        res.status = GMX_READ_TOO_MANY; res.top_score = 999999; res.denominator = 0; res.n_groups = 0;
        for (uint32_t c = lo + lane; c < hi; c += 32) { O.leader[c] = -1; O.slot[c] = -1; }
        if (lane == 0) O.results[r] = res;
        return;
    }

    // denominator: sum of exp(score) over accepted candidates in processing order (FP64, sequential)
    double denom = 0.0;
    for (uint32_t c0 = lo; c0 < hi; c0 += 32) {
        uint32_t c = c0 + lane;
        bool acc = (c < hi) && O.leader[c] >= 0;
        double e = acc ? O.expv[c] : 0.0;
        uint32_t m = __ballot_sync(0xffffffffu, acc);
        while (m) {
            int src = __ffs(m) - 1; m &= m - 1;
            double v = __shfl_sync(0xffffffffu, e, src);
            denom = __dadd_rn(denom, v);
        }
    }

    // best group: largest exp(score) with strict >, groups visited in key (lexicographic) order,
    // starting from the empty ScoredSeq whose score is -1  (reference src/Driver.cpp:636,672)
    float best_sc = -1.0f; int best_c = -1;
    for (uint32_t c0 = lo; c0 < hi; c0 += 32) {
        uint32_t c = c0 + lane;
        bool lead = (c < hi) && O.leader[c] == (int32_t)c;
        float sc = lead ? score[c] : -1.0f;
        if (lead && exp((double)sc) > exp(-1.0)) {
            // the reference compares exp(score) (a double that saturates at +inf for scores above ~709.78: very long
            // reads), so scores are compared through it; equal values fall back to the key order
            bool better = (best_c < 0) || exp((double)sc) > exp((double)best_sc);
            if (!better && exp((double)sc) == exp((double)best_sc)) {
                uint32_t ta, ra, da, tb, rb, db; gmx_decode_key(keys[c], ta, ra, da); gmx_decode_key(keys[best_c], tb, rb, db);
                better = gmx_key_compare(ix, da, (int)(ta & 1), db, (int)(tb & 1), n) < 0;
            }
            if (better) { best_sc = sc; best_c = (int)c; }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        float osc = __shfl_xor_sync(0xffffffffu, best_sc, o);
        int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
        bool better = false;
        if (oc >= 0) {
            if (best_c < 0 || exp((double)osc) > exp((double)best_sc)) better = true;
            else if (exp((double)osc) == exp((double)best_sc) && oc != best_c) {
                uint32_t ta, ra, da, tb, rb, db; gmx_decode_key(keys[oc], ta, ra, da); gmx_decode_key(keys[best_c], tb, rb, db);
                int cmp = gmx_key_compare(ix, da, (int)(ta & 1), db, (int)(tb & 1), n);
                better = cmp < 0;
            }
        }
        if (better) { best_sc = osc; best_c = oc; }
    }

    // leader ranks within the read (k_assign_slots turns them into slots)
    uint32_t rank_base = 0;
    for (uint32_t c0 = lo; c0 < hi; c0 += 32) {
        uint32_t c = c0 + lane;
        bool lead = (c < hi) && O.leader[c] == (int32_t)c;
        uint32_t m = __ballot_sync(0xffffffffu, lead);
        if (lead) O.slot[c] = (int32_t)(rank_base + (uint32_t)__popc(m & ((1u << lane) - 1u)));
        rank_base += (uint32_t)__popc(m);
    }
    if (lane == 0) O.groups_per_read[r] = (uint32_t)n_groups;

    res.status = GMX_READ_MAPPED;
    res.top_score = top; res.denominator = denom; res.n_groups = n_groups;
    res.hit_begin = (int32_t)lo; res.hit_end = (int32_t)hi;       // candidate range; remapped to hit indices on the host
    if (best_c >= 0) {
        // members of the best group: count and smallest (pos, strand)
        int cnt = 0; uint64_t first = ~0ull;
        for (uint32_t c0 = lo; c0 < hi; c0 += 32) {
            uint32_t c = c0 + lane;
            bool mem = (c < hi) && O.leader[c] == best_c;
            if (mem) { uint32_t t, rr, d; gmx_decode_key(keys[c], t, rr, d); uint64_t v = ((uint64_t)d << 1) | (t & 1); if (v < first) first = v; }
            cnt += __popc(__ballot_sync(0xffffffffu, mem));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { uint64_t t = __shfl_xor_sync(0xffffffffu, first, o); if (t < first) first = t; }
        uint32_t tb, rb, db; gmx_decode_key(keys[best_c], tb, rb, db);
        res.best_group = best_c - (int)lo;              // label: candidate index of the leader within the read
        res.best_score = best_sc;
        res.best_posterior = (float)(exp((double)best_sc) / denom);
        res.best_n_positions = cnt;
        res.best_first_strand = (int)(tb & 1);
        res.best_first_pos = first >> 1;
    }
    if (lane == 0) O.results[r] = res;
}

// ---- K2b per group leader ------------------------------------------------------------------------
struct LeaderStore {
    const uint32_t *lead_cand;   // [n_leaders]
    int32_t *alen;               // [n_leaders] length of the gapped `aligned` string
    uint8_t *aligned;            // [n_leaders][a_stride] (BS mode, and for gmx_get_best_alignments)
    char    *cigar;              // [n_leaders][c_stride]
    float   *hmm;                // [n_leaders][max_len][5]   (SNP mode)
    int a_stride, c_stride, max_len;
};

// moves: one word per (row, group) in a global scratch laid out [row][group], so the lanes of a warp write
// adjacent words (404 MB per million 100-bp groups: noise next to the ALU work)
__global__ void __launch_bounds__(128) k_traceback(DevIndex ix, DevReads R, DevTables T, DevParams P,
                                                   const unsigned long long *keys, uint32_t n_leaders, LeaderStore L,
                                                   uint32_t *moves, int want_aligned, uint32_t *truncated, const uint32_t *live)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= gmx_live(live, n_leaders)) return;
    uint32_t task, round, diag;
    gmx_decode_key(keys[L.lead_cand[s]], task, round, diag);
    ReadView rd = gmx_read_view(R, (int)(task >> 1), (int)(task & 1));
    WindowView win; win.pac = ix.pac; win.pos = diag; win.chars = nullptr;
    ConsView cons; cons.explicit_chars = nullptr;          // score(): max_char consensus of the oriented PWM
    TracebackOut out;
    out.aligned = want_aligned ? L.aligned + (size_t)s * L.a_stride : nullptr; out.aligned_cap = L.a_stride;
    out.cigar = L.cigar + (size_t)s * L.c_stride; out.cigar_cap = L.c_stride; out.fix_deletions = 1; out.truncated = truncated;
    L.alen[s] = gmx_nw_traceback(rd, win, cons, T, P.gap, P.max_gap, moves + s, (int64_t)n_leaders, out);
}

// ---- K3: posterior scatter -----------------------------------------------------------------------
// One warp per accepted (position, strand).  total = exp(score_of_group) / denominator as a float
// (reference src/NormalScoredSeq.cpp:29,71: double -> `const float&`).  Lanes cover consecutive
// genome positions; positions falling into the same accumulator bin are combined with
// __match_any_sync before the atomic, so Normal mode (8 bases per bin) issues one RED per bin.
struct Accum {
    float *amount;          // [n_amount]
    float *planes[5];       // [l_pac] each, BS / SNP
    uint64_t n_amount;
};

__global__ void __launch_bounds__(128) k_scatter(DevIndex ix, DevReads R, DevParams P, const unsigned long long *keys,
                                                 const float *score, const int32_t *leader, const int32_t *slot, uint32_t n_cand,
                                                 const gmx_read_result *results, LeaderStore L, Accum A, const uint32_t *live)
{
    n_cand = gmx_live(live, n_cand);
    // a warp owns 32 consecutive candidates and scatters the accepted ones in turn with all its lanes (most
    // candidates are not accepted: a warp per candidate would launch four idle warps for every working one)
    const uint32_t c_lane = (blockIdx.x * blockDim.x + threadIdx.x);
    const int lane = threadIdx.x & 31;
    const int32_t ld_lane = c_lane < n_cand ? leader[c_lane] : -1;
    for (uint32_t todo = __ballot_sync(0xffffffffu, ld_lane >= 0); todo; todo &= todo - 1) {
    const int src = __ffs(todo) - 1;
    const uint32_t c = (c_lane & ~31u) + (uint32_t)src;
    const int32_t ld = __shfl_sync(0xffffffffu, ld_lane, src);
    uint32_t task, round, diag, tl, rl, dl;
    gmx_decode_key(keys[c], task, round, diag);
    gmx_decode_key(keys[ld], tl, rl, dl);
    int r = (int)(task >> 1);
    int s = slot[ld];
    double denom = results[r].denominator;
    float total = (float)(exp((double)score[ld]) / denom);
    int same = ((task & 1) == (tl & 1));                 // strand == firstStrand of the group
    if (P.mode == GMX_MODE_SNP) {
        int n = gmx_read_len(R, r);
        const float *hmm = L.hmm + (size_t)s * L.max_len * 5;
        for (int i = lane; i < n; i += 32) {
            int64_t p = (int64_t)diag + i;
            if (p >= ix.l_pac) continue;
            atomicAdd(&A.amount[p / P.gen_size], total);
            // other strand: reverse_comp_cpy_phmm (reference inc/SequenceOperations.h:164-181)
            const float *h = hmm + 5 * (same ? i : n - 1 - i);
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                float v = same ? h[b] : (b < 4 ? h[3 - b] : h[4]);
                atomicAdd(&A.planes[b][p / P.gen_size], __fmul_rn(v, total));
            }
        }
        continue;
    }
    int alen = L.alen[s];
    const uint8_t *al = L.aligned + (size_t)s * L.a_stride;
    for (int i0 = 0; i0 < alen; i0 += 32) {
        int i = i0 + lane;
        int64_t p = (int64_t)diag + i;
        bool ok = i < alen && p < ix.l_pac;
        uint32_t bin = ok ? (uint32_t)(p / P.gen_size) : 0xffffffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (ok && (__ffs(peers) - 1) == lane) atomicAdd(&A.amount[bin], __fmul_rn((float)__popc(peers), total));
        if (ok && P.mode == GMX_MODE_BS) {
            // g_gen_CONVERSION of aligned[i] (same strand) or of reverse_comp(aligned)[i]
            uint8_t ch = same ? al[i] : al[alen - 1 - i];
            int code;
            switch (ch) { case 'a': code = 0; break; case 'c': code = 1; break; case 'g': code = 2; break; case 't': code = 3; break;
                          case 0: code = same ? 6 : 4; break; default: code = 4; break; }
            if (!same && code < 4) code = 3 - code;
            if (code < 5) atomicAdd(&A.planes[code][bin], total);
        }
    }
    }
}

// ---- best alignment per read (fast download path) ------------------------------------------------
// One thread per read: copies the CIGAR of the best group next to the per-read result so that only
// [n_reads] fixed-size records leave the device when the caller does not ask for the hit list.
__global__ void __launch_bounds__(128) k_gather_best(const gmx_read_result *results, gmx_read_result *out, int n_reads, const int32_t *slot,
                                                     LeaderStore L, char *best_cigar, int stride, int have_traceback, const uint32_t *live_lead)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    if (live_lead && __ldg(live_lead) == 0) have_traceback = 0;
    uint4 *dst = reinterpret_cast<uint4 *>(best_cigar + (size_t)r * stride);
    gmx_read_result res = results[r];
    const bool has = have_traceback && res.status == GMX_READ_MAPPED && res.best_group >= 0;
    if (has) {
        int s = slot[res.hit_begin + res.best_group];
        res.best_aligned_len = L.alen[s];
        const uint4 *src = reinterpret_cast<const uint4 *>(L.cigar + (size_t)s * L.c_stride);
        for (int k = 0; k < stride / 16; ++k) dst[k] = src[k];
    } else {
        for (int k = 0; k < stride / 16; ++k) dst[k] = make_uint4(0, 0, 0, 0);
    }
    // the candidate range / leader index are device-internal (and stay intact in `results` for a later PHASE B)
    res.hit_begin = 0; res.hit_end = 0; res.best_group = -1;
    out[r] = res;
}

// ---- SAM row: every (position, strand) of the best group ------------------------------
// get_SAM prints one record per element of the best ScoredSeq's position set (reference inc/ScoredSeq.h:293-404,
// src/Driver.cpp:700-716).  The per-read record already carries the smallest one; groups with more than one
// position (repeats) append all of theirs here -- a short list, so the fast download path stays fixed-size.
struct MultiPos { uint64_t pos; int32_t read; int32_t strand; };

__global__ void __launch_bounds__(256) k_gather_multi(const unsigned long long *keys, const int32_t *leader, uint32_t n_cand,
                                                      const gmx_read_result *results, int32_t read_base, MultiPos *out, uint32_t *count, uint32_t cap,
                                                      const uint32_t *live)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= gmx_live(live, n_cand)) return;
    const int32_t ld = leader[c];
    if (ld < 0) return;
    uint32_t task, round, diag;
    gmx_decode_key(keys[c], task, round, diag);
    const int r = (int)(task >> 1);
    const gmx_read_result &res = results[r];
    if (res.status != GMX_READ_MAPPED || res.best_n_positions <= 1 || res.best_group < 0) return;
    if (ld != res.hit_begin + res.best_group) return;
    const uint32_t at = atomicAdd(count, 1u);
    if (at < cap) { MultiPos m; m.pos = diag; m.read = read_base + r; m.strand = (int32_t)(task & 1); out[at] = m; }
}
