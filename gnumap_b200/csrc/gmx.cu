// gmx.cu -- host side of the C ABI declared in include/gmx.h: context, kernel-level entry points
// and the batched PHASE A / PHASE B pipeline.  sm_100a only; there is no CPU fallback -- without a
// CUDA device every entry point returns GMX_ERR_NO_DEVICE.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pipeline.cuh"
#include "pair_hmm.cuh"
#include "fastq.cuh"
#include "gmp_out.cuh"
#include "sam_out.cuh"

struct gmx_comm;

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// function-local scratch: freed on every return path (the context's own buffers are released by gmx_destroy)
struct ScratchBuf : DevBuf {
    ScratchBuf() = default;
    ScratchBuf(const ScratchBuf &) = delete;
    ScratchBuf &operator=(const ScratchBuf &) = delete;
    ~ScratchBuf() { release(); }
};

struct HostBuf {                                   // pinned host staging
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

#define GMX_CIGAR_STRIDE (ctx->cigar_stride)     // bytes per CIGAR slot (GMX_OPT_CIGAR_STRIDE, default 64)

enum Stage { ST_UPLOAD = 0, ST_PREP, ST_SEED, ST_CLASSIFY, ST_VOTE, ST_SORT, ST_NW, ST_FINALIZE, ST_TRACEBACK, ST_PHMM, ST_SCATTER, ST_DOWNLOAD, ST_COUNT };
static const char *kStageNames[ST_COUNT] = {"upload", "prep_reads", "seed_walk", "classify", "locate_vote", "sort_candidates",
                                            "nw_score", "finalize_reads", "nw_traceback", "pair_hmm", "scatter", "download"};

struct gmx_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    gmx_params params;
    DevParams dparams;
    DevIndex ix;
    DevTables tab;
    // index storage
    DevBuf d_bwt, d_sa_full, d_sa_samp, d_pac, d_seq_offset, d_tables, d_kmer_tab;
    // accumulators
    Accum acc;
    DevBuf d_amount, d_planes;
    uint64_t n_plane = 0;
    // reads of the current chunk
    DevBuf d_offsets[2], d_seq[2], d_qual[2], d_pwm[2];     // double-buffered: chunk i+1 uploads while chunk i computes
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t up_ev[2] = {nullptr, nullptr};
    cudaEvent_t done_ev[2] = {nullptr, nullptr};          // last kernel that reads buffer set [slot] has been issued
    DevReads up_view[2];
    int32_t up_max_len[2] = {0, 0};
    struct PendingScan { const gmx_reads *reads = nullptr; int32_t lo = 0, hi = 0; } up_scan[2];   // longest-read scan owed for a buffer set
    DevReads dreads;
    // pipeline buffers
    DevBuf d_fq_text, d_fq_nl, d_fq_tmp, d_fq_seq_off, d_fq_qual_off, d_fq_len, d_fq_recs, d_fq_flags, d_fq_count;   // FASTQ indexer
    DevBuf d_prep, d_seed_rank, d_seed_off, d_seed_n, d_seed_hits, d_cls_list, d_cls_meta;
    DevBuf d_keys, d_keys_alt, d_sort_tmp, d_score, d_leader, d_slot, d_lead_cand, d_hashes, d_expv, d_counters;
    DevBuf d_results, d_alen, d_aligned, d_cigar, d_hmm, d_moves, d_arena, d_phmm_scratch;
    size_t cand_cap = 0;
    // last-batch bookkeeping (whole batch = concatenation of chunks)
    bool mapped = false, scored = false;
    int32_t last_n_reads = 0;
    int32_t last_max_len = 0;
    std::vector<gmx_read_result> h_results;
    std::vector<gmx_hit> h_hits;
    int cigar_stride = 64;                     // GMX_OPT_CIGAR_STRIDE
    HostBuf h_best_cigar;                      // pinned [n_reads][cigar_stride]
    HostBuf h_batch_counts;                    // pinned: batch-level device counters (publish_batch_counts)
    HostBuf h_counters;                        // pinned staging of the per-chunk device counters, one per chunk parity
    std::vector<uint8_t> h_best_aligned;       // [n_reads][a_stride]  (collect_hits only)
    int h_a_stride = 0;
    // state of the chunk whose PHASE A results are resident on the device
    struct ChunkState {
        bool valid = false;
        int32_t lo = 0, n = 0, max_len = 0;
        int64_t total_bases = 0;
        uint32_t n_cand = 0, n_leaders = 0, n_accepted = 0;     // optimistic chunk: the bounds its grids and buffers cover
        const uint32_t *live_cand = nullptr, *live_lead = nullptr;   // optimistic chunk: device words holding the real counts
        bool opt = false;
        unsigned long long *keys = nullptr;    // sorted candidate keys (d_keys or d_keys_alt)
        LeaderStore L;
    } cs;
    // Optimistic chunks (phase_a): issued end to end without a host wait, over bounds predicted from the last settled
    // chunk; the host looks at a chunk's counters one chunk later (settle_chunk) and runs it again, synchronously, if it
    // did not fit.  Two may be in flight: the one just issued and its predecessor.
    struct Pending {
        bool active = false;
        int32_t lo = 0, hi = 0, max_len = 0; int slot = 0;
        gmx_reads reads; gmx_read_result *results = nullptr;
    } pend[2];
    cudaEvent_t settle_ev[2] = {nullptr, nullptr};
    double pred_cand = -1, pred_lead = -1;     // candidates / group leaders per read of the last settled chunk (< 0: none yet)
    double pred_margin = 1.0 / 16;             // head room over the prediction; doubled by every re-run
    bool stage_timing = true;                  // GMX_OPT_STAGE_TIMING: CUDA-event pairs around every stage
    int optimistic = 1;                        // GMX_OPT_OPTIMISTIC: 0 off, 1 on, 2 on with bounds that are too small (tests)
    uint64_t n_optimistic = 0, n_rerun = 0;
    // instrumentation (event pairs per chunk parity: a chunk's times are folded in when it is settled)
    int par = 0;
    cudaEvent_t ev[2][ST_COUNT][2];
    bool ev_used[2][ST_COUNT];
    float stage_ms[ST_COUNT];
    uint64_t stage_units[ST_COUNT], stage_bytes[ST_COUNT];
    int32_t stage_launches[ST_COUNT];
    std::string err;
    size_t chunk_reads = 1 << 19;              // measured: 524288 beats 262144 by 2.5 % and 1048576 (less transfer overlap)
    bool collect_hits = true;
    bool use_filter = true;                    // GMX_OPT_VOTE_FILTER
    int filter_shift = 0;                      // GMX_OPT_FILTER_SHIFT
    int vote_slots = GMX_VOTE_UNROLL;          // GMX_OPT_VOTE_SLOTS: 32-hit slots per step of the vote kernel
    int vote_compact = 2;                      // GMX_OPT_VOTE_COMPACT: 0 off, 1 two-bit variant, 2 three-bit variants (default)
    uint32_t class_hint = 0xfffu;              // vote classes (6 filter + 6 exact) that held tasks in the previous chunk
    int n_sm = 148;
    DevBuf d_ranges;                           // candidate range per read
    DevBuf d_groups, d_read_base;              // groups per read and their exclusive scan (leader slots)
    DevBuf d_multi, d_multi_count;             // (read, pos, strand) of multi-position best groups (fast path), whole batch
    uint32_t multi_cap = 0; bool multi_overflow = false;
    std::vector<MultiPos> h_multi;
    std::vector<int64_t> h_seq_offset;         // host copy of the sequence offsets (+ l_pac) for pos -> chromosome
    DevBuf d_batch_cigar, d_batch_results;     // fast download path: per-read records + best CIGARs of the WHOLE batch (they also
                                               // feed the device SAM formatter)
    bool batch_dev_valid = false;              // ... hold the batch last scored
    bool fq_batch = false;                     // the batch last scored came from gmx_process_fastq: text + record index are resident
    const char *fq_d_text = nullptr; int64_t fq_len = 0;
    bool sam_on_device = true;                 // GMX_OPT_SAM_DEVICE
    int64_t fq_piece_bytes = 96ll << 20;       // GMX_OPT_FASTQ_PIECE: bytes per piece of a pipelined host FASTQ text (0 = whole text)
    DevBuf d_sam_pieces, d_sam_cigar, d_sam_lens, d_sam_offs, d_sam_extra, d_sam_out, d_sam_names, d_sam_tmp;
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t gather_ev[2] = {nullptr, nullptr}, dl_ev[2] = {nullptr, nullptr};
    int dl_slot = 0;
    // input of the last multi-chunk gmx_map_batch (gmx_score_batch re-runs the batch from it)
    gmx_comm *comm = nullptr;                  // set by gmx_comm_create: this context's accumulators are one term of a sum
    gmx_reads keep; bool keep_valid = false;
    std::vector<int64_t> keep_offsets; std::vector<uint8_t> keep_seq, keep_qual; std::vector<float> keep_pwm;
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char b__[512];                                                                           \
            snprintf(b__, sizeof(b__), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            ctx->err = b__;                                                                          \
            return (e__ == cudaErrorMemoryAllocation) ? GMX_ERR_NOMEM : GMX_ERR_CUDA;                \
        }                                                                                            \
    } while (0)

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

static void stage_begin(gmx_ctx *ctx, int st) { if (ctx->stage_timing) cudaEventRecord(ctx->ev[ctx->par][st][0], ctx->stream); }
static void stage_end(gmx_ctx *ctx, int st, uint64_t units, uint64_t bytes, int launches)
{
    if (ctx->stage_timing) { cudaEventRecord(ctx->ev[ctx->par][st][1], ctx->stream); ctx->ev_used[ctx->par][st] = true; }
    ctx->stage_units[st] += units; ctx->stage_bytes[st] += bytes; ctx->stage_launches[st] += launches;
}
// fold the event pairs of a finished chunk (parity `par`, default: the current one) into the running totals
static void stage_collect(gmx_ctx *ctx, int par = -1)
{
    if (par < 0) par = ctx->par;
    for (int s = 0; s < ST_COUNT; ++s)
        if (ctx->ev_used[par][s]) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->ev[par][s][0], ctx->ev[par][s][1]) == cudaSuccess) ctx->stage_ms[s] += ms;
            ctx->ev_used[par][s] = false;
        }
}
static void stage_reset(gmx_ctx *ctx)
{
    for (int s = 0; s < ST_COUNT; ++s) {
        ctx->stage_ms[s] = 0; ctx->stage_units[s] = 0; ctx->stage_bytes[s] = 0; ctx->stage_launches[s] = 0;
        ctx->ev_used[0][s] = ctx->ev_used[1][s] = false;
    }
}

// ------------------------------------------------------------------------------------------------
// defaults + LUTs  (host arithmetic must not be contracted: compiled with -ffp-contract=off)
// ------------------------------------------------------------------------------------------------
static void fill_table(float T[256][4], float match, float transition, float transversion)
{
    static const char *lo = "acgt", *up = "ACGT";
    for (int i = 0; i < 256; ++i) for (int j = 0; j < 4; ++j) T[i][j] = transversion;
    for (int g = 0; g < 4; ++g)
        for (int b = 0; b < 4; ++b) {
            float v = (g == b) ? match : ((g ^ b) == 2 ? transition : transversion);
            T[(int)lo[g]][b] = T[(int)up[g]][b] = v;
        }
}

extern "C" void gmx_default_params(gmx_params *p)
{   // reference inc/const_define.h:46-107, inc/a_matrices.c:59-126
    memset(p, 0, sizeof(*p));
    float adjust = 0.25f, match = 3, transition = -2, transversion = -3, gap = -4;
    match *= adjust; transition *= adjust; transversion *= adjust; gap *= adjust;
    fill_table(p->align_scores, match, transition, transversion);
    fill_table(p->phmm_scores, 0.98f, 0.01f, 0.005f);
    p->gap = gap; p->max_gap = 3; p->mer = 10; p->jump = 5; p->min_seed_hits = 2;
    p->max_kmer_hits = 0; p->max_matches = 1000; p->gen_size = 8;
    p->align_score = 0.9f; p->perc = 1; p->cutoff = 0.0f;
    p->match_pos = 1; p->match_neg = 1; p->unique_only = 0; p->fast = 0; p->use_nw = 1;
    p->mode = GMX_MODE_NORMAL; p->illumina = 0; p->adjust = adjust;
}

// FASTQ (base, quality char) -> PWM row: reference src/SeqReader.cpp:618-627,1155-1216,1268-1271
static void pwm_row_host(int code, int qchar, int illumina, float out[4])
{
    int Q = qchar;
    double max_prb;
    if (illumina) { Q -= 64; double a = 1.0 - 1.0 / (pow(10.0, ((double)Q / 10.0))); max_prb = a > 1.0 ? 1.0 : a; }
    else { Q -= 33; double a = 1 - exp((-(double)Q / 10.0) * log(10.0)); max_prb = a > 1.0 ? 1.0 : a; }
    double other = (1 - max_prb) / 3;
    for (int b = 0; b < 4; ++b) out[b] = (float)((b == code) ? max_prb : other);
}

static float get_val_host(const float a[4], const float s[4])
{   // reference src/bin_seq.cpp:975-987
    volatile float t0 = a[0] * s[0], t1 = a[1] * s[1], t2 = a[2] * s[2], t3 = a[3] * s[3];
    volatile float r = t0 + t1; r = r + t2; r = r + t3;
    return r;
}

static float p_seq_host(const float x[4], const float P[4])
{   // reference src/bin_seq.cpp:41-57
    volatile float sum = 0;
    volatile float t;
    t = x[0] * P[0]; sum = sum + t;
    t = x[1] * P[1]; sum = sum + t;
    t = x[2] * P[2]; sum = sum + t;
    t = x[3] * P[3]; sum = sum + t;
    volatile float r = 3 * sum;
    return r;
}

static const size_t kLutFloats = 5 * GMX_NQ * 4;

static int build_tables(gmx_ctx *ctx)
{
    const gmx_params &p = ctx->params;
    // layout: sub_pos | sub_neg | pwm_lut | phmm_pos | phmm_neg | S | P | self
    std::vector<float> h(5 * kLutFloats + 2 * 1024 + 256 * GMX_NQ);
    float *sub_pos = h.data(), *sub_neg = sub_pos + kLutFloats, *pwm_lut = sub_neg + kLutFloats;
    float *phmm_pos = pwm_lut + kLutFloats, *phmm_neg = phmm_pos + kLutFloats, *S = phmm_neg + kLutFloats, *P = S + 1024;
    for (int code = 0; code < 5; ++code)
        for (int q = 0; q < GMX_NQ; ++q) {
            float row[4], rc[4];
            pwm_row_host(code, q + GMX_QMIN, p.illumina, row);
            rc[0] = row[3]; rc[1] = row[2]; rc[2] = row[1]; rc[3] = row[0];
            size_t o = ((size_t)code * GMX_NQ + q) * 4;
            for (int g = 0; g < 4; ++g) {
                int ch = "acgt"[g];
                pwm_lut[o + g] = row[g];
                sub_pos[o + g] = get_val_host(row, p.align_scores[ch]);
                sub_neg[o + g] = get_val_host(rc, p.align_scores[ch]);
                phmm_pos[o + g] = p_seq_host(row, p.phmm_scores[ch]);
                phmm_neg[o + g] = p_seq_host(rc, p.phmm_scores[ch]);
            }
        }
    memcpy(S, p.align_scores, sizeof(float) * 1024);
    memcpy(P, p.phmm_scores, sizeof(float) * 1024);
    float *self = P + 1024;
    for (int ch = 0; ch < 256; ++ch) {
        int code = 4;
        switch (ch) { case 'A': case 'a': code = 0; break; case 'C': case 'c': code = 1; break; case 'G': case 'g': code = 2; break; case 'T': case 't': code = 3; break; }
        for (int q = 0; q < GMX_NQ; ++q) self[ch * GMX_NQ + q] = get_val_host(pwm_lut + ((size_t)code * GMX_NQ + q) * 4, p.align_scores[ch]);
    }
    CK(ctx->d_tables.ensure(h.size() * sizeof(float)));
    CK(cudaMemcpyAsync(ctx->d_tables.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float *d = ctx->d_tables.as<float>();
    ctx->tab.sub_pos = d; ctx->tab.sub_neg = d + kLutFloats; ctx->tab.pwm_lut = d + 2 * kLutFloats;
    ctx->tab.phmm_pos = d + 3 * kLutFloats; ctx->tab.phmm_neg = d + 4 * kLutFloats;
    ctx->tab.S = d + 5 * kLutFloats; ctx->tab.P = d + 5 * kLutFloats + 1024; ctx->tab.self = d + 5 * kLutFloats + 2048;
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" const char *gmx_strerror(int code)
{
    switch (code) {
        case GMX_OK: return "ok";
        case GMX_ERR_INVALID: return "invalid argument";
        case GMX_ERR_CUDA: return "CUDA error";
        case GMX_ERR_NOMEM: return "out of memory";
        case GMX_ERR_UNSUPPORTED: return "unsupported option combination";
        case GMX_ERR_OVERFLOW: return "device work list overflow";
        case GMX_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case GMX_ERR_STATE: return "invalid call sequence";
        case GMX_ERR_FORMAT: return "malformed FASTQ text";
        default: return "unknown error";
    }
}
extern "C" const char *gmx_last_error(const gmx_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
extern "C" int gmx_abi_version(void) { return GMX_ABI_VERSION; }

static int validate_params(const gmx_params *p, std::string &why)
{
    if (!p->use_nw) { why = "use_nw = 0 (--no_nw) is not implemented on the device path"; return GMX_ERR_UNSUPPORTED; }
    if (p->mer < 1 || p->mer > 31) { why = "mer must be in 1..31"; return GMX_ERR_INVALID; }
    if (p->jump < 1) { why = "jump must be >= 1"; return GMX_ERR_INVALID; }
    if (p->max_gap < 1 || p->max_gap > 7) { why = "max_gap must be in 1..7"; return GMX_ERR_UNSUPPORTED; }
    if (p->min_seed_hits < 1 || p->min_seed_hits > 200) { why = "min_seed_hits must be in 1..200"; return GMX_ERR_UNSUPPORTED; }
    if (p->gen_size < 1) { why = "gen_size must be >= 1"; return GMX_ERR_INVALID; }
    if (p->mode < 0 || p->mode > 2) { why = "bad mode"; return GMX_ERR_INVALID; }
    return GMX_OK;
}

extern "C" int gmx_create(gmx_ctx **out, const gmx_index *index, const gmx_params *params, int device)
{
    if (!out || !index || !params) return GMX_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return GMX_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return GMX_ERR_INVALID;
    std::string why;
    int vr = validate_params(params, why);
    gmx_ctx *ctx = new gmx_ctx();
    *out = ctx;                                    // returned even on failure so gmx_last_error works
    ctx->device = device;
    ctx->params = *params;
    if (vr != GMX_OK) { ctx->err = why; return vr; }
    if (index->seq_len >= 0xFFFFFFF0ull || (uint64_t)index->l_pac != index->seq_len) {
        ctx->err = "index must be forward-only with seq_len < 2^32 - 16"; return GMX_ERR_UNSUPPORTED;
    }
    if (index->sa_intv <= 0 || (index->sa_intv & (index->sa_intv - 1))) { ctx->err = "sa_intv must be a power of two"; return GMX_ERR_INVALID; }
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) { CK(cudaEventCreateWithFlags(&ctx->up_ev[b], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->done_ev[b], cudaEventDisableTiming)); }
    CK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) { CK(cudaEventCreateWithFlags(&ctx->gather_ev[b], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->dl_ev[b], cudaEventDisableTiming)); }
    for (int b = 0; b < 2; ++b) {
        for (int s = 0; s < ST_COUNT; ++s) { CK(cudaEventCreate(&ctx->ev[b][s][0])); CK(cudaEventCreate(&ctx->ev[b][s][1])); }
        CK(cudaEventCreateWithFlags(&ctx->settle_ev[b], cudaEventDisableTiming));
    }
    stage_reset(ctx);
    cudaDeviceGetAttribute(&ctx->n_sm, cudaDevAttrMultiProcessorCount, device);

    DevParams &dp = ctx->dparams;
    dp.gap = params->gap; dp.align_score = params->align_score; dp.cutoff = params->cutoff; dp.max_gap = params->max_gap;
    dp.mer = params->mer; dp.jump = params->jump; dp.kmin = params->min_seed_hits; dp.perc = params->perc;
    dp.match_pos = params->match_pos; dp.match_neg = params->match_neg; dp.unique_only = params->unique_only;
    dp.fast = params->fast; dp.mode = params->mode; dp.max_kmer_hits = params->max_kmer_hits;
    dp.max_matches = params->max_matches; dp.gen_size = params->gen_size;

    // index upload
    size_t bwt_bytes = index->bwt_words * 4;
    CK(ctx->d_bwt.ensure(bwt_bytes + 64));
    CK(cudaMemcpyAsync(ctx->d_bwt.p, index->bwt, bwt_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_sa_samp.ensure(index->n_sa * 8));
    CK(cudaMemcpyAsync(ctx->d_sa_samp.p, index->sa, index->n_sa * 8, cudaMemcpyHostToDevice, ctx->stream));
    size_t pac_bytes = ((size_t)index->l_pac + 3) / 4;
    CK(ctx->d_pac.ensure(pac_bytes + 16));
    CK(cudaMemsetAsync(ctx->d_pac.p, 0, pac_bytes + 16, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_pac.p, index->pac, pac_bytes, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int64_t> offs(index->n_seqs + 1);
    for (int i = 0; i < index->n_seqs; ++i) offs[i] = index->seq_offset[i];
    offs[index->n_seqs] = index->l_pac;
    ctx->h_seq_offset = offs;
    CK(ctx->d_seq_offset.ensure(offs.size() * 8));
    CK(cudaMemcpyAsync(ctx->d_seq_offset.p, offs.data(), offs.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->d_sa_full.ensure((index->seq_len + 1) * 4));
    CK(cudaStreamSynchronize(ctx->stream));

    DevIndex &ix = ctx->ix;
    ix.bwt = ctx->d_bwt.as<uint32_t>(); ix.sa_full = ctx->d_sa_full.as<uint32_t>(); ix.sa_samp = ctx->d_sa_samp.as<uint64_t>();
    ix.pac = ctx->d_pac.as<uint8_t>(); ix.seq_offset = ctx->d_seq_offset.as<int64_t>();
    ix.primary = index->primary; ix.seq_len = index->seq_len;
    for (int i = 0; i < 5; ++i) ix.L2[i] = index->L2[i];
    ix.l_pac = index->l_pac; ix.sa_intv = index->sa_intv; ix.n_seqs = index->n_seqs;

    // de-sample the suffix array: sa_full[k] == bwt_sa(k) for every rank
    k_desample_sa<<<nblk((int64_t)index->seq_len + 1, 256), 256, 0, ctx->stream>>>(ix, ctx->d_sa_full.as<uint32_t>());
    CK(cudaGetLastError());

    // memoise the first min(mer, 12) backward-search steps of every k-mer lookup
    ix.kmer_tab = nullptr; ix.tab_len = 0;
    {
        const int T = std::min<int>(params->mer, GMX_KMER_TAB_MAX);
        const size_t n_tab = (size_t)1 << (2 * T);
        CK(ctx->d_kmer_tab.ensure(n_tab * sizeof(uint2)));
        k_build_kmer_table<<<nblk((int64_t)n_tab, 256), 256, 0, ctx->stream>>>(ix, T, ctx->d_kmer_tab.as<uint2>());
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
        ix.kmer_tab = ctx->d_kmer_tab.as<uint2>(); ix.tab_len = T;
    }

    int r = build_tables(ctx);
    if (r != GMX_OK) return r;
    {   // vote kernel: slots per step from the expected SA hits of one k-mer
        const double per_kmer = (double)index->seq_len / pow(4.0, (double)std::min(params->mer, 15));
        ctx->vote_slots = per_kmer * 1.15 > 32.0 * GMX_VOTE_UNROLL ? 6 : GMX_VOTE_UNROLL;
    }

    // accumulators (zeroed: neutralises the reference's un-initialised malloc)
    ctx->acc.n_amount = ((uint64_t)index->l_pac + params->gen_size - 1) / params->gen_size;
    CK(ctx->d_amount.ensure(ctx->acc.n_amount * 4));
    ctx->acc.amount = ctx->d_amount.as<float>();
    for (int b = 0; b < 5; ++b) ctx->acc.planes[b] = nullptr;
    if (params->mode != GMX_MODE_NORMAL) {
        ctx->n_plane = ctx->acc.n_amount;
        CK(ctx->d_planes.ensure(ctx->n_plane * 4 * 5));
        for (int b = 0; b < 5; ++b) ctx->acc.planes[b] = ctx->d_planes.as<float>() + (size_t)b * ctx->n_plane;
    }
    r = gmx_reset_accumulators(ctx);
    if (r != GMX_OK) return r;
    CK(cudaStreamSynchronize(ctx->stream));
    return GMX_OK;
}

static void comm_detach(gmx_ctx *ctx);

extern "C" void gmx_destroy(gmx_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->comm) comm_detach(ctx);
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_bwt, &ctx->d_sa_full, &ctx->d_sa_samp, &ctx->d_pac, &ctx->d_seq_offset, &ctx->d_tables, &ctx->d_amount,
                      &ctx->d_planes, &ctx->d_offsets[0], &ctx->d_seq[0], &ctx->d_qual[0], &ctx->d_pwm[0], &ctx->d_offsets[1], &ctx->d_seq[1], &ctx->d_qual[1], &ctx->d_pwm[1], &ctx->d_prep, &ctx->d_seed_rank,
                      &ctx->d_seed_off, &ctx->d_seed_n, &ctx->d_seed_hits, &ctx->d_cls_list, &ctx->d_cls_meta,
                      &ctx->d_keys, &ctx->d_keys_alt, &ctx->d_sort_tmp, &ctx->d_score, &ctx->d_leader, &ctx->d_slot, &ctx->d_lead_cand,
                      &ctx->d_hashes, &ctx->d_expv, &ctx->d_counters, &ctx->d_results, &ctx->d_alen, &ctx->d_aligned, &ctx->d_cigar,
                      &ctx->d_hmm, &ctx->d_moves, &ctx->d_arena, &ctx->d_phmm_scratch, &ctx->d_batch_cigar, &ctx->d_batch_results, &ctx->d_sam_pieces, &ctx->d_sam_cigar, &ctx->d_sam_lens, &ctx->d_sam_offs, &ctx->d_sam_extra, &ctx->d_sam_out, &ctx->d_sam_names, &ctx->d_sam_tmp, &ctx->d_kmer_tab, &ctx->d_multi, &ctx->d_multi_count, &ctx->d_ranges, &ctx->d_groups, &ctx->d_read_base, &ctx->d_fq_text, &ctx->d_fq_nl, &ctx->d_fq_tmp, &ctx->d_fq_seq_off, &ctx->d_fq_qual_off, &ctx->d_fq_len, &ctx->d_fq_recs, &ctx->d_fq_flags, &ctx->d_fq_count};
    for (DevBuf *b : bufs) b->release();
    ctx->h_best_cigar.release();
    ctx->h_counters.release();
    ctx->h_batch_counts.release();
    for (int b = 0; b < 2; ++b) {
        for (int s = 0; s < ST_COUNT; ++s) { if (ctx->ev[b][s][0]) cudaEventDestroy(ctx->ev[b][s][0]); if (ctx->ev[b][s][1]) cudaEventDestroy(ctx->ev[b][s][1]); }
        if (ctx->settle_ev[b]) cudaEventDestroy(ctx->settle_ev[b]);
    }
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->d2h_stream) { cudaStreamSynchronize(ctx->d2h_stream); cudaStreamDestroy(ctx->d2h_stream); }
    for (int b = 0; b < 2; ++b) { if (ctx->done_ev[b]) cudaEventDestroy(ctx->done_ev[b]); if (ctx->up_ev[b]) cudaEventDestroy(ctx->up_ev[b]); if (ctx->gather_ev[b]) cudaEventDestroy(ctx->gather_ev[b]); if (ctx->dl_ev[b]) cudaEventDestroy(ctx->dl_ev[b]); }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int gmx_set_stream(gmx_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return GMX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)cuda_stream;
    return GMX_OK;
}

extern "C" int gmx_synchronize(gmx_ctx *ctx)
{
    if (!ctx) return GMX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return GMX_OK;
}

extern "C" int gmx_reset_accumulators(gmx_ctx *ctx)
{
    if (!ctx) return GMX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->acc.amount, 0, ctx->acc.n_amount * 4, ctx->stream));
    if (ctx->acc.planes[0]) CK(cudaMemsetAsync(ctx->acc.planes[0], 0, ctx->n_plane * 4 * 5, ctx->stream));
    return GMX_OK;
}

extern "C" int gmx_accumulators_device(gmx_ctx *ctx, void **amount, uint64_t *n_amount, void *planes[5], uint64_t *n_plane)
{
    if (!ctx) return GMX_ERR_INVALID;
    if (amount) *amount = ctx->acc.amount;
    if (n_amount) *n_amount = ctx->acc.n_amount;
    if (planes) for (int b = 0; b < 5; ++b) planes[b] = ctx->acc.planes[b];
    if (n_plane) *n_plane = ctx->acc.planes[0] ? ctx->n_plane : 0;
    return GMX_OK;
}

static int comm_reduce_for_finish(gmx_ctx *ctx);

extern "C" int gmx_finish(gmx_ctx *ctx, float *amount_genome, float *const planes[5])
{
    if (!ctx || !amount_genome) return GMX_ERR_INVALID;
    if (ctx->comm) { int r = comm_reduce_for_finish(ctx); if (r != GMX_OK) return r; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(amount_genome, ctx->acc.amount, ctx->acc.n_amount * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (planes && ctx->acc.planes[0])
        for (int b = 0; b < 5; ++b)
            if (planes[b]) CK(cudaMemcpyAsync(planes[b], ctx->acc.planes[b], ctx->n_plane * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// reads upload
// ------------------------------------------------------------------------------------------------
static int scan_max_len(gmx_ctx *ctx, const gmx_reads *reads, int32_t lo, int32_t hi, int32_t *max_len_out)
{
    int32_t max_len = reads->max_len;
    if (!reads->on_device) {
        max_len = 0;
        for (int32_t i = lo; i < hi; ++i) {
            const int64_t len = reads->offsets[i + 1] - reads->offsets[i];
            if (len < 0) { ctx->err = "gmx_reads.offsets must be non-decreasing"; return GMX_ERR_INVALID; }
            if (len > GMX_MAX_READ_LEN) { ctx->err = "read longer than GMX_MAX_READ_LEN"; return GMX_ERR_UNSUPPORTED; }
            max_len = std::max<int32_t>(max_len, (int32_t)len);
        }
    } else if (max_len <= 0) { ctx->err = "device-resident gmx_reads need max_len"; return GMX_ERR_INVALID; }
    if (max_len > GMX_MAX_READ_LEN) { ctx->err = "read longer than GMX_MAX_READ_LEN"; return GMX_ERR_UNSUPPORTED; }
    *max_len_out = max_len;
    return GMX_OK;
}

// Make reads [lo, hi) visible to the kernels through buffer set `slot`.  Offsets stay absolute (relative to the
// start of the batch's seq/qual arrays); for host batches only the chunk's slice is copied and the device base
// pointers are biased so that the same offsets index it.  Copies are issued on `stream`; the view is left in
// ctx->up_view[slot] and ctx->up_ev[slot] is recorded behind them.
static int issue_upload(gmx_ctx *ctx, const gmx_reads *reads, int32_t lo, int32_t hi, int slot, cudaStream_t stream)
{
    int32_t n = hi - lo;
    if (!reads->seq) { ctx->err = "gmx_reads.seq is required (the consensus string for raw-PWM reads)"; return GMX_ERR_INVALID; }
    if (!reads->qual && !reads->pwm) { ctx->err = "gmx_reads needs qual or pwm"; return GMX_ERR_INVALID; }
    DevReads &v = ctx->up_view[slot];
    v.n_reads = n; v.qbase = ctx->params.illumina ? 64 : 33;
    v.qual = nullptr; v.pwm = nullptr; v.qoffsets = nullptr; v.lens = nullptr; v.max_len = 0;
    if (reads->on_device) {
        v.max_len = reads->max_len;
        v.offsets = reads->offsets + lo;
        v.seq = reads->seq; v.qual = reads->qual; v.pwm = reads->pwm;
        v.qoffsets = reads->qual_offsets ? reads->qual_offsets + lo : nullptr;
        v.lens = reads->lens ? reads->lens + lo : nullptr;
    } else {
        int64_t base = reads->offsets[lo], total = reads->offsets[hi] - base;
        CK(ctx->d_offsets[slot].ensure(((size_t)n + 1) * 8));
        CK(ctx->d_seq[slot].ensure((size_t)total + 16));
        CK(cudaMemcpyAsync(ctx->d_offsets[slot].p, reads->offsets + lo, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(ctx->d_seq[slot].p, reads->seq + base, (size_t)total, cudaMemcpyHostToDevice, stream));
        v.offsets = ctx->d_offsets[slot].as<int64_t>();
        v.seq = ctx->d_seq[slot].as<uint8_t>() - base;
        if (reads->qual) {
            CK(ctx->d_qual[slot].ensure((size_t)total + 16));
            CK(cudaMemcpyAsync(ctx->d_qual[slot].p, reads->qual + base, (size_t)total, cudaMemcpyHostToDevice, stream));
            v.qual = ctx->d_qual[slot].as<uint8_t>() - base;
        }
        if (reads->pwm) {
            CK(ctx->d_pwm[slot].ensure((size_t)total * 16 + 16));
            CK(cudaMemcpyAsync(ctx->d_pwm[slot].p, reads->pwm + 4 * base, (size_t)total * 16, cudaMemcpyHostToDevice, stream));
            v.pwm = ctx->d_pwm[slot].as<float>() - 4 * base;
        }
    }
    CK(cudaEventRecord(ctx->up_ev[slot], stream));
    // the chunk's longest read is found on the host later, at a point where the GPU has work queued (finish_scan)
    ctx->up_scan[slot].reads = reads; ctx->up_scan[slot].lo = lo; ctx->up_scan[slot].hi = hi;
    return GMX_OK;
}

static int finish_scan(gmx_ctx *ctx, int slot)
{
    gmx_ctx::PendingScan &p = ctx->up_scan[slot];
    if (!p.reads) return GMX_OK;
    int32_t max_len = 0;
    int r = scan_max_len(ctx, p.reads, p.lo, p.hi, &max_len);
    p.reads = nullptr;
    if (r != GMX_OK) return r;
    ctx->up_max_len[slot] = max_len;
    return GMX_OK;
}

// kernel-level entry points: upload on the compute stream into slot 0
static int upload_reads(gmx_ctx *ctx, const gmx_reads *reads, int32_t lo, int32_t hi, int32_t *max_len_out)
{
    int r = issue_upload(ctx, reads, lo, hi, 0, ctx->stream);
    if (r == GMX_OK) r = finish_scan(ctx, 0);
    if (r != GMX_OK) return r;
    ctx->dreads = ctx->up_view[0];
    if (max_len_out) *max_len_out = ctx->up_max_len[0];
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel-level entry points
// ------------------------------------------------------------------------------------------------
extern "C" int gmx_fm_search(gmx_ctx *ctx, const uint8_t *kmers, int32_t len, int64_t n, uint64_t *k_out, uint64_t *l_out)
{
    if (!ctx || !kmers || !k_out || !l_out || len < 1 || len > 64 || n < 0) return GMX_ERR_INVALID;
    if (n == 0) return GMX_OK;
    CK(cudaSetDevice(ctx->device));
    ScratchBuf in, ko, lo;
    CK(in.ensure((size_t)n * len)); CK(ko.ensure((size_t)n * 8)); CK(lo.ensure((size_t)n * 8));
    CK(cudaMemcpyAsync(in.p, kmers, (size_t)n * len, cudaMemcpyHostToDevice, ctx->stream));
    k_fm_search<<<nblk(n, 128), 128, 0, ctx->stream>>>(ctx->ix, in.as<uint8_t>(), len, n, ko.as<uint64_t>(), lo.as<uint64_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(k_out, ko.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(l_out, lo.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    in.release(); ko.release(); lo.release();
    return GMX_OK;
}

extern "C" int gmx_sa_locate(gmx_ctx *ctx, const uint64_t *ranks, int64_t n, int32_t mode, uint64_t *pos_out)
{
    if (!ctx || !ranks || !pos_out || n < 0 || (mode != 0 && mode != 1)) return GMX_ERR_INVALID;
    if (n == 0) return GMX_OK;
    for (int64_t i = 0; i < n; ++i) if (ranks[i] > ctx->ix.seq_len) { ctx->err = "rank out of range"; return GMX_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    ScratchBuf in, po;
    CK(in.ensure((size_t)n * 8)); CK(po.ensure((size_t)n * 8));
    CK(cudaMemcpyAsync(in.p, ranks, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_sa_locate<<<nblk(n, 128), 128, 0, ctx->stream>>>(ctx->ix, in.as<uint64_t>(), n, mode, po.as<uint64_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pos_out, po.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    in.release(); po.release();
    return GMX_OK;
}

extern "C" int gmx_get_windows(gmx_ctx *ctx, const uint64_t *begin, int64_t n, int32_t size, uint8_t *chars_out, int32_t *len_out)
{
    if (!ctx || !begin || !chars_out || !len_out || n < 0 || size < 1) return GMX_ERR_INVALID;
    if (n == 0) return GMX_OK;
    CK(cudaSetDevice(ctx->device));
    ScratchBuf in, ch, ln;
    CK(in.ensure((size_t)n * 8)); CK(ch.ensure((size_t)n * size)); CK(ln.ensure((size_t)n * 4));
    CK(cudaMemcpyAsync(in.p, begin, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_get_windows<<<nblk(n, 128), 128, 0, ctx->stream>>>(ctx->ix, in.as<uint64_t>(), n, size, ch.as<uint8_t>(), ln.as<int32_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(chars_out, ch.p, (size_t)n * size, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(len_out, ln.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    in.release(); ch.release(); ln.release();
    return GMX_OK;
}

extern "C" int gmx_self_score(gmx_ctx *ctx, const gmx_reads *reads, float *score_out)
{
    if (!ctx || !reads || !score_out || reads->n_reads < 0) return GMX_ERR_INVALID;
    if (reads->n_reads == 0) return GMX_OK;
    CK(cudaSetDevice(ctx->device));
    int r = upload_reads(ctx, reads, 0, reads->n_reads, nullptr);
    if (r != GMX_OK) return r;
    int n = reads->n_reads;
    CK(ctx->d_prep.ensure((size_t)n * sizeof(ReadPrep)));
    DevParams dp = ctx->dparams; dp.mer = 0;             // score every read regardless of length
    dp.cutoff = -INFINITY;
    k_prep_reads<<<nblk(n, GMX_PREP_THREADS), GMX_PREP_THREADS, 0, ctx->stream>>>(ctx->dreads, ctx->tab, dp, ctx->d_prep.as<ReadPrep>(), nullptr);
    CK(cudaGetLastError());
    std::vector<ReadPrep> h(n);
    CK(cudaMemcpyAsync(h.data(), ctx->d_prep.p, (size_t)n * sizeof(ReadPrep), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i) score_out[i] = h[i].max_align;
    return GMX_OK;
}

// explicit-window task kernels -------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_nw_score_tasks(DevReads R, DevTables T, DevParams P, int64_t n_tasks, const int32_t *read_idx,
                                                        const uint8_t *strand, const uint8_t *windows, int win_stride, float *out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    ReadView rd = gmx_read_view(R, read_idx[t], strand ? strand[t] : 0);
    WindowView win; win.pac = nullptr; win.pos = 0; win.chars = windows + t * win_stride;
    out[t] = gmx_nw_score_dispatch(rd, win, T, P.gap, P.max_gap);
}

__global__ void __launch_bounds__(128) k_traceback_tasks(DevReads R, DevTables T, DevParams P, int64_t n_tasks, const int32_t *read_idx,
                                                         const uint8_t *strand, const uint8_t *windows, int win_stride, const uint8_t *consensus,
                                                         uint8_t *aligned, int a_stride, int32_t *alen, char *cigar, int c_stride, uint32_t *moves)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    ReadView rd = gmx_read_view(R, read_idx[t], strand ? strand[t] : 0);
    WindowView win; win.pac = nullptr; win.pos = 0; win.chars = windows + t * win_stride;
    ConsView cons; cons.explicit_chars = consensus ? consensus + t * win_stride : nullptr;
    TracebackOut o; o.aligned = aligned + t * a_stride; o.aligned_cap = a_stride; o.cigar = cigar + t * c_stride; o.cigar_cap = c_stride; o.fix_deletions = 0;
    o.truncated = nullptr;
    alen[t] = gmx_nw_traceback(rd, win, cons, T, P.gap, P.max_gap, moves + t, n_tasks, o);
}

struct TaskUpload {
    ScratchBuf ridx, strand, windows, cons;
};

static int upload_tasks(gmx_ctx *ctx, TaskUpload &u, const gmx_reads *reads, int64_t n_tasks, const int32_t *read_idx, const uint8_t *strand,
                        const uint8_t *windows, int32_t win_stride, const uint8_t *consensus, int32_t *max_len)
{
    for (int64_t t = 0; t < n_tasks; ++t) {
        if (read_idx[t] < 0 || read_idx[t] >= reads->n_reads) { ctx->err = "read_idx out of range"; return GMX_ERR_INVALID; }
        int64_t len = reads->offsets[read_idx[t] + 1] - reads->offsets[read_idx[t]];
        if (len > win_stride) { ctx->err = "win_stride shorter than a read"; return GMX_ERR_INVALID; }
    }
    int r = upload_reads(ctx, reads, 0, reads->n_reads, max_len);
    if (r != GMX_OK) return r;
    CK(u.ridx.ensure((size_t)n_tasks * 4)); CK(u.windows.ensure((size_t)n_tasks * win_stride + 16));
    CK(cudaMemcpyAsync(u.ridx.p, read_idx, (size_t)n_tasks * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(u.windows.p, windows, (size_t)n_tasks * win_stride, cudaMemcpyHostToDevice, ctx->stream));
    if (strand) { CK(u.strand.ensure((size_t)n_tasks)); CK(cudaMemcpyAsync(u.strand.p, strand, (size_t)n_tasks, cudaMemcpyHostToDevice, ctx->stream)); }
    if (consensus) { CK(u.cons.ensure((size_t)n_tasks * win_stride + 16)); CK(cudaMemcpyAsync(u.cons.p, consensus, (size_t)n_tasks * win_stride, cudaMemcpyHostToDevice, ctx->stream)); }
    return GMX_OK;
}

extern "C" int gmx_nw_score(gmx_ctx *ctx, const gmx_reads *reads, int64_t n_tasks, const int32_t *read_idx, const uint8_t *strand,
                            const uint8_t *windows, int32_t win_stride, float *score_out)
{
    if (!ctx || !reads || !read_idx || !windows || !score_out || n_tasks < 0 || win_stride < 1) return GMX_ERR_INVALID;
    if (n_tasks == 0) return GMX_OK;
    CK(cudaSetDevice(ctx->device));
    TaskUpload u; ScratchBuf out;
    int r = upload_tasks(ctx, u, reads, n_tasks, read_idx, strand, windows, win_stride, nullptr, nullptr);
    if (r != GMX_OK) return r;
    CK(out.ensure((size_t)n_tasks * 4));
    k_nw_score_tasks<<<nblk(n_tasks, 128), 128, 0, ctx->stream>>>(ctx->dreads, ctx->tab, ctx->dparams, n_tasks, u.ridx.as<int32_t>(),
                                                                 strand ? u.strand.as<uint8_t>() : nullptr, u.windows.as<uint8_t>(), win_stride, out.as<float>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(score_out, out.p, (size_t)n_tasks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    u.ridx.release(); u.strand.release(); u.windows.release(); u.cons.release(); out.release();
    return GMX_OK;
}

extern "C" int gmx_nw_traceback(gmx_ctx *ctx, const gmx_reads *reads, int64_t n_tasks, const int32_t *read_idx, const uint8_t *strand,
                                const uint8_t *windows, int32_t win_stride, const uint8_t *consensus, uint8_t *aligned_out,
                                int32_t aligned_stride, int32_t *aligned_len_out, char *cigar_out, int32_t cigar_stride)
{
    if (!ctx || !reads || !read_idx || !windows || !aligned_out || !aligned_len_out || !cigar_out || n_tasks < 0 || win_stride < 1 ||
        aligned_stride < 1 || cigar_stride < 2)
        return GMX_ERR_INVALID;
    if (n_tasks == 0) return GMX_OK;
    CK(cudaSetDevice(ctx->device));
    TaskUpload u; ScratchBuf al, ln, cg, mv;
    int32_t max_len = 0;
    int r = upload_tasks(ctx, u, reads, n_tasks, read_idx, strand, windows, win_stride, consensus, &max_len);
    if (r != GMX_OK) return r;
    CK(al.ensure((size_t)n_tasks * aligned_stride)); CK(ln.ensure((size_t)n_tasks * 4)); CK(cg.ensure((size_t)n_tasks * cigar_stride));
    CK(mv.ensure((size_t)n_tasks * ((size_t)max_len + 1) * 4));
    CK(cudaMemsetAsync(al.p, 0, (size_t)n_tasks * aligned_stride, ctx->stream));
    k_traceback_tasks<<<nblk(n_tasks, 128), 128, 0, ctx->stream>>>(ctx->dreads, ctx->tab, ctx->dparams, n_tasks, u.ridx.as<int32_t>(),
                                                                  strand ? u.strand.as<uint8_t>() : nullptr, u.windows.as<uint8_t>(), win_stride,
                                                                  consensus ? u.cons.as<uint8_t>() : nullptr, al.as<uint8_t>(), aligned_stride,
                                                                  ln.as<int32_t>(), cg.as<char>(), cigar_stride, mv.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(aligned_out, al.p, (size_t)n_tasks * aligned_stride, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(aligned_len_out, ln.p, (size_t)n_tasks * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cigar_out, cg.p, (size_t)n_tasks * cigar_stride, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    u.ridx.release(); u.strand.release(); u.windows.release(); u.cons.release(); al.release(); ln.release(); cg.release(); mv.release();
    return GMX_OK;
}

extern "C" int gmx_pair_hmm(gmx_ctx *ctx, const gmx_reads *reads, int64_t n_tasks, const int32_t *read_idx, const uint8_t *strand,
                            const uint8_t *windows, int32_t win_stride, float *post_out)
{
    if (!ctx || !reads || !read_idx || !windows || !post_out || n_tasks < 0 || win_stride < 1) return GMX_ERR_INVALID;
    if (n_tasks == 0) return GMX_OK;
    CK(cudaSetDevice(ctx->device));
    TaskUpload u; ScratchBuf out, scratch;
    int32_t max_len = 0;
    int r = upload_tasks(ctx, u, reads, n_tasks, read_idx, strand, windows, win_stride, nullptr, &max_len);
    if (r != GMX_OK) return r;
    if (max_len > 32 * GMX_PHMM_LONGC) { ctx->err = "reads longer than 1024 bp are not supported by the pair-HMM kernel"; return GMX_ERR_UNSUPPORTED; }
    CK(out.ensure((size_t)n_tasks * win_stride * 5 * 4));
    CK(cudaMemsetAsync(out.p, 0, (size_t)n_tasks * win_stride * 5 * 4, ctx->stream));
    size_t per_task = gmx_phmm_scratch_doubles(max_len);
    int64_t wave = std::min<int64_t>(n_tasks, 148 * 16);
    wave = std::max<int64_t>(1, std::min<int64_t>(wave, (int64_t)((4ull << 30) / (per_task * 8))));      // long reads: 17 MB of scratch per task at 1024 bp
    CK(scratch.ensure((size_t)wave * per_task * 8));
    for (int64_t t0 = 0; t0 < n_tasks; t0 += wave) {
        int64_t cnt = std::min<int64_t>(wave, n_tasks - t0);
        k_pair_hmm_tasks<<<(unsigned)cnt, GMX_PHMM_THREADS, 0, ctx->stream>>>(ctx->dreads, ctx->tab, t0, cnt, u.ridx.as<int32_t>(),
                                                                              strand ? u.strand.as<uint8_t>() : nullptr, u.windows.as<uint8_t>(),
                                                                              win_stride, out.as<float>(), scratch.as<double>(), per_task);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(post_out, out.p, (size_t)n_tasks * win_stride * 5 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    u.ridx.release(); u.strand.release(); u.windows.release(); u.cons.release(); out.release(); scratch.release();
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// batch pipeline
// ------------------------------------------------------------------------------------------------
struct Counters {            // device-resident scalars, reset before every vote attempt
    uint32_t n_cand, cand_overflow, n_leaders, n_accepted, arena_overflow;
    uint32_t live_cand, live_lead, bad;                                  // optimistic chunks: k_seal_candidates / k_seal_leaders
    unsigned long long arena_used;
    uint32_t cls_count[GMX_N_CLASSES], cls_cursor[GMX_N_CLASSES];        // exact classes
    uint32_t fcls_count[GMX_N_CLASSES], fcls_cursor[GMX_N_CLASSES];      // filter classes
};
struct ChunkStats {          // device-resident, reset once per chunk: lookups, search steps, SA hits
    unsigned long long v[4];
};
struct HostCounters { Counters c; ChunkStats s; };

__global__ void k_publish_words(const uint32_t *src, uint32_t *dst, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

static int settle_chunk(gmx_ctx *ctx, int par);
static int settle_all(gmx_ctx *ctx) { for (int b = 0; b < 2; ++b) { int r = settle_chunk(ctx, b); if (r != GMX_OK) return r; } return GMX_OK; }

// algorithmic work of the seed / vote stages (DESIGN.md "Roofline"): every backward-search step reads two 64-byte occ
// blocks; every SA hit reads one 4-byte entry of the de-sampled suffix array
static void account_seed_vote(gmx_ctx *ctx, const ChunkStats &s)
{
    ctx->stage_units[ST_SEED] += s.v[0]; ctx->stage_bytes[ST_SEED] += s.v[1] * 128ull + (ctx->ix.tab_len > 0 ? s.v[0] * 8ull : 0ull);
    ctx->stage_units[ST_VOTE] += s.v[2]; ctx->stage_bytes[ST_VOTE] += s.v[2] * 4ull;
}
static uint64_t scatter_bytes_per_hit(const DevParams &P, int max_len)
{
    // per accepted (position, strand): one float RMW per aligned base and plane, Normal mode after bin
    // aggregation: ceil(len / gen_size) RMWs
    return P.mode == GMX_MODE_NORMAL ? 8ull * ((uint64_t)(max_len + P.gen_size - 1) / P.gen_size) : (P.mode == GMX_MODE_BS ? 16ull : 48ull) * (uint64_t)max_len;
}

template <int SL, int WARPS>
static cudaError_t launch_vote(gmx_ctx *ctx, const SeedStore &S, const ClassLists &C, int cls, const CandSink &sink, int n_sm)
{
    size_t smem = (size_t)WARPS * ((1u << SL) + (1u << SL) / 4) * 4;
    cudaError_t e = cudaFuncSetAttribute(k_vote_smem<SL, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vote_smem<SL, WARPS>, WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    k_vote_smem<SL, WARPS><<<n_sm * per_sm, WARPS * 32, smem, ctx->stream>>>(ctx->ix, S, C, cls, ctx->dparams.kmin, sink);
    return cudaGetLastError();
}

template <int FL, int WARPS, bool BITS, int U = GMX_VOTE_UNROLL, int COMPACT = 0>
static cudaError_t launch_filter(gmx_ctx *ctx, const SeedStore &S, const ClassLists &F, const ClassLists &E, int cls, const CandSink &sink, int n_sm,
                                 uint32_t pac_words)
{
    size_t smem = (size_t)WARPS * (COMPACT ? gmx_filter_warp_bytes_compact(COMPACT) : gmx_filter_warp_bytes(FL));
    cudaError_t e = cudaFuncSetAttribute(k_vote_filter<FL, WARPS, BITS, U, COMPACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vote_filter<FL, WARPS, BITS, U, COMPACT>, WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    k_vote_filter<FL, WARPS, BITS, U, COMPACT><<<n_sm * per_sm, WARPS * 32, smem, ctx->stream>>>(ctx->ix, pac_words, S, F, E, cls, ctx->dparams.kmin,
                                                                                          ctx->dparams.mer, sink);
    return cudaGetLastError();
}

// PHASE A for reads [lo, hi) of the batch; leaves its results resident on the device (ctx->cs).
//
// `opt`: issue the chunk without waiting for its counts.  Needs a prediction (ctx->pred_*) from a settled chunk; the
// candidate list, the sort and every grid behind it are sized by the predicted bound, the kernels read the real counts on
// the device, and settle_chunk() checks one chunk later that the bounds held.
static int phase_a(gmx_ctx *ctx, const gmx_reads *reads, int32_t lo, int32_t hi, int slot, bool opt = false)
{
    gmx_ctx::ChunkState &cs = ctx->cs;
    cs.valid = false;
    if (!opt) { int r = settle_all(ctx); if (r != GMX_OK) return r; }      // a synchronous chunk reads the counters in place
    ctx->par = ctx->pend[0].active ? 1 : 0;
    if (ctx->pend[ctx->par].active) { int r = settle_chunk(ctx, ctx->par); if (r != GMX_OK) return r; }
    const int32_t n = hi - lo;
    const int64_t n_tasks = 2 * (int64_t)n;
    int32_t max_len = 0;
    const int n_sm = ctx->n_sm;

    // the chunk's reads were put in flight by run_batch (copy stream, buffer set `slot`); this stage is the wait
    stage_begin(ctx, ST_UPLOAD);
    CK(cudaStreamWaitEvent(ctx->stream, ctx->up_ev[slot], 0));
    ctx->dreads = ctx->up_view[slot];
    { int r = finish_scan(ctx, slot); if (r != GMX_OK) return r; }           // only the batch's first chunk still owes it here
    max_len = ctx->up_max_len[slot];
    int64_t total_bases = reads->on_device ? (int64_t)n * max_len : reads->offsets[hi] - reads->offsets[lo];
    uint64_t up_bytes = reads->on_device ? 0 : (uint64_t)total_bases * (reads->qual ? 2 : 1) + (reads->pwm ? 16ull * total_bases : 0) + 8ull * n;
    stage_end(ctx, ST_UPLOAD, (uint64_t)n, up_bytes, 0);
    ctx->last_max_len = std::max(ctx->last_max_len, max_len);

    const DevParams &P = ctx->dparams;
    int max_seeds = 1;
    if (max_len > P.mer) max_seeds = (max_len - P.mer + P.jump - 1) / P.jump + 1;
    if (max_seeds > GMX_MAX_SEEDS) { ctx->err = "too many k-mer rounds per read (read too long for this jump)"; return GMX_ERR_UNSUPPORTED; }

    // buffers
    CK(ctx->d_prep.ensure((size_t)n * sizeof(ReadPrep)));
    CK(ctx->d_seed_rank.ensure((size_t)max_seeds * n_tasks * 16));        // SeedStore::rec
    CK(ctx->d_seed_off.ensure((size_t)max_seeds * n_tasks * 2));
    CK(ctx->d_seed_n.ensure((size_t)n_tasks));
    CK(ctx->d_seed_hits.ensure((size_t)n_tasks * 4));
    CK(ctx->d_cls_list.ensure((size_t)2 * GMX_N_CLASSES * n_tasks * 4));
    CK(ctx->d_counters.ensure(sizeof(Counters) + sizeof(ChunkStats)));
    CK(ctx->d_results.ensure((size_t)n * sizeof(gmx_read_result)));
    if (ctx->cand_cap == 0) ctx->cand_cap = std::max<size_t>(1 << 16, (size_t)n * 16);

    Counters *dc = ctx->d_counters.as<Counters>();
    ChunkStats *ds = reinterpret_cast<ChunkStats *>(dc + 1);
    SeedStore S;
    S.rec = ctx->d_seed_rank.as<uint4>(); S.offset = ctx->d_seed_off.as<uint16_t>();
    S.n_seeds = ctx->d_seed_n.as<uint8_t>(); S.hits = ctx->d_seed_hits.as<uint32_t>(); S.max_seeds = max_seeds; S.n_tasks = n_tasks;
    ClassLists C;
    C.list = ctx->d_cls_list.as<uint32_t>(); C.count = dc->cls_count; C.cursor = dc->cls_cursor; C.n_tasks = n_tasks;
    ClassLists F;
    F.list = C.list + (size_t)GMX_N_CLASSES * n_tasks; F.count = dc->fcls_count; F.cursor = dc->fcls_cursor; F.n_tasks = n_tasks;
    const int use_filter = (P.kmin >= 2 && ctx->use_filter) ? 1 : 0;
    const uint32_t pac_words = (uint32_t)((((size_t)ctx->ix.l_pac + 3) / 4 + 16) / 4);

    // a3 + status
    stage_begin(ctx, ST_PREP);
    CK(cudaMemsetAsync(ds, 0, sizeof(ChunkStats), ctx->stream));
    k_prep_reads<<<nblk(n, GMX_PREP_THREADS), GMX_PREP_THREADS, 0, ctx->stream>>>(ctx->dreads, ctx->tab, P, ctx->d_prep.as<ReadPrep>(), &ds->v[3]);
    CK(cudaGetLastError());
    stage_end(ctx, ST_PREP, (uint64_t)n, (uint64_t)total_bases * 2, 1);

    // K1: k-mer walk + backward search
    stage_begin(ctx, ST_SEED);
    k_seed_walk<<<nblk(n_tasks, 128), 128, 0, ctx->stream>>>(ctx->ix, ctx->dreads, P, ctx->d_prep.as<ReadPrep>(), S, ds->v);
    CK(cudaGetLastError());
    stage_end(ctx, ST_SEED, 0, 0, 1);

    CK(ctx->h_counters.ensure(2 * sizeof(HostCounters)));
    HostCounters &hc = ctx->h_counters.as<HostCounters>()[ctx->par];   // pinned: the small D2H copies per chunk stay asynchronous
    uint32_t n_cand = 0, lead_cap = 0;
    if (opt) {
        const double room = ctx->optimistic == 2 ? 0.5 : 1.0 + ctx->pred_margin;
        const double slack = ctx->optimistic == 2 ? 0 : 16384;
        n_cand = (uint32_t)std::min<double>(0x7ffffff0, (double)n * ctx->pred_cand * room + slack + 2);
        lead_cap = (uint32_t)std::min<double>(n_cand, (double)n * ctx->pred_lead * room + slack + 1);
        ctx->cand_cap = std::max<size_t>(ctx->cand_cap, n_cand);
    }
    for (int attempt = 0;; ++attempt) {
        CK(cudaMemsetAsync(dc, 0, sizeof(Counters), ctx->stream));
        CK(ctx->d_keys.ensure(ctx->cand_cap * 8));
        CK(ctx->d_keys_alt.ensure(ctx->cand_cap * 8));
        stage_begin(ctx, ST_CLASSIFY);
        k_classify<<<nblk(n_tasks, 256), 256, 0, ctx->stream>>>(S.hits, F, C, use_filter);
        CK(cudaGetLastError());
        stage_end(ctx, ST_CLASSIFY, (uint64_t)n_tasks, (uint64_t)n_tasks * 8, 1);

        // K1b + K1c.  Twelve kernels cover the task classes (6 filter + 6 exact); a uniform workload fills one or two of
        // them.  Only the classes that held tasks in the previous chunk are launched up front.  A synchronous chunk then
        // reads the counters back and launches any class that turned out non-empty without having been launched; an
        // optimistic chunk lets k_seal_candidates void it in that case.
        CandSink sink; sink.keys = ctx->d_keys.as<unsigned long long>(); sink.count = &dc->n_cand; sink.overflow = &dc->cand_overflow; sink.cap = (uint32_t)ctx->cand_cap;
        auto launch_class = [&](int k) -> cudaError_t {         // k: 0..5 filter classes, 6..11 exact classes
            if (k < GMX_N_CLASSES) {
                const int c = k;
                if (!use_filter) return cudaSuccess;
                if (P.kmin == 2) {       // blocked Bloom filter over bits: (4 << filter_shift) bytes of filter per SA hit of the class
                    // measured on B200: occupancy beats a sparse filter -- 8 KB (20 warps/SM) up to 8k hits per task
                    static const int kBitsLog2[GMX_N_CLASSES] = {11, 12, 13, 13, 13, 15};
                    switch (std::min(std::max(kBitsLog2[c] + ctx->filter_shift, 10), 16)) {
                        case 10: return launch_filter<10, 8, true>(ctx, S, F, C, c, sink, n_sm, pac_words);
                        case 11: return launch_filter<11, 8, true>(ctx, S, F, C, c, sink, n_sm, pac_words);
                        case 12: return launch_filter<12, 8, true>(ctx, S, F, C, c, sink, n_sm, pac_words);
                        case 13:
                            // 32-hit slots per step for the expected hits of one k-mer (seq_len / 4^mer on a random genome:
                            // 95 at 100 Mb, 149 at 156 Mb for mer 10) plus head room
                            if (ctx->vote_compact && S.max_seeds <= 32) {
                                // tasks of at most 32 k-mers: the occupancy variants.  Up to 2 k hits per task (class 2) a 5 KB
                                // filter with three bits per diagonal (32 warps / SM); above, 7.4 KB (24 warps / SM)
                                const int variant = ctx->vote_compact >= 2 ? (c <= 2 ? 2 : 3) : 1;
                                const bool six = ctx->vote_slots >= 6;
                                if (variant == 2) return six ? launch_filter<13, 4, true, 6, 2>(ctx, S, F, C, c, sink, n_sm, pac_words) : launch_filter<13, 4, true, 4, 2>(ctx, S, F, C, c, sink, n_sm, pac_words);
                                if (variant == 3) return six ? launch_filter<13, 4, true, 6, 3>(ctx, S, F, C, c, sink, n_sm, pac_words) : launch_filter<13, 4, true, 4, 3>(ctx, S, F, C, c, sink, n_sm, pac_words);
                                return six ? launch_filter<13, 4, true, 6, 1>(ctx, S, F, C, c, sink, n_sm, pac_words) : launch_filter<13, 4, true, 4, 1>(ctx, S, F, C, c, sink, n_sm, pac_words);
                            }
                            if (ctx->vote_slots >= 6) return launch_filter<13, 4, true, 6>(ctx, S, F, C, c, sink, n_sm, pac_words);
                            return launch_filter<13, 4, true>(ctx, S, F, C, c, sink, n_sm, pac_words);
                        case 14: return launch_filter<14, 2, true>(ctx, S, F, C, c, sink, n_sm, pac_words);
                        case 15: return launch_filter<15, 1, true>(ctx, S, F, C, c, sink, n_sm, pac_words);
                        default: return launch_filter<16, 1, true>(ctx, S, F, C, c, sink, n_sm, pac_words);
                    }
                }
                switch (c) {             // byte counters: 8 bytes of filter per SA hit of the class
                    case 0: return launch_filter<12, 8, false>(ctx, S, F, C, 0, sink, n_sm, pac_words);
                    case 1: return launch_filter<13, 4, false>(ctx, S, F, C, 1, sink, n_sm, pac_words);
                    case 2: return launch_filter<14, 2, false>(ctx, S, F, C, 2, sink, n_sm, pac_words);
                    case 3: return launch_filter<15, 1, false>(ctx, S, F, C, 3, sink, n_sm, pac_words);
                    case 4: return launch_filter<16, 1, false>(ctx, S, F, C, 4, sink, n_sm, pac_words);
                    default: return launch_filter<17, 1, false>(ctx, S, F, C, 5, sink, n_sm, pac_words);
                }
            }
            switch (k - GMX_N_CLASSES) {
                case 0: return launch_vote<10, 2>(ctx, S, C, 0, sink, n_sm);
                case 1: return launch_vote<11, 1>(ctx, S, C, 1, sink, n_sm);
                case 2: return launch_vote<12, 1>(ctx, S, C, 2, sink, n_sm);
                case 3: return launch_vote<13, 1>(ctx, S, C, 3, sink, n_sm);
                case 4: return launch_vote<14, 1>(ctx, S, C, 4, sink, n_sm);
                default: {
                    if (ctx->d_arena.cap == 0) { cudaError_t e = ctx->d_arena.ensure((size_t)256 << 20); if (e != cudaSuccess) return e; }
                    GlobalTableArena A; A.words = ctx->d_arena.as<uint32_t>(); A.used = &dc->arena_used; A.cap = ctx->d_arena.cap / 4; A.overflow = &dc->arena_overflow;
                    k_vote_gmem<<<n_sm * 4, 128, 0, ctx->stream>>>(ctx->ix, S, C, 5, P.kmin, sink, A);
                    return cudaGetLastError();
                }
            }
        };
        stage_begin(ctx, ST_VOTE);
        uint32_t launched = 0;
        int n_launch = 0;
        for (int k = 0; k < 2 * GMX_N_CLASSES; ++k)
            if ((ctx->class_hint >> k) & 1u) { CK(launch_class(k)); launched |= 1u << k; n_launch++; }
        stage_end(ctx, ST_VOTE, 0, 0, n_launch);
        if (opt) {
            SealIn in;
            in.n_cand = &dc->n_cand; in.cand_overflow = &dc->cand_overflow; in.arena_overflow = &dc->arena_overflow;
            in.cls_count = dc->cls_count; in.fcls_count = dc->fcls_count; in.launched = launched; in.bound = n_cand;
            k_seal_candidates<<<nblk(n_cand, 256), 256, 0, ctx->stream>>>(in, sink.keys, &dc->live_cand, &dc->bad);
            CK(cudaGetLastError());
            break;
        }
        CK(cudaMemcpyAsync(&hc, dc, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream));
        { int r = finish_scan(ctx, slot ^ 1); if (r != GMX_OK) return r; }      // host work for the next chunk while the vote runs
        CK(cudaStreamSynchronize(ctx->stream));
        stage_collect(ctx);
        // classes that hold tasks but were not launched (the filter kernels also hand tasks to the exact classes as they run)
        for (int pass = 0; pass < 3; ++pass) {
            uint32_t need = 0;
            for (int c = 0; c < GMX_N_CLASSES; ++c) {
                if (hc.c.fcls_count[c]) need |= 1u << c;
                if (hc.c.cls_count[c]) need |= 1u << (GMX_N_CLASSES + c);
            }
            if (pass == 0) ctx->class_hint = need ? need : ctx->class_hint;
            const uint32_t todo = need & ~launched;
            if (!todo) break;
            stage_begin(ctx, ST_VOTE);
            n_launch = 0;
            for (int k = 0; k < 2 * GMX_N_CLASSES; ++k)
                if ((todo >> k) & 1u) { CK(launch_class(k)); launched |= 1u << k; n_launch++; }
            stage_end(ctx, ST_VOTE, 0, 0, n_launch);
            CK(cudaMemcpyAsync(&hc, dc, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            stage_collect(ctx);
        }
        if (hc.c.arena_overflow) {
            if (attempt >= 4) { ctx->err = "global vote-table arena overflow"; return GMX_ERR_OVERFLOW; }
            size_t want = ctx->d_arena.cap * 4;
            ctx->d_arena.release();
            CK(ctx->d_arena.ensure(want));
            continue;
        }
        if (hc.c.cand_overflow || hc.c.n_cand > ctx->cand_cap) {
            if (attempt >= 6) { ctx->err = "candidate list overflow"; return GMX_ERR_OVERFLOW; }
            ctx->cand_cap = std::max<size_t>(ctx->cand_cap * 2, (size_t)hc.c.n_cand + 1024);
            continue;
        }
        n_cand = hc.c.n_cand;
        break;
    }
    const uint32_t *live_c = opt ? &dc->live_cand : nullptr, *live_l = opt ? &dc->live_lead : nullptr;
    if (!opt) {
        if (hc.s.v[3]) { ctx->err = "device-resident gmx_reads: a read is longer than gmx_reads.max_len (or has a negative length)"; return GMX_ERR_INVALID; }
        account_seed_vote(ctx, hc.s);
    }

    // restore the reference's processing order: (task, round, position)
    unsigned long long *keys = ctx->d_keys.as<unsigned long long>();
    if (n_cand > 1) {
        stage_begin(ctx, ST_SORT);
        size_t tmp_bytes = 0;
        int task_bits = 1; while ((1ll << task_bits) < n_tasks) task_bits++;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, ctx->d_keys_alt.as<unsigned long long>(), (int)n_cand, 0, 40 + task_bits, ctx->stream));
        CK(ctx->d_sort_tmp.ensure(tmp_bytes));
        CK(cub::DeviceRadixSort::SortKeys(ctx->d_sort_tmp.p, tmp_bytes, keys, ctx->d_keys_alt.as<unsigned long long>(), (int)n_cand, 0, 40 + task_bits, ctx->stream));
        keys = ctx->d_keys_alt.as<unsigned long long>();
        stage_end(ctx, ST_SORT, opt ? 0 : n_cand, opt ? 0 : (uint64_t)n_cand * 16, 1);
    }

    size_t nc = std::max<uint32_t>(n_cand, 1);
    CK(ctx->d_score.ensure(nc * 4)); CK(ctx->d_leader.ensure(nc * 4)); CK(ctx->d_slot.ensure(nc * 4)); CK(ctx->d_lead_cand.ensure(nc * 4));
    CK(ctx->d_hashes.ensure(nc * 8)); CK(ctx->d_expv.ensure(nc * 8));

    // GetString + K2a
    if (n_cand) {
        stage_begin(ctx, ST_NW);
        k_cand_score<<<nblk(n_cand, 128), 128, 0, ctx->stream>>>(ctx->ix, ctx->dreads, ctx->tab, P, keys, n_cand, ctx->d_score.as<float>(), live_c);
        CK(cudaGetLastError());
        // per candidate: 2-bit window (n/4 B) + read bases and qualities (2n B) + key (8 B) + score (4 B)
        const uint32_t nc_acc = opt ? 0 : n_cand;
        stage_end(ctx, ST_NW, (uint64_t)nc_acc * (uint64_t)std::max(7 * max_len - 12, 0), (uint64_t)nc_acc * (uint64_t)(max_len / 4 + 2 * max_len + 12), 1);
    }

    // acceptance, grouping, denominator, best group
    FinalizeOut O;
    O.results = ctx->d_results.as<gmx_read_result>(); O.leader = ctx->d_leader.as<int32_t>(); O.slot = ctx->d_slot.as<int32_t>();
    O.lead_cand = ctx->d_lead_cand.as<uint32_t>(); O.n_leaders = &dc->n_leaders; O.n_accepted = &dc->n_accepted;
    O.hashes = ctx->d_hashes.as<uint64_t>(); O.expv = ctx->d_expv.as<double>();
    CK(ctx->d_ranges.ensure((size_t)n * 8));
    O.range = ctx->d_ranges.as<uint32_t>();
    CK(ctx->d_groups.ensure((size_t)n * 4 + 16)); CK(ctx->d_read_base.ensure((size_t)n * 4 + 16));
    O.groups_per_read = ctx->d_groups.as<uint32_t>();
    stage_begin(ctx, ST_FINALIZE);
    CK(cudaMemsetAsync(ctx->d_groups.p, 0, (size_t)n * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_ranges.p, 0, (size_t)n * 8, ctx->stream));
    if (n_cand) { k_cand_ranges<<<nblk(n_cand, 256), 256, 0, ctx->stream>>>(keys, n_cand, ctx->d_ranges.as<uint32_t>(), live_c); CK(cudaGetLastError()); }
    k_finalize_reads<<<nblk((int64_t)n * 32, 128), 128, 0, ctx->stream>>>(ctx->ix, ctx->dreads, P, ctx->d_prep.as<ReadPrep>(), keys, ctx->d_score.as<float>(), n_cand, O);
    CK(cudaGetLastError());
    if (n_cand) {
        size_t scan_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, ctx->d_groups.as<uint32_t>(), ctx->d_read_base.as<uint32_t>(), (int)n, ctx->stream));
        CK(ctx->d_sort_tmp.ensure(scan_bytes));
        CK(cub::DeviceScan::ExclusiveSum(ctx->d_sort_tmp.p, scan_bytes, ctx->d_groups.as<uint32_t>(), ctx->d_read_base.as<uint32_t>(), (int)n, ctx->stream));
        k_assign_slots<<<nblk(n_cand, 256), 256, 0, ctx->stream>>>(keys, ctx->d_leader.as<int32_t>(), ctx->d_slot.as<int32_t>(), ctx->d_read_base.as<uint32_t>(),
                                                                  ctx->d_lead_cand.as<uint32_t>(), n_cand, &dc->n_leaders, &dc->n_accepted,
                                                                  live_c, opt ? lead_cap : 0xffffffffu);
        CK(cudaGetLastError());
    }
    if (opt) { k_seal_leaders<<<1, 1, 0, ctx->stream>>>(&dc->n_leaders, lead_cap, &dc->live_cand, &dc->live_lead, &dc->bad); CK(cudaGetLastError()); }
    stage_end(ctx, ST_FINALIZE, (uint64_t)n, 0, 4);

    if (!opt) {
        CK(cudaMemcpyAsync(&hc.c, dc, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        stage_collect(ctx);
        if (n > 0) { ctx->pred_cand = (double)n_cand / n; ctx->pred_lead = (double)hc.c.n_leaders / n; }
    }

    cs.lo = lo; cs.n = n; cs.max_len = max_len; cs.total_bases = total_bases;
    cs.n_cand = n_cand; cs.n_leaders = opt ? lead_cap : hc.c.n_leaders; cs.n_accepted = opt ? 0 : hc.c.n_accepted; cs.keys = keys;
    cs.opt = opt; cs.live_cand = live_c; cs.live_lead = live_l;
    LeaderStore &L = cs.L;
    L.lead_cand = ctx->d_lead_cand.as<uint32_t>();
    L.a_stride = max_len + 2 * P.max_gap + 8; L.c_stride = GMX_CIGAR_STRIDE; L.max_len = max_len;
    size_t nl = std::max<uint32_t>(cs.n_leaders, 1);
    CK(ctx->d_alen.ensure(nl * 4)); CK(ctx->d_aligned.ensure(nl * L.a_stride)); CK(ctx->d_cigar.ensure(nl * L.c_stride));
    L.alen = ctx->d_alen.as<int32_t>(); L.aligned = ctx->d_aligned.as<uint8_t>(); L.cigar = ctx->d_cigar.as<char>(); L.hmm = nullptr;
    cs.valid = true;
    return GMX_OK;
}

// PHASE B for the resident chunk: K2b (+ K2c) per group leader, then K3.
static int phase_b(gmx_ctx *ctx)
{
    gmx_ctx::ChunkState &cs = ctx->cs;
    if (!cs.valid) { ctx->err = "no mapped chunk is resident on the device"; return GMX_ERR_STATE; }
    const DevParams &P = ctx->dparams;
    LeaderStore &L = cs.L;
    const uint32_t n_leaders = cs.n_leaders, n_cand = cs.n_cand;
    const int max_len = cs.max_len;
    if (!n_leaders) return GMX_OK;
    // traceback is needed in every mode for the CIGAR of the best hit (get_SAM, reference inc/ScoredSeq.h:314-372)
    CK(ctx->d_moves.ensure((size_t)n_leaders * ((size_t)max_len + 1) * 4));
    // the gapped read strings themselves are only consumed by the BS scatter and by gmx_get_best_alignments
    const int want_aligned = (P.mode == GMX_MODE_BS || ctx->collect_hits) ? 1 : 0;
    stage_begin(ctx, ST_TRACEBACK);
    k_traceback<<<nblk(n_leaders, 128), 128, 0, ctx->stream>>>(ctx->ix, ctx->dreads, ctx->tab, P, cs.keys, n_leaders, L,
                                                              ctx->d_moves.as<uint32_t>(), want_aligned, ctx->d_multi_count.as<uint32_t>() + 1, cs.live_lead);
    CK(cudaGetLastError());
    const uint32_t nl_acc = cs.opt ? 0 : n_leaders;        // an optimistic chunk's work is accounted when it is settled
    stage_end(ctx, ST_TRACEBACK, (uint64_t)nl_acc * (uint64_t)std::max(7 * max_len - 12, 0), 0, 1);
    if (P.mode == GMX_MODE_SNP) {
        if (max_len > 32 * GMX_PHMM_LONGC) { ctx->err = "SNP mode: reads longer than 1024 bp are not supported by the pair-HMM kernel"; return GMX_ERR_UNSUPPORTED; }
        CK(ctx->d_hmm.ensure((size_t)n_leaders * max_len * 5 * 4));
        L.hmm = ctx->d_hmm.as<float>();
        size_t per_task = gmx_phmm_scratch_doubles(max_len);
        const int C = (max_len + 31) / 32;
        // persistent grid: as many one-warp CTAs as the SMs hold at once (register-bound), each with its own scratch slot
        int per_sm = 1;
#define GMX_PHMM_OCC(CT) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pair_hmm_leaders<CT>, GMX_PHMM_THREADS, 0))
        switch (C) { case 1: GMX_PHMM_OCC(1); break; case 2: GMX_PHMM_OCC(2); break; case 3: GMX_PHMM_OCC(3); break; case 4: GMX_PHMM_OCC(4); break;
                     case 5: GMX_PHMM_OCC(5); break; default: if (C <= GMX_PHMM_MAXC) GMX_PHMM_OCC(0); else GMX_PHMM_OCC(-1); break; }
#undef GMX_PHMM_OCC
        uint32_t grid = std::min<uint32_t>(n_leaders, (uint32_t)ctx->n_sm * (uint32_t)std::max(per_sm, 1));
        grid = std::max<uint32_t>(1, std::min<uint32_t>(grid, (uint32_t)((4ull << 30) / (per_task * 8))));      // long reads: at most 4 GB of scratch
        CK(ctx->d_phmm_scratch.ensure((size_t)grid * per_task * 8 + 16));
        uint32_t *cursor = reinterpret_cast<uint32_t *>(ctx->d_phmm_scratch.as<char>() + (size_t)grid * per_task * 8);
        stage_begin(ctx, ST_PHMM);
        CK(cudaMemsetAsync(cursor, 0, 4, ctx->stream));
        int launches = 1;
        double *scr = ctx->d_phmm_scratch.as<double>();
#define GMX_PHMM_LAUNCH(CT) k_pair_hmm_leaders<CT><<<grid, GMX_PHMM_THREADS, 0, ctx->stream>>>(ctx->ix, ctx->dreads, ctx->tab, cs.keys, L, n_leaders, cursor, scr, per_task, cs.live_lead)
        switch (C) {
            case 1: GMX_PHMM_LAUNCH(1); break;
            case 2: GMX_PHMM_LAUNCH(2); break;
            case 3: GMX_PHMM_LAUNCH(3); break;
            case 4: GMX_PHMM_LAUNCH(4); break;
            case 5: GMX_PHMM_LAUNCH(5); break;
            default: if (C <= GMX_PHMM_MAXC) GMX_PHMM_LAUNCH(0); else GMX_PHMM_LAUNCH(-1); break;
        }
#undef GMX_PHMM_LAUNCH
        CK(cudaGetLastError());
        stage_end(ctx, ST_PHMM, (uint64_t)nl_acc * (uint64_t)max_len * (uint64_t)max_len, 0, launches);
    }
    stage_begin(ctx, ST_SCATTER);
    k_scatter<<<nblk(n_cand, 128), 128, 0, ctx->stream>>>(ctx->ix, ctx->dreads, P, cs.keys, ctx->d_score.as<float>(), ctx->d_leader.as<int32_t>(),
                                                                        ctx->d_slot.as<int32_t>(), n_cand, ctx->d_results.as<gmx_read_result>(), L, ctx->acc, cs.live_cand);
    CK(cudaGetLastError());
    stage_end(ctx, ST_SCATTER, cs.n_accepted, (uint64_t)cs.n_accepted * scatter_bytes_per_hit(P, max_len), 1);
    return GMX_OK;
}

// Per-read results (+ hit list / best alignments) of the resident chunk -> host.
static int download_chunk(gmx_ctx *ctx, bool scored, gmx_read_result *results_out)
{
    gmx_ctx::ChunkState &cs = ctx->cs;
    if (!cs.valid) { ctx->err = "no mapped chunk is resident on the device"; return GMX_ERR_STATE; }
    const int32_t n = cs.n, lo = cs.lo;
    const uint32_t n_cand = cs.n_cand, n_leaders = cs.n_leaders;
    LeaderStore &L = cs.L;
    // the library's own result storage is the destination when the caller passes none and the staging area of the
    // hit-collecting path: size it here, since a gmx_score_batch may follow a gmx_map_batch that did not need it
    if ((!results_out || ctx->collect_hits) && ctx->h_results.size() < (size_t)ctx->last_n_reads) ctx->h_results.resize((size_t)ctx->last_n_reads);
    if (ctx->collect_hits && ctx->h_best_aligned.size() < (size_t)ctx->last_n_reads * (size_t)ctx->h_a_stride) {
        ctx->h_best_aligned.assign((size_t)ctx->last_n_reads * (size_t)ctx->h_a_stride, 0);
        memset(ctx->h_best_cigar.p, 0, (size_t)std::max(ctx->last_n_reads, 1) * GMX_CIGAR_STRIDE);
    }
    stage_begin(ctx, ST_DOWNLOAD);
    if (!ctx->collect_hits) {
        // fast path: fixed-size records only, gathered into the batch's device arrays and copied out on a second stream
        // while the next chunk computes
        gmx_read_result *d_res = ctx->d_batch_results.as<gmx_read_result>() + lo;
        char *d_cig = ctx->d_batch_cigar.as<char>() + (size_t)lo * GMX_CIGAR_STRIDE;
        if (scored && n_cand && ctx->multi_cap) {                           // positions of multi-position best groups (SAM row)
            k_gather_multi<<<nblk(n_cand, 256), 256, 0, ctx->stream>>>(cs.keys, ctx->d_leader.as<int32_t>(), n_cand, ctx->d_results.as<gmx_read_result>(),
                                                                      lo, ctx->d_multi.as<MultiPos>(), ctx->d_multi_count.as<uint32_t>(), ctx->multi_cap, cs.live_cand);
            CK(cudaGetLastError());
        }
        k_gather_best<<<nblk(n, 128), 128, 0, ctx->stream>>>(ctx->d_results.as<gmx_read_result>(), d_res, n,
                                                           ctx->d_slot.as<int32_t>(), L, d_cig, GMX_CIGAR_STRIDE,
                                                           scored && n_leaders ? 1 : 0, cs.live_lead);
        CK(cudaGetLastError());
        const int sl = ctx->dl_slot; ctx->dl_slot ^= 1;
        CK(cudaEventRecord(ctx->gather_ev[sl], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->d2h_stream, ctx->gather_ev[sl], 0));
        gmx_read_result *dst = results_out ? results_out + lo : ctx->h_results.data() + lo;
        CK(cudaMemcpyAsync(dst, d_res, (size_t)n * sizeof(gmx_read_result), cudaMemcpyDeviceToHost, ctx->d2h_stream));
        CK(cudaMemcpyAsync(ctx->h_best_cigar.as<char>() + (size_t)lo * GMX_CIGAR_STRIDE, d_cig, (size_t)n * GMX_CIGAR_STRIDE,
                           cudaMemcpyDeviceToHost, ctx->d2h_stream));
        stage_end(ctx, ST_DOWNLOAD, (uint64_t)n, (uint64_t)n * (sizeof(gmx_read_result) + GMX_CIGAR_STRIDE), 1);
        return GMX_OK;                                                      // run_batch / gmx_score_batch drain d2h_stream
    }
    gmx_read_result *hres = ctx->h_results.data() + lo;
    CK(cudaMemcpyAsync(hres, ctx->d_results.p, (size_t)n * sizeof(gmx_read_result), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<unsigned long long> h_keys(n_cand);
    std::vector<int32_t> h_leader(n_cand), h_slot(n_cand), h_alen(n_leaders);
    std::vector<float> h_score(n_cand);
    std::vector<char> h_cigar((size_t)n_leaders * L.c_stride);
    std::vector<uint8_t> h_aligned((size_t)n_leaders * L.a_stride);
    if (n_cand) {
        CK(cudaMemcpyAsync(h_keys.data(), cs.keys, (size_t)n_cand * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_leader.data(), ctx->d_leader.p, (size_t)n_cand * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_slot.data(), ctx->d_slot.p, (size_t)n_cand * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_score.data(), ctx->d_score.p, (size_t)n_cand * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (scored && n_leaders) {
        CK(cudaMemcpyAsync(h_alen.data(), ctx->d_alen.p, (size_t)n_leaders * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_cigar.data(), ctx->d_cigar.p, h_cigar.size(), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(h_aligned.data(), ctx->d_aligned.p, h_aligned.size(), cudaMemcpyDeviceToHost, ctx->stream));
    }
    stage_end(ctx, ST_DOWNLOAD, (uint64_t)n, (uint64_t)n * sizeof(gmx_read_result) + (uint64_t)n_cand * 20, 0);
    CK(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx);

    // host-side assembly of the variable-length outputs (hit list, best CIGAR): pure bookkeeping
    const int a_out = ctx->h_a_stride;
    char *best_cigar = ctx->h_best_cigar.as<char>();
    for (int32_t i = 0; i < n; ++i) {
        gmx_read_result &res = hres[i];
        int32_t c_lo = res.hit_begin, c_hi = res.hit_end;
        res.hit_begin = res.hit_end = (int32_t)ctx->h_hits.size();
        if (res.status != GMX_READ_MAPPED) continue;
        // group labels: order of first appearance of the leader (processing order)
        std::vector<int32_t> leaders;
        int32_t best_leader = res.best_group >= 0 ? c_lo + res.best_group : -1;
        int best_label = -1;
        for (int32_t c = c_lo; c < c_hi; ++c) {
            if (h_leader[c] < 0) continue;
            int label = -1;
            for (size_t k = 0; k < leaders.size(); ++k) if (leaders[k] == h_leader[c]) { label = (int)k; break; }
            if (label < 0) { label = (int)leaders.size(); leaders.push_back(h_leader[c]); }
            if (h_leader[c] == best_leader) best_label = label;
            gmx_hit h;
            h.pos = (uint32_t)h_keys[c]; h.score = h_score[h_leader[c]]; h.read = lo + i; h.group = (int16_t)label;
            h.strand = (uint8_t)((h_keys[c] >> 40) & 1); h.first_strand = (uint8_t)((h_keys[h_leader[c]] >> 40) & 1);
            ctx->h_hits.push_back(h);
        }
        res.hit_end = (int32_t)ctx->h_hits.size();
        std::sort(ctx->h_hits.begin() + res.hit_begin, ctx->h_hits.end(), [](const gmx_hit &a, const gmx_hit &b) {
            if (a.group != b.group) return a.group < b.group;
            if (a.pos != b.pos) return a.pos < b.pos;
            return a.strand < b.strand;
        });
        res.best_group = best_label;
        if (scored && best_leader >= 0) {
            int s = h_slot[best_leader];
            res.best_aligned_len = h_alen[s];
            memcpy(best_cigar + (size_t)(lo + i) * GMX_CIGAR_STRIDE, &h_cigar[(size_t)s * L.c_stride], GMX_CIGAR_STRIDE);
            int cp = std::min(a_out, L.a_stride);
            memcpy(&ctx->h_best_aligned[(size_t)(lo + i) * a_out], &h_aligned[(size_t)s * L.a_stride], (size_t)cp);
        }
    }
    if (results_out) memcpy(results_out + lo, hres, (size_t)n * sizeof(gmx_read_result));
    return GMX_OK;
}

// ---- optimistic chunks ----------------------------------------------------------------------------
static bool chunk_can_be_optimistic(const gmx_ctx *ctx, bool do_score)
{
    return do_score && ctx->optimistic && !ctx->collect_hits && ctx->pred_cand >= 0;
}

// behind the last kernel of an optimistic chunk: its counters leave for the host, the chunk waits to be settled
static int chunk_mark_pending(gmx_ctx *ctx, const gmx_reads *reads, int slot, gmx_read_result *results)
{
    const int par = ctx->par;
    gmx_ctx::Pending &pd = ctx->pend[par];
    // stored by a kernel straight into the pinned host words: a copy would queue on the D2H engine behind the chunk's own
    // result download, and the next chunk's kernels behind the copy (measured: 0.6 ms of idle GPU per chunk)
    static_assert(sizeof(HostCounters) % 4 == 0, "word copy");
    k_publish_words<<<1, 64, 0, ctx->stream>>>(ctx->d_counters.as<uint32_t>(), reinterpret_cast<uint32_t *>(&ctx->h_counters.as<HostCounters>()[par]),
                                              (int)(sizeof(HostCounters) / 4));
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->settle_ev[par], ctx->stream));
    pd.active = true; pd.lo = ctx->cs.lo; pd.hi = ctx->cs.lo + ctx->cs.n; pd.max_len = ctx->cs.max_len; pd.slot = slot;
    pd.reads = *reads; pd.results = results;
    ctx->n_optimistic++;
    return GMX_OK;
}

// one chunk, PHASE A + B + download, in whichever way is possible
static int run_chunk(gmx_ctx *ctx, const gmx_reads *reads, int32_t lo, int32_t hi, int slot, gmx_read_result *results, bool opt)
{
    int r = phase_a(ctx, reads, lo, hi, slot, opt);
    if (r == GMX_OK) r = phase_b(ctx);
    if (r == GMX_OK) r = download_chunk(ctx, true, results);
    if (r == GMX_OK && opt) r = chunk_mark_pending(ctx, reads, slot, results);
    return r;
}

// Wait for the optimistic chunk of parity `par`, fold its counters into the statistics and the next prediction -- or, when
// k_seal_* voided it (nothing of it reached the accumulators or the multi-position list), run it again the synchronous
// way.  Its reads are still where they were: the caller issues the upload that reuses their buffers after this.
static int settle_chunk(gmx_ctx *ctx, int par)
{
    gmx_ctx::Pending &pd = ctx->pend[par];
    if (!pd.active) return GMX_OK;
    pd.active = false;
    CK(cudaEventSynchronize(ctx->settle_ev[par]));
    stage_collect(ctx, par);
    const HostCounters hc = ctx->h_counters.as<HostCounters>()[par];
    if (hc.s.v[3]) { ctx->err = "device-resident gmx_reads: a read is longer than gmx_reads.max_len (or has a negative length)"; return GMX_ERR_INVALID; }
    const int32_t n = pd.hi - pd.lo;
    if (hc.c.bad) {
        ctx->n_rerun++;
        ctx->pred_margin = std::min(4.0, ctx->pred_margin * 2);
        uint32_t need = 0;
        for (int c = 0; c < GMX_N_CLASSES; ++c) { if (hc.c.fcls_count[c]) need |= 1u << c; if (hc.c.cls_count[c]) need |= 1u << (GMX_N_CLASSES + c); }
        ctx->class_hint |= need;
        const gmx_reads again = pd.reads;
        return run_chunk(ctx, &again, pd.lo, pd.hi, pd.slot, pd.results, false);
    }
    account_seed_vote(ctx, hc.s);
    const DevParams &P = ctx->dparams;
    const uint64_t nc = hc.c.n_cand, nl = hc.c.n_leaders, na = hc.c.n_accepted, L = (uint64_t)pd.max_len;
    ctx->stage_units[ST_SORT] += nc; ctx->stage_bytes[ST_SORT] += nc * 16;
    ctx->stage_units[ST_NW] += nc * (uint64_t)std::max(7 * pd.max_len - 12, 0); ctx->stage_bytes[ST_NW] += nc * (L / 4 + 2 * L + 12);
    ctx->stage_units[ST_TRACEBACK] += nl * (uint64_t)std::max(7 * pd.max_len - 12, 0);
    if (P.mode == GMX_MODE_SNP) ctx->stage_units[ST_PHMM] += nl * L * L;
    ctx->stage_units[ST_SCATTER] += na; ctx->stage_bytes[ST_SCATTER] += na * scatter_bytes_per_hit(P, pd.max_len);
    if (n > 0) { ctx->pred_cand = (double)nc / n; ctx->pred_lead = (double)nl / n; }
    uint32_t need = 0;
    for (int c = 0; c < GMX_N_CLASSES; ++c) { if (hc.c.fcls_count[c]) need |= 1u << c; if (hc.c.cls_count[c]) need |= 1u << (GMX_N_CLASSES + c); }
    if (need) ctx->class_hint = need;
    return GMX_OK;
}

static int batch_max_len(gmx_ctx *ctx, const gmx_reads *reads, int32_t *out)
{
    return scan_max_len(ctx, reads, 0, reads->n_reads, out);
}

static int run_batch_impl(gmx_ctx *ctx, const gmx_reads *reads, gmx_read_result *results, bool do_score);

// behind the batch's last kernel: [0] multi-position entries, [1] truncated CIGARs -> pinned host words
static int publish_batch_counts(gmx_ctx *ctx)
{
    CK(ctx->h_batch_counts.ensure(16));
    k_publish_words<<<1, 32, 0, ctx->stream>>>(ctx->d_multi_count.as<uint32_t>(), ctx->h_batch_counts.as<uint32_t>(), 2);
    CK(cudaGetLastError());
    return GMX_OK;
}

// end of a batch (streams idle): the multi-position list of the fast download path and the truncated-CIGAR count
static int collect_batch_counters(gmx_ctx *ctx)
{
    const uint32_t *c = ctx->h_batch_counts.as<uint32_t>();            // published by batch_end / gmx_score_batch before their drain
    if (ctx->multi_cap) {
        uint32_t nm = c[0];
        if (nm > ctx->multi_cap) { ctx->multi_overflow = true; nm = ctx->multi_cap; }
        ctx->h_multi.resize(nm);
        if (nm) CK(cudaMemcpy(ctx->h_multi.data(), ctx->d_multi.p, (size_t)nm * sizeof(MultiPos), cudaMemcpyDeviceToHost));
    }
    if (c[1]) {
        char b[160];
        snprintf(b, sizeof(b), "%u alignment(s) need a CIGAR longer than the %d-byte slot: raise GMX_OPT_CIGAR_STRIDE", c[1], ctx->cigar_stride);
        ctx->err = b;
        return GMX_ERR_OVERFLOW;
    }
    return GMX_OK;
}

// PHASE A (+ PHASE B when do_score).  Every error leaves through one cleanup path: copies on the upload / download
// streams may still be reading the caller's reads or writing the caller's results, so all three streams are drained
// and the batch state is invalidated before the error code goes back.
static int run_batch(gmx_ctx *ctx, const gmx_reads *reads, gmx_read_result *results, bool do_score)
{
    if (!ctx || !reads || reads->n_reads < 0 || (reads->n_reads > 0 && !reads->offsets)) return GMX_ERR_INVALID;
    const int r = run_batch_impl(ctx, reads, results, do_score);
    if (r != GMX_OK) {
        const std::string why = ctx->err;
        cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->d2h_stream);
        cudaGetLastError();
        ctx->cs.valid = false; ctx->mapped = false; ctx->scored = false; ctx->keep_valid = false;
        ctx->pend[0].active = ctx->pend[1].active = false;
        ctx->up_scan[0].reads = nullptr; ctx->up_scan[1].reads = nullptr;
        ctx->err = why;
    }
    return r;
}

// per-batch state before the first chunk runs: `n` = reads of the batch (an upper bound for a FASTQ text that is indexed
// piece by piece), `max_len` = longest read where it is known up front
static int batch_begin(gmx_ctx *ctx, int32_t n, int32_t max_len, gmx_read_result *results, bool do_score)
{
    ctx->h_results.assign(ctx->collect_hits || !results ? (size_t)n : 0, gmx_read_result());
    ctx->h_hits.clear();
    ctx->h_multi.clear();
    ctx->multi_overflow = false; ctx->multi_cap = 0;
    CK(ctx->d_multi_count.ensure(16));                                       // [0] multi-position entries, [1] truncated CIGARs
    CK(cudaMemsetAsync(ctx->d_multi_count.p, 0, 16, ctx->stream));
    if (!ctx->collect_hits && do_score && n > 0) {
        ctx->multi_cap = (uint32_t)std::min<int64_t>(4ll * n + 65536, 0x7fffffffll);
        CK(ctx->d_multi.ensure((size_t)ctx->multi_cap * sizeof(MultiPos)));
    }
    ctx->h_a_stride = max_len + 2 * ctx->params.max_gap + 8;
    ctx->batch_dev_valid = false; ctx->fq_batch = false;
    if (!ctx->collect_hits) {
        CK(ctx->d_batch_results.ensure((size_t)std::max(n, 1) * sizeof(gmx_read_result)));
        CK(ctx->d_batch_cigar.ensure((size_t)std::max(n, 1) * GMX_CIGAR_STRIDE));
    }
    CK(ctx->h_best_cigar.ensure((size_t)std::max(n, 1) * GMX_CIGAR_STRIDE));
    if (ctx->collect_hits) {
        memset(ctx->h_best_cigar.p, 0, (size_t)std::max(n, 1) * GMX_CIGAR_STRIDE);
        ctx->h_best_aligned.assign((size_t)n * ctx->h_a_stride, 0);
    }
    ctx->last_n_reads = n; ctx->last_max_len = 0;
    ctx->mapped = false; ctx->scored = false; ctx->cs.valid = false;
    stage_reset(ctx);
    return GMX_OK;
}

// after the last chunk has been issued: drain, fold the counters, mark the batch
static int batch_end(gmx_ctx *ctx, bool do_score)
{
    { int r = settle_all(ctx); if (r != GMX_OK) return r; }
    { int r = publish_batch_counts(ctx); if (r != GMX_OK) return r; }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->d2h_stream));
    stage_collect(ctx);
    int r = collect_batch_counters(ctx);
    if (r != GMX_OK) return r;
    ctx->mapped = true; ctx->scored = do_score;
    ctx->batch_dev_valid = do_score && !ctx->collect_hits;
    return GMX_OK;
}

static int run_batch_impl(gmx_ctx *ctx, const gmx_reads *reads, gmx_read_result *results, bool do_score)
{
    CK(cudaSetDevice(ctx->device));
    const int32_t n = reads->n_reads;
    int32_t max_len = 0;       // of the whole batch: only the hit-collecting download needs it before the first chunk runs
    if (n > 0 && (ctx->collect_hits || reads->on_device)) { int r = batch_max_len(ctx, reads, &max_len); if (r != GMX_OK) return r; }
    { int r = batch_begin(ctx, n, max_len, results, do_score); if (r != GMX_OK) return r; }
    // Chunk schedule.  A chunk's upload rides under its predecessor's kernels and its download under its successor's,
    // so what a batch exposes is the first upload and the last download.  When PHASE A and B run together, a host
    // batch therefore starts and every batch ends with a short chunk of 1/8 of the regular size (measured at 1 M reads:
    // end to end 19.3 ms per step; a finer ramp 1/16, 3/16, 9/16 costs 19.8 ms -- small chunks run the kernels at a
    // worse rate than the PCIe time they hide).
    const int32_t step = (int32_t)ctx->chunk_reads;
    const int32_t edge = std::max<int32_t>(step / 8, 1);
    std::vector<int32_t> cuts(1, 0);
    if (do_score && n > 4 * edge) {
        int32_t at = 0;
        if (!reads->on_device) { at = edge; cuts.push_back(at); }
        const int32_t body = n - at - edge;
        const int32_t parts = std::max<int32_t>(1, (int32_t)(((int64_t)body - edge + step - 1) / step));     // a part may exceed the regular size by 1/8
        for (int32_t k = 1; k <= parts; ++k) cuts.push_back(at + (int32_t)((int64_t)body * k / parts));
        cuts.push_back(n);
    } else {
        for (int32_t lo = 0; lo < n; lo += step) cuts.push_back((int32_t)std::min<int64_t>(n, (int64_t)lo + step));
    }
    if (n > 0) {   // first chunk's upload; every later one is issued while its predecessor computes
        int r = issue_upload(ctx, reads, 0, cuts[1], 0, ctx->copy_stream);
        if (r != GMX_OK) return r;
    }
    int slot = 0;
    for (size_t c = 0; c + 1 < cuts.size(); ++c, slot ^= 1) {
        const int32_t lo = cuts[c], hi = cuts[c + 1];
        auto next_upload = [&]() -> int {
            if (hi >= n) return GMX_OK;
            // buffer set slot^1 was last read by chunk c-1: its kernels must have drained first
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->done_ev[slot ^ 1], 0));
            return issue_upload(ctx, reads, hi, cuts[c + 2], slot ^ 1, ctx->copy_stream);
        };
        // chunk c-1 may still be waiting to be settled, and a chunk that has to be run again needs its reads where they
        // are: then the upload that overwrites them is issued after the settling, below
        const int prev = ctx->pend[0].active ? 0 : (ctx->pend[1].active ? 1 : -1);
        if (prev < 0) { int r = next_upload(); if (r != GMX_OK) return r; }
        const bool opt = chunk_can_be_optimistic(ctx, do_score);
        int r = phase_a(ctx, reads, lo, hi, slot, opt);
        if (r != GMX_OK) return r;
        const bool last_and_split = !do_score && hi == n && lo == 0;       // single-chunk map_batch: PHASE B may follow
        if (do_score) { r = phase_b(ctx); if (r != GMX_OK) return r; }
        r = download_chunk(ctx, do_score, results);
        if (r != GMX_OK) return r;
        CK(cudaEventRecord(ctx->done_ev[slot], ctx->stream));
        if (opt) { r = chunk_mark_pending(ctx, reads, slot, results); if (r != GMX_OK) return r; }
        if (prev >= 0) {                                                   // the GPU has chunk c queued while the host waits for c-1
            r = settle_chunk(ctx, prev);
            if (r == GMX_OK) r = next_upload();
            if (r != GMX_OK) return r;
        }
        if (!last_and_split && !do_score) ctx->cs.valid = false;
    }
    if (do_score) ctx->cs.valid = false;
    return batch_end(ctx, do_score);
}

extern "C" int gmx_process_batch(gmx_ctx *ctx, const gmx_reads *reads, gmx_read_result *results) { return run_batch(ctx, reads, results, true); }

// PHASE A only.  When the batch fits one internal chunk (GMX_OPT_CHUNK_READS, default 524288 -- the reference
// hands its workers 2048 reads at a time) its candidates stay resident and gmx_score_batch continues from them;
// larger batches are remembered by value (host input) or by reference (device input) and re-run.
extern "C" int gmx_map_batch(gmx_ctx *ctx, const gmx_reads *reads, gmx_read_result *results)
{
    int r = run_batch(ctx, reads, results, false);
    if (r != GMX_OK) return r;
    ctx->keep_valid = false;
    if (!ctx->cs.valid && reads->n_reads > 0) {
        ctx->keep = *reads;
        if (!reads->on_device) {
            const int32_t n = reads->n_reads;
            const int64_t total = reads->offsets[n];
            ctx->keep_offsets.assign(reads->offsets, reads->offsets + n + 1);
            ctx->keep_seq.assign(reads->seq, reads->seq + total);
            ctx->keep.offsets = ctx->keep_offsets.data(); ctx->keep.seq = ctx->keep_seq.data();
            if (reads->qual) { ctx->keep_qual.assign(reads->qual, reads->qual + total); ctx->keep.qual = ctx->keep_qual.data(); }
            if (reads->pwm) { ctx->keep_pwm.assign(reads->pwm, reads->pwm + 4 * total); ctx->keep.pwm = ctx->keep_pwm.data(); }
        }
        ctx->keep_valid = true;
    }
    return GMX_OK;
}

extern "C" int gmx_score_batch(gmx_ctx *ctx, gmx_read_result *results)
{
    if (!ctx) return GMX_ERR_INVALID;
    if (!ctx->mapped || ctx->scored) { ctx->err = "gmx_score_batch needs a preceding gmx_map_batch"; return GMX_ERR_STATE; }
    CK(cudaSetDevice(ctx->device));
    if (ctx->last_n_reads == 0) { ctx->scored = true; return GMX_OK; }
    if (ctx->cs.valid) {
        ctx->h_hits.clear();
        ctx->h_multi.clear(); ctx->multi_overflow = false; ctx->multi_cap = 0;
        CK(ctx->d_multi_count.ensure(16));
        CK(cudaMemsetAsync(ctx->d_multi_count.p, 0, 16, ctx->stream));
        ctx->batch_dev_valid = false; ctx->fq_batch = false;
        if (!ctx->collect_hits) {
            ctx->multi_cap = (uint32_t)std::min<int64_t>(4ll * ctx->last_n_reads + 65536, 0x7fffffffll);
            CK(ctx->d_multi.ensure((size_t)ctx->multi_cap * sizeof(MultiPos)));
            CK(ctx->d_batch_results.ensure((size_t)std::max(ctx->last_n_reads, 1) * sizeof(gmx_read_result)));
            CK(ctx->d_batch_cigar.ensure((size_t)std::max(ctx->last_n_reads, 1) * GMX_CIGAR_STRIDE));
        }
        int r = phase_b(ctx);
        if (r != GMX_OK) return r;
        r = download_chunk(ctx, true, results);
        if (r != GMX_OK) return r;
        r = publish_batch_counts(ctx);
        if (r != GMX_OK) return r;
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->d2h_stream));
        stage_collect(ctx);
        r = collect_batch_counters(ctx);
        if (r != GMX_OK) return r;
        ctx->cs.valid = false; ctx->scored = true;
        ctx->batch_dev_valid = !ctx->collect_hits;
        return GMX_OK;
    }
    if (!ctx->keep_valid) { ctx->err = "the mapped batch is no longer available"; return GMX_ERR_STATE; }
    gmx_reads again = ctx->keep;
    ctx->keep_valid = false;
    return run_batch(ctx, &again, results, true);
}

extern "C" int gmx_chunk_stats(gmx_ctx *ctx, uint64_t *optimistic, uint64_t *rerun)
{
    if (!ctx) return GMX_ERR_INVALID;
    if (optimistic) *optimistic = ctx->n_optimistic;
    if (rerun) *rerun = ctx->n_rerun;
    return GMX_OK;
}

extern "C" int gmx_set_option(gmx_ctx *ctx, int option, int64_t value)
{
    if (!ctx) return GMX_ERR_INVALID;
    switch (option) {
        case GMX_OPT_COLLECT_HITS: ctx->collect_hits = value != 0; return GMX_OK;
        case GMX_OPT_CHUNK_READS:
            if (value < 1 || value > (1 << 22)) { ctx->err = "chunk_reads must be in 1..4194304"; return GMX_ERR_INVALID; }
            ctx->chunk_reads = (size_t)value; return GMX_OK;
        case GMX_OPT_VOTE_FILTER: ctx->use_filter = value != 0; return GMX_OK;
        case GMX_OPT_VOTE_SLOTS:
            if (value != 4 && value != 6) { ctx->err = "vote_slots must be 4 or 6"; return GMX_ERR_INVALID; }
            ctx->vote_slots = (int)value; return GMX_OK;
        case GMX_OPT_SAM_DEVICE: ctx->sam_on_device = value != 0; return GMX_OK;
        case GMX_OPT_FASTQ_PIECE:
            if (value < 0) { ctx->err = "fastq piece bytes must be >= 0"; return GMX_ERR_INVALID; }
            ctx->fq_piece_bytes = value; return GMX_OK;
        case GMX_OPT_STAGE_TIMING: ctx->stage_timing = value != 0; return GMX_OK;
        case GMX_OPT_OPTIMISTIC: ctx->optimistic = value < 0 ? 0 : (value > 2 ? 2 : (int)value); return GMX_OK;
        case GMX_OPT_VOTE_COMPACT: ctx->vote_compact = value < 0 ? 0 : (value > 2 ? 2 : (int)value); return GMX_OK;
        case GMX_OPT_CIGAR_STRIDE:
            if (value < 16 || value > 2048 || (value & 15)) { ctx->err = "cigar_stride must be a multiple of 16 in 16..2048"; return GMX_ERR_INVALID; }
            ctx->cigar_stride = (int)value; ctx->mapped = false; ctx->scored = false; ctx->cs.valid = false; return GMX_OK;
        case GMX_OPT_FILTER_SHIFT:
            if (value < -2 || value > 2) { ctx->err = "filter_shift must be in -2..2"; return GMX_ERR_INVALID; }
            ctx->filter_shift = (int)value; return GMX_OK;
        default: ctx->err = "unknown option"; return GMX_ERR_INVALID;
    }
}

extern "C" int gmx_get_hits(gmx_ctx *ctx, gmx_hit *hits, int64_t capacity, int64_t *n_hits)
{
    if (!ctx || !n_hits) return GMX_ERR_INVALID;
    if (!ctx->mapped) return GMX_ERR_STATE;
    if (!ctx->collect_hits) { ctx->err = "hit collection is disabled (GMX_OPT_COLLECT_HITS = 0)"; return GMX_ERR_STATE; }
    *n_hits = (int64_t)ctx->h_hits.size();
    if (!hits) return GMX_OK;
    if (capacity < *n_hits) return GMX_ERR_OVERFLOW;
    memcpy(hits, ctx->h_hits.data(), ctx->h_hits.size() * sizeof(gmx_hit));
    return GMX_OK;
}

extern "C" int gmx_get_best_alignments(gmx_ctx *ctx, char *cigar_out, int32_t cigar_stride, uint8_t *aligned_out, int32_t aligned_stride)
{
    if (!ctx) return GMX_ERR_INVALID;
    if (!ctx->scored) return GMX_ERR_STATE;
    if (aligned_out && !ctx->collect_hits) { ctx->err = "aligned strings need GMX_OPT_COLLECT_HITS = 1"; return GMX_ERR_STATE; }
    const char *best_cigar = ctx->h_best_cigar.as<char>();
    for (int32_t i = 0; i < ctx->last_n_reads; ++i) {
        if (cigar_out) {
            memset(cigar_out + (size_t)i * cigar_stride, 0, (size_t)cigar_stride);
            strncpy(cigar_out + (size_t)i * cigar_stride, best_cigar + (size_t)i * GMX_CIGAR_STRIDE, (size_t)std::min(cigar_stride - 1, GMX_CIGAR_STRIDE - 1));
        }
        if (aligned_out) {
            memset(aligned_out + (size_t)i * aligned_stride, 0, (size_t)aligned_stride);
            memcpy(aligned_out + (size_t)i * aligned_stride, &ctx->h_best_aligned[(size_t)i * ctx->h_a_stride], (size_t)std::min(aligned_stride, ctx->h_a_stride));
        }
    }
    return GMX_OK;
}

extern "C" int gmx_get_stage_stats(gmx_ctx *ctx, gmx_stage_stats *out)
{
    if (!ctx || !out) return GMX_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    out->n_stages = ST_COUNT;
    for (int s = 0; s < ST_COUNT && s < GMX_N_STAGES; ++s) {
        out->name[s] = kStageNames[s]; out->ms[s] = ctx->stage_ms[s]; out->units[s] = ctx->stage_units[s];
        out->bytes[s] = ctx->stage_bytes[s]; out->launches[s] = ctx->stage_launches[s];
    }
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// next row: FASTQ text -> reads
// ------------------------------------------------------------------------------------------------
// SeqReader::get_more_fastq, reference src/SeqReader.cpp:1023-1292, line by line (READ_BUFFER batching, adaptor
// trimming and PWM construction aside: the PWM is built on the device from (base, quality char)).
extern "C" int gmx_fastq_scan_host(const char *text, int64_t len, int illumina, gmx_fastq_rec *recs, int64_t capacity, int64_t *n_recs)
{
    if (!text || len < 0 || !n_recs || (capacity > 0 && !recs)) return GMX_ERR_INVALID;
    FastqLines in{text, len, 0, false};
    int64_t n = 0;
    int rc = GMX_OK;
    int qbase = illumina ? 64 : 33;
    // the four line buffers live across records, as the reference's strings do (:1049): a getline on a stream that
    // already hit end-of-file leaves its string untouched, so a truncated last record can pick up the previous
    // record's lines -- the reference emits such reads, and so does this scan
    int64_t no = 0, nn = 0, so = 0, sn = 0, po = 0, pn = 0, qo = 0, qn = 0;
    while (true) {
        if (in.eof) break;                                         // :1060
        in.getline(no, nn);                                        // :1070
        while (nn == 0 && !in.eof) in.getline(no, nn);             // :1073-1076 blank lines
        if (in.eof) break;                                         // :1079
        in.getline(so, sn); in.getline(po, pn); in.getline(qo, qn);                      // :1086-1088
        bool ended = false;
        while (in.first(no, nn) != '@' || in.first(po, pn) != '+' || sn > qn) {           // :1091
            if (in.first(no, nn) != '@' || in.first(po, pn) != '+') {                      // :1095
                while ((in.first(no, nn) != '@' || in.first(po, pn) != '+') && !in.eof) {
                    no = so; nn = sn; so = po; sn = pn; po = qo; pn = qn;               // shift the four lines up
                    in.getline(qo, qn);
                    if (in.eof) { ended = true; break; }                                  // :1106
                }
                if (ended || in.eof) { ended = true; break; }                             // :1114
            }
            if (sn > qn) {                                                                // :1125
                in.getline(no, nn); in.getline(so, sn); in.getline(po, pn); in.getline(qo, qn);
                if (in.eof) { ended = true; break; }                                      // :1135
            }
        }
        if (ended) break;
        // quality characters below the offset give max_prb < 0: with --illumina the reference turns the flag off and
        // restarts the read at offset 33 (:1180-1188), otherwise it throws (:1189-1196)
        for (int64_t i = 0; i < sn; ++i) {
            int Q = (int)(unsigned char)text[qo + i] - qbase;
            if (Q < 0) {
                if (qbase == 64) { qbase = 33; i = -1; continue; }
                rc = GMX_ERR_FORMAT;
                break;
            }
        }
        if (rc != GMX_OK) break;
        if (n < capacity) {
            gmx_fastq_rec &r = recs[n];
            r.name_off = no + 1; r.name_len = (int32_t)(nn - 1); r.seq_off = so; r.seq_len = (int32_t)sn;
            r.qual_off = qo; r.qual_len = (int32_t)qn; r.pad = 0;
        }
        ++n;
    }
    *n_recs = n;
    if (rc != GMX_OK) return rc;
    return n > capacity ? GMX_ERR_OVERFLOW : GMX_OK;
}

// device indexer; leaves seq_off / qual_off / seq_len / recs in ctx buffers and the text on the device
static int fastq_scan_device(gmx_ctx *ctx, const char *text, int64_t len, int text_on_device, int64_t *n_recs, int32_t *max_len, const char **d_text_out)
{
    if (len >= 0x7fffffffll) { ctx->err = "FASTQ text of 2 GiB or more per call: split it at a record boundary"; return GMX_ERR_UNSUPPORTED; }
    const char *d_text = text;
    if (!text_on_device) {
        CK(ctx->d_fq_text.ensure((size_t)len + 16));
        CK(cudaMemcpyAsync(ctx->d_fq_text.p, text, (size_t)len, cudaMemcpyHostToDevice, ctx->stream));
        d_text = ctx->d_fq_text.as<char>();
    }
    *d_text_out = d_text;
    *n_recs = 0; *max_len = 0;
    if (len == 0) return GMX_OK;
    // newline positions, in order: count them first so that the position list is sized exactly
    CK(ctx->d_fq_count.ensure(16));
    thrust::counting_iterator<uint32_t> idx(0);
    IsNewline pred{d_text};
    CK(cudaMemsetAsync(ctx->d_fq_count.p, 0, 16, ctx->stream));
    k_count_newlines<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(d_text, len, ctx->d_fq_count.as<unsigned long long>() + 1);
    uint32_t n_expect = 0;
    {
        unsigned long long h = 0;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&h, ctx->d_fq_count.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        n_expect = (uint32_t)h;
    }
    CK(ctx->d_fq_nl.ensure(((size_t)n_expect + 16) * 4));
    size_t tmp_bytes = 0;
    CK(cub::DeviceSelect::If(nullptr, tmp_bytes, idx, ctx->d_fq_nl.as<uint32_t>(), ctx->d_fq_count.as<uint32_t>(), (int)len, pred, ctx->stream));
    CK(ctx->d_fq_tmp.ensure(tmp_bytes));
    CK(cub::DeviceSelect::If(ctx->d_fq_tmp.p, tmp_bytes, idx, ctx->d_fq_nl.as<uint32_t>(), ctx->d_fq_count.as<uint32_t>(), (int)len, pred, ctx->stream));
    uint32_t n_nl = 0;
    char last = 0;
    CK(cudaMemcpyAsync(&n_nl, ctx->d_fq_count.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&last, d_text + len - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint64_t n_lines = (uint64_t)n_nl + (last != '\n' ? 1 : 0);
    if (n_lines % 4 != 0) { ctx->err = "FASTQ text is not a whole number of 4-line records (blank or missing lines): use gmx_fastq_scan_host"; return GMX_ERR_FORMAT; }
    const uint32_t n = (uint32_t)(n_lines / 4);
    if (n == 0) return GMX_OK;
    CK(ctx->d_fq_seq_off.ensure((size_t)n * 8)); CK(ctx->d_fq_qual_off.ensure((size_t)n * 8)); CK(ctx->d_fq_len.ensure((size_t)n * 4));
    CK(ctx->d_fq_recs.ensure((size_t)n * sizeof(gmx_fastq_rec))); CK(ctx->d_fq_flags.ensure(16));
    const uint32_t init[4] = {0u, 0u, 0xffffffffu, 0u};
    CK(cudaMemcpyAsync(ctx->d_fq_flags.p, init, 16, cudaMemcpyHostToDevice, ctx->stream));
    FastqDev out;
    out.seq_off = ctx->d_fq_seq_off.as<int64_t>(); out.qual_off = ctx->d_fq_qual_off.as<int64_t>(); out.seq_len = ctx->d_fq_len.as<int32_t>();
    out.recs = ctx->d_fq_recs.as<gmx_fastq_rec>(); out.flags = ctx->d_fq_flags.as<uint32_t>();
    k_fastq_records<<<nblk(n, 256), 256, 0, ctx->stream>>>(d_text, len, ctx->d_fq_nl.as<uint32_t>(), n_nl, n, ctx->params.illumina ? 64 : 33, out, 0);
    CK(cudaGetLastError());
    uint32_t flags[4];
    CK(cudaMemcpyAsync(flags, ctx->d_fq_flags.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (flags[0]) {
        char b[160];
        snprintf(b, sizeof(b), "%u malformed FASTQ record(s), first at record %u: use gmx_fastq_scan_host", flags[0], flags[2]);
        ctx->err = b;
        return GMX_ERR_FORMAT;
    }
    *n_recs = n; *max_len = (int32_t)flags[1];
    return GMX_OK;
}

extern "C" int gmx_fastq_scan(gmx_ctx *ctx, const char *text, int64_t len, int text_on_device, gmx_fastq_rec *recs, int64_t capacity, int64_t *n_recs)
{
    if (!ctx || !text || len < 0 || !n_recs) return GMX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int32_t max_len = 0; const char *d_text = nullptr;
    int r = fastq_scan_device(ctx, text, len, text_on_device, n_recs, &max_len, &d_text);
    if (r != GMX_OK) return r;
    if (recs) {
        if (capacity < *n_recs) return GMX_ERR_OVERFLOW;
        CK(cudaMemcpyAsync(recs, ctx->d_fq_recs.p, (size_t)*n_recs * sizeof(gmx_fastq_rec), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return GMX_OK;
}

// ---- FASTQ text from the host, pipelined ---------------------------------------------------------------
// The text is cut at record boundaries into pieces; piece p + 1 crosses PCIe and is indexed while piece p is mapped, so
// that only the first piece's upload + index stay exposed (the whole-text path below exposes all of it: 3.9 ms of PCIe
// + the indexer per 1 M x 100 bp reads, a quarter of the path's own time).

// First record start at or after `from`: a line that starts with '@' whose next-but-one line starts with '+'.  A quality
// line may start with '@' too, but the line two below a quality line is a sequence line, which never starts with '+'.
static int64_t fastq_record_cut(const char *text, int64_t len, int64_t from)
{
    if (from <= 0) return 0;
    const char *q = (const char *)memchr(text + from - 1, '\n', (size_t)(len - (from - 1)));
    for (int tries = 0; tries < 16 && q; ++tries) {
        const int64_t l0 = (q - text) + 1;
        if (l0 >= len) return len;
        const char *q1 = (const char *)memchr(text + l0, '\n', (size_t)(len - l0));
        if (!q1) return -1;
        const char *q2 = (const char *)memchr(q1 + 1, '\n', (size_t)(len - (q1 + 1 - text)));
        if (!q2) return -1;
        const int64_t l2 = (q2 - text) + 1;
        if (text[l0] == '@' && l2 < len && text[l2] == '+') return l0;
        q = q1;
    }
    return -1;
}

// index one piece [at, at + plen) of the device text on `st`; appends its reads at `first`.  Host-synchronous on `st`.
static int fastq_index_piece(gmx_ctx *ctx, const char *d_text, int64_t at, int64_t plen, int64_t first, int64_t n_cap, cudaStream_t st,
                             int64_t *n_out, int32_t *max_len_out)
{
    *n_out = 0; *max_len_out = 0;
    if (plen <= 0) return GMX_OK;
    if (plen >= 0x7fffffffll) { ctx->err = "FASTQ piece of 2 GiB or more"; return GMX_ERR_UNSUPPORTED; }
    const char *pt = d_text + at;
    CK(ctx->d_fq_count.ensure(16));
    CK(cudaMemsetAsync(ctx->d_fq_count.p, 0, 16, st));
    k_count_newlines<<<ctx->n_sm * 8, 256, 0, st>>>(pt, plen, ctx->d_fq_count.as<unsigned long long>() + 1);
    CK(cudaGetLastError());
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, ctx->d_fq_count.as<unsigned long long>() + 1, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_expect = (uint32_t)h;
    CK(ctx->d_fq_nl.ensure(((size_t)n_expect + 16) * 4));
    thrust::counting_iterator<uint32_t> idx(0);
    IsNewline pred{pt};
    size_t tmp_bytes = 0;
    CK(cub::DeviceSelect::If(nullptr, tmp_bytes, idx, ctx->d_fq_nl.as<uint32_t>(), ctx->d_fq_count.as<uint32_t>(), (int)plen, pred, st));
    CK(ctx->d_fq_tmp.ensure(tmp_bytes));
    CK(cub::DeviceSelect::If(ctx->d_fq_tmp.p, tmp_bytes, idx, ctx->d_fq_nl.as<uint32_t>(), ctx->d_fq_count.as<uint32_t>(), (int)plen, pred, st));
    char last = 0;
    CK(cudaMemcpyAsync(&last, pt + plen - 1, 1, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t n_nl = n_expect;
    const uint64_t n_lines = (uint64_t)n_nl + (last != '\n' ? 1 : 0);
    if (n_lines % 4 != 0) { ctx->err = "FASTQ text is not a whole number of 4-line records (blank or missing lines): use gmx_fastq_scan_host"; return GMX_ERR_FORMAT; }
    const uint32_t n = (uint32_t)(n_lines / 4);
    if (n == 0) return GMX_OK;
    if (first + (int64_t)n > n_cap) { *n_out = n; return GMX_ERR_OVERFLOW; }
    CK(ctx->d_fq_flags.ensure(16));
    const uint32_t init[4] = {0u, 0u, 0xffffffffu, 0u};
    CK(cudaMemcpyAsync(ctx->d_fq_flags.p, init, 16, cudaMemcpyHostToDevice, st));
    FastqDev out;
    out.seq_off = ctx->d_fq_seq_off.as<int64_t>() + first; out.qual_off = ctx->d_fq_qual_off.as<int64_t>() + first; out.seq_len = ctx->d_fq_len.as<int32_t>() + first;
    out.recs = ctx->d_fq_recs.as<gmx_fastq_rec>() + first; out.flags = ctx->d_fq_flags.as<uint32_t>();
    k_fastq_records<<<nblk(n, 256), 256, 0, st>>>(pt, plen, ctx->d_fq_nl.as<uint32_t>(), n_nl, n, ctx->params.illumina ? 64 : 33, out, at);
    CK(cudaGetLastError());
    uint32_t flags[4];
    CK(cudaMemcpyAsync(flags, ctx->d_fq_flags.p, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (flags[0]) {
        char b[200];
        snprintf(b, sizeof(b), "%u malformed FASTQ record(s), first at record %lld: use gmx_fastq_scan_host", flags[0], (long long)first + flags[2]);
        ctx->err = b;
        return GMX_ERR_FORMAT;
    }
    *n_out = n; *max_len_out = (int32_t)flags[1];
    return GMX_OK;
}

// GMX_ERR_STATE: the text cannot be cut (caller takes the whole-text path)
static int process_fastq_pipelined(gmx_ctx *ctx, const char *text, int64_t len, gmx_read_result *results, int64_t capacity, int64_t *n_reads, gmx_fastq_rec *recs)
{
    const int64_t piece = ctx->fq_piece_bytes;
    std::vector<int64_t> cuts(1, 0);
    while (true) {
        // a short first piece: its upload and index are the exposed ones.  (A longer ramp -- 1/6, 1/3, 2/3 of the regular
        // piece -- measured slower, 47.6 against 49.8 M reads/s: small chunks run the kernels at a worse rate than the
        // transfer they hide, as in run_batch.)
        const int64_t want = cuts.back() + (cuts.size() == 1 ? std::max<int64_t>(piece / 4, 1) : piece);
        if (want >= len - piece / 8) break;
        const int64_t c = fastq_record_cut(text, len, want);
        if (c < 0) return GMX_ERR_STATE;
        if (c >= len || c <= cuts.back()) break;
        cuts.push_back(c);
    }
    cuts.push_back(len);
    const size_t P = cuts.size() - 1;
    if (P < 2) return GMX_ERR_STATE;
    const int64_t n_cap = std::min<int64_t>((results || recs) ? capacity : len / 6 + 1, 0x7ffffff0ll);
    if (n_cap <= 0) return GMX_ERR_STATE;
    CK(ctx->d_fq_text.ensure((size_t)len + 16));
    CK(ctx->d_fq_seq_off.ensure((size_t)n_cap * 8)); CK(ctx->d_fq_qual_off.ensure((size_t)n_cap * 8)); CK(ctx->d_fq_len.ensure((size_t)n_cap * 4));
    CK(ctx->d_fq_recs.ensure((size_t)n_cap * sizeof(gmx_fastq_rec)));
    char *d_text = ctx->d_fq_text.as<char>();
    { int r = batch_begin(ctx, (int32_t)n_cap, 0, results, true); if (r != GMX_OK) return r; }
    CK(cudaStreamSynchronize(ctx->stream));               // batch_begin's resets precede everything issued on the copy stream
    auto upload = [&](size_t p) -> int {
        CK(cudaMemcpyAsync(d_text + cuts[p], text + cuts[p], (size_t)(cuts[p + 1] - cuts[p]), cudaMemcpyHostToDevice, ctx->copy_stream));
        return GMX_OK;
    };
    int64_t first = 0, n_p = 0;
    int32_t max_p = 0;
    int rc = upload(0);
    if (rc == GMX_OK) rc = fastq_index_piece(ctx, d_text, cuts[0], cuts[1] - cuts[0], 0, n_cap, ctx->copy_stream, &n_p, &max_p);
    if (rc != GMX_OK) { *n_reads = rc == GMX_ERR_OVERFLOW ? n_p : 0; return rc; }
    // the record index of a piece leaves for the host as soon as the piece is indexed (the host has waited for its kernels)
    auto send_recs = [&](int64_t at, int64_t cnt) -> int {
        if (recs && cnt) CK(cudaMemcpyAsync(recs + at, ctx->d_fq_recs.as<gmx_fastq_rec>() + at, (size_t)cnt * sizeof(gmx_fastq_rec), cudaMemcpyDeviceToHost, ctx->d2h_stream));
        return GMX_OK;
    };
    rc = send_recs(0, n_p);
    if (rc != GMX_OK) return rc;
    gmx_reads in;
    memset(&in, 0, sizeof(in));
    in.offsets = ctx->d_fq_seq_off.as<int64_t>(); in.qual_offsets = ctx->d_fq_qual_off.as<int64_t>(); in.lens = ctx->d_fq_len.as<int32_t>();
    in.seq = reinterpret_cast<const uint8_t *>(d_text); in.qual = reinterpret_cast<const uint8_t *>(d_text);
    in.on_device = 1;
    int slot = 0;
    for (size_t p = 0; p < P; ++p) {
        if (p + 1 < P) { rc = upload(p + 1); if (rc != GMX_OK) return rc; }       // crosses PCIe while piece p is mapped
        in.n_reads = (int32_t)(first + n_p); in.max_len = std::max(max_p, 1);
        const int64_t step = (int64_t)ctx->chunk_reads, end = first + n_p;
        // like run_batch, the batch ends with a short chunk: the last download is the one nothing hides
        const int64_t edge = std::max<int64_t>(step / 8, 1);
        const bool tail = p + 1 == P && n_p > 4 * edge;
        for (int64_t lo = first, hi = first; lo < end; lo = hi, slot ^= 1) {
            hi = std::min(end, lo + step);
            if (tail && lo < end - edge) hi = std::min(hi, end - edge);
            const int prev = ctx->pend[0].active ? 0 : (ctx->pend[1].active ? 1 : -1);
            rc = issue_upload(ctx, &in, (int32_t)lo, (int32_t)hi, slot, ctx->stream);      // device-resident: a view, no copy
            if (rc == GMX_OK) rc = run_chunk(ctx, &in, (int32_t)lo, (int32_t)hi, slot, results, chunk_can_be_optimistic(ctx, true));
            if (rc == GMX_OK && prev >= 0) rc = settle_chunk(ctx, prev);                   // before the next view takes its slot
            if (rc != GMX_OK) return rc;
        }
        first += n_p;
        if (p + 1 < P) {                                   // its kernels and host waits run beside piece p's PHASE B and download
            rc = fastq_index_piece(ctx, d_text, cuts[p + 1], cuts[p + 2] - cuts[p + 1], first, n_cap, ctx->copy_stream, &n_p, &max_p);
            if (rc != GMX_OK) {
                // the reads before this piece have been mapped AND scored: the caller continues from there
                ctx->cs.valid = false;
                const std::string why = ctx->err;
                ctx->last_n_reads = (int32_t)first;
                int r2 = batch_end(ctx, true);
                *n_reads = first;
                ctx->err = why + " (the reads before it have been mapped and scored)";
                if (r2 == GMX_OK) { ctx->fq_batch = ctx->batch_dev_valid; ctx->fq_d_text = d_text; ctx->fq_len = len; }
                return rc;
            }
            rc = send_recs(first, n_p);
            if (rc != GMX_OK) return rc;
        }
    }
    ctx->cs.valid = false;
    ctx->last_n_reads = (int32_t)first;
    rc = batch_end(ctx, true);
    if (rc != GMX_OK) return rc;
    *n_reads = first;
    if (ctx->batch_dev_valid) { ctx->fq_batch = true; ctx->fq_d_text = d_text; ctx->fq_len = len; }
    return GMX_OK;
}

extern "C" int gmx_process_fastq(gmx_ctx *ctx, const char *text, int64_t len, int text_on_device, gmx_read_result *results, int64_t capacity,
                                 int64_t *n_reads, gmx_fastq_rec *recs)
{
    if (!ctx || !text || len < 0 || !n_reads) return GMX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (!text_on_device && !ctx->collect_hits && ctx->fq_piece_bytes > 0 && len >= 2 * ctx->fq_piece_bytes) {
        int r = process_fastq_pipelined(ctx, text, len, results, capacity, n_reads, recs);
        if (r != GMX_OK && r != GMX_ERR_STATE) {           // every error leaves through the drain of run_batch's error exit
            const std::string why = ctx->err;
            cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->d2h_stream);
            cudaGetLastError();
            ctx->cs.valid = false;
            ctx->pend[0].active = ctx->pend[1].active = false;
            ctx->err = why;
        }
        if (r != GMX_ERR_STATE) return r;
    }
    int32_t max_len = 0; const char *d_text = nullptr;
    int r = fastq_scan_device(ctx, text, len, text_on_device, n_reads, &max_len, &d_text);
    if (r != GMX_OK) return r;
    if ((results || recs) && capacity < *n_reads) return GMX_ERR_OVERFLOW;
    if (*n_reads > 0x7fffffffll) return GMX_ERR_UNSUPPORTED;
    if (recs && *n_reads) CK(cudaMemcpyAsync(recs, ctx->d_fq_recs.p, (size_t)*n_reads * sizeof(gmx_fastq_rec), cudaMemcpyDeviceToHost, ctx->d2h_stream));
    gmx_reads in;
    memset(&in, 0, sizeof(in));
    in.n_reads = (int32_t)*n_reads;
    in.offsets = ctx->d_fq_seq_off.as<int64_t>();
    in.seq = reinterpret_cast<const uint8_t *>(d_text); in.qual = reinterpret_cast<const uint8_t *>(d_text);
    in.qual_offsets = ctx->d_fq_qual_off.as<int64_t>(); in.lens = ctx->d_fq_len.as<int32_t>();
    in.on_device = 1; in.max_len = std::max(max_len, 1);
    r = run_batch(ctx, &in, results, true);
    if (r == GMX_OK && ctx->batch_dev_valid) { ctx->fq_batch = true; ctx->fq_d_text = d_text; ctx->fq_len = len; }
    return r;
}

// ------------------------------------------------------------------------------------------------
// next row: SAM emission
// ------------------------------------------------------------------------------------------------
namespace {
struct SamPos { uint64_t pos; int strand; };

inline char sam_rc(char c)
{   // reverse_comp, reference inc/SequenceOperations.h:56-96: anything that is not acgtACGT- becomes 'n'
    switch (c) {
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case '-': return '-'; default: return 'n';
    }
}

// reverse_CIGAR, reference inc/SequenceOperations.h:109-123 (its digit test is 48..58)
inline void sam_reverse_cigar(const char *c, std::string &out)
{
    out.clear();
    std::string num, rev;
    for (; *c; ++c) {
        if (*c >= 48 && *c <= 58) num += *c;
        else { rev = num + *c + rev; num.clear(); }
    }
    out = rev;
}

inline void sam_put_int(std::string &out, int64_t v)
{
    char b[24]; int n = 0;
    const bool neg = v < 0; uint64_t u = neg ? (uint64_t)(-v) : (uint64_t)v;
    do { b[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) out += '-';
    while (n) out += b[--n];
}

inline int sam_mapq(double total)
{   // reference inc/ScoredSeq.h:302-309
    int q;
    if (total == 1) q = 30;
    else {
        double v = 1 - total;
        q = v <= 0 ? 30 : (int)round(-10 * log(v) / log(10.0));
    }
    return q > 30 ? 30 : q;
}

void sam_format_range(const gmx_ctx *ctx, const char *text, const gmx_fastq_rec *recs, const gmx_read_result *results, int64_t lo, int64_t hi,
                      const char *const *chrom_names, const std::vector<std::vector<SamPos>> *multi_of, const std::vector<int64_t> &multi_index,
                      std::string &out)
{
    const std::vector<int64_t> &off = ctx->h_seq_offset;
    const char *best_cigar = ctx->h_best_cigar.as<char>();
    char num[64];
    std::string rcig, seq_rc, qual_rv;
    for (int64_t r = lo; r < hi; ++r) {
        const gmx_read_result &res = results[r];
        if (!GMX_READ_PRINTS_SAM(res)) continue;                            // unmapped reads print nothing (Driver.cpp:620-629); nor
                                                                            // does a best group below top - SAME_DIFF (:695)
        const gmx_fastq_rec &rec = recs[r];
        const char *cigar = best_cigar + (size_t)r * GMX_CIGAR_STRIDE;
        const double total = exp((double)res.best_score) / res.denominator;
        const int q = sam_mapq(total);
        const double xa = (double)res.best_score * (1.0 / (double)ctx->params.adjust);     // XA prints score / gADJUST (Driver.cpp:2202)
        // positions of the best group, ascending (pos, strand) as std::set iterates them
        SamPos one{res.best_first_pos, res.best_first_strand};
        const SamPos *ps = &one; size_t np = 1;
        if (res.best_n_positions > 1) {
            const int64_t mi = multi_index[r];
            if (mi >= 0) { ps = (*multi_of)[mi].data(); np = (*multi_of)[mi].size(); }
        }
        bool have_rc = false;
        for (size_t k = 0; k < np; ++k) {
            const uint64_t pos = ps[k].pos; const bool neg = ps[k].strand == GMX_NEG_STRAND;
            size_t rid = std::upper_bound(off.begin(), off.end() - 1, (int64_t)pos) - off.begin() - 1;
            out.append(text + rec.name_off, (size_t)rec.name_len);
            out += neg ? "\t16\t" : "\t0\t";
            out += chrom_names[rid];
            out += '\t'; sam_put_int(out, (int64_t)pos - off[rid] + 1); out += '\t'; sam_put_int(out, q); out += '\t';
            if (neg) { sam_reverse_cigar(cigar, rcig); out += rcig; } else out += cigar;
            out += "\t*\t0\t0\t";
            if (neg) {
                if (!have_rc) {
                    seq_rc.assign((size_t)rec.seq_len, 'n'); qual_rv.assign((size_t)rec.qual_len, '!');
                    for (int i = 0; i < rec.seq_len; ++i) seq_rc[i] = sam_rc(text[rec.seq_off + rec.seq_len - 1 - i]);
                    for (int i = 0; i < rec.qual_len; ++i) qual_rv[i] = text[rec.qual_off + rec.qual_len - 1 - i];
                    have_rc = true;
                }
                out += seq_rc; out += '\t'; out += qual_rv;
            } else {
                out.append(text + rec.seq_off, (size_t)rec.seq_len); out += '\t'; out.append(text + rec.qual_off, (size_t)rec.qual_len);
            }
            if (k == 0) snprintf(num, sizeof(num), "\tXA:f:%g\tXP:f:%g\tX0:i:%d\n", xa, (double)res.best_posterior, res.best_n_positions);
            out += num;
        }
    }
}
}  // namespace

// The batch last run by gmx_process_fastq, formatted where its text, record index, results and CIGARs already are
// (sam_out.cuh); reads whose best group holds several positions are written by the host formatter into the place the
// device reserved for them.
static int format_sam_device(gmx_ctx *ctx, const char *text, const gmx_fastq_rec *recs, const gmx_read_result *results, int64_t n_reads,
                             const char *const *chrom_names, char *out, int64_t cap, int64_t *len)
{
    CK(cudaSetDevice(ctx->device));
    const int n = (int)n_reads;
    *len = 0;
    if (n == 0) return GMX_OK;
    const int n_seqs = ctx->ix.n_seqs;
    // chromosome names: chars | offsets | lengths
    std::vector<int32_t> noff((size_t)n_seqs), nlen((size_t)n_seqs);
    std::string chars;
    for (int i = 0; i < n_seqs; ++i) { noff[i] = (int32_t)chars.size(); nlen[i] = (int32_t)strlen(chrom_names[i]); chars += chrom_names[i]; }
    const size_t chars_pad = (chars.size() + 15) & ~(size_t)15;
    CK(ctx->d_sam_names.ensure(chars_pad + (size_t)n_seqs * 8 + 16));
    CK(cudaMemcpyAsync(ctx->d_sam_names.p, chars.data(), chars.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_sam_names.as<char>() + chars_pad, noff.data(), (size_t)n_seqs * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_sam_names.as<char>() + chars_pad + (size_t)n_seqs * 4, nlen.data(), (size_t)n_seqs * 4, cudaMemcpyHostToDevice, ctx->stream));
    SamNames names;
    names.chars = ctx->d_sam_names.as<char>(); names.off = reinterpret_cast<const int32_t *>(ctx->d_sam_names.as<char>() + chars_pad);
    names.len = names.off + n_seqs; names.n = n_seqs;

    CK(ctx->d_sam_pieces.ensure((size_t)n * sizeof(SamPiece)));
    CK(ctx->d_sam_cigar.ensure((size_t)n * GMX_CIGAR_STRIDE));
    CK(ctx->d_sam_lens.ensure(((size_t)n + 1) * 8)); CK(ctx->d_sam_offs.ensure(((size_t)n + 1) * 8));
    CK(ctx->d_sam_extra.ensure((size_t)n * 8 + 16));
    uint32_t *d_unc = reinterpret_cast<uint32_t *>(ctx->d_sam_extra.as<char>() + (size_t)n * 8);
    CK(cudaMemsetAsync(ctx->d_sam_extra.p, 0, (size_t)n * 8 + 16, ctx->stream));
    long long *d_lens = ctx->d_sam_lens.as<long long>(), *d_offs = ctx->d_sam_offs.as<long long>();
    const gmx_fastq_rec *d_recs = ctx->d_fq_recs.as<gmx_fastq_rec>();
    k_sam_measure<<<nblk(n, 128), 128, 0, ctx->stream>>>(ctx->d_batch_results.as<gmx_read_result>(), d_recs, ctx->d_batch_cigar.as<char>(), GMX_CIGAR_STRIDE, n,
                                                        ctx->ix, names, 1.0 / (double)ctx->params.adjust, ctx->d_sam_pieces.as<SamPiece>(),
                                                        ctx->d_sam_cigar.as<char>(), d_lens, d_unc);
    CK(cudaGetLastError());
    const uint32_t n_multi = (uint32_t)ctx->h_multi.size();
    if (n_multi) {
        k_sam_multi_len<<<nblk(n_multi, 256), 256, 0, ctx->stream>>>(ctx->d_multi.as<MultiPos>(), n_multi, ctx->ix, names, d_lens,
                                                                    ctx->d_sam_extra.as<unsigned long long>());
        CK(cudaGetLastError());
    }
    k_sam_fix_lens<<<nblk(n, 256), 256, 0, ctx->stream>>>(d_lens, ctx->d_sam_extra.as<unsigned long long>(), n);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(d_lens + n, 0, 8, ctx->stream));
    size_t tmp_bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_lens, d_offs, n + 1, ctx->stream));
    CK(ctx->d_sam_tmp.ensure(tmp_bytes));
    CK(cub::DeviceScan::ExclusiveSum(ctx->d_sam_tmp.p, tmp_bytes, d_lens, d_offs, n + 1, ctx->stream));
    long long total = 0; uint32_t unc = 0;
    CK(cudaMemcpyAsync(&total, d_offs + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&unc, d_unc, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (unc) return GMX_ERR_STATE;
    *len = total;
    if (total > cap) return GMX_ERR_OVERFLOW;
    CK(ctx->d_sam_out.ensure((size_t)total + 16));
    k_sam_write<<<nblk((int64_t)n * 32, 256), 256, 0, ctx->stream>>>(ctx->fq_d_text, d_recs, ctx->d_sam_pieces.as<SamPiece>(), ctx->d_sam_cigar.as<char>(),
                                                                    GMX_CIGAR_STRIDE, d_offs, n, names, ctx->d_sam_out.as<char>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->d_sam_out.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    // the reads with several positions: their records from the host formatter, into the reserved places
    std::vector<int64_t> multi_index((size_t)n_reads, -1);
    std::vector<std::vector<SamPos>> multi_of;
    for (const MultiPos &m : ctx->h_multi)
        if (m.read >= 0 && m.read < n_reads) {
            if (multi_index[m.read] < 0) { multi_index[m.read] = (int64_t)multi_of.size(); multi_of.emplace_back(); }
            multi_of[multi_index[m.read]].push_back(SamPos{m.pos, m.strand});
        }
    for (auto &v : multi_of) std::sort(v.begin(), v.end(), [](const SamPos &a, const SamPos &b) { return a.pos != b.pos ? a.pos < b.pos : a.strand < b.strand; });
    std::vector<long long> offs;
    if (!multi_of.empty()) {
        offs.resize((size_t)n + 1);
        CK(cudaMemcpyAsync(offs.data(), d_offs, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    std::string one;
    for (int64_t r = 0; r < n_reads; ++r) {
        if (multi_index[r] < 0 || !GMX_READ_PRINTS_SAM(results[r])) continue;
        one.clear();
        sam_format_range(ctx, text, recs, results, r, r + 1, chrom_names, &multi_of, multi_index, one);
        if ((long long)one.size() != offs[r + 1] - offs[r]) { ctx->err = "device SAM formatter: a multi-position read's size differs from the host's"; return GMX_ERR_CUDA; }
        memcpy(out + offs[r], one.data(), one.size());
    }
    return GMX_OK;
}

// "%g" as the device SAM formatter writes it (sam_out.cuh), callable on the host
extern "C" int gmx_format_g(double v, char *out, int cap)
{
    char b[40];
    const int n = gmx_fmt_g6(v, b);
    if (n < 0) return GMX_ERR_UNSUPPORTED;
    if (!out || n + 1 > cap) return GMX_ERR_OVERFLOW;
    memcpy(out, b, (size_t)n); out[n] = 0;
    return n;
}

extern "C" int gmx_format_sam(gmx_ctx *ctx, const char *text, const gmx_fastq_rec *recs, const gmx_read_result *results, int64_t n_reads,
                              const char *const *chrom_names, char *out, int64_t cap, int64_t *len)
{
    if (!ctx || !text || !recs || !results || !chrom_names || !len || n_reads < 0 || (cap > 0 && !out)) return GMX_ERR_INVALID;
    if (!ctx->scored || n_reads != ctx->last_n_reads) { ctx->err = "gmx_format_sam formats the batch last scored"; return GMX_ERR_STATE; }
    if (!ctx->collect_hits && ctx->multi_overflow) { ctx->err = "more multi-position hits than the fast path keeps (4 per read): set GMX_OPT_COLLECT_HITS"; return GMX_ERR_OVERFLOW; }
    if (ctx->sam_on_device && ctx->fq_batch && ctx->batch_dev_valid && !ctx->collect_hits) {
        int r = format_sam_device(ctx, text, recs, results, n_reads, chrom_names, out, cap, len);
        if (r != GMX_ERR_STATE) return r;                  // GMX_ERR_STATE: a number the device formatter does not cover -- host path
    }
    // positions of the multi-position best groups: from the hit list (collect mode) or from the device list (fast path)
    std::vector<int64_t> multi_index((size_t)n_reads, -1);
    std::vector<std::vector<SamPos>> multi_of;
    auto add = [&](int64_t r, uint64_t pos, int strand) {
        if (multi_index[r] < 0) { multi_index[r] = (int64_t)multi_of.size(); multi_of.emplace_back(); }
        multi_of[multi_index[r]].push_back(SamPos{pos, strand});
    };
    if (ctx->collect_hits) {
        for (int64_t r = 0; r < n_reads; ++r) {
            const gmx_read_result &res = results[r];
            if (res.status != GMX_READ_MAPPED || res.best_n_positions <= 1) continue;
            for (int32_t h = res.hit_begin; h < res.hit_end; ++h)
                if (ctx->h_hits[h].group == res.best_group) add(r, ctx->h_hits[h].pos, ctx->h_hits[h].strand);
        }
    } else {
        for (const MultiPos &m : ctx->h_multi) if (m.read >= 0 && m.read < n_reads) add(m.read, m.pos, m.strand);
    }
    for (auto &v : multi_of) std::sort(v.begin(), v.end(), [](const SamPos &a, const SamPos &b) { return a.pos != b.pos ? a.pos < b.pos : a.strand < b.strand; });
    // format in parallel slices into pre-sized buffers, then copy them into place in parallel (read order is kept)
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (n_reads < 4096) nt = 1;
    std::vector<std::string> parts(nt);
    auto run = [&](auto &&fn) {
        if (nt == 1) { fn(0u); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(fn, t);
        for (auto &x : th) x.join();
    };
    run([&](unsigned t) {
        const int64_t a = n_reads * t / nt, b = n_reads * (t + 1) / nt;
        size_t est = 0;
        for (int64_t r = a; r < b; ++r)
            if (GMX_READ_PRINTS_SAM(results[r])) est += (size_t)(recs[r].name_len + recs[r].seq_len + recs[r].qual_len + 96) * (size_t)std::max(results[r].best_n_positions, 1);
        parts[t].reserve(est);
        sam_format_range(ctx, text, recs, results, a, b, chrom_names, &multi_of, multi_index, parts[t]);
    });
    std::vector<int64_t> at(nt + 1, 0);
    for (unsigned t = 0; t < nt; ++t) at[t + 1] = at[t] + (int64_t)parts[t].size();
    *len = at[nt];
    if (at[nt] > cap) return GMX_ERR_OVERFLOW;
    run([&](unsigned t) { memcpy(out + at[t], parts[t].data(), parts[t].size()); });
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// next row: .sgr output
// ------------------------------------------------------------------------------------------------
extern "C" int gmx_format_sgr(gmx_ctx *ctx, const char *const *chrom_names, double min_print, char *out, int64_t cap, int64_t *len)
{
    if (!ctx || !chrom_names || !len || (cap > 0 && !out)) return GMX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const uint64_t nb = ctx->acc.n_amount;
    if (nb >= 0x7fffffffull) { ctx->err = "more than 2^31 accumulator bins"; return GMX_ERR_UNSUPPORTED; }
    // printable bins, in order, selected on the device
    ScratchBuf d_idx, d_val, d_cnt, d_tmp;
    CK(d_idx.ensure((size_t)nb * 4 + 16)); CK(d_cnt.ensure(16));
    thrust::counting_iterator<uint32_t> it(0);
    SgrRowSelect pred{ctx->acc.amount, min_print};            // float > double literal, compared in double as the reference does
    size_t tmp_bytes = 0;
    CK(cub::DeviceSelect::If(nullptr, tmp_bytes, it, d_idx.as<uint32_t>(), d_cnt.as<uint32_t>(), (int)nb, pred, ctx->stream));
    CK(d_tmp.ensure(tmp_bytes));
    CK(cub::DeviceSelect::If(d_tmp.p, tmp_bytes, it, d_idx.as<uint32_t>(), d_cnt.as<uint32_t>(), (int)nb, pred, ctx->stream));
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, d_cnt.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> idx; std::vector<float> val;
    if (n) {
        CK(d_val.ensure((size_t)n * 4));
        k_gather_f32<<<nblk(n, 256), 256, 0, ctx->stream>>>(ctx->acc.amount, d_idx.as<uint32_t>(), n, d_val.as<float>());
        CK(cudaGetLastError());
        // the lines themselves on the device: size, place, write (gmp_out.cuh); only the finished text crosses to the host
        const int n_seqs = ctx->ix.n_seqs;
        std::vector<int32_t> noff((size_t)n_seqs), nlen((size_t)n_seqs);
        std::string chars;
        for (int i = 0; i < n_seqs; ++i) { noff[i] = (int32_t)chars.size(); nlen[i] = (int32_t)strlen(chrom_names[i]); chars += chrom_names[i]; }
        const size_t chars_pad = (chars.size() + 15) & ~(size_t)15;
        ScratchBuf d_names, d_lens, d_offs, d_text;
        CK(d_names.ensure(chars_pad + (size_t)n_seqs * 8 + 16)); CK(d_lens.ensure(((size_t)n + 1) * 8 + 16)); CK(d_offs.ensure(((size_t)n + 1) * 8));
        CK(cudaMemcpyAsync(d_names.p, chars.data(), chars.size(), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_names.as<char>() + chars_pad, noff.data(), (size_t)n_seqs * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_names.as<char>() + chars_pad + (size_t)n_seqs * 4, nlen.data(), (size_t)n_seqs * 4, cudaMemcpyHostToDevice, ctx->stream));
        SgrNames names;
        names.chars = d_names.as<char>(); names.off = reinterpret_cast<const int32_t *>(d_names.as<char>() + chars_pad); names.len = names.off + n_seqs;
        long long *dl = d_lens.as<long long>(), *dof = d_offs.as<long long>();
        uint32_t *d_unc = reinterpret_cast<uint32_t *>(dl + n + 1);
        CK(cudaMemsetAsync(dl + n, 0, 16, ctx->stream));
        k_sgr_measure<<<nblk(n, 256), 256, 0, ctx->stream>>>(d_idx.as<uint32_t>(), d_val.as<float>(), n, ctx->params.gen_size, ctx->ix.seq_offset, n_seqs, names, dl, d_unc);
        CK(cudaGetLastError());
        size_t scan_bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, dl, dof, (int)n + 1, ctx->stream));
        CK(d_tmp.ensure(scan_bytes));
        CK(cub::DeviceScan::ExclusiveSum(d_tmp.p, scan_bytes, dl, dof, (int)n + 1, ctx->stream));
        long long total = 0; uint32_t unc = 0;
        CK(cudaMemcpyAsync(&total, dof + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(&unc, d_unc, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (!unc) {
            *len = total;
            if (total > cap) return GMX_ERR_OVERFLOW;
            CK(d_text.ensure((size_t)total + 16));
            k_sgr_write<<<nblk(n, 256), 256, 0, ctx->stream>>>(d_idx.as<uint32_t>(), d_val.as<float>(), n, ctx->params.gen_size, ctx->ix.seq_offset, n_seqs, names, dof, d_text.as<char>());
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(out, d_text.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            return GMX_OK;
        }
        // a value outside the fixed-point writer's range: the host formats the file (snprintf)
        idx.resize(n); val.resize(n);
        CK(cudaMemcpyAsync(idx.data(), d_idx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(val.data(), d_val.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    d_idx.release(); d_val.release(); d_cnt.release(); d_tmp.release();
    // the reference's loop counter walks the genome in steps of gen_size from 0 and prints bin count / gen_size under the
    // sequence that contains `count`
    const std::vector<int64_t> &off = ctx->h_seq_offset;
    const uint64_t gs = ctx->params.gen_size;
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (n < 65536) nt = 1;
    std::vector<std::string> parts(nt);
    auto run = [&](auto &&fn) {
        if (nt == 1) { fn(0u); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(fn, t);
        for (auto &x : th) x.join();
    };
    run([&](unsigned t) {
        const size_t a = (size_t)n * t / nt, b = (size_t)n * (t + 1) / nt;
        std::string &o = parts[t];
        o.reserve((b - a) * 32);
        char num[96];
        for (size_t k = a; k < b; ++k) {
            const int64_t count = (int64_t)((uint64_t)idx[k] * gs);
            if (count >= off.back()) continue;
            const size_t rid = std::upper_bound(off.begin(), off.end() - 1, count) - off.begin() - 1;
            o += chrom_names[rid];
            char *e = num; *e++ = '\t';
            e = gmx_put_int(e, (long long)(count - off[rid] + 1)); *e++ = '\t';
            e = gmx_put_fixed(e, val[k], 5); *e++ = '\n';
            o.append(num, (size_t)(e - num));
        }
    });
    std::vector<int64_t> at(nt + 1, 0);
    for (unsigned t = 0; t < nt; ++t) at[t + 1] = at[t] + (int64_t)parts[t].size();
    *len = at[nt];
    if (at[nt] > cap) return GMX_ERR_OVERFLOW;
    run([&](unsigned t) { memcpy(out + at[t], parts[t].data(), parts[t].size()); });
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// next row: .gmp output with the SNP call (SNP / bisulfite / A->G part)
// ------------------------------------------------------------------------------------------------
extern "C" int gmx_format_gmp(gmx_ctx *ctx, const char *const *chrom_names, int target_base, double min_print, float snp_pval, int snp_monoploid,
                              char *out, int64_t cap, int64_t *len)
{
    if (!ctx || !chrom_names || !len || (cap > 0 && !out)) return GMX_ERR_INVALID;
    const bool snp = ctx->params.mode == GMX_MODE_SNP;
    if (ctx->params.mode == GMX_MODE_NORMAL || !ctx->acc.planes[0]) { ctx->err = "the .gmp file exists in SNP / bisulfite / A->G mode only (gmx_format_sgr prints Normal mode)"; return GMX_ERR_STATE; }
    if (!snp && (target_base < 0 || target_base > 3)) { ctx->err = "bisulfite / A->G rows need the genome base to report (0..3 = a,c,g,t)"; return GMX_ERR_INVALID; }
    CK(cudaSetDevice(ctx->device));
    const uint64_t nb = ctx->acc.n_amount;
    if (nb >= 0x7fffffffull) { ctx->err = "more than 2^31 accumulator bins"; return GMX_ERR_UNSUPPORTED; }
    const uint64_t gs = ctx->params.gen_size;
    const uint64_t l_pac = (uint64_t)ctx->h_seq_offset.back();
    // printable rows, in genome order, selected and gathered on the device
    ScratchBuf d_idx, d_rows, d_base, d_cnt, d_tmp;
    CK(d_idx.ensure((size_t)nb * 4 + 16)); CK(d_cnt.ensure(16));
    thrust::counting_iterator<uint32_t> it(0);
    GmpRowSelect pred{ctx->acc.amount, ctx->ix.pac, gs, l_pac, min_print, snp ? -1 : target_base};
    size_t tmp_bytes = 0;
    CK(cub::DeviceSelect::If(nullptr, tmp_bytes, it, d_idx.as<uint32_t>(), d_cnt.as<uint32_t>(), (int)nb, pred, ctx->stream));
    CK(d_tmp.ensure(tmp_bytes));
    CK(cub::DeviceSelect::If(d_tmp.p, tmp_bytes, it, d_idx.as<uint32_t>(), d_cnt.as<uint32_t>(), (int)nb, pred, ctx->stream));
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, d_cnt.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> idx(n); std::vector<float> rows((size_t)n * 6); std::vector<uint8_t> base(n);
    if (n) {
        CK(d_rows.ensure((size_t)n * 24)); CK(d_base.ensure(n));
        k_gmp_gather<<<nblk(n, 256), 256, 0, ctx->stream>>>(ctx->acc.amount, ctx->acc.planes[0], ctx->acc.planes[1], ctx->acc.planes[2], ctx->acc.planes[3],
                                                            ctx->acc.planes[4], ctx->ix.pac, gs, l_pac, d_idx.as<uint32_t>(), n, d_rows.as<float>(), d_base.as<uint8_t>());
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(idx.data(), d_idx.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(rows.data(), d_rows.p, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(base.data(), d_base.p, n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    d_idx.release(); d_rows.release(); d_base.release(); d_cnt.release(); d_tmp.release();
    const std::vector<int64_t> &off = ctx->h_seq_offset;
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (n < 16384) nt = 1;
    std::vector<std::string> parts(nt);
    auto run = [&](auto &&fn) {
        if (nt == 1) { fn(0u); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(fn, t);
        for (auto &x : th) x.join();
    };
    run([&](unsigned t) {
        const size_t a = (size_t)n * t / nt, b = (size_t)n * (t + 1) / nt;
        std::string &o = parts[t];
        o.reserve((b - a) * 80);
        char line[256];
        for (size_t k = a; k < b; ++k) {
            const int64_t count = (int64_t)((uint64_t)idx[k] * gs);
            if (count >= off.back()) continue;
            const size_t rid = std::upper_bound(off.begin(), off.end() - 1, count) - off.begin() - 1;
            o += chrom_names[rid];
            char *e = line; *e++ = '\t';
            e = gmx_put_int(e, (long long)(count - off[rid] + 1)); *e++ = '\t';
            e = gmx_put_fixed(e, rows[k], snp ? 5 : 6);                        // "%.5f" in PrintFinalSNP, "%f" in PrintFinalBisulfite
            float counts[5];
            for (int c = 0; c < 5; ++c) { counts[c] = rows[(size_t)(c + 1) * n + k]; *e++ = '\t'; e = gmx_put_fixed(e, counts[c], 5); }
            if (snp) e = gmx_put_snp_call(e, counts, base[k], snp_monoploid != 0, snp_pval);
            *e++ = '\n';
            o.append(line, (size_t)(e - line));
        }
    });
    std::vector<int64_t> at(nt + 1, 0);
    for (unsigned t = 0; t < nt; ++t) at[t + 1] = at[t] + (int64_t)parts[t].size();
    *len = at[nt];
    if (at[nt] > cap) return GMX_ERR_OVERFLOW;
    run([&](unsigned t) { memcpy(out + at[t], parts[t].data(), parts[t].size()); });
    return GMX_OK;
}

// GenomeBwt::is_snp + the call column of PrintSNPCall for one position; host arithmetic only (no context, no device)
extern "C" int gmx_snp_call(const float counts[5], int genome_base, int snp_monoploid, float snp_pval, int *first, int *second, int *diploid,
                            double *pval, char *text, int text_cap)
{
    if (!counts || genome_base < 0 || genome_base > 4) return GMX_ERR_INVALID;
    const GmpCall c = gmx_snp_call(counts, snp_monoploid != 0);
    if (first) *first = c.first;
    if (second) *second = c.second;
    if (diploid) *diploid = c.diploid ? 1 : 0;
    if (pval) *pval = c.pval;
    if (text) {
        char buf[96];
        char *e = gmx_put_snp_call(buf, counts, genome_base, snp_monoploid != 0, snp_pval);
        const int n = (int)(e - buf);
        if (n + 1 > text_cap) return GMX_ERR_OVERFLOW;
        memcpy(text, buf, (size_t)n); text[n] = 0;
    }
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// next row: index construction
// ------------------------------------------------------------------------------------------------
#include "index_build.cuh"

extern "C" int gmx_index_sizes(int64_t l_pac, uint64_t *bwt_words, uint64_t *n_sa, uint64_t *pac_bytes)
{
    if (l_pac < 1) return GMX_ERR_INVALID;
    const uint64_t n = (uint64_t)l_pac, n_blocks = (n + GMX_IDX_OCC - 1) / GMX_IDX_OCC;
    if (bwt_words) *bwt_words = ((n + 15) >> 4) + (n_blocks + 1) * 8;
    if (n_sa) *n_sa = (n + GMX_IDX_SA_INTV) / GMX_IDX_SA_INTV;
    if (pac_bytes) *pac_bytes = (n + 3) / 4;
    return GMX_OK;
}

#define ICK(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (e__ == cudaErrorMemoryAllocation) ? GMX_ERR_NOMEM : GMX_ERR_CUDA;                \
        }                                                                                            \
    } while (0)

extern "C" int gmx_index_build(const uint8_t *codes, int64_t l_pac, int device, uint32_t *bwt, uint64_t *primary_out, uint64_t L2[5],
                               uint64_t *sa, uint8_t *pac, int32_t *rounds, char *err, int err_cap)
{
    if (!codes || l_pac < 1 || !bwt || !primary_out || !L2 || !sa || !pac) return GMX_ERR_INVALID;
    if (l_pac >= 0x7fffff00ll) { if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "genomes of 2^31 bases or more are not supported by the builder"); return GMX_ERR_UNSUPPORTED; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return GMX_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return GMX_ERR_INVALID;
    ICK(cudaSetDevice(device));
    const int64_t n = l_pac;
    cudaStream_t st = nullptr;
    ICK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct Guard { cudaStream_t s; ~Guard() { if (s) cudaStreamDestroy(s); } } guard{st};
    ScratchBuf d_codes, d_key[2], d_idx[2], d_rank, d_flag, d_tmp;
    ICK(d_codes.ensure((size_t)n + 32)); ICK(d_key[0].ensure((size_t)n * 8)); ICK(d_key[1].ensure((size_t)n * 8));
    ICK(d_idx[0].ensure((size_t)n * 4)); ICK(d_idx[1].ensure((size_t)n * 4)); ICK(d_rank.ensure((size_t)n * 4)); ICK(d_flag.ensure((size_t)n * 4));
    ICK(cudaMemcpyAsync(d_codes.p, codes, (size_t)n, cudaMemcpyHostToDevice, st));
    cub::DoubleBuffer<unsigned long long> keys(d_key[0].as<unsigned long long>(), d_key[1].as<unsigned long long>());
    cub::DoubleBuffer<uint32_t> vals(d_idx[0].as<uint32_t>(), d_idx[1].as<uint32_t>());
    size_t sort_bytes = 0, scan_bytes = 0;
    ICK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys, vals, (int)n, 0, 64, st));
    ICK(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, d_flag.as<uint32_t>(), d_flag.as<uint32_t>(), (int)n, st));
    ICK(d_tmp.ensure(std::max(sort_bytes, scan_bytes)));
    int rank_bits = 1; while ((1ll << rank_bits) <= n) rank_bits++;
    int n_rounds = 0;
    int64_t h = 16;
    uint32_t n_ranks = 0;
    k_idx_key16<<<nblk(n, 256), 256, 0, st>>>(d_codes.as<uint8_t>(), n, keys.Current(), vals.Current());
    ICK(cudaGetLastError());
    while (true) {
        size_t tb = d_tmp.cap;
        ICK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tb, keys, vals, (int)n, 0, n_rounds == 0 ? 48 : 32 + rank_bits, st));
        k_idx_flags<<<nblk(n, 256), 256, 0, st>>>(keys.Current(), n, d_flag.as<uint32_t>());
        ICK(cudaGetLastError());
        tb = d_tmp.cap;
        ICK(cub::DeviceScan::InclusiveSum(d_tmp.p, tb, d_flag.as<uint32_t>(), d_flag.as<uint32_t>(), (int)n, st));
        k_idx_scatter_rank<<<nblk(n, 256), 256, 0, st>>>(vals.Current(), d_flag.as<uint32_t>(), n, d_rank.as<uint32_t>());
        ICK(cudaGetLastError());
        ICK(cudaMemcpyAsync(&n_ranks, d_flag.as<uint32_t>() + (n - 1), 4, cudaMemcpyDeviceToHost, st));
        ICK(cudaStreamSynchronize(st));
        n_rounds++;
        if ((int64_t)n_ranks >= n) break;
        if (h >= n) { if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "suffix sort did not converge"); return GMX_ERR_CUDA; }
        k_idx_pair_key<<<nblk(n, 256), 256, 0, st>>>(d_rank.as<uint32_t>(), n, h, keys.Current(), vals.Current());
        ICK(cudaGetLastError());
        h *= 2;
    }
    if (rounds) *rounds = n_rounds;
    const uint32_t *d_sa = vals.Current();                       // suffix array of the n suffixes (augmented rank R >= 1 <-> d_sa[R - 1])
    uint32_t rank0 = 0;
    ICK(cudaMemcpyAsync(&rank0, d_rank.p, 4, cudaMemcpyDeviceToHost, st));
    ICK(cudaStreamSynchronize(st));
    const int64_t primary = (int64_t)rank0;                       // rank of suffix 0 among the '$'-augmented suffixes
    const int64_t n_blocks = (n + GMX_IDX_OCC - 1) / GMX_IDX_OCC, n_words = (n + 15) >> 4;
    uint64_t bwt_words = 0, n_sa = 0, pac_bytes = 0;
    gmx_index_sizes(n, &bwt_words, &n_sa, &pac_bytes);
    // counts per block -> exclusive prefix sums (the key buffers are free now: reuse them)
    d_key[0].release(); d_key[1].release();
    ScratchBuf d_c, d_out, d_sas, d_pac;
    ICK(d_c.ensure((size_t)(n_blocks + 1) * 8 * 4)); ICK(d_out.ensure((size_t)bwt_words * 4)); ICK(d_sas.ensure((size_t)n_sa * 8)); ICK(d_pac.ensure((size_t)pac_bytes));
    unsigned long long *c[4];
    for (int k = 0; k < 4; ++k) c[k] = d_c.as<unsigned long long>() + (size_t)k * (n_blocks + 1);
    ICK(cudaMemsetAsync(d_c.p, 0, (size_t)(n_blocks + 1) * 8 * 4, st));
    ICK(cudaMemsetAsync(d_out.p, 0, (size_t)bwt_words * 4, st));
    k_idx_block_counts<<<nblk(n_blocks, 4), 128, 0, st>>>(d_codes.as<uint8_t>(), d_sa, n, primary, n_blocks, c[0], c[1], c[2], c[3]);
    ICK(cudaGetLastError());
    size_t sb = 0;
    ICK(cub::DeviceScan::ExclusiveSum(nullptr, sb, c[0], c[0], (int)(n_blocks + 1), st));
    ICK(d_tmp.ensure(sb));
    for (int k = 0; k < 4; ++k) { size_t tb = d_tmp.cap; ICK(cub::DeviceScan::ExclusiveSum(d_tmp.p, tb, c[k], c[k], (int)(n_blocks + 1), st)); }
    k_idx_interleave<<<nblk(n_words + 1, 256), 256, 0, st>>>(d_codes.as<uint8_t>(), d_sa, n, primary, n_blocks, c[0], c[1], c[2], c[3], d_out.as<uint32_t>());
    ICK(cudaGetLastError());
    k_idx_sample_sa<<<nblk((int64_t)n_sa, 256), 256, 0, st>>>(d_sa, (int64_t)n_sa, d_sas.as<unsigned long long>());
    ICK(cudaGetLastError());
    k_idx_pack_pac<<<nblk((int64_t)pac_bytes, 256), 256, 0, st>>>(d_codes.as<uint8_t>(), n, d_pac.as<uint8_t>(), (int64_t)pac_bytes);
    ICK(cudaGetLastError());
    unsigned long long tot[4];
    for (int k = 0; k < 4; ++k) ICK(cudaMemcpyAsync(&tot[k], c[k] + n_blocks, 8, cudaMemcpyDeviceToHost, st));
    ICK(cudaMemcpyAsync(bwt, d_out.p, (size_t)bwt_words * 4, cudaMemcpyDeviceToHost, st));
    ICK(cudaMemcpyAsync(sa, d_sas.p, (size_t)n_sa * 8, cudaMemcpyDeviceToHost, st));
    ICK(cudaMemcpyAsync(pac, d_pac.p, (size_t)pac_bytes, cudaMemcpyDeviceToHost, st));
    ICK(cudaStreamSynchronize(st));
    *primary_out = (uint64_t)primary;
    L2[0] = 0; L2[1] = tot[0]; L2[2] = L2[1] + tot[1]; L2[3] = L2[2] + tot[2]; L2[4] = L2[3] + tot[3];
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// measured ALU ceilings for the roofline of the NW kernels ("measure with a mul/add micro-kernel")
// ------------------------------------------------------------------------------------------------
// K2a / K2b are bit-exact FP32 with separate multiplies and adds (no FMA); K2c runs on the FP64 pipe.  Eight independent
// chains per thread keep the pipes full; the result is stored so that nothing is optimised away.
template <int KIND>
__global__ void __launch_bounds__(256) k_alu_peak(float *out, int iters, float x, float y)
{
    float a[8]; double d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = (float)(threadIdx.x + k) * 1e-3f; d[k] = (double)a[k]; }
    const double dx = (double)x, dy = (double)y;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (KIND == 0) a[k] = __fadd_rn(__fmul_rn(a[k], x), y);              // FMUL + FADD: 2 flops, 2 instructions
            else if (KIND == 1) a[k] = fmaxf(__fadd_rn(a[k], x), y);             // FADD + FMNMX: the max-plus step of the NW cell
            else d[k] = __dadd_rn(__dmul_rn(d[k], dx), dy);                       // DMUL + DADD
        }
    }
    float s = 0; double sd = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += a[k]; sd += d[k]; }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + (float)sd;
}

// kind 0: FP32 multiply + add without FMA (what K2a / K2b issue); 1: FP32 add + max; 2: FP64 multiply + add without FMA.
// *tera_ops = 1e-12 x arithmetic instructions x 32 lanes / second (one flop per lane per instruction), best of five.
extern "C" int gmx_measure_alu_peak(gmx_ctx *ctx, int kind, double *tera_ops)
{
    if (!ctx || !tera_ops || kind < 0 || kind > 2) return GMX_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const int blocks = ctx->n_sm * 16, iters = kind == 2 ? 1 << 12 : 1 << 14;
    ScratchBuf out;
    CK(out.ensure((size_t)blocks * 256 * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0, ctx->stream));
        if (kind == 0) k_alu_peak<0><<<blocks, 256, 0, ctx->stream>>>(out.as<float>(), iters, 0.999f, 0.001f);
        else if (kind == 1) k_alu_peak<1><<<blocks, 256, 0, ctx->stream>>>(out.as<float>(), iters, 0.999f, 0.001f);
        else k_alu_peak<2><<<blocks, 256, 0, ctx->stream>>>(out.as<float>(), iters, 0.999f, 0.001f);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tera_ops = (double)blocks * 256.0 * (double)iters * 8.0 * 2.0 / ((double)best * 1e-3) * 1e-12;
    return GMX_OK;
}

// ------------------------------------------------------------------------------------------------
// several GPUs in one process: accumulator reduce
// ------------------------------------------------------------------------------------------------
#include "comm.cuh"

#define CCK(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char b__[512];                                                                           \
            snprintf(b__, sizeof(b__), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            comm->err = b__; comm->ctx[0]->err = b__;                                                \
            return GMX_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

extern "C" int gmx_comm_create(gmx_comm **out, gmx_ctx *const *ctxs, int n, int backend)
{
    if (!out || !ctxs || n < 1 || n > GMX_COMM_MAX || backend < GMX_COMM_AUTO || backend > GMX_COMM_NCCL) return GMX_ERR_INVALID;
    *out = nullptr;
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i] || ctxs[i]->comm) return GMX_ERR_INVALID;
        for (int j = 0; j < i; ++j) if (ctxs[j] == ctxs[i]) return GMX_ERR_INVALID;
        // every context must accumulate the same arrays
        if (ctxs[i]->acc.n_amount != ctxs[0]->acc.n_amount || ctxs[i]->n_plane != ctxs[0]->n_plane || ctxs[i]->params.mode != ctxs[0]->params.mode) {
            ctxs[0]->err = "gmx_comm_create: the contexts do not share one accumulator layout (genome, mode, gen_size)";
            return GMX_ERR_INVALID;
        }
    }
    bool distinct = true;
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) if (ctxs[i]->device == ctxs[j]->device) distinct = false;
    gmx_comm *comm = new gmx_comm();
    comm->n = n;
    for (int i = 0; i < n; ++i) comm->ctx[i] = ctxs[i];
    if (backend == GMX_COMM_NCCL && (!distinct || n < 2)) { ctxs[0]->err = "GMX_COMM_NCCL needs two or more contexts on distinct devices"; delete comm; return GMX_ERR_INVALID; }
    if (backend == GMX_COMM_AUTO) backend = GMX_COMM_PEER;
    if (backend == GMX_COMM_NCCL) {
        if (!comm->nccl.load()) { ctxs[0]->err = "libnccl.so.2 could not be opened"; delete comm; return GMX_ERR_UNSUPPORTED; }
        int devs[GMX_COMM_MAX];
        for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
        ncclResult_t r = comm->nccl.CommInitAll(comm->comms, n, devs);
        if (r != ncclSuccess) { ctxs[0]->err = std::string("ncclCommInitAll: ") + comm->nccl.GetErrorString(r); delete comm; return GMX_ERR_CUDA; }
        comm->have_comms = true;
    } else {
        // peer access between every pair of distinct devices (NVLink / NVSwitch on a B200 box)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                const int a = ctxs[i]->device, b = ctxs[j]->device;
                if (a == b) continue;
                int can = 0;
                CCK(cudaDeviceCanAccessPeer(&can, a, b));
                if (!can) { ctxs[0]->err = "GMX_COMM_PEER: the devices cannot access each other's memory (use GMX_COMM_NCCL)"; delete comm; return GMX_ERR_UNSUPPORTED; }
                CCK(cudaSetDevice(a));
                cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { CCK(e); }
                cudaGetLastError();
            }
    }
    comm->backend = backend;
    CCK(cudaSetDevice(ctxs[0]->device));
    CCK(cudaEventCreate(&comm->ev[0])); CCK(cudaEventCreate(&comm->ev[1]));
    for (int i = 0; i < n; ++i) ctxs[i]->comm = comm;
    *out = comm;
    return GMX_OK;
}

extern "C" int gmx_comm_reduce(gmx_comm *comm, int all)
{
    if (!comm) return GMX_ERR_INVALID;
    const int n = comm->n;
    for (int i = 0; i < n; ++i) if (!comm->ctx[i]) return GMX_ERR_STATE;
    gmx_ctx *root = comm->ctx[0];
    // every context's kernels must have finished scattering
    for (int i = 0; i < n; ++i) { CCK(cudaSetDevice(comm->ctx[i]->device)); CCK(cudaStreamSynchronize(comm->ctx[i]->stream)); }
    const uint64_t n_amount = root->acc.n_amount, n_planes = root->acc.planes[0] ? 5ull * root->n_plane : 0ull;
    CCK(cudaSetDevice(root->device));
    CCK(cudaEventRecord(comm->ev[0], root->stream));
    if (n > 1 && comm->backend == GMX_COMM_NCCL) {
        ncclResult_t r = comm->nccl.GroupStart();
        for (int i = 0; i < n && r == ncclSuccess; ++i) {
            gmx_ctx *c = comm->ctx[i];
            CCK(cudaSetDevice(c->device));
            // amount_genome is all-reduced in the reference (src/Driver.cpp:1672), the planes reduced to rank 0 (:1719-1767)
            r = all ? comm->nccl.AllReduce(c->acc.amount, c->acc.amount, n_amount, ncclFloat, ncclSum, comm->comms[i], c->stream)
                    : comm->nccl.Reduce(c->acc.amount, c->acc.amount, n_amount, ncclFloat, ncclSum, 0, comm->comms[i], c->stream);
            if (r == ncclSuccess && n_planes)
                r = all ? comm->nccl.AllReduce(c->acc.planes[0], c->acc.planes[0], n_planes, ncclFloat, ncclSum, comm->comms[i], c->stream)
                        : comm->nccl.Reduce(c->acc.planes[0], c->acc.planes[0], n_planes, ncclFloat, ncclSum, 0, comm->comms[i], c->stream);
        }
        ncclResult_t r2 = comm->nccl.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess) { comm->err = std::string("nccl reduce: ") + comm->nccl.GetErrorString(r != ncclSuccess ? r : r2); root->err = comm->err; return GMX_ERR_CUDA; }
    } else if (n > 1) {
        // GPU g owns slice g of both arrays
        for (int pass = 0; pass < 2; ++pass) {
            const uint64_t total = pass == 0 ? n_amount : n_planes;
            if (!total) continue;
            PeerBufs B; B.n = n;
            for (int i = 0; i < n; ++i) B.buf[i] = pass == 0 ? comm->ctx[i]->acc.amount : comm->ctx[i]->acc.planes[0];
            for (int g = 0; g < n; ++g) {
                gmx_ctx *c = comm->ctx[g];
                uint64_t lo = total * (uint64_t)g / (uint64_t)n, hi = total * (uint64_t)(g + 1) / (uint64_t)n;
                lo &= ~3ull; if (g + 1 < n) hi &= ~3ull;                        // slice borders on float4 boundaries
                if (hi <= lo) continue;
                CCK(cudaSetDevice(c->device));
                const unsigned blocks = (unsigned)std::min<uint64_t>((uint64_t)c->n_sm * 8, ((hi - lo) / 4 + 255) / 256 + 1);
                k_reduce_slice<<<blocks, 256, 0, c->stream>>>(B, lo, hi, all ? 1 : 0);
                CCK(cudaGetLastError());
            }
        }
    }
    // the root's stream waits for every slice owner
    for (int i = 1; i < n; ++i) { CCK(cudaSetDevice(comm->ctx[i]->device)); CCK(cudaStreamSynchronize(comm->ctx[i]->stream)); }
    CCK(cudaSetDevice(root->device));
    CCK(cudaEventRecord(comm->ev[1], root->stream));
    CCK(cudaStreamSynchronize(root->stream));
    float ms = 0; cudaEventElapsedTime(&ms, comm->ev[0], comm->ev[1]);
    comm->last_ms = ms; comm->last_bytes = 4ull * (n_amount + n_planes);
    if (!all)      // the sum now lives in the root: the other terms start again from zero, so a later reduce adds only what is new
        for (int i = 1; i < n; ++i) { int r = gmx_reset_accumulators(comm->ctx[i]); if (r != GMX_OK) return r; CCK(cudaStreamSynchronize(comm->ctx[i]->stream)); }
    return GMX_OK;
}

static void comm_detach(gmx_ctx *ctx)
{   // a context destroyed before its communicator: the communicator is unusable from then on
    gmx_comm *comm = ctx->comm;
    for (int i = 0; i < comm->n; ++i) if (comm->ctx[i] == ctx) comm->ctx[i] = nullptr;
    ctx->comm = nullptr;
}

static int comm_reduce_for_finish(gmx_ctx *ctx)
{
    if (ctx->comm->ctx[0] != ctx) { ctx->err = "gmx_finish: only the first context of a communicator holds the sum"; return GMX_ERR_STATE; }
    return gmx_comm_reduce(ctx->comm, 0);
}

extern "C" int gmx_comm_stats(const gmx_comm *comm, float *ms, uint64_t *bytes, int *backend)
{
    if (!comm) return GMX_ERR_INVALID;
    if (ms) *ms = comm->last_ms;
    if (bytes) *bytes = comm->last_bytes;
    if (backend) *backend = comm->backend;
    return GMX_OK;
}

extern "C" void gmx_comm_destroy(gmx_comm *comm)
{
    if (!comm) return;
    for (int i = 0; i < comm->n; ++i) if (comm->ctx[i]) comm->ctx[i]->comm = nullptr;
    if (comm->have_comms) for (int i = 0; i < comm->n; ++i) comm->nccl.CommDestroy(comm->comms[i]);
    if (comm->ev[0]) { if (comm->ctx[0]) cudaSetDevice(comm->ctx[0]->device); cudaEventDestroy(comm->ev[0]); cudaEventDestroy(comm->ev[1]); }
    delete comm;
}
