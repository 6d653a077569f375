// nw.cuh -- banded PWM Needleman-Wunsch (score, traceback) and the pair-HMM forward/backward
// on the device.
//
//   K2a  bin_seq::get_align_score(read, gen)          reference src/bin_seq.cpp:761-850
//   K2b  bin_seq::get_align_score_w_traceback          reference src/bin_seq.cpp:445-718
//   K2c  bin_seq::pairHMM                              reference src/bin_seq.cpp:60-244
//
// FP32 arithmetic follows the reference operation by operation (separate multiply / add, left to
// right; the file is compiled with -fmad=false), so scores are bit-identical to the CPU and the
// acceptance test `score >= min_align_score` (reference inc/align_seq2_raw.cpp:102) cannot flip.
// The max-plus recurrence has no contraction to feed a tensor core: this is ALU work.
#pragma once

#include "gmx_common.cuh"

// ---- operand sources -------------------------------------------------------------------------

// One (read, strand) as the alignment kernels see it: row i of the strand-oriented PWM.
struct ReadView {
    const uint8_t *seq;    // raw ASCII of the read as given (forward)
    const uint8_t *qual;
    const float   *pwm;    // optional raw PWM rows of the read as given
    int n;
    int neg;               // 1: reverse-complement orientation (reference SequenceOperations.h:149-161)

    __device__ __forceinline__ int src(int i) const { return neg ? n - 1 - i : i; }

    // get_val(pwm[i], g) for g = a,c,g,t   (reference src/bin_seq.cpp:975-987)
    __device__ __forceinline__ float4 sub_row(const DevTables &T, int i) const
    {
        int ii = src(i);
        if (pwm) {
            float4 p = reinterpret_cast<const float4 *>(pwm)[ii];
            if (neg) p = make_float4(p.w, p.z, p.y, p.x);
            float r[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float *s = T.S + 4 * (int)("acgt"[g]);
                r[g] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, s[0]), __fmul_rn(p.y, s[1])), __fmul_rn(p.z, s[2])), __fmul_rn(p.w, s[3]));
            }
            return make_float4(r[0], r[1], r[2], r[3]);
        }
        int code = gmx_nt4(seq[ii]);
        int q = gmx_qidx(qual[ii], 0);
        const float4 *lut = reinterpret_cast<const float4 *>(neg ? T.sub_neg : T.sub_pos);
        return __ldg(lut + code * GMX_NQ + q);
    }

    // get_val(pwm[i], ch) for a window character that is not a/c/g/t.  Every such row of
    // gALIGN_SCORES is identical (reference inc/a_matrices.c:65-67); the row of 'n' stands for all.
    __device__ __forceinline__ float sub_other(const DevTables &T, int i) const
    {
        float4 p = pwm_row(T, i);
        const float *s = T.S + 4 * (int)'n';
        return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, s[0]), __fmul_rn(p.y, s[1])), __fmul_rn(p.z, s[2])), __fmul_rn(p.w, s[3]));
    }

    // the strand-oriented PWM row itself
    __device__ __forceinline__ float4 pwm_row(const DevTables &T, int i) const
    {
        int ii = src(i);
        float4 p;
        if (pwm) p = reinterpret_cast<const float4 *>(pwm)[ii];
        else p = __ldg(reinterpret_cast<const float4 *>(T.pwm_lut) + gmx_nt4(seq[ii]) * GMX_NQ + gmx_qidx(qual[ii], 0));
        if (neg) p = make_float4(p.w, p.z, p.y, p.x);
        return p;
    }

    // 3 * sum_b pwm[i][b] * P[g][b]  (reference src/bin_seq.cpp:41-57 p_seq) for g = a,c,g,t
    __device__ __forceinline__ float4 phmm_row(const DevTables &T, int i) const
    {
        int ii = src(i);
        if (pwm) {
            float4 p = reinterpret_cast<const float4 *>(pwm)[ii];
            if (neg) p = make_float4(p.w, p.z, p.y, p.x);
            float r[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float *s = T.P + 4 * (int)("acgt"[g]);
                float sum = 0.f;
                sum = __fadd_rn(sum, __fmul_rn(p.x, s[0]));
                sum = __fadd_rn(sum, __fmul_rn(p.y, s[1]));
                sum = __fadd_rn(sum, __fmul_rn(p.z, s[2]));
                sum = __fadd_rn(sum, __fmul_rn(p.w, s[3]));
                r[g] = __fmul_rn(3.f, sum);
            }
            return make_float4(r[0], r[1], r[2], r[3]);
        }
        const float4 *lut = reinterpret_cast<const float4 *>(neg ? T.phmm_neg : T.phmm_pos);
        return __ldg(lut + gmx_nt4(seq[ii]) * GMX_NQ + gmx_qidx(qual[ii], 0));
    }
};

__device__ __forceinline__ char gmx_max_char(float4 c)
{   // reference inc/ScoredSeq.h:71-103
    if ((c.x == c.y) && (c.x == c.z) && (c.x == c.w)) return 'n';
    if (c.x >= c.y) {
        if (c.x >= c.z) return (c.x >= c.w) ? 'a' : 't';
        return (c.z >= c.w) ? 'g' : 't';
    }
    if (c.y >= c.z) return (c.y >= c.w) ? 'c' : 't';
    return (c.z >= c.w) ? 'g' : 't';
}

// Genome window: either the packed genome at `pos`, or an explicit ASCII string.
struct WindowView {
    const uint8_t *pac;   // non-null: window = genome[pos, pos+n)
    int64_t pos;
    const uint8_t *chars; // else explicit "acgt..." characters
    // base code 0..3, or 4 for a non-acgt character of an explicit window
    __device__ __forceinline__ int base(int j) const
    {
        if (pac) return gmx_pac_base(pac, pos + j);
        return gmx_nt4(chars[j]);
    }
};

__device__ __forceinline__ float gmx_sel4(float4 v, int g, float other)
{
    return g == 0 ? v.x : (g == 1 ? v.y : (g == 2 ? v.z : (g == 3 ? v.w : other)));
}

// ---- K2a: score only -------------------------------------------------------------------------
// The matrix is filled from (n,n) back to (0,0) inside the band |j-i| <= G; only two band rows
// (2G+3 floats each, including the NEG_INF / boundary guard cells at |j-i| = G+1) are live.
template <int G>
__device__ float gmx_nw_band_score(const ReadView &rd, const WindowView &win, const DevTables &T, float gap)
{
    constexpr int W = 2 * G + 3;
    const int n = rd.n;
    float prev[W], cur[W];
    int gb[W];                                    // genome bases gen[i+d], d = -G-1..G+1 (guards unused)
    // transversion-row value for a non-acgt window character of an explicit window: every
    // gALIGN_SCORES row that is not a/c/g/t (reference a_matrices.c:65-67) -- resolved by caller LUT
    // row n:  nm[n][j] = gap * (n - j)  for j >= n-G-1   (reference src/bin_seq.cpp:805-807)
#pragma unroll
    for (int d = -G - 1; d <= G + 1; ++d) prev[d + G + 1] = (d <= 0) ? __fmul_rn(gap, (float)(-d)) : GMX_NEG_INF;
#pragma unroll
    for (int d = -G - 1; d <= G + 1; ++d) { int j = n - 1 + d; gb[d + G + 1] = (j >= 0 && j < n) ? win.base(j) : 0; }

    for (int i = n - 1; i >= 0; --i) {
        float4 sub = rd.sub_row(T, i);
        const float other = win.pac ? 0.f : rd.sub_other(T, i);
        cur[2 * G + 2] = (i + G + 1 == n) ? __fmul_rn(gap, (float)(n - i)) : GMX_NEG_INF;
#pragma unroll
        for (int d = G; d >= -G; --d) {
            int j = i + d;
            float v;
            if (j >= n) v = (j == n) ? __fmul_rn(gap, (float)(n - i)) : GMX_NEG_INF;
            else if (j < 0) v = GMX_NEG_INF;
            else {
                float m_mm = __fadd_rn(prev[d + G + 1], gmx_sel4(sub, gb[d + G + 1], other));
                float gap1 = __fadd_rn(prev[d - 1 + G + 1], gap);
                float gap2 = __fadd_rn(cur[d + 1 + G + 1], gap);
                v = gmx_max3(m_mm, gap1, gap2);
            }
            cur[d + G + 1] = v;
        }
        cur[0] = GMX_NEG_INF;
#pragma unroll
        for (int d = 0; d < W; ++d) prev[d] = cur[d];
        // slide the genome window: gen[(i-1)+d] = gen[i+(d-1)]
#pragma unroll
        for (int d = W - 1; d > 0; --d) gb[d] = gb[d - 1];
        { int j = i - 1 - G - 1; gb[0] = (j >= 0) ? win.base(j) : 0; }
    }
    return prev[G + 1];
}

// Same recurrence for band half-widths without a specialisation (local-memory band rows).
__device__ float gmx_nw_band_score_any(const ReadView &rd, const WindowView &win, const DevTables &T, float gap, int G)
{
    const int n = rd.n;
    float prev[2 * GMX_MAX_GAP + 3], cur[2 * GMX_MAX_GAP + 3];
    const int W = 2 * G + 3;
    for (int d = -G - 1; d <= G + 1; ++d) prev[d + G + 1] = (d <= 0) ? __fmul_rn(gap, (float)(-d)) : GMX_NEG_INF;
    for (int i = n - 1; i >= 0; --i) {
        float4 sub = rd.sub_row(T, i);
        const float other = win.pac ? 0.f : rd.sub_other(T, i);
        cur[2 * G + 2] = (i + G + 1 == n) ? __fmul_rn(gap, (float)(n - i)) : GMX_NEG_INF;
        for (int d = G; d >= -G; --d) {
            int j = i + d;
            float v;
            if (j >= n) v = (j == n) ? __fmul_rn(gap, (float)(n - i)) : GMX_NEG_INF;
            else if (j < 0) v = GMX_NEG_INF;
            else {
                float m_mm = __fadd_rn(prev[d + G + 1], gmx_sel4(sub, win.base(j), other));
                float gap1 = __fadd_rn(prev[d + G], gap);
                float gap2 = __fadd_rn(cur[d + G + 2], gap);
                v = gmx_max3(m_mm, gap1, gap2);
            }
            cur[d + G + 1] = v;
        }
        cur[0] = GMX_NEG_INF;
        for (int d = 0; d < W; ++d) prev[d] = cur[d];
    }
    return prev[G + 1];
}

// ---- K2a fast path: G == 3, FASTQ read (LUT rows), window in the packed genome ---------------------
// Same cell arithmetic (one __fadd_rn for the substitution, one per gap, max of three), restructured so that no
// register array is shifted and no branch sits in the interior rows: the value of column j lives in slot j & 7
// (7 live columns + the one entering), the 8 row phases are unrolled so every slot index is a compile-time
// constant, and the 4-way choice of the substitution value by the genome base is three byte-permutes with
// selectors kept per slot.
// eight bytes starting at any address, from the one or two aligned 64-bit words that hold them
__device__ __forceinline__ unsigned long long gmx_load8_unaligned(const uint8_t *p)
{
    const uintptr_t a = (uintptr_t)p;
    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(a & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(a & 7u) * 8u;
    unsigned long long v = __ldg(w);
    if (sh) v = (v >> sh) | (__ldg(w + 1) << (64u - sh));
    return v;
}

struct NwFastState {
    float S[8];            // nm[i+1][j] of the live columns, slot j & 7
    uint32_t A[8], B[8];   // PRMT selectors of the slot's genome base: bit 0 / bit 1
};

__device__ __forceinline__ float gmx_nw_pick(const float4 &sub, uint32_t selA, uint32_t selB)
{
    const uint32_t lo = __byte_perm(__float_as_uint(sub.x), __float_as_uint(sub.y), selA);
    const uint32_t hi = __byte_perm(__float_as_uint(sub.z), __float_as_uint(sub.w), selA);
    return __uint_as_float(__byte_perm(lo, hi, selB));
}

// one row i with i & 7 == P.  INTERIOR: 3 <= i <= n - 5 (every band column inside [0, n-1], right guard outside)
// g_in: the genome base of the entering column when the caller has it already (interior rows), else -1
template <int P, bool INTERIOR>
__device__ __forceinline__ void gmx_nw_fast_row(NwFastState &st, int i, int n, const float4 &sub, const uint8_t *pac, int64_t pos, float gap, int g_in = -1)
{
    // the column entering the band on the left: j = i - 3, slot (P + 5) & 7
    {
        constexpr int sn = (P + 5) & 7;
        const int jn = i - 3;
        int g = 0;
        if (INTERIOR) g = g_in;
        else if (jn >= 0) g = gmx_pac_base(pac, pos + jn);
        st.A[sn] = (g & 1) ? 0x7654u : 0x3210u;
        st.B[sn] = (g & 2) ? 0x7654u : 0x3210u;
        st.S[sn] = (!INTERIOR && i == n - 1) ? __fmul_rn(gap, 4.f) : GMX_NEG_INF;       // nm[n][n-4] is a border cell
    }
    float right = (!INTERIOR && i + 4 == n) ? __fmul_rn(gap, (float)(n - i)) : GMX_NEG_INF;   // nm[i][i+4]
    float diag = st.S[(P + 4) & 7];                                                      // nm[i+1][i+4]
#pragma unroll
    for (int d = 3; d >= -3; --d) {
        constexpr int dummy = 0; (void)dummy;
        const int s = (P + d + 8) & 7;
        const float up = st.S[s];
        float v = gmx_max3(__fadd_rn(diag, gmx_nw_pick(sub, st.A[s], st.B[s])), __fadd_rn(up, gap), __fadd_rn(right, gap));
        if (!INTERIOR) {
            const int j = i + d;
            if (j >= n) v = (j == n) ? __fmul_rn(gap, (float)(n - i)) : GMX_NEG_INF;
            else if (j < 0) v = GMX_NEG_INF;
        }
        diag = up; st.S[s] = v; right = v;
    }
}

__device__ float gmx_nw_band_score_fast3(const ReadView &rd, const uint8_t *pac, int64_t pos, const DevTables &T, float gap)
{
    const int n = rd.n;
    NwFastState st;
    // row n: nm[n][j] = gap * (n - j) for j <= n, NEG_INF beyond (reference src/bin_seq.cpp:805-807); columns n-3 .. n+3
#pragma unroll
    for (int s = 0; s < 8; ++s) { st.S[s] = GMX_NEG_INF; st.A[s] = 0x3210u; st.B[s] = 0x3210u; }
    for (int j = n - 3; j <= n + 3; ++j) {
        const float v = j <= n ? __fmul_rn(gap, (float)(n - j)) : GMX_NEG_INF;
        const int g = (j >= 0 && j < n) ? gmx_pac_base(pac, pos + j) : 0;
        const uint32_t a = (g & 1) ? 0x7654u : 0x3210u, b = (g & 2) ? 0x7654u : 0x3210u;
#pragma unroll
        for (int s = 0; s < 8; ++s) if (s == (j & 7)) { st.S[s] = v; st.A[s] = a; st.B[s] = b; }
    }
    int i = n - 1;
    while (i >= 0) {
        if ((i & 7) == 7 && i <= n - 5 && i - 7 >= 3) {
            // Eight interior rows.  The lanes of a warp sit on different reads and genome positions, so every byte
            // load costs a wavefront per lane: the 8 sequence bytes, the 8 quality bytes and the 8 entering genome
            // bases of the block come from two aligned 64-bit words each (two 32-bit words for the packed genome).
            const uint8_t *a_seq = rd.seq + (rd.neg ? n - 1 - i : i - 7), *a_qual = rd.qual + (rd.neg ? n - 1 - i : i - 7);
            const unsigned long long sq = gmx_load8_unaligned(a_seq), ql = gmx_load8_unaligned(a_qual);
            const int64_t x0 = pos + i - 10;                                   // first (lowest) entering column of the block
            const uintptr_t pa = (uintptr_t)(pac + (x0 >> 2));
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(pa & ~(uintptr_t)3);
            const unsigned long long gw = ((unsigned long long)__ldg(pw + 1) << 32) | __ldg(pw);
            const int64_t xb = 4 * ((x0 >> 2) - (int64_t)(pa & 3));               // genome index of the first base of byte 0 of gw
            const float4 *lut = reinterpret_cast<const float4 *>(rd.neg ? T.sub_neg : T.sub_pos);
            auto row_sub = [&](int k) -> float4 {                               // row i - k
                const int b = rd.neg ? k : 7 - k;
                return __ldg(lut + gmx_nt4((uint8_t)(sq >> (8 * b))) * GMX_NQ + gmx_qidx((uint8_t)(ql >> (8 * b)), 0));
            };
            auto row_g = [&](int k) -> int {                                    // genome base of column i - k - 3
                const int64_t x = pos + i - k - 3;
                const int rel = (int)(x - xb);                                  // bases from the start of gw: 4 per byte, first base in the top bits
                return (int)((gw >> (8 * (rel >> 2) + 2 * (3 - (rel & 3)))) & 3ull);
            };
            gmx_nw_fast_row<7, true>(st, i, n, row_sub(0), pac, pos, gap, row_g(0));
            gmx_nw_fast_row<6, true>(st, i - 1, n, row_sub(1), pac, pos, gap, row_g(1));
            gmx_nw_fast_row<5, true>(st, i - 2, n, row_sub(2), pac, pos, gap, row_g(2));
            gmx_nw_fast_row<4, true>(st, i - 3, n, row_sub(3), pac, pos, gap, row_g(3));
            gmx_nw_fast_row<3, true>(st, i - 4, n, row_sub(4), pac, pos, gap, row_g(4));
            gmx_nw_fast_row<2, true>(st, i - 5, n, row_sub(5), pac, pos, gap, row_g(5));
            gmx_nw_fast_row<1, true>(st, i - 6, n, row_sub(6), pac, pos, gap, row_g(6));
            gmx_nw_fast_row<0, true>(st, i - 7, n, row_sub(7), pac, pos, gap, row_g(7));
            i -= 8;
            continue;
        }
        const float4 sub = rd.sub_row(T, i);
        switch (i & 7) {
            case 7: gmx_nw_fast_row<7, false>(st, i, n, sub, pac, pos, gap); break;
            case 6: gmx_nw_fast_row<6, false>(st, i, n, sub, pac, pos, gap); break;
            case 5: gmx_nw_fast_row<5, false>(st, i, n, sub, pac, pos, gap); break;
            case 4: gmx_nw_fast_row<4, false>(st, i, n, sub, pac, pos, gap); break;
            case 3: gmx_nw_fast_row<3, false>(st, i, n, sub, pac, pos, gap); break;
            case 2: gmx_nw_fast_row<2, false>(st, i, n, sub, pac, pos, gap); break;
            case 1: gmx_nw_fast_row<1, false>(st, i, n, sub, pac, pos, gap); break;
            default: gmx_nw_fast_row<0, false>(st, i, n, sub, pac, pos, gap); break;
        }
        --i;
    }
    return st.S[0];                                           // nm[0][0]
}

__device__ __forceinline__ float gmx_nw_score_dispatch(const ReadView &rd, const WindowView &win, const DevTables &T, float gap, int G)
{
    if (G == 3 && win.pac && !rd.pwm && rd.n >= 8) return gmx_nw_band_score_fast3(rd, win.pac, win.pos, T, gap);
    if (G == 3) return gmx_nw_band_score<3>(rd, win, T, gap);
    return gmx_nw_band_score_any(rd, win, T, gap, G);
}

// ---- K2b: forward fill with moves + traceback -------------------------------------------------
// Moves: 2 bits per band cell (0 = D, 1 = U, 2 = L), one uint32 per row (2G+1 <= 15 cells), kept
// in a global scratch laid out [row][task] so that the lanes of a warp write adjacent words.
#define GMX_MV_D 0u
#define GMX_MV_U 1u
#define GMX_MV_L 2u

struct TracebackOut {
    uint8_t *aligned;      // this task's gapped read string (capacity aligned_cap), may be null
    int aligned_cap;
    char *cigar;           // this task's CIGAR text (capacity cigar_cap, NUL terminated), may be null
    int cigar_cap;
    int fix_deletions;     // apply fix_CIGAR_for_deletions (reference SequenceOperations.h:32-42) + "*" for empty
    uint32_t *truncated;   // may be null: counts the tasks whose CIGAR did not fit (more runs than GMX_CIGAR_MAX_OPS or more
                           // text than cigar_cap) -- the reference builds the string unbounded (MAX_CIGAR_SZ 1024 in TopReadOutput)
};
#define GMX_CIGAR_MAX_OPS 256

// consensus character of oriented row i; i == n yields the std::string terminator the reference
// reads at src/bin_seq.cpp:607,660
struct ConsView {
    const uint8_t *explicit_chars;   // non-null: explicit consensus (length n)
    __device__ __forceinline__ uint8_t at(const ReadView &rd, const DevTables &T, int i) const
    {
        if (i >= rd.n) return 0;
        if (explicit_chars) return explicit_chars[i];
        return (uint8_t)gmx_max_char(rd.pwm_row(T, i));
    }
};

// forward fill of the band, one move word per row.  G_T > 0: band half-width known at compile time (rows live in
// registers); G_T == 0: run-time G (local-memory rows).
template <int G_T>
__device__ __forceinline__ void gmx_nw_fill_moves(const ReadView &rd, const WindowView &win, const DevTables &T, float gap, int G_rt,
                                                  uint32_t *moves, int64_t mv_stride)
{
    constexpr int GMAX = G_T > 0 ? G_T : GMX_MAX_GAP;
    const int G = G_T > 0 ? G_T : G_rt;
    const int n = rd.n, m = rd.n;
    float prev[2 * GMAX + 3], cur[2 * GMAX + 3];
    // row 0: nm[0][j] = gap * j for j <= G+2  (reference src/bin_seq.cpp:508-511)
#pragma unroll
    for (int x = 0; x < 2 * GMAX + 3; ++x) { int d = x - G - 1; prev[x] = (x < 2 * G + 3 && d >= 0) ? __fmul_rn(gap, (float)d) : GMX_NEG_INF; }
    // window bases of the band cells of row i: gb[x] = base(i + (x - G - 1) - 1), slid by one per row
    int gb[2 * GMAX + 3];
#pragma unroll
    for (int x = 0; x < 2 * GMAX + 3; ++x) { int j = 1 + (x - G - 1) - 1; gb[x] = (x <= 2 * G + 1 && j >= 0 && j < m) ? win.base(j) : 0; }
    // one row of the fill; g_enter: the genome base of the column entering the band on the right if the caller has
    // it already, else -1
    auto do_row = [&](int i, const float4 &sub, float other, int g_enter) {
        uint32_t mv = 0;
        // guard cell left of the band: column 0 carries gap*i for i <= G+2, otherwise NEG_INF
        cur[0] = (i - G - 1 == 0) ? __fmul_rn(gap, (float)i) : GMX_NEG_INF;
#pragma unroll
        for (int x = 1; x <= 2 * GMAX + 1; ++x) {
            if (x <= 2 * G + 1) {
                const int d = x - G - 1;
                const int j = i + d;
                float v;
                if (j <= 0) v = (j == 0) ? __fmul_rn(gap, (float)i) : GMX_NEG_INF;
                else if (j > m) v = GMX_NEG_INF;
                else {
                    float diag = __fadd_rn(prev[x], gmx_sel4(sub, gb[x], other));
                    float upgap = __fadd_rn(prev[x + 1], gap);
                    float leftgap = __fadd_rn(cur[x - 1], gap);
                    uint32_t path;                         // reference src/bin_seq.cpp:989-1011
                    if (diag >= upgap) { if (diag >= leftgap) { path = GMX_MV_D; v = diag; } else { path = GMX_MV_L; v = leftgap; } }
                    else               { if (upgap >= leftgap) { path = GMX_MV_U; v = upgap; } else { path = GMX_MV_L; v = leftgap; } }
                    mv |= path << (2 * (x - 1));
                }
                cur[x] = v;
            }
        }
#pragma unroll
        for (int x = 0; x < 2 * GMAX + 3; ++x) { if (x == 2 * G + 2) cur[x] = GMX_NEG_INF; if (x <= 2 * G + 2) prev[x] = cur[x]; }
#pragma unroll
        for (int x = 0; x < 2 * GMAX + 2; ++x) gb[x] = gb[x + 1];
        { int j = (i + 1) + G - 1; if (2 * G + 1 < 2 * GMAX + 3) gb[2 * G + 1] = j < m ? (g_enter >= 0 ? g_enter : win.base(j)) : 0; }
        moves[(int64_t)i * mv_stride] = mv;
    };
    int i = 1;
    while (i <= n) {
        if (G_T == 3 && win.pac && !rd.pwm && i + 6 <= n - 1) {
            // Eight rows whose sequence / quality bytes and entering genome bases are fetched together (the lanes of a
            // warp sit on different reads and genome positions: a byte load is a wavefront per lane), as in K2a.
            const int r = i - 1;                                                 // read rows r .. r + 7
            const uint8_t *a_seq = rd.seq + (rd.neg ? n - 1 - r - 7 : r), *a_qual = rd.qual + (rd.neg ? n - 1 - r - 7 : r);
            const unsigned long long sq = gmx_load8_unaligned(a_seq), ql = gmx_load8_unaligned(a_qual);
            const int64_t x0 = win.pos + i + G;                                  // entering column of row i (the lowest of the block)
            const uintptr_t pa = (uintptr_t)(win.pac + (x0 >> 2));
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(pa & ~(uintptr_t)3);
            const unsigned long long gw = ((unsigned long long)__ldg(pw + 1) << 32) | __ldg(pw);
            const int64_t xb = 4 * ((x0 >> 2) - (int64_t)(pa & 3));               // genome index of the first base of byte 0 of gw
            const float4 *lut = reinterpret_cast<const float4 *>(rd.neg ? T.sub_neg : T.sub_pos);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int bsel = rd.neg ? 7 - k : k;
                const float4 sub = __ldg(lut + gmx_nt4((uint8_t)(sq >> (8 * bsel))) * GMX_NQ + gmx_qidx((uint8_t)(ql >> (8 * bsel)), 0));
                const int rel = (int)(x0 + k - xb);
                const int g = (int)((gw >> (8 * (rel >> 2) + 2 * (3 - (rel & 3)))) & 3ull);
                do_row(i + k, sub, 0.f, g);
            }
            i += 8;
            continue;
        }
        do_row(i, rd.sub_row(T, i - 1), win.pac ? 0.f : rd.sub_other(T, i - 1), -1);
        ++i;
    }
}

__device__ int gmx_nw_traceback(const ReadView &rd, const WindowView &win, const ConsView &cons, const DevTables &T,
                                float gap, int G, uint32_t *moves, int64_t mv_stride, TracebackOut out)
{
    const int n = rd.n, m = rd.n;
    if (G == 3) gmx_nw_fill_moves<3>(rd, win, T, gap, G, moves, mv_stride);
    else gmx_nw_fill_moves<0>(rd, win, T, gap, G, moves, mv_stride);

    // pass 1: path length and run-length ops, walking back from (n, m)  (reference :571-698)
    uint16_t ops[GMX_CIGAR_MAX_OPS];                   // (count << 2) | type, in backward order
    int n_ops = 0, alen = 0;
    bool cut = false;
    {
        int i = n, j = m, c_type = 0, c_counter = 0;
        auto push = [&](int type) {
            if (c_type == type) c_counter++;
            else {
                if (c_counter) { if (n_ops < GMX_CIGAR_MAX_OPS) ops[n_ops++] = (uint16_t)((c_counter << 2) | c_type); else cut = true; }
                c_type = type; c_counter = 1;
            }
        };
        // the walk visits the rows in descending order, at most one new row per step: eight move words are fetched
        // at once (independent loads) instead of one dependent load per step
        while (i != 0 && j != 0) {
            uint32_t rows[8];
            const int top = i;
#pragma unroll
            for (int k = 0; k < 8; ++k) rows[k] = top - k >= 1 ? moves[(int64_t)(top - k) * mv_stride] : 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                while (i == top - k && i != 0 && j != 0) {
                    uint32_t mv = (rows[k] >> (2 * (j - i + G))) & 3u;
                    if (mv == GMX_MV_D) { push(0); i--; j--; }
                    else if (mv == GMX_MV_U) { push(1); i--; }
                    else { push(2); j--; }
                    alen++;
                }
            }
        }
        while (i > 0) { push(1); i--; alen++; }
        while (j > 0) { push(2); j--; alen++; }
        if (c_counter > 0) { if (n_ops < GMX_CIGAR_MAX_OPS) ops[n_ops++] = (uint16_t)((c_counter << 2) | c_type); else cut = true; }
    }
    // pass 2: the gapped read string, written at its final (reversed) positions
    if (out.aligned) {
        int i = n, j = m, k = 0;
        auto put = [&](uint8_t ch) { int at = alen - 1 - k; if (at < out.aligned_cap) out.aligned[at] = ch; k++; };
        while (i != 0 && j != 0) {
            uint32_t rows[8];
            const int top = i;
#pragma unroll
            for (int k = 0; k < 8; ++k) rows[k] = top - k >= 1 ? moves[(int64_t)(top - k) * mv_stride] : 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                while (i == top - k && i != 0 && j != 0) {
                    uint32_t mv = (rows[k] >> (2 * (j - i + G))) & 3u;
                    if (mv == GMX_MV_D) { put(cons.at(rd, T, i - 1)); i--; j--; }
                    else if (mv == GMX_MV_U) { put(cons.at(rd, T, i)); i--; }      // sic: consense[i]
                    else { put('-'); j--; }
                }
            }
        }
        while (i > 0) { put(cons.at(rd, T, i)); i--; }
        while (j > 0) { put('-'); j--; }
        if (alen < out.aligned_cap) out.aligned[alen] = 0;
    }
    // CIGAR text, forward order = ops reversed
    if (out.cigar) {
        int last = 0;                                  // ops[0] is the LAST op of the forward CIGAR
        if (out.fix_deletions && n_ops > 0 && (ops[0] & 3) == 2) last = 1;   // strip a trailing D run
        int p = 0;
        for (int o = n_ops - 1; o >= last; --o) {
            int cnt = ops[o] >> 2, type = ops[o] & 3;
            char digits[8]; int nd = 0;
            do { digits[nd++] = (char)('0' + cnt % 10); cnt /= 10; } while (cnt);
            if (p + nd + 1 > out.cigar_cap - 1) { cut = true; break; }
            while (nd) out.cigar[p++] = digits[--nd];
            out.cigar[p++] = type == 0 ? 'M' : (type == 1 ? 'I' : 'D');
        }
        if (cut && out.truncated) atomicAdd(out.truncated, 1u);
        if (p == 0 && out.fix_deletions && n_ops == 0 && p < out.cigar_cap - 1) out.cigar[p++] = '*';
        out.cigar[p] = 0;
    }
    return alen;
}
