// pair_hmm.cuh -- K2c: the 3-state (M, X, Y) pair-HMM forward / backward posterior of
// bin_seq::pairHMM (reference src/bin_seq.cpp:60-244), FP64 state with FP32 emissions exactly as
// the reference mixes them (inc/bin_seq.h:48-69: the parameters are floats, promoted inside the
// recurrences).  All products and sums use the reference's association order without FMA
// contraction (__dmul_rn / __dadd_rn), so forward and backward values are bit-identical to the
// CPU; only the order of the final per-column float accumulation differs.
//
// Mapping: one warp per alignment.  Lane l owns a strip of C = ceil(m/32) genome columns and the
// warp sweeps the rows as a skewed wavefront: at step s lane l computes row (s - l) of its strip,
// receiving the boundary cell of the strip on its left (forward) / right (backward) through warp
// shuffles.  The forward matrix is parked in a per-warp global scratch (3 doubles per cell) and
// re-read by the backward sweep, which forms the posteriors on the fly and reduces them per
// genome column.  This is FP64-pipe work; there is no tensor-core formulation of this recurrence
// that keeps FP64 state.
#pragma once

#include "pipeline.cuh"

#define GMX_PHMM_THREADS 32
#define GMX_PHMM_MAXC 8                       // columns per lane of the generic path's register strips: reads <= 256 bp
#define GMX_PHMM_LONGC 32                     // ... of its long-read instantiation (strips in local memory): reads <= 1024 bp

// forward matrix parked in global memory: (fM, fY) per cell -- the X state never enters the posterior
__host__ __device__ inline size_t gmx_phmm_scratch_doubles(int max_len) { return (size_t)(max_len + 33) * (size_t)(((max_len + 31) / 32) * 32) * 2; }

struct PhmmConst {
    double Tmm, Tgm, Tmg, Tgg, q, t;          // floats promoted to double
    float  fTmm, fTgm, qTmg, qTgg;            // the float products the backward pass forms first
};

__device__ __forceinline__ PhmmConst gmx_phmm_const()
{   // reference inc/bin_seq.h:60-69
    float q = 0.25f, t = 0.05f, d = 0.0025f, e = 0.5f;
    float Tmm = __fsub_rn(__fsub_rn(1.0f, __fmul_rn(2.0f, d)), t);
    float Tgm = __fsub_rn(__fsub_rn(1.0f, d), t);
    float Tmg = d, Tgg = e;
    PhmmConst c;
    c.Tmm = Tmm; c.Tgm = Tgm; c.Tmg = Tmg; c.Tgg = Tgg; c.q = q; c.t = t;
    c.fTmm = Tmm; c.fTgm = Tgm; c.qTmg = __fmul_rn(q, Tmg); c.qTgg = __fmul_rn(q, Tgg);
    return c;
}

__device__ __forceinline__ double gmx_shfl_d(double v, int src)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}

// emission p_seq(pwm[i], genome base) for a strand-oriented read row; g == 4 -> non-acgt window char
__device__ __forceinline__ float gmx_phmm_emit(const ReadView &rd, const DevTables &T, float4 row, int i, int g)
{
    if (g < 4) return gmx_sel4(row, g, 0.f);
    float4 p = rd.pwm_row(T, i);
    const float *s = T.P + 4 * (int)'n';
    float sum = 0.f;
    sum = __fadd_rn(sum, __fmul_rn(p.x, s[0])); sum = __fadd_rn(sum, __fmul_rn(p.y, s[1]));
    sum = __fadd_rn(sum, __fmul_rn(p.z, s[2])); sum = __fadd_rn(sum, __fmul_rn(p.w, s[3]));
    return __fmul_rn(3.f, sum);
}

// One warp: posteriors of read `rd` against window `win` (m == n).  post: float[m][5], zeroed by the caller.
// MAXC = columns per lane the register arrays are sized for (5: reads <= 160 bp, 8: <= 256 bp).
template <int MAXC>
__device__ void gmx_pair_hmm_warp(const ReadView &rd, const WindowView &win, const DevTables &T, double *F, float *post)
{
    const int lane = threadIdx.x & 31;
    const int n = rd.n, m = rd.n;
    const int C = (m + 31) >> 5;
    const PhmmConst K = gmx_phmm_const();
    const int j0 = lane * C;                                   // first 0-based genome column of the strip
    int gb[MAXC], gbn[MAXC];                  // genome base of column j, and of column j+1
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        int j = j0 + c;
        gb[c] = (c < C && j < m) ? win.base(j) : 0;
        gbn[c] = (c < C && j + 1 < m) ? win.base(j + 1) : 0;
    }

    // ---------------- forward (reference :141-157); f index (i, j) 1-based, strip column c <-> j = j0 + c + 1
    double pM[MAXC], pX[MAXC], pY[MAXC];     // row i-1 of the strip
#pragma unroll
    for (int c = 0; c < MAXC; ++c) { pM[c] = 0; pX[c] = 0; pY[c] = 0; }
    // boundary cells on the left of the strip: (i-1, j0) "diag" and (i, j0) "left"
    double dM = 0, dX = 0, dY = 0, lM = 0, lY = 0;
    double outM = 0, outX = 0, outY = 0;                         // last column of the row just computed
    double fE_M = 0, fE_X = 0, fE_Y = 0;
    float4 row_nxt = (1 - lane >= 1 && 1 - lane <= n) ? rd.phmm_row(T, 0 - lane) : make_float4(0, 0, 0, 0);
    for (int s = 1; s <= n + 31; ++s) {
        // receive the left neighbour's last column of the row it computed in the previous step (= my row i)
        double rM = gmx_shfl_d(outM, lane - 1), rX = gmx_shfl_d(outX, lane - 1), rY = gmx_shfl_d(outY, lane - 1);
        int i = s - lane;                                      // 1-based read row
        const float4 row_cur = row_nxt;
        row_nxt = (i + 1 >= 1 && i + 1 <= n) ? rd.phmm_row(T, i) : make_float4(0, 0, 0, 0);   // emission row of the next step
        if (lane == 0) { rM = 0; rX = 0; rY = 0; }             // column 0: f[i][0] = 0 for i >= 1
        // previous step's received row is now the diagonal row (i-1); for lane 0, (i-1, 0) is (0,0) when i == 1
        if (i >= 1 && i <= n) {
            if (lane == 0) { dM = (i == 1) ? 1.0 : 0.0; dX = 0; dY = 0; }
            lM = rM; lY = rY;
            const float4 row = row_cur;
            double cM_prev = dM, cX_prev = dX, cY_prev = dY;   // (i-1, j-1) for the first column
            double leftM = lM, leftY = lY;                     // (i, j-1)
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
                int j = j0 + c;                                // 0-based genome column
                if (c < C && j < m) {
                    float e = gmx_phmm_emit(rd, T, row, i - 1, gb[c]);
                    double sum = __dadd_rn(__dadd_rn(__dmul_rn(K.Tmm, cM_prev), __dmul_rn(K.Tgm, cX_prev)), __dmul_rn(K.Tgm, cY_prev));
                    double fM = __dmul_rn((double)e, sum);
                    double fX = __dmul_rn(K.q, __dadd_rn(__dmul_rn(K.Tmg, pM[c]), __dmul_rn(K.Tgg, pX[c])));
                    double fY = __dmul_rn(K.q, __dadd_rn(__dmul_rn(K.Tmg, leftM), __dmul_rn(K.Tgg, leftY)));
                    cM_prev = pM[c]; cX_prev = pX[c]; cY_prev = pY[c];      // becomes (i-1, j) = diag of the next column
                    pM[c] = fM; pX[c] = fX; pY[c] = fY;
                    leftM = fM; leftY = fY;
                    reinterpret_cast<double2 *>(F)[(size_t)(i - 1) * m + j] = make_double2(fM, fY);
                    if (i == n && j == m - 1) { fE_M = fM; fE_X = fX; fE_Y = fY; }
                    outM = fM; outX = fX; outY = fY;
                }
            }
            // the row received now is the diagonal boundary for the next row
            dM = rM; dX = rX; dY = rY;
        }
    }
    // fE lives in the lane that owns column m-1
    int owner = (m - 1) / C;
    fE_M = gmx_shfl_d(fE_M, owner); fE_X = gmx_shfl_d(fE_X, owner); fE_Y = gmx_shfl_d(fE_Y, owner);
    const double fE = __dmul_rn(K.t, __dadd_rn(__dadd_rn(fE_M, fE_X), fE_Y));                  // reference :160

    // ---------------- backward (reference :164-204) + posterior (:206-241); b index (i, j) 0-based
    double qM[MAXC], qX[MAXC];               // row i+1 of the strip (bM, bX)
    float acc[MAXC][5];
#pragma unroll
    for (int c = 0; c < MAXC; ++c) { qM[c] = 0; qX[c] = 0;
#pragma unroll
        for (int b = 0; b < 5; ++b) acc[c][b] = 0.f; }
    double eM = 0;                                             // bM of (i+1, strip_end+1): diagonal boundary on the right
    double sndM = 0, sndY = 0;                                 // first column of the row just computed
    const int last_lane = (m - 1) / C;
    for (int s = 0; s <= n - 1 + 31; ++s) {
        double rM = gmx_shfl_d(sndM, lane + 1), rY = gmx_shfl_d(sndY, lane + 1);
        int i = (n - 1) - (s - (last_lane - lane));            // lanes right of last_lane own no columns
        bool live = lane <= last_lane && i >= 0 && i <= n - 1;
        if (live) {
            if (lane == last_lane) { rM = 0; rY = 0; }
            float4 row = (i + 1 <= n - 1) ? rd.phmm_row(T, i + 1) : make_float4(0, 0, 0, 0);
            int code = 4;
            { char ch = gmx_max_char(rd.pwm_row(T, i)); code = ch == 'a' ? 0 : ch == 'c' ? 1 : ch == 'g' ? 2 : ch == 't' ? 3 : 4; }
            double rightM_next = eM;                           // bM (i+1, j+1) for the last column of the strip
            double rightY = rY;                                // bY (i, j+1)
            double firstM = 0, firstY = 0;
            double2 fv[MAXC];                                  // f[i+1][j+1] of the strip, all loads in flight before the chain
#pragma unroll
            for (int c = 0; c < MAXC; ++c) {
                int j = j0 + c;
                fv[c] = (c < C && j < m) ? reinterpret_cast<const double2 *>(F)[(size_t)i * m + j] : make_double2(0, 0);
            }
#pragma unroll
            for (int c = MAXC - 1; c >= 0; --c) {
                int j = j0 + c;
                if (c < C && j < m) {
                    double bM, bX, bY;
                    // (i+1, j+1): from the strip itself unless c is its last column
                    double diagM = rightM_next;
                    if (j == m - 1 && i == n - 1) { bM = K.t; bX = K.t; bY = K.t; }
                    else if (j == m - 1) {
                        bM = __dmul_rn((double)K.qTmg, qX[c]);
                        bX = __dmul_rn((double)K.qTgg, qX[c]);
                        bY = 0;
                    } else if (i == n - 1) {
                        bM = __dmul_rn((double)K.qTmg, rightY);
                        bY = __dmul_rn((double)K.qTgg, rightY);
                        bX = 0;
                    } else {
                        float e = gmx_phmm_emit(rd, T, row, i + 1, gbn[c]);
                        float eTmm = __fmul_rn(e, K.fTmm), eTgm = __fmul_rn(e, K.fTgm);
                        bM = __dadd_rn(__dadd_rn(__dmul_rn((double)eTmm, diagM), __dmul_rn((double)K.qTmg, qX[c])), __dmul_rn((double)K.qTmg, rightY));
                        bX = __dadd_rn(__dmul_rn((double)eTgm, diagM), __dmul_rn((double)K.qTgg, qX[c]));
                        bY = __dadd_rn(__dmul_rn((double)eTgm, diagM), __dmul_rn((double)K.qTgg, rightY));
                    }
                    // posterior of (read i, genome j): f[i+1][j+1] * b[i][j] / fE, M and Y states
                    double pMv = __ddiv_rn(__dmul_rn(fv[c].x, bM), fE);
                    double pYv = __ddiv_rn(__dmul_rn(fv[c].y, bY), fE);
                    double add = __dadd_rn(pYv, pMv);
#pragma unroll
                    for (int b = 0; b < 5; ++b) if (b == code) acc[c][b] = (float)__dadd_rn((double)acc[c][b], add);
                    rightM_next = qM[c];                       // (i+1, j) = diag of the next column to the left
                    qM[c] = bM; qX[c] = bX;
                    rightY = bY;
                    firstM = bM; firstY = bY;
                }
            }
            eM = rM;                                           // received (i, strip_end+1).M is next row's diagonal
            sndM = firstM; sndY = firstY;
        }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        int j = j0 + c;
        if (c < C && j < m) {
#pragma unroll
            for (int b = 0; b < 5; ++b) post[(size_t)j * 5 + b] = acc[c][b];
        }
    }
}

// ---- fast path: packed-genome windows, C columns per lane known at compile time -------------------------
// Same recurrences, but the strip loops carry no per-cell range checks: the matrix is padded to 32*C columns.
// Padding is inert: forward values flow left->right / top->bottom (padded columns never feed real ones), and in the
// backward sweep a padded cell only ever combines zeros, so real cells of the last column see exactly the
// reference's special-cased formulas (x + 0.0 is exact).  Row n-1 of the backward sweep (reference :164-183) is
// peeled.  Posteriors are accumulated in shared memory ([column slot][code][lane]: conflict-free) and the two
// divisions by fE become multiplications by its reciprocal and the recurrences use fused multiply-adds (K2c is a
// tolerance kernel: 1e-5 relative; the generic path below keeps the reference's operation order).
#define GMX_PHMM_FAST_MAXC 5

// Parked forward values: (fM, fY) of the C cells a lane computes in one step, as two float bit patterns per cell scaled
// by a power of two chosen per (lane, step) so that the largest of the 2C values lands in [1, 2), plus that exponent
// offset (one int per lane and step).  The forward matrix spans 1e-300 .. 1e+66 over a 150 x 150 alignment, far outside float range, but
// within one lane-step the values that matter sit within a few decades of the largest: a cell more than 2^-126 below
// it flushes to zero, and its posterior f * b / fE is then below 1e-25 (b varies by at most (1 / (q * Tmg))^(C-1) ~ 1e13
// across the C adjacent columns while f * b / fE <= 1 holds for the largest) -- far under the 1e-7 absolute tolerance of
// this kernel.  The stored mantissa keeps 24 bits (relative 6e-8 against the kernel's 1e-5).  This halves the one
// HBM-bound stream of the kernel (ncu before: DRAM 4.06 TB/s = 50 % of peak next to 49 % issue utilisation) and frees
// twenty registers of prefetched forward values.
// scaled float bits of a non-negative double: the float of x * 2^(1023 - eb) with the mantissa truncated to 24 bits, by
// integer arithmetic on the bit pattern (hi word minus the exponent offset, then a funnel shift by 3) -- the conversion
// unit (F2F, a quarter-rate pipe) was the busiest pipe of this kernel (ncu: xu 46 %) while it packed with DMUL + F2F.
// Values more than 2^-127 below the scale come out as (sub)normal-range garbage below 1e-38 of it: inert.
__device__ __forceinline__ uint32_t gmx_phmm_pack(double x, int c_off)
{
    const int t = max(__double2hiint(x) - c_off, 0);
    return __funnelshift_l((uint32_t)__double2loint(x), (uint32_t)t, 3);
}
__device__ __forceinline__ double gmx_phmm_unpack(uint32_t f, int c_off)
{
    return __hiloint2double((int)(f >> 3) + c_off, (int)(f << 29));
}

#define GMX_PHMM_ESTRIDE(C) (32 * (C) + 4)      // doubles per genome base in the emission table (rows 0 .. n, padded)

template <int C>
__device__ void gmx_pair_hmm_warp_fast(const ReadView &rd, const uint8_t *pac, int64_t pos, const DevTables &T, uint2 *F, int *E, float *post,
                                       float *acc_s /* [C][5][32] */, double *e_s /* [4][ESTRIDE] */, uint8_t *code_s /* [32*C] */)
{
    const int lane = threadIdx.x & 31;
    const int n = rd.n, m = rd.n;
    constexpr int MP = 32 * C;                                   // padded row length
    constexpr int ES = GMX_PHMM_ESTRIDE(C);
    // per-row operands of the read, staged once: the emission p_seq(pwm[i], g) as a DOUBLE per genome base g (one
    // shared-memory load per cell instead of a four-way select and a conversion; rows past the read are zero, so the
    // backward sweep's "row i + 1" needs no range test) and the consensus code
    for (int i = lane; i < ES; i += 32) {
        float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            r4 = rd.phmm_row(T, i);
            const char ch = gmx_max_char(rd.pwm_row(T, i));
            code_s[i] = (uint8_t)(ch == 'a' ? 0 : ch == 'c' ? 1 : ch == 'g' ? 2 : ch == 't' ? 3 : 4);
        }
        e_s[i] = (double)r4.x; e_s[ES + i] = (double)r4.y; e_s[2 * ES + i] = (double)r4.z; e_s[3 * ES + i] = (double)r4.w;
    }
    __syncwarp();
    const PhmmConst K = gmx_phmm_const();
    const int j0 = lane * C;
    const double *ep[C], *epn[C];             // emission column of genome base j, and of base j + 1 (zero column when past the window)
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j = j0 + c;
        ep[c] = e_s + (j < m ? gmx_pac_base(pac, pos + j) : 0) * ES;
        epn[c] = e_s + (j + 1 < m ? gmx_pac_base(pac, pos + j + 1) : 0) * ES;
    }
#pragma unroll
    for (int x = 0; x < C * 5; ++x) acc_s[x * 32 + lane] = 0.f;

    // ---------------- forward
    const double dqTmg = (double)K.qTmg, dqTgg = (double)K.qTgg;
    double pM[C], pX[C], pY[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { pM[c] = 0; pX[c] = 0; pY[c] = 0; }
    double dM = 0, dX = 0, dY = 0;
    double outM = 0, outX = 0, outY = 0;
    for (int s = 1; s <= n + 31; ++s) {
        double rM = gmx_shfl_d(outM, lane - 1), rX = gmx_shfl_d(outX, lane - 1), rY = gmx_shfl_d(outY, lane - 1);
        const int i = s - lane;                                  // 1-based read row
        if (lane == 0) { rM = 0; rX = 0; rY = 0; }
        if (i >= 1 && i <= n) {
            if (lane == 0) { dM = (i == 1) ? 1.0 : 0.0; dX = 0; dY = 0; }
            double cM = dM, cX = dX, cY = dY;                    // (i-1, j-1)
            double leftM = rM, leftY = rY;                       // (i, j-1)
            // parked by STEP, slot-major: at any step the 32 lanes write 32 adjacent pairs (the backward sweep reads the
            // block of forward step n + 31 - s_b at its step s_b: the skew cancels, both sides are coalesced)
            uint2 *frow = F + (size_t)s * MP + lane;
            int eh = 0;                                           // largest high word = largest value (all are >= 0)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double e = ep[c][i - 1];
                // fused multiply-adds: this is the tolerance kernel (posteriors within 1e-5), and FMA only removes roundings
                const double sum = fma(K.Tmm, cM, fma(K.Tgm, cX, K.Tgm * cY));
                const double fM = e * sum;
                const double fX = fma(dqTmg, pM[c], dqTgg * pX[c]);
                const double fY = fma(dqTmg, leftM, dqTgg * leftY);
                cM = pM[c]; cX = pX[c]; cY = pY[c];
                pM[c] = fM; pX[c] = fX; pY[c] = fY;
                leftM = fM; leftY = fY;
                eh = max(eh, max(__double2hiint(fM), __double2hiint(fY)));
            }
            int eb = eh >> 20;                                    // biased exponent of the largest value of this lane-step
            if (eb == 0) eb = 1023;                               // all zero (or denormal): any scale will do
            const int c_off = (eb - 127) << 20;                   // double exponent field -> float exponent field of x * 2^(1023 - eb)
#pragma unroll
            for (int c = 0; c < C; ++c) frow[c * 32] = make_uint2(gmx_phmm_pack(pM[c], c_off), gmx_phmm_pack(pY[c], c_off));
            E[s * 32 + lane] = c_off;
            outM = pM[C - 1]; outX = pX[C - 1]; outY = pY[C - 1];
            dM = rM; dX = rX; dY = rY;
        }
    }
    // fE from row n, column m-1: held by lane `owner`, strip slot `c_last`
    const int owner = (m - 1) / C, c_last = (m - 1) - owner * C;
    double eMv = pM[0], eXv = pX[0], eYv = pY[0];
#pragma unroll
    for (int c = 1; c < C; ++c) if (c == c_last) { eMv = pM[c]; eXv = pX[c]; eYv = pY[c]; }
    eMv = gmx_shfl_d(eMv, owner); eXv = gmx_shfl_d(eXv, owner); eYv = gmx_shfl_d(eYv, owner);
    const double fE = __dmul_rn(K.t, __dadd_rn(__dadd_rn(eMv, eXv), eYv));
    const double inv_fE = 1.0 / fE;

    // ---------------- backward + posterior
    double qM[C], qX[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { qM[c] = 0; qX[c] = 0; }
    double eM = 0, sndM = 0, sndY = 0;
    uint2 fv_nxt[C];                                             // parked forward values of the row handled in the next step
    int e_nxt = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) fv_nxt[c] = make_uint2(0u, 0u);
    if (lane == 31) {                                            // row n-1 of lane 31 was written at forward step n + 31
#pragma unroll
        for (int c = 0; c < C; ++c) fv_nxt[c] = F[(size_t)(n + 31) * MP + c * 32 + lane];
        e_nxt = E[(n + 31) * 32 + lane];
    }
    for (int s = 0; s <= n - 1 + 31; ++s) {
        double rM = gmx_shfl_d(sndM, lane + 1), rY = gmx_shfl_d(sndY, lane + 1);
        if (lane == 31) { rM = 0; rY = 0; }
        const int i = (n - 1) - (s - (31 - lane));               // 0-based read row
        uint2 fv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) fv[c] = fv_nxt[c];
        const int c_off = e_nxt;
        if (i - 1 >= 0 && i - 1 <= n - 1) {                      // request the next step's row before this one's chain
            const uint2 *fnext = F + (size_t)(n + 31 - (s + 1)) * MP + lane;
#pragma unroll
            for (int c = 0; c < C; ++c) fv_nxt[c] = fnext[c * 32];
            e_nxt = E[(n + 31 - (s + 1)) * 32 + lane];
        }
        if (i >= 0 && i <= n - 1) {
            const int code = code_s[i];
            float *acc = acc_s + code * 32 + lane;
            double diagM = eM, rightY = rY;
            if (i == n - 1) {                                    // last read row (reference :164-183)
#pragma unroll
                for (int c = C - 1; c >= 0; --c) {
                    const int j = j0 + c;
                    double bM = __dmul_rn(dqTmg, rightY), bX = 0, bY = __dmul_rn(dqTgg, rightY);
                    if (j == m - 1) { bM = K.t; bX = K.t; bY = K.t; }
                    if (j > m - 1) { bM = 0; bX = 0; bY = 0; }
                    const double fMv = gmx_phmm_unpack(fv[c].x, c_off), fYv = gmx_phmm_unpack(fv[c].y, c_off);
                    const double add = __dadd_rn(__dmul_rn(__dmul_rn(fYv, bY), inv_fE), __dmul_rn(__dmul_rn(fMv, bM), inv_fE));
                    if (j <= m - 1) acc[c * 160] = (float)__dadd_rn((double)acc[c * 160], add);
                    qM[c] = bM; qX[c] = bX; rightY = bY;
                }
            } else {
#pragma unroll
                for (int c = C - 1; c >= 0; --c) {
                    const double e = epn[c][i + 1];              // zero past the read's last row / the window's last column
                    const double eTmm = e * K.Tmm, eTgm = e * K.Tgm;
                    const double bM = fma(eTmm, diagM, fma(dqTmg, qX[c], dqTmg * rightY));
                    const double gd = eTgm * diagM;
                    const double bX = fma(dqTgg, qX[c], gd);
                    const double bY = fma(dqTgg, rightY, gd);
                    const double fMv = gmx_phmm_unpack(fv[c].x, c_off), fYv = gmx_phmm_unpack(fv[c].y, c_off);
                    const double add = fma(fYv, bY, fMv * bM) * inv_fE;
                    acc[c * 160] = (float)__dadd_rn((double)acc[c * 160], add);
                    diagM = qM[c];
                    qM[c] = bM; qX[c] = bX; rightY = bY;
                }
            }
            eM = rM;
            sndM = qM[0]; sndY = rightY;
        }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j = j0 + c;
        if (j < m) {
#pragma unroll
            for (int b = 0; b < 5; ++b) post[(size_t)j * 5 + b] = acc_s[(c * 5 + b) * 32 + lane];
        }
    }
}

// explicit-window tasks (kernel-level entry point gmx_pair_hmm)
__global__ void __launch_bounds__(GMX_PHMM_THREADS) k_pair_hmm_tasks(DevReads R, DevTables T, int64_t t0, int64_t cnt, const int32_t *read_idx,
                                                                     const uint8_t *strand, const uint8_t *windows, int win_stride,
                                                                     float *post, double *scratch, size_t per_task)
{
    int64_t t = t0 + blockIdx.x;
    if ((int64_t)blockIdx.x >= cnt) return;
    ReadView rd = gmx_read_view(R, read_idx[t], strand ? strand[t] : 0);
    WindowView win; win.pac = nullptr; win.pos = 0; win.chars = windows + t * win_stride;
    if (rd.n <= 160) gmx_pair_hmm_warp<5>(rd, win, T, scratch + (size_t)blockIdx.x * per_task, post + (size_t)t * win_stride * 5);
    else if (rd.n <= 32 * GMX_PHMM_MAXC) gmx_pair_hmm_warp<GMX_PHMM_MAXC>(rd, win, T, scratch + (size_t)blockIdx.x * per_task, post + (size_t)t * win_stride * 5);
    else gmx_pair_hmm_warp<GMX_PHMM_LONGC>(rd, win, T, scratch + (size_t)blockIdx.x * per_task, post + (size_t)t * win_stride * 5);
}

// one group leader per warp (SNPScoredSeq::score, reference src/SNPScoredSeq.cpp:44-67).  C_T > 0: every read of the
// chunk fits 32 * C_T columns (fast path, padded); C_T == 0: generic path, reads up to 256 bp; C_T < 0: generic path with
// 32 columns per lane, reads up to 1024 bp (bin_seq::pairHMM has no length limit; GMX_MAX_READ_LEN is the library's).  Persistent: the grid is what the SMs hold
// at once, every CTA (one warp) owns one scratch slot and takes group leaders from a shared cursor until none is left
// (one launch per chunk instead of one per wave: no tail of half-empty SMs between waves).
template <int C_T>
__global__ void __launch_bounds__(GMX_PHMM_THREADS) k_pair_hmm_leaders(DevIndex ix, DevReads R, DevTables T, const unsigned long long *keys,
                                                                       LeaderStore L, uint32_t n_leaders, uint32_t *cursor, double *scratch, size_t per_task,
                                                                       const uint32_t *live)
{
    constexpr int CS = C_T > 0 ? C_T : 1;
    n_leaders = gmx_live(live, n_leaders);
    __shared__ float acc_s[CS * 5 * 32];
    __shared__ double e_s[4 * GMX_PHMM_ESTRIDE(CS)];
    __shared__ uint8_t code_s[CS * 32 + 32];
    double *my = scratch + (size_t)blockIdx.x * per_task;
    while (true) {
        uint32_t s = 0;
        if (threadIdx.x == 0) s = atomicAdd(cursor, 1u);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s >= n_leaders) break;
        uint32_t task, round, diag;
        gmx_decode_key(keys[L.lead_cand[s]], task, round, diag);
        ReadView rd = gmx_read_view(R, (int)(task >> 1), (int)(task & 1));
        float *out = L.hmm + (size_t)s * L.max_len * 5;
        if (C_T > 0) {
            // scratch slot: packed pairs [n + 32][32 * C] followed by int [n + 32][32]
            uint2 *F = reinterpret_cast<uint2 *>(my);
            int *E = reinterpret_cast<int *>(F + (size_t)(L.max_len + 32) * (32 * CS));
            gmx_pair_hmm_warp_fast<CS>(rd, ix.pac, diag, T, F, E, out, acc_s, e_s, code_s);
        } else {
            WindowView win; win.pac = ix.pac; win.pos = diag; win.chars = nullptr;
            gmx_pair_hmm_warp<(C_T < 0 ? GMX_PHMM_LONGC : GMX_PHMM_MAXC)>(rd, win, T, my, out);
        }
        __syncwarp();
    }
}
