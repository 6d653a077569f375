// fm_index.cuh -- FM-index primitives on the device: occ, backward search, LF / locate.
//
// Bit-exact device restatements of the vendored BWA routines the reference calls
// (reference src/bwt.c:53-59 bwt_invPsi, :86-97 bwt_sa, :98-130 __occ_aux / bwt_occ,
//  :132-163 bwt_2occ, :222-239 bwt_match_exact) over the reference's own BWT layout:
// every 128 symbols one 64-byte block = 4 x uint64 running counts + 8 x uint32 words of sixteen
// 2-bit symbols, most significant first (reference src/bwtindex.c:128-150, inc/bwt.h:72-78).
// One block is exactly two 32-byte sectors, fetched as 4 x 128-bit loads.
#pragma once

#include "gmx_common.cuh"

// number of symbols equal to c among the sixteen 2-bit symbols of w
__device__ __forceinline__ uint32_t gmx_popc_sym(uint32_t w, uint32_t c)
{
    uint32_t x = w ^ (c * 0x55555555u);           // equal symbols become 00
    return __popc(~(x | (x >> 1)) & 0x55555555u);
}

// bwt_occ(bwt, k, c): occurrences of c in B[0..k] (reference src/bwt.c:107-130)
__device__ __forceinline__ uint64_t gmx_bwt_occ(const DevIndex &ix, uint64_t k, uint32_t c)
{
    if (k == ix.seq_len) return ix.L2[c + 1] - ix.L2[c];
    if (k == ~0ull) return 0;
    k -= (k >= ix.primary);
    const uint4 *blk = reinterpret_cast<const uint4 *>(ix.bwt + ((k >> 7) << 4));
    uint4 cnt = __ldg(blk + (c >> 1));            // counts of symbols {0,1} or {2,3}
    uint64_t n = (c & 1) ? ((uint64_t)cnt.w << 32 | cnt.z) : ((uint64_t)cnt.y << 32 | cnt.x);
    uint4 w0 = __ldg(blk + 2), w1 = __ldg(blk + 3);
    uint32_t r = (uint32_t)k & 127u;
    uint32_t nfull = r >> 4;                       // whole words before the one holding k
    uint32_t rem = r & 15u;                        // k is symbol `rem` of word `nfull`
    uint32_t keep_mask = ~((1u << ((15u - rem) << 1)) - 1u);
    uint32_t ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    uint32_t acc = 0;
#pragma unroll
    for (uint32_t w = 0; w < 8; ++w) {
        uint32_t full = gmx_popc_sym(ws[w], c);
        uint32_t part = gmx_popc_sym(ws[w] & keep_mask, c);
        acc += (w < nfull) ? full : ((w == nfull) ? part : 0u);
    }
    n += acc;
    if (c == 0) n -= (15u - rem);                  // masked-out symbols read as 'A'
    return n;
}

// symbol of the $-removed BWT at (already primary-adjusted) position x: bwt_B0, inc/bwt.h:72,78
__device__ __forceinline__ uint32_t gmx_bwt_B0(const DevIndex &ix, uint64_t x)
{
    uint32_t w = __ldg(ix.bwt + ((x >> 7) << 4) + 8 + ((x & 0x7f) >> 4));
    return (w >> ((~x & 0xf) << 1)) & 3u;
}

// bwt_invPsi (reference src/bwt.c:53-59): one LF step
__device__ __forceinline__ uint64_t gmx_inv_psi(const DevIndex &ix, uint64_t k)
{
    uint64_t x = k - (k > ix.primary);
    uint32_t c = gmx_bwt_B0(ix, x);
    uint64_t r = ix.L2[c] + gmx_bwt_occ(ix, k, c);
    return k == ix.primary ? 0 : r;
}

// bwt_sa over the SAMPLED suffix array (reference src/bwt.c:86-97)
__device__ __forceinline__ uint64_t gmx_bwt_sa_sampled(const DevIndex &ix, uint64_t k)
{
    uint64_t sa = 0, mask = (uint64_t)ix.sa_intv - 1;
    while (k & mask) { ++sa; k = gmx_inv_psi(ix, k); }
    return sa + ix.sa_samp[k / (uint64_t)ix.sa_intv];
}

// bwt_match_exact (reference src/bwt.c:222-239) with the symbols supplied by `sym(i)`, i in [0,len).
// Returns true and the inclusive interval [k,l] on a hit.
template <class SymFn>
__device__ __forceinline__ bool gmx_match_exact(const DevIndex &ix, int len, SymFn sym, uint64_t &k_out, uint64_t &l_out, uint32_t &n_steps)
{
    uint64_t k = 0, l = ix.seq_len;
    for (int i = len - 1; i >= 0; --i) {
        uint32_t c = sym(i);
        if (c > 3) return false;
        n_steps++;
        uint64_t ok = gmx_bwt_occ(ix, k - 1, c);
        uint64_t ol = gmx_bwt_occ(ix, l, c);
        k = ix.L2[c] + ok + 1;
        l = ix.L2[c] + ol;
        if (k > l) return false;
    }
    k_out = k; l_out = l;
    return true;
}

// bwt_match_exact with its first `tab_len` steps (the LAST tab_len symbols of the string) looked up in the
// memoised table; identical intervals by construction (the table is filled by gmx_match_exact itself).
template <class SymFn>
__device__ __forceinline__ bool gmx_match_exact_tab(const DevIndex &ix, int len, SymFn sym, uint64_t &k_out, uint64_t &l_out, uint32_t &n_steps)
{
    const int T = ix.tab_len;
    if (T <= 0 || len < T) return gmx_match_exact(ix, len, sym, k_out, l_out, n_steps);
    uint32_t code = 0; bool ok = true;
    for (int i = len - T; i < len; ++i) { uint32_t c = sym(i); ok = ok && c <= 3; code = (code << 2) | (c & 3u); }
    // the reference walks from the last symbol backwards and stops at the first non-ACGT symbol or empty interval;
    // either way the answer is "absent"
    for (int i = 0; i < len - T; ++i) ok = ok && sym(i) <= 3;
    if (!ok) return false;
    const uint2 kl = __ldg(ix.kmer_tab + code);
    if (kl.x > kl.y) return false;
    uint64_t k = kl.x, l = kl.y;
    for (int i = len - T - 1; i >= 0; --i) {
        uint32_t c = sym(i);
        n_steps++;
        uint64_t ok2 = gmx_bwt_occ(ix, k - 1, c);
        uint64_t ol = gmx_bwt_occ(ix, l, c);
        k = ix.L2[c] + ok2 + 1;
        l = ix.L2[c] + ol;
        if (k > l) return false;
    }
    k_out = k; l_out = l;
    return true;
}

// The same search for a k-mer already packed 2 bits per base (first base most significant) with a mask of its
// non-ACGT positions (bit t set <=> base t is not ACGT): the seed walk keeps both in a rolling window.
__device__ __forceinline__ bool gmx_match_exact_packed(const DevIndex &ix, int len, unsigned long long code, unsigned long long bad,
                                                       uint64_t &k_out, uint64_t &l_out, uint32_t &n_steps)
{
    if (bad) return false;                       // the reference stops at the first non-ACGT symbol: absent
    const int T = ix.tab_len;
    uint64_t k, l;
    int i;
    if (T > 0 && len >= T) {
        const uint2 kl = __ldg(ix.kmer_tab + (uint32_t)(code & ((1ull << (2 * T)) - 1ull)));
        if (kl.x > kl.y) return false;
        k = kl.x; l = kl.y; i = len - T - 1;
    } else { k = 0; l = ix.seq_len; i = len - 1; }
    for (; i >= 0; --i) {
        const uint32_t c = (uint32_t)(code >> (2 * (len - 1 - i))) & 3u;
        n_steps++;
        const uint64_t ok = gmx_bwt_occ(ix, k - 1, c);
        const uint64_t ol = gmx_bwt_occ(ix, l, c);
        k = ix.L2[c] + ok + 1;
        l = ix.L2[c] + ol;
        if (k > l) return false;
    }
    k_out = k; l_out = l;
    return true;
}

// ---- kernels ---------------------------------------------------------------------------------

// fill the memoised table: entry `code` = interval of the T-mer whose symbols are the base-4 digits of `code`
__global__ void k_build_kmer_table(DevIndex ix, int T, uint2 *tab)
{
    uint32_t code = blockIdx.x * blockDim.x + threadIdx.x;
    if (code >= (1u << (2 * T))) return;
    uint64_t k = 0, l = 0; uint32_t steps = 0;
    bool hit = gmx_match_exact(ix, T, [&](int i) { return (code >> (2 * (T - 1 - i))) & 3u; }, k, l, steps);
    tab[code] = hit ? make_uint2((uint32_t)k, (uint32_t)l) : make_uint2(1u, 0u);
}

// K1 (primitive form): one k-mer per thread.  GenomeBwt::get_sa_int (reference src/GenomeBwt.cpp:438-474)
__global__ void k_fm_search(DevIndex ix, const uint8_t *kmers, int len, int64_t n, uint64_t *k_out, uint64_t *l_out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint8_t *s = kmers + t * len;
    uint64_t k = 0, l = 0;
    uint32_t steps = 0;
    bool hit = gmx_match_exact(ix, len, [&](int i) { uint8_t ch = s[i]; return (uint32_t)(ch < 4 ? ch : gmx_nt4(ch)); }, k, l, steps);
    k_out[t] = hit ? k : 0;
    l_out[t] = hit ? l : 0;
}

// K1b (primitive form): GenomeBwt::get_sa_coord (reference src/GenomeBwt.cpp:431-436)
__global__ void k_sa_locate(DevIndex ix, const uint64_t *ranks, int64_t n, int mode, uint64_t *pos_out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t k = ranks[t];
    uint64_t v;
    if (mode == 1) v = gmx_bwt_sa_sampled(ix, k);
    else v = (k == 0) ? ~0ull : (uint64_t)ix.sa_full[k];      // bwt_sa(0) = sa[0] = (bwtint_t)-1
    pos_out[t] = v;
}

// De-sample the suffix array once at load: sa_full[k] = bwt_sa(k) for every rank.  Each thread
// walks LF until it reaches a sampled rank (<= sa_intv - 1 steps, 15.5 on average); every step
// touches one 64-byte BWT block, which stays L2-resident (the whole BWT is 0.5 byte / base).
__global__ void k_desample_sa(DevIndex ix, uint32_t *sa_full)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > ix.seq_len) return;
    sa_full[k] = (k == 0) ? (uint32_t)ix.seq_len : (uint32_t)gmx_bwt_sa_sampled(ix, k);
}

// GenomeBwt::GetString (reference src/GenomeBwt.cpp:384-415)
__global__ void k_get_windows(DevIndex ix, const uint64_t *begin, int64_t n, int size, uint8_t *chars, int32_t *len_out)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint64_t b = begin[t];
    bool ok = gmx_window_valid(ix, b, size);
    len_out[t] = ok ? size : 0;
    uint8_t *o = chars + t * size;
    for (int i = 0; i < size; ++i) o[i] = ok ? (uint8_t)("acgt"[gmx_pac_base(ix.pac, (int64_t)b + i)]) : 0;
}
