// index_build.cuh -- SURVEY.md §8(f) rank 4: index construction on the GPU.
//
//   reference src/bwtindex.c:187-293   bwa_index: pac -> BWT (bwt_pac2bwt / is.c or bwt_gen) -> occ interleave
//                                      (bwt_bwtupdate_core :128-150) -> sampled suffix array (bwt_cal_sa, src/bwt.c:62-84)
//   reference src/bntseq.c:224-225     2-bit pac, four bases per byte, first base most significant
//
// The BWT of a text is unique, so any correct suffix sorter reproduces the reference's files bit for bit.  The sorter here
// is prefix doubling over CUB radix sorts: round 0 sorts every suffix by its first 16 symbols (3 bits each, 0 = past the
// end, so a suffix that is a proper prefix of another sorts first, as '$' makes it), every later round by the pair (rank of
// the suffix, rank of the suffix h symbols further) with h doubling; ranks come from a flag + scan over the sorted keys.
// A random 100 Mb genome is sorted after round 1 (4^16 >> 1e8); repeats cost one round per doubling of their length.
// Everything downstream is one streaming kernel each: BWT symbols, per-128-symbol counts + scan, the occ-interleaved
// layout, the 1/32 suffix-array sample, the pac.
#pragma once

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gmx_common.cuh"

#define GMX_IDX_OCC 128          // OCC_INTERVAL, reference inc/bwt.h
#define GMX_IDX_SA_INTV 32       // reference src/bwtindex.c:286

__global__ void __launch_bounds__(256) k_idx_key16(const uint8_t *codes, int64_t n, unsigned long long *key, uint32_t *idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) k = (k << 3) | (unsigned long long)(i + t < n ? codes[i + t] + 1 : 0);
    key[i] = k; idx[i] = (uint32_t)i;
}

// flag[r] = 1 where a new key starts in the sorted order
__global__ void __launch_bounds__(256) k_idx_flags(const unsigned long long *skey, int64_t n, uint32_t *flag)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    flag[r] = (r == 0 || skey[r] != skey[r - 1]) ? 1u : 0u;
}

// rank of suffix idx[r] = number of distinct keys up to and including its own (1-based)
__global__ void __launch_bounds__(256) k_idx_scatter_rank(const uint32_t *idx, const uint32_t *grp, int64_t n, uint32_t *rank)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) rank[idx[r]] = grp[r];
}

__global__ void __launch_bounds__(256) k_idx_pair_key(const uint32_t *rank, int64_t n, int64_t h, unsigned long long *key, uint32_t *idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = ((unsigned long long)rank[i] << 32) | (unsigned long long)(i + h < n ? rank[i + h] : 0u);
    idx[i] = (uint32_t)i;
}

// stored BWT symbol j (the '$' entry at `primary` removed, reference src/bwtindex.c bwt_pac2bwt): augmented rank R = j or j + 1
__device__ __forceinline__ uint32_t gmx_idx_bwt_sym(const uint8_t *codes, const uint32_t *sa, int64_t n, int64_t primary, int64_t j)
{
    const int64_t R = j < primary ? j : j + 1;
    if (R == 0) return codes[n - 1];
    const uint32_t p = sa[R - 1];
    return p ? codes[p - 1] : 0u;
}

// per block of 128 stored symbols: the four symbol counts
__global__ void __launch_bounds__(128) k_idx_block_counts(const uint8_t *codes, const uint32_t *sa, int64_t n, int64_t primary, int64_t n_blocks,
                                                          unsigned long long *c0, unsigned long long *c1, unsigned long long *c2, unsigned long long *c3)
{
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n_blocks) return;
    uint32_t cnt[4] = {0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int64_t j = b * GMX_IDX_OCC + t * 32 + lane;
        if (j < n) cnt[gmx_idx_bwt_sym(codes, sa, n, primary, j)]++;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        for (int o = 16; o; o >>= 1) cnt[c] += __shfl_xor_sync(0xffffffffu, cnt[c], o);
    if (lane == 0) { c0[b] = cnt[0]; c1[b] = cnt[1]; c2[b] = cnt[2]; c3[b] = cnt[3]; }
}

// the occ-interleaved layout (bwt_bwtupdate_core, reference src/bwtindex.c:128-150): per 128 symbols 4 x u64 running counts
// then 8 x u32 of sixteen 2-bit symbols, first symbol most significant; the last block holds only the words that exist;
// one more block of counts closes the array.  One thread per output symbol word.
__global__ void __launch_bounds__(256) k_idx_interleave(const uint8_t *codes, const uint32_t *sa, int64_t n, int64_t primary, int64_t n_blocks,
                                                        const unsigned long long *c0, const unsigned long long *c1, const unsigned long long *c2,
                                                        const unsigned long long *c3, uint32_t *out)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // symbol word index
    const int64_t n_words = (n + 15) >> 4;
    if (w < n_words) {
        uint32_t v = 0;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int64_t j = w * 16 + t;
            v = (v << 2) | (j < n ? gmx_idx_bwt_sym(codes, sa, n, primary, j) : 0u);
        }
        const int64_t b = w >> 3;
        out[b * 16 + 8 + (w & 7)] = v;
    }
    if (w <= n_blocks) {                                                   // counts in front of block w (w == n_blocks: the closing block)
        const int64_t at = w < n_blocks ? w * 16 : n_blocks * 8 + n_words;
        const unsigned long long c[4] = {c0[w], c1[w], c2[w], c3[w]};      // exclusive prefix sums, [n_blocks] = totals
#pragma unroll
        for (int k = 0; k < 4; ++k) { out[at + 2 * k] = (uint32_t)c[k]; out[at + 2 * k + 1] = (uint32_t)(c[k] >> 32); }
    }
}

// bwt_cal_sa with intv 32: sa_s[k] = SA of augmented rank 32 k; sa_s[0] = (bwtint_t)-1 (reference src/bwt.c:83)
__global__ void __launch_bounds__(256) k_idx_sample_sa(const uint32_t *sa, int64_t n_sa, unsigned long long *out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_sa) return;
    out[k] = k == 0 ? ~0ull : (unsigned long long)sa[k * GMX_IDX_SA_INTV - 1];
}

__global__ void __launch_bounds__(256) k_idx_pack_pac(const uint8_t *codes, int64_t n, uint8_t *pac, int64_t n_bytes)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bytes) return;
    uint32_t v = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) v = (v << 2) | (uint32_t)(4 * b + t < n ? codes[4 * b + t] : 0);
    pac[b] = (uint8_t)v;
}
