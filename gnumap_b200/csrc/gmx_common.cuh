// gmx_common.cuh -- shared device-side types for the B200 (sm_100a) GNUMAP hot path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gmx.h"

#define GMX_NEG_INF (-100000.0f)          // reference inc/bin_seq.h:37
#define GMX_EMPTY_KEY 0xFFFFFFFFu
#define GMX_QMIN 33                        // lowest FASTQ quality char handled by the LUTs
#define GMX_NQ 94                          // quality chars 33..126
#define GMX_MAX_READ_LEN 1024
#define GMX_MAX_GAP 8                      // compile-time cap of the band half-width
#define GMX_KMER_TAB_MAX 12                // 4^12 x 8 B = 128 MB at most
#define GMX_MAX_SEEDS 255                  // seeds (rounds) per (read, strand): 8 bits of the sort key

// Device view of the index (kernel parameter, passed by value).
struct DevIndex {
    const uint32_t *bwt;        // occ-interleaved BWT, reference layout (64-byte blocks)
    const uint32_t *sa_full;    // de-sampled suffix array, [seq_len + 1], == bwt_sa(k) for k >= 1
    const uint64_t *sa_samp;    // the reference's sampled SA (validation path)
    const uint8_t  *pac;        // 2-bit packed genome
    const int64_t  *seq_offset; // [n_seqs + 1], last = l_pac
    const uint2    *kmer_tab;   // [4^tab_len] SA interval (k, l) of every tab_len-mer, k > l when absent: the first
                                // tab_len backward-search steps of bwt_match_exact, memoised at load
    int32_t  tab_len;           // min(mer, GMX_KMER_TAB_MAX)
    uint64_t primary, seq_len;
    uint64_t L2[5];
    int64_t  l_pac;
    int32_t  sa_intv, n_seqs;
};

// Scoring tables resident in global memory (read through L1; staged to shared where hot).
struct DevTables {
    const float *sub_pos;    // [5][GMX_NQ][4]  get_val(pwm(base,q), genome g) for the read as given
    const float *sub_neg;    // [5][GMX_NQ][4]  same for the reverse-complemented PWM row
    const float *pwm_lut;    // [5][GMX_NQ][4]  FASTQ -> PWM row
    const float *phmm_pos;   // [5][GMX_NQ][4]  p_seq(pwm(base,q), genome g) (pair HMM emission)
    const float *phmm_neg;   // [5][GMX_NQ][4]
    const float *S;          // [256][4] gALIGN_SCORES
    const float *P;          // [256][4] gPHMM_ALIGN_SCORES
    const float *self;       // [256][GMX_NQ]  get_val(pwm(nt4(ch), q), ch): one base's term of the read's self score
};

// Device view of one batch of reads.
struct DevReads {
    const int64_t *offsets;  // [n_reads + 1]
    const uint8_t *seq;      // raw ASCII
    const uint8_t *qual;     // raw ASCII (may be null when pwm is given)
    const float   *pwm;      // optional [total][4]
    const int64_t *qoffsets; // optional: quality string of read r starts at qual[qoffsets[r]] (FASTQ text used in place)
    const int32_t *lens;     // optional: explicit read lengths (reads not contiguous in `seq`)
    int32_t n_reads;
    int32_t qbase;           // 33, or 64 with --illumina
    int32_t max_len;         // > 0: the caller's bound for device-resident reads (checked by k_prep_reads); 0: lengths were
                             // validated on the host
};

__device__ __forceinline__ int gmx_read_len(const DevReads &R, int r)
{
    return R.lens ? R.lens[r] : (int)(R.offsets[r + 1] - R.offsets[r]);
}

__device__ __forceinline__ int gmx_nt4(uint8_t c)
{   // reference src/bntseq.c:47-64 nst_nt4_table, folded to 0..3 / 4
    switch (c) {
        case 'A': case 'a': return 0; case 'C': case 'c': return 1;
        case 'G': case 'g': return 2; case 'T': case 't': return 3;
        default: return 4;
    }
}

__device__ __forceinline__ int gmx_qidx(uint8_t q, int qbase)
{   // index into the [GMX_NQ] LUT axis: the LUT is built for Q = char - qbase, stored at char - 33
    int v = (int)q - GMX_QMIN;
    return v < 0 ? 0 : (v >= GMX_NQ ? GMX_NQ - 1 : v);
}

__device__ __forceinline__ float gmx_max3(float a, float b, float c)
{   // reference src/bin_seq.cpp:1013-1026 (value only; ties give the same value)
    return fmaxf(fmaxf(a, b), c);
}

__device__ __forceinline__ int gmx_pac_base(const uint8_t *pac, int64_t k)
{   // reference src/bntseq.c:225 _get_pac
    return (pac[k >> 2] >> ((~k & 3) << 1)) & 3;
}

// GenomeBwt::GetString validity (reference src/GenomeBwt.cpp:384-415, bns_intv2rid bntseq.c:365-373):
// the window [begin, begin+size) must lie inside one sequence and inside the genome.
__device__ __forceinline__ bool gmx_window_valid(const DevIndex &ix, uint64_t begin, int size)
{
    int64_t rb = (int64_t)begin, re = rb + size;
    if (re > ix.l_pac || rb >= ix.l_pac) return false;
    if (ix.n_seqs == 1) return true;
    int lo = 0, hi = ix.n_seqs;                    // last sequence whose offset <= rb
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ix.seq_offset[mid] <= rb) lo = mid; else hi = mid; }
    return re <= ix.seq_offset[lo + 1];
}
