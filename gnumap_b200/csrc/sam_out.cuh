// sam_out.cuh -- SURVEY.md §8(f) rank 2: SAM emission on the device.
//
//   reference inc/ScoredSeq.h:293-404   ScoredSeq::get_SAM: MAPQ, CIGAR, one record per (position, strand) of the best group
//   reference src/Driver.cpp:2166-2205  the SAM writer: name, flag, chromosome, position, MAPQ, CIGAR, "*\t0\t0", sequence,
//                                       quality, XA:f / XP:f (ostream default float format == "%g"), X0:i
//
// After gmx_process_fastq everything a record needs is resident on the GPU: the FASTQ text (names, sequences, qualities),
// the record index, the per-read results and the CIGARs of the best groups.  Three passes:
//   k_sam_measure   one thread per read: the read's small text pieces (flag, position, MAPQ, the CIGAR as printed, the
//                   "%g" numbers) into a per-read slot, and the byte length of its record
//   (cub scan)      record offsets
//   k_sam_write     one warp per read: copies name / pieces / sequence (reverse-complemented on the - strand) / quality
//                   to its place in the output -- 270 B per read written once, coalesced per piece
// Reads whose best group holds several positions (repeats) get their place reserved (k_sam_multi_len) and are written
// by the host formatter afterwards; so is the rare read whose numbers the exact 128-bit "%g" below does not cover.
#pragma once

#include "pipeline.cuh"

// "%g" (precision 6) of a finite double, digit for digit as printf rounds it (ties to even on the EXACT value): the
// significand times a power of ten is formed in 128-bit integers, so the six digits come from an exact quotient.
// Returns the length, or -1 when the value needs more than 128 bits (|v| below ~1e-17 or above ~1e21): the caller
// then leaves the record to the host's snprintf.
__host__ __device__ inline int gmx_fmt_g6(double v, char *out)
{
    typedef unsigned __int128 u128;
    unsigned long long bits;
    memcpy(&bits, &v, 8);
    int n = 0;
    if (bits >> 63) out[n++] = '-';
    bits &= 0x7fffffffffffffffull;
    if (bits == 0) { out[n++] = '0'; return n; }
    const int be = (int)(bits >> 52);
    if (be == 0x7ff) return -1;
    unsigned long long m = bits & 0xfffffffffffffull;
    int e2;
    if (be == 0) e2 = -1074; else { m |= 1ull << 52; e2 = be - 1075; }
    const double a = v < 0 ? -v : v;
    int X = (int)floor(log10(a));
    unsigned long long D = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int k = 5 - X;                                   // D = round(a * 10^k)
        u128 num = m, den = 1;
        if (k > 22 || k < -18) return -1;
        for (int i = 0; i < (k > 0 ? k : -k); ++i) { if (k > 0) num *= 10; else den *= 10; }
        if (e2 >= 0) { if (e2 > 60) return -1; num <<= e2; }
        else { if (-e2 > 120) return -1; if (den >> (127 + e2)) return -1; den <<= -e2; }
        u128 q = num / den, r = num - q * den;
        const u128 twice = r * 2;
        if (twice > den || (twice == den && (q & 1))) q += 1;
        if (q < 100000) { X -= 1; continue; }
        if (q >= 1000000) {
            // either the estimate of X was one too small, or rounding carried 999999.5 up to 1000000
            u128 q10 = num / (den * 10);
            if (q10 >= 100000) { X += 1; continue; }
            X += 1; D = 100000; break;
        }
        D = (unsigned long long)q;
        break;
    }
    if (D == 0) return -1;
    char dg[6];
    for (int i = 5; i >= 0; --i) { dg[i] = (char)('0' + D % 10); D /= 10; }
    int nd = 6;
    while (nd > 1 && dg[nd - 1] == '0') nd--;                  // %g strips trailing zeros
    if (X < -4 || X >= 6) {
        out[n++] = dg[0];
        if (nd > 1) { out[n++] = '.'; for (int i = 1; i < nd; ++i) out[n++] = dg[i]; }
        out[n++] = 'e';
        int ex = X;
        if (ex < 0) { out[n++] = '-'; ex = -ex; } else out[n++] = '+';
        if (ex >= 100) { out[n++] = (char)('0' + ex / 100); ex %= 100; }
        out[n++] = (char)('0' + ex / 10); out[n++] = (char)('0' + ex % 10);
    } else if (X >= 0) {
        for (int i = 0; i <= X; ++i) out[n++] = i < nd ? dg[i] : '0';
        if (nd > X + 1) { out[n++] = '.'; for (int i = X + 1; i < nd; ++i) out[n++] = dg[i]; }
    } else {
        out[n++] = '0'; out[n++] = '.';
        for (int i = 0; i < -X - 1; ++i) out[n++] = '0';
        for (int i = 0; i < nd; ++i) out[n++] = dg[i];
    }
    return n;
}

__host__ __device__ inline int gmx_fmt_uint(unsigned long long v, char *out)
{
    char t[24]; int k = 0;
    do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < k; ++i) out[i] = t[k - 1 - i];
    return k;
}

#define GMX_SAM_HOST 0xffffffffu     // SamPiece::len of a read the host formats (several positions, or an uncovered number)

struct SamPiece {                    // the small text pieces of one read's record
    uint32_t len;                    // bytes of the read's record(s); 0: prints nothing; GMX_SAM_HOST: left to the host
    int32_t chrom;                   // index of the chromosome name
    uint8_t head1_len, head2_len, tail_len, cigar_len;
    uint8_t neg, pad[3];
    char head1[8];                   // "\t0\t" | "\t16\t"
    char head2[20];                  // "\t<pos>\t<mapq>\t"
    char tail[60];                   // "\tXA:f:<g>\tXP:f:<g>\tX0:i:<n>\n"
};

struct SamNames {                    // chromosome names on the device
    const char *chars; const int32_t *off; const int32_t *len; int32_t n;
};

__device__ __forceinline__ int gmx_sam_mapq(double total)
{   // reference inc/ScoredSeq.h:302-309
    int q;
    if (total == 1) q = 30;
    else {
        const double vv = 1 - total;
        q = vv <= 0 ? 30 : (int)round(-10 * log(vv) / log(10.0));
    }
    return q > 30 ? 30 : q;
}

// reverse_CIGAR (reference inc/SequenceOperations.h:109-123; its digit test is 48..58): the runs in reverse order, digits
// that no operator follows are dropped.  Returns the length written.
__device__ inline int gmx_sam_reverse_cigar(const char *cg, int cl, char *out)
{
    int w = 0, end = cl;
    while (end > 0 && cg[end - 1] >= 48 && cg[end - 1] <= 58) end--;          // trailing digits without an operator
    while (end > 0) {
        int b = end - 1;                                                       // cg[b] is the run's operator
        while (b > 0 && cg[b - 1] >= 48 && cg[b - 1] <= 58) b--;
        for (int j = b; j < end; ++j) out[w++] = cg[j];
        end = b;
    }
    return w;
}

// one thread per read
__global__ void __launch_bounds__(128) k_sam_measure(const gmx_read_result *res, const gmx_fastq_rec *recs, const char *cigars, int cigar_stride,
                                                     int n_reads, DevIndex ix, SamNames names, double inv_adjust, SamPiece *pieces, char *cigar_out,
                                                     long long *lens, uint32_t *n_uncovered)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const gmx_read_result R = res[r];
    SamPiece P;
    P.len = 0; P.chrom = 0; P.head1_len = P.head2_len = P.tail_len = P.cigar_len = 0; P.neg = 0;
    long long total_len = 0;
    if (GMX_READ_PRINTS_SAM(R)) {
        // the numbers every record of the read ends with
        int t = 0;
        const char xa[] = "\tXA:f:", xp[] = "\tXP:f:", x0[] = "\tX0:i:";
        for (int i = 0; i < 6; ++i) P.tail[t++] = xa[i];
        const int a = gmx_fmt_g6((double)R.best_score * inv_adjust, P.tail + t);
        t += a > 0 ? a : 0;
        for (int i = 0; i < 6; ++i) P.tail[t++] = xp[i];
        const int b = gmx_fmt_g6((double)R.best_posterior, P.tail + t);
        t += b > 0 ? b : 0;
        for (int i = 0; i < 6; ++i) P.tail[t++] = x0[i];
        t += gmx_fmt_uint((unsigned long long)R.best_n_positions, P.tail + t);
        P.tail[t++] = '\n';
        P.tail_len = (uint8_t)t;
        if (a < 0 || b < 0) atomicAdd(n_uncovered, 1u);            // the whole batch then goes through the host formatter
        const gmx_fastq_rec rec = recs[r];
        const char *cg = cigars + (size_t)r * cigar_stride;
        int cl = 0;
        while (cl < cigar_stride && cg[cl]) cl++;
        const int mapq = gmx_sam_mapq(exp((double)R.best_score) / R.denominator);
        char mq[4]; const int mql = gmx_fmt_uint((unsigned long long)mapq, mq);
        if (R.best_n_positions != 1) {
            // several positions: the host writes these records; k_sam_multi_len adds one record's bytes per listed position.
            // Until then lens[r] holds minus the position-independent bytes of ONE record: name, mapq + tab, cigar,
            // "\t*\t0\t0\t", sequence, tab, quality, tail
            P.len = GMX_SAM_HOST;
            total_len = -((long long)rec.name_len + mql + 1 + cl + 7 + rec.seq_len + 1 + rec.qual_len + t);
        } else {
            const uint64_t pos = R.best_first_pos;
            int lo = 0, hi = ix.n_seqs;                            // last sequence whose offset <= pos
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((uint64_t)ix.seq_offset[mid] <= pos) lo = mid; else hi = mid; }
            P.chrom = lo;
            P.neg = (uint8_t)(R.best_first_strand == GMX_NEG_STRAND);
            int h1 = 0, h2 = 0;
            P.head1[h1++] = '\t';
            if (P.neg) { P.head1[h1++] = '1'; P.head1[h1++] = '6'; } else P.head1[h1++] = '0';
            P.head1[h1++] = '\t';
            P.head1_len = (uint8_t)h1;
            P.head2[h2++] = '\t';
            h2 += gmx_fmt_uint((unsigned long long)(pos - (uint64_t)ix.seq_offset[lo] + 1), P.head2 + h2);
            P.head2[h2++] = '\t';
            for (int i = 0; i < mql; ++i) P.head2[h2++] = mq[i];
            P.head2[h2++] = '\t';
            P.head2_len = (uint8_t)h2;
            char *co = cigar_out + (size_t)r * cigar_stride;       // the CIGAR as printed
            if (!P.neg) { for (int i = 0; i < cl; ++i) co[i] = cg[i]; }
            else cl = gmx_sam_reverse_cigar(cg, cl, co);
            P.cigar_len = (uint8_t)cl;
            total_len = (long long)rec.name_len + h1 + names.len[lo] + h2 + cl + 7 + rec.seq_len + 1 + rec.qual_len + t;
            P.len = (uint32_t)total_len;
        }
    }
    pieces[r] = P;
    lens[r] = total_len;
}

// per (position, strand) of a multi-position best group: the bytes of one more record of that read
__global__ void __launch_bounds__(256) k_sam_multi_len(const MultiPos *multi, uint32_t n_multi, DevIndex ix, SamNames names,
                                                       const long long *fixed_neg, unsigned long long *extra)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_multi) return;
    const MultiPos m = multi[k];
    const long long f = fixed_neg[m.read];
    if (f >= 0) return;                                            // not a host-formatted read (cannot happen for listed reads)
    int lo = 0, hi = ix.n_seqs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((uint64_t)ix.seq_offset[mid] <= m.pos) lo = mid; else hi = mid; }
    char tmp[24];
    const int pd = gmx_fmt_uint(m.pos - (uint64_t)ix.seq_offset[lo] + 1, tmp);
    // "\t0\t" | "\t16\t", chromosome, "\t<pos>\t" (+ the mapq and its tab are in the fixed part)
    const long long bytes = -f + (m.strand == GMX_NEG_STRAND ? 4 : 3) + names.len[lo] + 1 + pd + 1;
    atomicAdd(&extra[m.read], (unsigned long long)bytes);
}

// lens[r] < 0 (host-formatted reads) -> the summed size of their records
__global__ void __launch_bounds__(256) k_sam_fix_lens(long long *lens, const unsigned long long *extra, int n_reads)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads && lens[r] < 0) lens[r] = (long long)extra[r];
}

__device__ __forceinline__ char gmx_sam_rc(char c)
{   // reverse_comp, reference inc/SequenceOperations.h:56-96: anything that is not acgtACGT- becomes 'n'
    switch (c) {
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case '-': return '-'; default: return 'n';
    }
}

// one warp per read
__global__ void __launch_bounds__(256) k_sam_write(const char *text, const gmx_fastq_rec *recs, const SamPiece *pieces, const char *cigar_txt,
                                                   int cigar_stride, const long long *offs, int n_reads, SamNames names, char *out)
{
    const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const SamPiece *P = pieces + r;
    const uint32_t len = P->len;
    if (len == 0 || len == GMX_SAM_HOST) return;
    const gmx_fastq_rec rec = recs[r];
    char *o = out + offs[r];
    const bool neg = P->neg;
    auto copy = [&](const char *src, int n) { for (int i = lane; i < n; i += 32) o[i] = src[i]; o += n; };
    copy(text + rec.name_off, rec.name_len);
    copy(P->head1, P->head1_len);
    copy(names.chars + names.off[P->chrom], names.len[P->chrom]);
    copy(P->head2, P->head2_len);
    copy(cigar_txt + (size_t)r * cigar_stride, P->cigar_len);
    {
        const char mid[7] = {'\t', '*', '\t', '0', '\t', '0', '\t'};
        if (lane < 7) o[lane] = mid[lane];
        o += 7;
    }
    if (!neg) {
        copy(text + rec.seq_off, rec.seq_len);
        if (lane == 0) o[0] = '\t';
        o += 1;
        copy(text + rec.qual_off, rec.qual_len);
    } else {
        const char *s = text + rec.seq_off;
        for (int i = lane; i < rec.seq_len; i += 32) o[i] = gmx_sam_rc(s[rec.seq_len - 1 - i]);
        o += rec.seq_len;
        if (lane == 0) o[0] = '\t';
        o += 1;
        const char *q = text + rec.qual_off;
        for (int i = lane; i < rec.qual_len; i += 32) o[i] = q[rec.qual_len - 1 - i];
        o += rec.qual_len;
    }
    copy(P->tail, P->tail_len);
}
