// Accumulator output: the .gmp row -- device-side row selection and gather, host-side
// likelihood-ratio SNP call and text formatting.
//
//   reference src/GenomeBwt.cpp:930-1005   PrintFinalSNP        "chrom\tpos\t%.5f" + 5 x "\t%.5f" + call
//   reference src/GenomeBwt.cpp:1011-1092  PrintSNPCall         "\tN" | "\t[YN]:g->a p_val=%.2e" | "\t[YN]:g->a/b p_val=%.2e"
//   reference src/GenomeBwt.cpp:739-755    LRT                  monoploid likelihood ratio
//   reference src/GenomeBwt.cpp:760-873    dipLRT               monoploid-vs-diploid likelihood ratio
//   reference src/GenomeBwt.cpp:1094-1205  PrintFinalBisulfite  "chrom\tpos\t%f" + 5 x "\t%.5f"
//
// A full-genome scan: 24 B of accumulators per position, HBM-bound, so the scan and the compaction run on the
// device (one pass over `amount`, the genome base read from the 2-bit pac for the bisulfite filter); only the
// printable rows travel to the host, where the call statistics use the host libm (pow / log / lgamma / exp are the
// reference's own calls, so the p-values agree to the printed digit) and many threads format the text.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "gmx_common.cuh"

struct GmpRowSelect {            // DeviceSelect predicate over accumulator bins
    const float *amount;
    const uint8_t *pac;
    uint64_t gen_size, l_pac;
    double min_print;            // SNP rows: amount > min_print, compared in double as the reference's literal is
    int target_base;             // bisulfite / A->G rows: genome base that must stand at the position, amount > 0
    __device__ bool operator()(uint32_t bin) const
    {
        const float a = amount[bin];
        if (target_base < 0) return (double)a > min_print;
        const uint64_t count = (uint64_t)bin * gen_size;
        if (count >= l_pac) return false;
        return gmx_pac_base(pac, (int64_t)count) == target_base && a > 0.0f;
    }
};

struct SgrRowSelect {
    const float *amount;
    double min_print;
    __device__ bool operator()(uint32_t bin) const { return (double)amount[bin] > min_print; }
};

// rows[k] = {amount, plane A, C, G, T, N} of selected bin k (structure of arrays: [6][n]); base[k] = genome base code
__global__ void k_gmp_gather(const float *amount, const float *p0, const float *p1, const float *p2, const float *p3, const float *p4,
                             const uint8_t *pac, uint64_t gen_size, uint64_t l_pac, const uint32_t *idx, uint32_t n, float *rows, uint8_t *base)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t bin = idx[k];
    rows[k] = amount[bin];
    rows[(size_t)n + k] = p0[bin]; rows[2 * (size_t)n + k] = p1[bin]; rows[3 * (size_t)n + k] = p2[bin];
    rows[4 * (size_t)n + k] = p3[bin]; rows[5 * (size_t)n + k] = p4[bin];
    const uint64_t count = (uint64_t)bin * gen_size;
    base[k] = count < l_pac ? (uint8_t)gmx_pac_base(pac, (int64_t)count) : 4;
}

// ---- .sgr rows formatted on the device -----------------------------------------------------------------
// GenomeBwt::PrintFinalSGR (reference src/GenomeBwt.cpp:1212-1273): "chrom\tpos\t%.5f\n" per selected bin.  A float times
// 10^5 is exact in a double, so rint() (round-half-even) gives printf's digits; one thread per row sizes it, a scan places
// it, one thread per row writes it.  8 M rows = 196 MB of text that the host used to format on all its threads.
struct SgrNames { const char *chars; const int32_t *off; const int32_t *len; };

__device__ __forceinline__ int gmx_dev_digits(unsigned long long v) { int d = 1; while (v >= 10) { v /= 10; d++; } return d; }
__device__ __forceinline__ char *gmx_dev_put_uint(char *o, unsigned long long v)
{
    const int d = gmx_dev_digits(v);
    for (int i = d - 1; i >= 0; --i) { o[i] = (char)('0' + v % 10); v /= 10; }
    return o + d;
}

// row k: bin idx[k] with value val[k]; lens[k] = bytes of its line (0 when the bin starts past the genome)
__global__ void __launch_bounds__(256) k_sgr_measure(const uint32_t *idx, const float *val, uint32_t n, uint64_t gen_size, const int64_t *seq_offset, int n_seqs,
                                                     SgrNames names, long long *lens, uint32_t *uncovered)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t count = (int64_t)((uint64_t)idx[k] * gen_size);
    long long len = 0;
    if (count < seq_offset[n_seqs]) {
        int lo = 0, hi = n_seqs;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seq_offset[mid] <= count) lo = mid; else hi = mid; }
        const float v = val[k];
        if (!(v >= 0.0f && v < 1.0e9f)) atomicAdd(uncovered, 1u);          // the host's snprintf takes the whole file then
        const unsigned long long q = (unsigned long long)rint((double)v * 100000.0);
        len = names.len[lo] + 1 + gmx_dev_digits((unsigned long long)(count - seq_offset[lo] + 1)) + 1 + gmx_dev_digits(q / 100000ull) + 1 + 5 + 1;
    }
    lens[k] = len;
}

__global__ void __launch_bounds__(256) k_sgr_write(const uint32_t *idx, const float *val, uint32_t n, uint64_t gen_size, const int64_t *seq_offset, int n_seqs,
                                                   SgrNames names, const long long *offs, char *out)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (offs[k + 1] == offs[k]) return;
    const int64_t count = (int64_t)((uint64_t)idx[k] * gen_size);
    int lo = 0, hi = n_seqs;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seq_offset[mid] <= count) lo = mid; else hi = mid; }
    char *o = out + offs[k];
    const char *nm = names.chars + names.off[lo];
    for (int i = 0; i < names.len[lo]; ++i) *o++ = nm[i];
    *o++ = '\t';
    o = gmx_dev_put_uint(o, (unsigned long long)(count - seq_offset[lo] + 1));
    *o++ = '\t';
    const unsigned long long q = (unsigned long long)rint((double)val[k] * 100000.0);
    o = gmx_dev_put_uint(o, q / 100000ull);
    *o++ = '.';
    unsigned long long f = q % 100000ull;
    for (int d = 4; d >= 0; --d) { o[d] = (char)('0' + f % 10); f /= 10; }
    o += 5;
    *o++ = '\n';
}

// ---- host side ------------------------------------------------------------------------------------

// "%.{dec}f" of a non-negative float below 1e9, digit for digit as printf rounds it (ties to even on the exact
// value): a 24-bit significand times 10^dec (<= 10^6 = 2^6 * 15625) is exact in a double, so rint() sees the exact
// product.  Anything else goes through snprintf.
static inline char *gmx_put_fixed(char *o, float v, int dec)
{
    static const double p10[7] = {1., 10., 100., 1000., 10000., 100000., 1000000.};
    if (!(v >= 0.0f && v < 1.0e9f) || dec > 6) return o + sprintf(o, "%.*f", dec, (double)v);
    const uint64_t n = (uint64_t)std::rint((double)v * p10[dec]);
    const uint64_t scale = (uint64_t)p10[dec];
    uint64_t ip = n / scale, fp = n % scale;
    char tmp[24]; int t = 0;
    do { tmp[t++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
    while (t) *o++ = tmp[--t];
    if (dec) {
        *o++ = '.';
        for (int d = dec - 1; d >= 0; --d) { o[d] = (char)('0' + fp % 10); fp /= 10; }
        o += dec;
    }
    return o;
}

static inline char *gmx_put_int(char *o, long long v)
{
    if (v < 0) { *o++ = '-'; v = -v; }
    char tmp[24]; int t = 0;
    do { tmp[t++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (t) *o++ = tmp[--t];
    return o;
}

// chi-square CDF P(x; nu) = regularised lower incomplete gamma P(nu/2, x/2): power series below a + 1, Lentz's
// continued fraction for the upper tail above it.  The reference takes this one function from GSL
// (gsl_cdf_chisq_P, src/GenomeBwt.cpp:749,776,803,824); 1 - P cancels for strong calls, so p-values below ~1e-12
// carry few digits in either implementation.
static inline double gmx_chisq_cdf(double x, double nu)
{
    const double a = nu / 2.0, h = x / 2.0;
    if (!(h > 0.0)) return 0.0;
    const double lead = std::exp(-h + a * std::log(h) - std::lgamma(a));
    if (h < a + 1.0) {
        double term = 1.0 / a, sum = term, ap = a;
        for (int it = 0; it < 100000; ++it) {
            ap += 1.0;
            term *= h / ap;
            sum += term;
            if (std::fabs(term) < std::fabs(sum) * 1e-16) break;
        }
        return sum * lead;
    }
    const double tiny = 1e-300;
    double b = h + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, f = d;
    for (int i = 1; i < 100000; ++i) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
        c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double delta = d * c;
        f *= delta;
        if (std::fabs(delta - 1.0) < 1e-16) break;
    }
    return 1.0 - lead * f;
}

static inline int gmx_first_max5(const float *x)
{
    int m = 0;
    for (int i = 1; i < 5; ++i) if (x[i] > x[m]) m = i;     // first of the largest, as std::max_element
    return m;
}

struct GmpCall { int first, second; bool diploid; double pval; };

// one-allele likelihood ratio against the uniform model (reference LRT, src/GenomeBwt.cpp:739-755)
static inline double gmx_lr_single(const float *x, int m, double sum)
{
    return std::pow(.2, sum) / (std::pow(x[m] / sum, (double)x[m]) * std::pow((sum - x[m]) / sum / 4, sum - x[m]));
}

static inline GmpCall gmx_snp_call(const float counts[5], bool monoploid)
{
    GmpCall r{0, -1, false, 0.0};
    float x[5];
    for (int i = 0; i < 5; ++i) x[i] = counts[i];
    r.first = gmx_first_max5(x);
    if (monoploid) {                                            // LRT: the sum is a float sum there
        const double sum = x[0] + x[1] + x[2] + x[3] + x[4];
        r.pval = 1 - gmx_chisq_cdf(-2 * std::log(gmx_lr_single(x, r.first, sum)), 1);
        r.second = 0;                                           // never printed for monoploid calls
        return r;
    }
    // dipLRT, src/GenomeBwt.cpp:760-873
    const int a = r.first;
    double sum = 0;
    for (int i = 0; i < 5; ++i) sum += x[i];
    double ratio1 = gmx_lr_single(x, a, sum);
    double pval1 = 1 - gmx_chisq_cdf(-2 * std::log(ratio1), 1);
    float rest[5];
    for (int i = 0; i < 5; ++i) rest[i] = x[i];
    rest[a] = 0;
    int b = gmx_first_max5(rest);
    double pval2, ratio2;
    if (x[a] / x[b] > 3.0f || a == b) {                         // MONO_DIP_RATIO: too lopsided for two alleles
        pval2 = (double)0.01f; ratio2 = 0.0; b = -1;            // MAX_PVAL, MAX_RATIO are floats there
    } else {
        for (int i = 0; i < 5; ++i) { x[i] = (float)(x[i] + 0.2); sum += 0.2; }      // DIFF_UNIF_PRIOR in, rounded to float
        ratio1 = gmx_lr_single(x, a, sum);
        pval1 = 1 - gmx_chisq_cdf(-2 * std::log(ratio1), 1);
        const double num = std::pow(.2, sum);
        const double den = std::pow(x[a] / sum, (double)x[a]) * std::pow(x[b] / sum, (double)x[b])
                         * std::pow(((sum - x[a] + x[b]) / sum) / 3, sum - x[a] - x[b]);      // "+ x[b]" as in the reference
        ratio2 = num / den;
        pval2 = 1 - gmx_chisq_cdf(-2 * std::log(ratio2), 2);
        for (int i = 0; i < 5; ++i) { x[i] = (float)(x[i] - 0.2); sum -= 0.2; }      // ... and out again (x is now re-rounded)
    }
    r.second = b;
    // x[b] with b == -1 reads the element in front of the array in the reference; that branch's ratio test is only
    // reached with diploid impossible (pval2 = 0.01 > any pval1 that matters, ratio2 = 0), see below
    const bool balanced = b >= 0 && x[a] / x[b] < 3.0f;
    if (pval2 == 0 && pval1 == 0) {                             // both beyond the chi-square's reach: compare the ratios
        r.diploid = ratio2 < ratio1 && balanced;
        r.pval = 0;
        return r;
    }
    if (pval2 < pval1 && balanced) { r.diploid = true; r.pval = pval2; }
    else { r.diploid = false; r.pval = pval1; }
    return r;
}

// the call column of one .gmp row (PrintSNPCall); `o` has room for 48 bytes
static inline char *gmx_put_snp_call(char *o, const float counts[5], int genome_base, bool monoploid, float snp_pval)
{
    static const char letters[] = "acgtn";
    const GmpCall c = gmx_snp_call(counts, monoploid);
    if (c.first == genome_base && !c.diploid) { *o++ = '\t'; *o++ = 'N'; return o; }
    const char yn = c.pval < snp_pval ? 'Y' : 'N';
    const char g = letters[genome_base];
    if (c.diploid) return o + sprintf(o, "\t%c:%c->%c/%c p_val=%.2e", yn, g, letters[c.first], c.second >= 0 ? letters[c.second] : '?', c.pval);
    return o + sprintf(o, "\t%c:%c->%c p_val=%.2e", yn, g, letters[c.first], c.pval);
}
