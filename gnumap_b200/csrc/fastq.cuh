// fastq.cuh -- SURVEY.md §8(f) rank 1: FASTQ text -> reads, the step immediately before the hot path.
//
// Reference: SeqReader::get_more_fastq (reference src/SeqReader.cpp:1023-1292).  Two implementations of the same
// record index (name / sequence / quality line of every read, as offsets into the text):
//   * gmx_fastq_scan_host  -- sequential C++ restatement with the reference's recovery from malformed records
//                             (blank lines, shifted '@' / '+' lines, quality shorter than the sequence);
//   * the device indexer   -- newline positions by stream compaction, one thread per 4-line record, for well-formed
//                             text; any record that would need the recovery path makes it return GMX_ERR_FORMAT and
//                             the caller falls back to the host scan.
// The reads are then used IN PLACE: DevReads.seq / .qual point into the text, with per-read offsets, quality offsets
// and lengths -- nothing is gathered or re-packed (the PWM of a FASTQ read is a function of (base, quality char)).
#pragma once

#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "gmx_common.cuh"

struct IsNewline {
    const char *text;
    __host__ __device__ bool operator()(uint32_t i) const { return text[i] == '\n'; }
};

// newlines of the text, 16 bytes per thread and step
__global__ void k_count_newlines(const char *text, int64_t len, unsigned long long *count)
{
    unsigned long long n = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i < len; i += stride) {
        if (i + 16 <= len && ((reinterpret_cast<uintptr_t>(text) + i) & 15) == 0) {
            const uint4 v = *reinterpret_cast<const uint4 *>(text + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t x = w[k] ^ 0x0a0a0a0au;                       // bytes equal to '\n' become 0
                n += __popc(((x - 0x01010101u) & ~x & 0x80808080u));
            }
        } else {
            for (int64_t j = i; j < len && j < i + 16; ++j) n += text[j] == '\n';
        }
    }
    for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, n);
}


__global__ void k_gather_f32(const float *v, const uint32_t *idx, uint32_t n, float *out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v[idx[i]];
}

struct FastqDev {
    int64_t *seq_off;      // [n]
    int64_t *qual_off;     // [n]
    int32_t *seq_len;      // [n]
    gmx_fastq_rec *recs;   // [n]
    uint32_t *flags;       // [0] = number of malformed records, [1] = max read length, [2] = index of the first malformed record
};

// line l of the text: [begin, end) without the '\n'
__device__ __forceinline__ void gmx_fastq_line(const uint32_t *nl, uint32_t n_nl, int64_t len, uint32_t l, int64_t &b, int64_t &e)
{
    b = l == 0 ? 0 : (int64_t)nl[l - 1] + 1;
    e = l < n_nl ? (int64_t)nl[l] : len;
}

// `base`: offset of `text` inside the whole text when a piece of it is indexed (all stored offsets are global)
__global__ void k_fastq_records(const char *text, int64_t len, const uint32_t *nl, uint32_t n_nl, uint32_t n_recs, int qmin, FastqDev out, int64_t base)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_recs) return;
    int64_t b0, e0, b1, e1, b2, e2, b3, e3;
    gmx_fastq_line(nl, n_nl, len, 4 * r, b0, e0);
    gmx_fastq_line(nl, n_nl, len, 4 * r + 1, b1, e1);
    gmx_fastq_line(nl, n_nl, len, 4 * r + 2, b2, e2);
    gmx_fastq_line(nl, n_nl, len, 4 * r + 3, b3, e3);
    bool ok = e0 > b0 && text[b0] == '@' && e2 > b2 && text[b2] == '+' && (e1 - b1) <= (e3 - b3);
    const int n = (int)(e1 - b1);
    if (ok) for (int i = 0; i < n; ++i) if ((int)(unsigned char)text[b3 + i] < qmin) { ok = false; break; }   // Q < 0: the reference throws
    if (!ok) { atomicAdd(&out.flags[0], 1u); atomicMin(&out.flags[2], r); }
    atomicMax(&out.flags[1], (uint32_t)n);
    out.seq_off[r] = base + b1; out.qual_off[r] = base + b3; out.seq_len[r] = n;
    gmx_fastq_rec rec;
    rec.name_off = base + b0 + 1; rec.seq_off = base + b1; rec.qual_off = base + b3;
    rec.name_len = (int32_t)(e0 - b0 - 1); rec.seq_len = n; rec.qual_len = (int32_t)(e3 - b3); rec.pad = 0;
    out.recs[r] = rec;
}

// ---- host restatement --------------------------------------------------------------------------------
struct FastqLines {            // std::getline over a memory buffer, with ifstream's eofbit behaviour
    const char *t; int64_t len, pos; bool eof;
    void getline(int64_t &off, int64_t &n)
    {
        if (eof) return;               // the stream is no longer good: std::getline leaves the string as it was
        off = pos; n = 0;
        const char *q = (const char *)memchr(t + pos, '\n', (size_t)(len - pos));
        if (q) { n = (q - t) - pos; pos = (q - t) + 1; }
        else { n = len - pos; pos = len; eof = true; }
    }
    char first(int64_t off, int64_t n) const { return n > 0 ? t[off] : '\0'; }
};
