/* gnumap_oracle.h -- CPU restatement of the GNUMAP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under oracle/ may be imported, linked or executed by the product (gnumap_b200/,
 * include/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle_*.py) against
 *   - the reference's own known-answer tests in src/bin_seq.cpp:1046-1277 (traceback strings,
 *     CIGARs, exact score, pair-HMM table), committed under tests/golden/;
 *   - outputs of the unmodified reference compiled here into oracle/_ref/ (function-level probe
 *     and whole-program SAM/SGR), fixtures committed under tests/golden/ with their generator.
 *
 * The struct types are the public ones of include/gmx.h so that tests can hand the same buffers
 * to the oracle and to the CUDA library.
 */
#ifndef GNUMAP_ORACLE_H
#define GNUMAP_ORACLE_H

#include "../include/gmx.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- read -> PWM ------------------------------------------------------------------------- */
void  orc_fastq_pwm(const uint8_t *seq, const uint8_t *qual, int n, int illumina, float *pwm);
void  orc_revcomp_pwm(const float *pwm, int n, float *out);
void  orc_revcomp_str(const uint8_t *s, int n, uint8_t *out);
char  orc_max_char(const float *row);

/* ---- alignment kernels ------------------------------------------------------------------- */
float orc_self_score(const float *pwm, const uint8_t *consensus, int n, const float S[256][4]);
float orc_nw_score(const float *pwm, int n, const uint8_t *gen, const float S[256][4],
                   float gap, int max_gap);
float orc_align_score_range(const float *pwm, int n, const uint8_t *gen, unsigned begin, unsigned end,
                            const float S[256][4], float gap, int max_gap);
int   orc_nw_traceback(const float *pwm, int n, const uint8_t *consense, const uint8_t *gen, int m,
                       const float S[256][4], float gap, int max_gap,
                       char *aligned_out, int aligned_cap, char *cigar_out, int cigar_cap);
void  orc_pair_hmm(const float *pwm, int n, const uint8_t *consensus, const uint8_t *gen, int m,
                   const float P[256][4], float *post_out);
void  orc_fix_cigar_for_deletions(char *cigar);

/* ---- FM index ---------------------------------------------------------------------------- */
uint64_t orc_bwt_occ(const gmx_index *ix, uint64_t k, int c);
void     orc_bwt_2occ(const gmx_index *ix, uint64_t k, uint64_t l, int c, uint64_t *ok, uint64_t *ol);
int      orc_match_exact(const gmx_index *ix, int len, const uint8_t *codes, uint64_t *k, uint64_t *l);
void     orc_get_sa_int(const gmx_index *ix, const uint8_t *ascii, int len, uint64_t *start, uint64_t *end);
uint64_t orc_bwt_sa(const gmx_index *ix, uint64_t k);
int      orc_get_string(const gmx_index *ix, uint64_t begin, int size, uint8_t *out);

/* ---- whole path --------------------------------------------------------------------------
 * PHASE A (+ PHASE B when do_score != 0) for a batch.  amount / planes are accumulated into
 * (caller zero-initialises).  hits: caller buffer of hits_cap entries; *n_hits receives the count
 * (function returns -5 if the buffer is too small).  cigar_out / aligned_out may be NULL. */
int orc_process_batch(const gmx_index *ix, const gmx_params *pr, const gmx_reads *reads, int do_score,
                      gmx_read_result *results, gmx_hit *hits, int64_t hits_cap, int64_t *n_hits,
                      char *cigar_out, int cigar_stride, uint8_t *aligned_out, int aligned_stride,
                      float *amount, float *const planes[5]);

void orc_default_params(gmx_params *p);

#ifdef __cplusplus
}
#endif
#endif
