/* ref_probe.cpp -- C-ABI shim around the UNMODIFIED reference objects.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiled by oracle/Makefile together with the reference's own sources (taken where they lie
 * under /root/reference, never copied) into oracle/_ref/libref_probe.so.  Nothing here restates an
 * algorithm: every function forwards to the reference's own classes so that tests can (a) pin the
 * C restatement in gnumap_oracle.c against the real thing and (b) generate the golden fixtures
 * committed under tests/golden/ (tests/golden/make_golden.py).
 *
 * Globals are defined exactly as the reference driver does it (src/Driver.cpp:43-44,963):
 * const_define.h provides the tunables, a_matrices.c the substitution tables.
 */
#include <pthread.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <set>

#include "const_include.h"
#include "const_define.h"
#include "GenomeBwt.h"
#include "bin_seq.h"
#include "ScoredSeq.h"
#include "NormalScoredSeq.h"
#include "BSScoredSeq.h"
#include "SNPScoredSeq.h"
#include "SequenceOperations.h"
#include "SeqReader.h"

const char *pos_matrix = NULL;      /* src/Driver.cpp:72, read by a_matrices.c */
#include "a_matrices.c"

static bool g_inited = false;
static GenomeBwt *g_gen = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static Read make_read(const float *pwm, int n, float **&rows)
{
    rows = new float *[n];
    for (int i = 0; i < n; ++i) {
        rows[i] = new float[4];
        for (int b = 0; b < 4; ++b) rows[i][b] = pwm[4 * i + b];
    }
    Read r(rows, n);
    r.name = 0;
    return r;
}
static void free_rows(float **rows, int n)
{
    for (int i = 0; i < n; ++i) delete[] rows[i];
    delete[] rows;
}

extern "C" {

void refp_init(void)
{
    if (g_inited) return;
    InitProg();                     /* const_define.h:127-164: conversion tables */
    gINT2BASE[5] = 'I'; gINT2BASE[6] = 'D';
    setup_alignment_matrices();     /* a_matrices.c:25 */
    gMER_SIZE = DEF_MER_SIZE;
    gJUMP_SIZE = gMER_SIZE / 2;
    gVERBOSE = 0;
    g_inited = true;
}

void refp_get_tables(float *align_scores, float *phmm_scores, float *scalars)
{
    memcpy(align_scores, gALIGN_SCORES, sizeof(gALIGN_SCORES));
    memcpy(phmm_scores, gPHMM_ALIGN_SCORES, sizeof(gPHMM_ALIGN_SCORES));
    scalars[0] = gGAP; scalars[1] = gMATCH; scalars[2] = gTRANSITION; scalars[3] = gTRANSVERSION; scalars[4] = gADJUST;
}

/* mode: 0 normal, 1 bisulfite (-b, both strands: src/Driver.cpp:1260-1268), 2 snp */
void refp_set_mode(int mode)
{
    gSNP = (mode == 2); gBISULFITE = (mode == 1);
    if (mode == 1) gALIGN_SCORES[(int)'c'][3] = gMATCH;
    if (mode != 0) gGEN_SIZE = 1;
}

int refp_bin_seq_test(void)
{
    unsigned int warnings = 0;
    bool ok = bin_seq::Test(std::cerr, warnings);
    return (ok ? 0 : 1) | (int)(warnings << 8);
}

float refp_self_score(const float *pwm, int n, const char *consensus)
{
    float **rows; Read r = make_read(pwm, n, rows);
    bin_seq bs; const Read &cr = r; const std::string cons(consensus, n);
    float v = bs.get_align_score(cr, cons, 0u, (unsigned)(n - 1));
    free_rows(rows, n);
    return v;
}

float refp_align_score_range(const float *pwm, int n, const char *gen, int glen, unsigned begin, unsigned end)
{
    float **rows; Read r = make_read(pwm, n, rows);
    bin_seq bs; const Read &cr = r; const std::string g(gen, glen);
    float v = bs.get_align_score(cr, g, begin, end);
    free_rows(rows, n);
    return v;
}

float refp_nw_score(const float *pwm, int n, const char *gen, int glen)
{
    float **rows; Read r = make_read(pwm, n, rows);
    bin_seq bs; const Read &cr = r; const std::string g(gen, glen);
    float v = bs.get_align_score(cr, g);
    free_rows(rows, n);
    return v;
}

int refp_nw_traceback(const float *pwm, int n, const char *consensus, const char *gen, int glen,
                      char *aligned_out, int aligned_cap, char *cigar_out, int cigar_cap)
{
    float **rows; Read r = make_read(pwm, n, rows);
    bin_seq bs; const Read &cr = r; const std::string g(gen, glen); const std::string cons(consensus, n);
    std::pair<std::string, std::string> res = bs.get_align_score_w_traceback(cr, cons, g);
    int len = (int)res.first.size();
    memset(aligned_out, 0, aligned_cap);
    memcpy(aligned_out, res.first.data(), len < aligned_cap ? len : aligned_cap);
    strncpy(cigar_out, res.second.c_str(), cigar_cap - 1); cigar_out[cigar_cap - 1] = 0;
    free_rows(rows, n);
    return len;
}

void refp_pair_hmm(const float *pwm, int n, const char *consensus, const char *gen, int glen, float *out)
{
    float **rows; Read r = make_read(pwm, n, rows);
    bin_seq bs; const Read &cr = r; const std::string g(gen, glen); const std::string cons(consensus, n);
    float **h = bs.pairHMM(cr, cons, g);
    for (int i = 0; i < n; ++i) { for (int b = 0; b < 5; ++b) out[5 * i + b] = h[i][b]; delete[] h[i]; }
    delete[] h;
    free_rows(rows, n);
}

/* ---- genome-backed probes -------------------------------------------------------------- */
int refp_load_genome(const char *fasta)
{
    refp_init();
    if (g_gen) { delete g_gen; g_gen = 0; }
    g_gen = new GenomeBwt();
    g_gen->use(fasta);
    g_gen->LoadGenome();
    return 0;
}

void refp_get_sa_int(const char *kmer, int len, uint64_t *start, uint64_t *end)
{
    std::string s(kmer, len);
    g_gen->get_sa_int(s, start, end);
}

uint64_t refp_get_sa_coord(uint64_t k) { return g_gen->get_sa_coord(k); }

int refp_get_string(uint64_t begin, unsigned size, char *out)
{
    std::string s = g_gen->GetString(begin, size);
    memcpy(out, s.data(), s.size());
    return (int)s.size();
}

/* One ScoredSeq::score() call on zeroed accumulators; returns aligned extent touched.
 * kind: 0 Normal, 1 BS, 2 SNP.  positions: n_pos x (pos, strand). */
int refp_score_once(int kind, const float *pwm, int n, const char *gen_string, double align_score,
                    const uint64_t *pos, const int *strand, int n_pos, double denom,
                    float *amount_out, uint64_t amount_cap, float *planes_out /*5 x cap or NULL*/)
{
    float **rows; Read r = make_read(pwm, n, rows);
    std::string g(gen_string, n);
    ScoredSeq *ss;
    if (kind == 2) ss = new SNPScoredSeq(g, align_score, pos[0], strand[0]);
    else if (kind == 1) ss = new BSScoredSeq(g, align_score, pos[0], strand[0]);
    else ss = new NormalScoredSeq(g, align_score, pos[0], strand[0]);
    for (int i = 1; i < n_pos; ++i) ss->add_spot(pos[i], strand[i]);
    uint64_t n_amt = g_gen->size() / gGEN_SIZE;
    float *amt = g_gen->GetGenomeAmtPtr();
    memset(amt, 0, sizeof(float) * n_amt);
    float *pl[5] = {0, 0, 0, 0, 0};
    if (kind != 0) {
        pl[0] = g_gen->GetGenomeAPtr(); pl[1] = g_gen->GetGenomeCPtr(); pl[2] = g_gen->GetGenomeGPtr();
        pl[3] = g_gen->GetGenomeTPtr(); pl[4] = g_gen->GetGenomeNPtr();
        for (int b = 0; b < 5; ++b) if (pl[b]) memset(pl[b], 0, sizeof(float) * g_gen->size());
    }
    ss->score(denom, *g_gen, 1, r, g_lock);
    uint64_t c = n_amt < amount_cap ? n_amt : amount_cap;
    memcpy(amount_out, amt, sizeof(float) * c);
    if (planes_out && kind != 0)
        for (int b = 0; b < 5; ++b) if (pl[b]) memcpy(planes_out + (uint64_t)b * amount_cap, pl[b], sizeof(float) * c);
    delete ss;
    free_rows(rows, n);
    return 0;
}

/* The call column GenomeBwt::PrintSNPCall (src/GenomeBwt.cpp:1011-1092) prints for the read counts `counts`
 * (A,C,G,T,N) placed at genome position `count`.  Needs refp_set_mode(2) before refp_load_genome (the five
 * read planes only exist in SNP / bisulfite mode). */
#ifdef GMX_REAL_GSL
/* Built against the reference's own GSL 1.9 (oracle/Makefile GSL=...): its default error handler aborts the program, e.g. in
 * gsl_cdf_chisq_P(inf, 1) when a likelihood ratio underflows to 0 (src/GenomeBwt.cpp:749).  The probe counts such calls
 * instead, so that the fixture generator can leave those inputs out: the reference has no output for them. */
#include <gsl/gsl_errno.h>
static int g_gsl_errors = 0;
static void refp_gsl_handler(const char *, const char *, int, int) { g_gsl_errors++; }
extern "C" int refp_gsl_errors(void) { static bool on = false; if (!on) { gsl_set_error_handler(&refp_gsl_handler); on = true; } int n = g_gsl_errors; g_gsl_errors = 0; return n; }
#else
extern "C" int refp_gsl_errors(void) { return -1; }          /* stand-in chi-square (oracle/gsl_stub): no GSL error handling to observe */
#endif

int refp_snp_call(uint64_t count, const float *counts, int monop, float pval, char *out, int cap)
{
    if (!g_gen || !g_gen->GetGenomeAPtr()) return -1;
    float *planes[5] = {g_gen->GetGenomeAPtr(), g_gen->GetGenomeCPtr(), g_gen->GetGenomeGPtr(), g_gen->GetGenomeTPtr(), g_gen->GetGenomeNPtr()};
    float total = 0;
    for (int b = 0; b < 5; ++b) { planes[b][count] = counts[b]; total += counts[b]; }
    g_gen->GetGenomeAmtPtr()[count] = total;
    gSNP_MONOP = monop != 0; gSNP_PVAL = pval;
    char *buf = 0; size_t len = 0;
    FILE *f = open_memstream(&buf, &len);
    g_gen->PrintSNPCall(count, f);
    fclose(f);
    int n = (int)len < cap - 1 ? (int)len : cap - 1;
    memcpy(out, buf, n); out[n] = 0;
    free(buf);
    return n;
}


/* SeqReader on a FASTQ file: one line "name\tseq\tfq\n" per Read, in order (reference src/SeqReader.cpp:1023-1292).
 * Returns the number of reads, or -1 when the reader threw (invalid quality character). */
int refp_read_fastq(const char *fn, char *out, int cap)
{
    refp_init();
    int n = 0, used = 0;
    try {
        SeqReader sr;
        sr.use(fn);
        Read *r;
        while ((r = sr.GetNextSequence()) != 0) {
            std::string line = std::string(r->name ? r->name : "") + "\t" + r->seq + "\t" + r->fq + "\n";
            if (used + (int)line.size() < cap) { memcpy(out + used, line.data(), line.size()); used += (int)line.size(); }
            ++n;
        }
    } catch (...) {
        return -1;
    }
    out[used] = 0;
    return n;
}

} /* extern "C" */
