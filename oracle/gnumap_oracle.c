/* gnumap_oracle.c -- plain-C CPU restatement of the GNUMAP 4.0 hot path.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path, never the thing shipped or measured
 * (see gnumap_oracle.h for who may call it and how parity is pinned).
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * checkout).  Arithmetic types, operation order and tie-breaks follow the reference exactly;
 * build with -ffp-contract=off (the reference objects contain no FMA).
 */
#include "gnumap_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#define ORC_NEG_INF (-100000.0f)      /* inc/bin_seq.h:37 */
#define ORC_SAME_DIFF 0.00001          /* inc/const_include.h:188 */

/* ------------------------------------------------------------------------------------------
 * defaults: inc/const_define.h:46-107 + setup_alignment_matrices() inc/a_matrices.c:25-126
 * ---------------------------------------------------------------------------------------- */
static void fill_table(float T[256][4], float match, float transition, float transversion)
{
    static const char *lo = "acgt", *up = "ACGT";
    for (int i = 0; i < 256; ++i)
        for (int j = 0; j < 4; ++j) T[i][j] = transversion;
    for (int g = 0; g < 4; ++g)
        for (int b = 0; b < 4; ++b) {
            float v = (g == b) ? match : ((g ^ b) == 2 ? transition : transversion);
            T[(int)lo[g]][b] = T[(int)up[g]][b] = v;   /* a<->g and c<->t are transitions */
        }
}

void orc_default_params(gmx_params *p)
{
    memset(p, 0, sizeof(*p));
    float adjust = 0.25f, match = 3, transition = -2, transversion = -3, gap = -4;
    match *= adjust; transition *= adjust; transversion *= adjust; gap *= adjust;  /* a_matrices.c:59-62 */
    fill_table(p->align_scores, match, transition, transversion);
    fill_table(p->phmm_scores, 0.98f, 0.01f, 0.005f);                              /* const_define.h:79-81 */
    p->gap = gap; p->max_gap = 3; p->mer = 10; p->jump = 5; p->min_seed_hits = 2;
    p->max_kmer_hits = 0; p->max_matches = 1000; p->gen_size = 8;
    p->align_score = 0.9f; p->perc = 1; p->cutoff = 0.0f;
    p->match_pos = 1; p->match_neg = 1; p->unique_only = 0; p->fast = 0; p->use_nw = 1;
    p->mode = GMX_MODE_NORMAL; p->illumina = 0; p->adjust = adjust;
}

/* ------------------------------------------------------------------------------------------
 * FASTQ -> PWM            src/SeqReader.cpp:618-627 (Q2Prb_*), :1155-1216, :1259-1272
 * ---------------------------------------------------------------------------------------- */
static double q2prb_std(double Q) { double a = 1 - exp((-Q / 10.0) * log(10.0)); return a > 1.0 ? 1.0 : a; }
static double q2prb_ill(double Q) { double a = 1.0 - 1.0 / (pow(10.0, (Q / 10.0))); return a > 1.0 ? 1.0 : a; }

void orc_fastq_pwm(const uint8_t *seq, const uint8_t *qual, int n, int illumina, float *pwm)
{
    for (int i = 0; i < n; ++i) {
        int Q = (int)qual[i];
        double max_prb;
        if (illumina) { Q -= 64; max_prb = q2prb_ill((double)Q); }
        else          { Q -= 33; max_prb = q2prb_std((double)Q); }
        double other = (1 - max_prb) / 3;
        double v[4] = {other, other, other, other};
        switch (tolower(seq[i])) {
            case 'a': v[0] = max_prb; break;
            case 'c': v[1] = max_prb; break;
            case 'g': v[2] = max_prb; break;
            case 't': v[3] = max_prb; break;
            default: break;                      /* 'n' and anything else: all equal */
        }
        for (int b = 0; b < 4; ++b) pwm[4 * i + b] = (float)v[b];
    }
}

/* inc/SequenceOperations.h:149-161 reverse_comp_cpy */
void orc_revcomp_pwm(const float *pwm, int n, float *out)
{
    for (int i = 0; i < n; ++i) {
        float *d = out + 4 * (n - 1 - i);
        d[0] = pwm[4 * i + 3]; d[1] = pwm[4 * i + 2]; d[2] = pwm[4 * i + 1]; d[3] = pwm[4 * i + 0];
    }
}

/* inc/SequenceOperations.h:56-96 reverse_comp(string&) */
void orc_revcomp_str(const uint8_t *s, int n, uint8_t *out)
{
    for (int i = 1; i <= n; ++i) {
        uint8_t c;
        switch (s[n - i]) {
            case 'a': c = 't'; break; case 'A': c = 'T'; break;
            case 't': c = 'a'; break; case 'T': c = 'A'; break;
            case 'c': c = 'g'; break; case 'C': c = 'G'; break;
            case 'g': c = 'c'; break; case 'G': c = 'C'; break;
            case '-': c = '-'; break;
            default:  c = 'n'; break;
        }
        out[i - 1] = c;
    }
}

/* inc/ScoredSeq.h:71-103 max_char */
char orc_max_char(const float *chr)
{
    if ((chr[0] == chr[1]) && (chr[0] == chr[2]) && (chr[0] == chr[3])) return 'n';
    if (chr[0] >= chr[1]) {
        if (chr[0] >= chr[2]) { if (chr[0] >= chr[3]) return 'a'; else return 't'; }
        else                  { if (chr[2] >= chr[3]) return 'g'; else return 't'; }
    } else {
        if (chr[1] >= chr[2]) { if (chr[1] >= chr[3]) return 'c'; else return 't'; }
        else                  { if (chr[2] >= chr[3]) return 'g'; else return 't'; }
    }
}

static int gen_conversion(uint8_t c)   /* g_gen_CONVERSION, src/Driver.cpp:912-923 */
{
    switch (c) {
        case 'a': case 'A': return 0; case 'c': case 'C': return 1;
        case 'g': case 'G': return 2; case 't': case 'T': return 3;
        case 'n': case 'N': return 4;
        case 10: case 11: case 12: case 13: return 5;
        case 0: return 6; case '>': return 7;
        default: return 4;
    }
}

/* ------------------------------------------------------------------------------------------
 * alignment kernels
 * ---------------------------------------------------------------------------------------- */

/* src/bin_seq.cpp:975-987 get_val */
static float get_val(const float *a, uint8_t b, const float S[256][4])
{
    const float *s = S[b];
    float score = (a[0] * s[0] + a[1] * s[1] + a[2] * s[2] + a[3] * s[3]);
    return score;
}

/* src/bin_seq.cpp:1013-1026 max_flt(a,b,c) */
static float max3(float a, float b, float c)
{
    if (a >= b) { if (a >= c) return a; else return c; }
    else        { if (b >= c) return b; else return c; }
}

/* src/bin_seq.cpp:989-1011 max_flt(path, diag, upgap, leftgap) */
static float max3_path(char *path, float diag, float upgap, float leftgap)
{
    if (diag >= upgap) {
        if (diag >= leftgap) { *path = 'D'; return diag; }
        else                 { *path = 'L'; return leftgap; }
    } else {
        if (upgap >= leftgap) { *path = 'U'; return upgap; }
        else                  { *path = 'L'; return leftgap; }
    }
}

/* src/bin_seq.cpp:739-759 get_align_score(read, consensus, 0, n-1): begin(…,0)=0, end(…,n-1)=0,
 * so the value is get_align_score_mid (:860-893) added into a float that starts at 0. */
float orc_self_score(const float *pwm, const uint8_t *consense, int n, const float S[256][4])
{
    float value = 0;
    value += 0.0f;
    float score = 0;
    for (int i = 0; i <= n - 1; ++i) {
        const float *p = pwm + 4 * i;
        const float *s = S[consense[i]];
        score += (p[0] * s[0]) + (p[1] * s[1]) + (p[2] * s[2]) + (p[3] * s[3]);
    }
    value += score;
    value += 0.0f;
    return value;
}

/* src/bin_seq.cpp:781-850 get_align_score_begin(read, gen, end).
 * A fresh bin_seq has def_arr filled with NEG_INF once it has grown (check_pointer_length :720-734);
 * every cell the recurrence reads is (re)initialised below, as in :791-807. */
static float nw_score_begin(const float *pwm, const uint8_t *gen, unsigned end, const float S[256][4], float gGAP, int G)
{
    if (end == 0) return 0;
    unsigned size = end + 1;
    float *nm = (float *)malloc(sizeof(float) * (size_t)size * size);
    for (size_t q = 0; q < (size_t)size * size; ++q) nm[q] = ORC_NEG_INF;

    for (unsigned i = 0; i <= end; i++)
        for (int j = (int)i - G - 1; j <= (int)i + G + 1 && j <= (int)end; j++) {
            if (j < 0) continue;
            nm[i * size + j] = ORC_NEG_INF;
        }
    for (int i = (int)end; i > (int)end - G - 2 && i >= 0; i--) nm[i * size + end] = gGAP * (end - i);
    for (int j = (int)end; j > (int)end - G - 2 && j >= 0; j--) nm[end * size + j] = gGAP * (end - j);

    for (int i = (int)end - 1; i >= 0; i--) {
        for (int j = i + G; j >= i - G; j--) {
            if (j >= (int)end) continue;
            if (j < 0) break;
            float m_mm1 = nm[(i + 1) * size + (j + 1)];
            float m_mm2 = get_val(pwm + 4 * i, gen[j], S);
            float m_mm = m_mm1 + m_mm2;
            float gap1 = nm[(i + 1) * size + j] + gGAP;
            float gap2 = nm[i * size + (j + 1)] + gGAP;
            nm[i * size + j] = max3(m_mm, gap1, gap2);
        }
    }
    float r = nm[0];
    free(nm);
    return r;
}

/* src/bin_seq.cpp:761-767 get_align_score(read, gen) == get_align_score_begin(read, gen, read.length) */
float orc_nw_score(const float *pwm, int n, const uint8_t *gen, const float S[256][4], float gGAP, int G)
{
    return nw_score_begin(pwm, gen, (unsigned)n, S, gGAP, G);
}

/* src/bin_seq.cpp:907-971 get_align_score_end(read, gen, start) */
static float nw_score_end(const float *pwm, int n, const uint8_t *gen, unsigned start, const float S[256][4], float gGAP, int G)
{
    if (start == (unsigned)n - 1) return 0;
    unsigned length = (unsigned)n - start;
    float *nm = (float *)malloc(sizeof(float) * (size_t)length * length);
    for (size_t q = 0; q < (size_t)length * length; ++q) nm[q] = ORC_NEG_INF;
    for (unsigned i = 0; i < length; i++)
        for (int j = (int)i - G - 1; j <= (int)i + G + 1 && j < (int)length; j++) {
            if (j < 0) continue;
            nm[i * length + j] = ORC_NEG_INF;
        }
    for (unsigned i = 0; i <= (unsigned)G + 1 && i < length; i++) nm[i * length + 0] = gGAP * i;
    for (unsigned j = 0; j <= (unsigned)G + 1 && j < length; j++) nm[0 * length + j] = gGAP * j;
    for (int i = 1; i < (int)length; i++)
        for (int j = i - G; j <= i + G && j < (int)length; j++) {
            if (j <= 0) continue;
            if (j >= (int)length) break;
            float m_mm1 = nm[(i - 1) * length + (j - 1)];
            float m_mm2 = get_val(pwm + 4 * (i + start), gen[j + start], S);
            float m_mm = m_mm1 + m_mm2;
            float gap1 = nm[i * length + (j - 1)] + gGAP;
            float gap2 = nm[(i - 1) * length + j] + gGAP;
            nm[i * length + j] = max3(m_mm, gap1, gap2);
        }
    float r = nm[(length - 1) * length + (length - 1)];
    free(nm);
    return r;
}

/* src/bin_seq.cpp:739-759 get_align_score(read, gen, begin, end): begin + mid + end pieces.
 * Only exercised by the reference's own known-answer test (src/bin_seq.cpp:1119-1127) and, with
 * (0, n-1), by the self score. */
float orc_align_score_range(const float *pwm, int n, const uint8_t *gen, unsigned begin, unsigned end,
                            const float S[256][4], float gGAP, int G)
{
    float value = 0;
    value += nw_score_begin(pwm, gen, begin, S, gGAP, G);
    float score = 0;
    for (unsigned i = begin; i <= end; ++i) {
        const float *p = pwm + 4 * i;
        const float *s = S[gen[i]];
        score += (p[0] * s[0]) + (p[1] * s[1]) + (p[2] * s[2]) + (p[3] * s[3]);
    }
    value += score;
    value += nw_score_end(pwm, n, gen, end, S, gGAP, G);
    return value;
}

/* prepend "<count><op>" to a CIGAR under construction (src/bin_seq.cpp:592-594 etc.) */
static void cigar_prepend(char *cigar, int cap, int count, int type)
{
    char temp[1024];
    strncpy(temp, cigar, sizeof(temp) - 1); temp[sizeof(temp) - 1] = 0;
    char op = (type == 0) ? 'M' : (type == 1) ? 'I' : 'D';    /* CIGAR_TO_STR inc/bin_seq.h:43 */
    snprintf(cigar, cap, "%d%c%s", count, op, temp);
}

/* src/bin_seq.cpp:445-718 get_align_score_w_traceback.  `consense` must be readable at index n
 * (the reference reads the std::string terminator there, :607/:660).  Returns strlen(aligned). */
int orc_nw_traceback(const float *pwm, int n, const uint8_t *consense, const uint8_t *gen, int m,
                     const float S[256][4], float gGAP, int G,
                     char *aligned_out, int aligned_cap, char *cigar_out, int cigar_cap)
{
    unsigned row_size = (unsigned)n + 1;
    size_t squareLen = (size_t)(n + 1) * (m + 1);
    float *nm = (float *)malloc(sizeof(float) * squareLen);
    char *moves = (char *)malloc(squareLen);
    for (size_t q = 0; q < squareLen; ++q) { nm[q] = ORC_NEG_INF; moves[q] = '-'; }   /* :465-483 */

    for (int i = 0; i <= n; i++)                                                       /* :493-500 */
        for (int j = i - G - 1; j <= i + G + 1; j++) {
            if (j < 0 || j > m) continue;
            nm[i * row_size + j] = ORC_NEG_INF; moves[i * row_size + j] = 'D';
        }
    for (int i = 0; i <= G + 2 && i <= n; i++) { nm[i * row_size] = gGAP * i; moves[i * row_size] = 'U'; }
    for (int j = 0; j <= G + 2 && j <= m; j++) { nm[j] = gGAP * j; moves[j] = 'L'; }
    moves[0] = 'D';

    for (int i = 1; i <= n; i++)                                                       /* :515-536 */
        for (int j = i - G; j <= i + G; j++) {
            if (j <= 0) continue;
            if (j > m) break;
            float m_mm1 = nm[(i - 1) * row_size + (j - 1)];
            float m_mm2 = get_val(pwm + 4 * (i - 1), gen[j - 1], S);
            float m_mm = m_mm1 + m_mm2;
            float gap1 = nm[(i - 1) * row_size + j] + gGAP;
            float gap2 = nm[i * row_size + (j - 1)] + gGAP;
            nm[i * row_size + j] = max3_path(&moves[i * row_size + j], m_mm, gap1, gap2);
        }

    char *aligned = (char *)malloc((size_t)n + m + 8);
    int alen = 0;
    char CIGAR[1024]; CIGAR[0] = 0;
    int i = n, j = m;
    int c_type = 0 /*MATCH*/, c_counter = 0;
    int err = 0;
    while ((i != 0) && (j != 0) && !err) {                                             /* :578-650 */
        switch (moves[i * row_size + j]) {
            case 'D':
                aligned[alen++] = (char)consense[i - 1];
                if (c_type == 0) c_counter++;
                else { if (c_counter) cigar_prepend(CIGAR, sizeof(CIGAR), c_counter, c_type); c_type = 0; c_counter = 1; }
                i -= 1; j -= 1; break;
            case 'U':
                aligned[alen++] = (char)consense[i];         /* sic: the reference's off-by-one */
                if (c_type == 1) c_counter++;
                else { if (c_counter) cigar_prepend(CIGAR, sizeof(CIGAR), c_counter, c_type); c_type = 1; c_counter = 1; }
                i -= 1; break;
            case 'L':
                aligned[alen++] = '-';
                if (c_type == 2) c_counter++;
                else { if (c_counter) cigar_prepend(CIGAR, sizeof(CIGAR), c_counter, c_type); c_type = 2; c_counter = 1; }
                j -= 1; break;
            default: err = 1; break;
        }
    }
    if (err) { free(nm); free(moves); free(aligned); if (aligned_cap) aligned_out[0] = 0; if (cigar_cap) cigar_out[0] = 0; return 0; }
    while (i > 0) {                                                                     /* :657-672 */
        aligned[alen++] = (char)consense[i];
        if (c_type == 1) c_counter++;
        else { cigar_prepend(CIGAR, sizeof(CIGAR), c_counter, c_type); c_type = 1; c_counter = 1; }
        i--;
    }
    while (j > 0) {                                                                     /* :675-690 */
        aligned[alen++] = '-';
        if (c_type == 2) c_counter++;
        else { cigar_prepend(CIGAR, sizeof(CIGAR), c_counter, c_type); c_type = 2; c_counter = 1; }
        j--;
    }
    if (c_counter > 0) cigar_prepend(CIGAR, sizeof(CIGAR), c_counter, c_type);          /* :693-698 */

    /* std::string::operator+= of '\0' keeps the NUL as a real character, so size() counts it. */
    for (int k = 0; k < alen && k < aligned_cap; ++k) aligned_out[k] = aligned[alen - 1 - k];  /* :701-704 */
    if (alen < aligned_cap) aligned_out[alen] = 0;
    if (cigar_cap) { strncpy(cigar_out, CIGAR, (size_t)cigar_cap - 1); cigar_out[cigar_cap - 1] = 0; }
    free(nm); free(moves); free(aligned);
    return alen;
}

/* inc/SequenceOperations.h:32-42 */
void orc_fix_cigar_for_deletions(char *cigar)
{
    int len = (int)strlen(cigar);
    if (len == 0) return;
    if (cigar[len - 1] == 'D') {
        int i;
        for (i = len - 2; i >= 0; i--) if (!isdigit((unsigned char)cigar[i])) break;
        cigar[i + 1] = 0;
    }
}

/* src/bin_seq.cpp:41-57 p_seq with pam_p :36-39 */
static float p_seq(const float *x, uint8_t y, const float P[256][4])
{
    float sum = 0;
    int yy = tolower(y);
    sum += x[0] * P[yy][0];
    sum += x[1] * P[yy][1];
    sum += x[2] * P[yy][2];
    sum += x[3] * P[yy][3];
    return 3 * sum;
}

/* src/bin_seq.cpp:60-244 pairHMM.  post_out is float[m][5] (A,C,G,T,N), zero-initialised here. */
void orc_pair_hmm(const float *pwm, int n, const uint8_t *consensus, const uint8_t *genome, int m,
                  const float P[256][4], float *post_out)
{
    /* inc/bin_seq.h:60-69 -- stored as float, promoted in the recurrences */
    float PHMM_q = 0.25f, PHMM_t = 0.05f, PHMM_d = 0.0025f, PHMM_e = 0.5f;
    float PHMM_Tmm = 1 - 2 * PHMM_d - PHMM_t;
    float PHMM_Tgm = 1 - PHMM_d - PHMM_t;
    float PHMM_Tmg = PHMM_d;
    float PHMM_Tgg = PHMM_e;

    int i, j;
    int col_size = (m + 1) * 3;
    int p_col_size = m * 3;
    size_t max_size = (size_t)(n + 1) * (m + 1) * 3;
    double *fMXY = (double *)calloc(max_size, sizeof(double));
    double *bMXY = (double *)calloc(max_size, sizeof(double));
    double *pMXY = (double *)calloc(max_size, sizeof(double));
    for (i = 0; i < m * 5; ++i) post_out[i] = 0;

    fMXY[0] = 1;
    for (i = 1; i < n + 1; i++)
        for (j = 1; j < m + 1; j++) {
            fMXY[i * col_size + 3 * j] = p_seq(pwm + 4 * (i - 1), genome[j - 1], P) *
                (PHMM_Tmm * fMXY[(i - 1) * col_size + 3 * (j - 1)]
                 + PHMM_Tgm * fMXY[(i - 1) * col_size + 3 * (j - 1) + 1]
                 + PHMM_Tgm * fMXY[(i - 1) * col_size + 3 * (j - 1) + 2]);
            fMXY[i * col_size + 3 * j + 1] = PHMM_q * (PHMM_Tmg * fMXY[(i - 1) * col_size + 3 * j] + PHMM_Tgg * fMXY[(i - 1) * col_size + 3 * j + 1]);
            fMXY[i * col_size + 3 * j + 2] = PHMM_q * (PHMM_Tmg * fMXY[i * col_size + 3 * (j - 1)] + PHMM_Tgg * fMXY[i * col_size + 3 * (j - 1) + 2]);
        }
    double fE = PHMM_t * (fMXY[n * col_size + 3 * m] + fMXY[n * col_size + 3 * m + 1] + fMXY[n * col_size + 3 * m + 2]);

    bMXY[(n - 1) * col_size + 3 * (m - 1)] = bMXY[(n - 1) * col_size + 3 * (m - 1) + 1] = bMXY[(n - 1) * col_size + 3 * (m - 1) + 2] = PHMM_t;
    for (i = n - 1; i >= 0; i--)
        for (j = m - 1; j >= 0; j--) {
            if (j == (m - 1) && i == (n - 1)) continue;
            if (j == m - 1) {
                bMXY[i * col_size + 3 * j] = PHMM_q * PHMM_Tmg * bMXY[(i + 1) * col_size + 3 * j + 1];
                bMXY[i * col_size + 3 * j + 1] = PHMM_q * PHMM_Tgg * bMXY[(i + 1) * col_size + 3 * j + 1];
                bMXY[i * col_size + 3 * j + 2] = 0;
                continue;
            }
            if (i == n - 1) {
                bMXY[i * col_size + 3 * j] = PHMM_q * PHMM_Tmg * bMXY[i * col_size + 3 * (j + 1) + 2];
                bMXY[i * col_size + 3 * j + 2] = PHMM_q * PHMM_Tgg * bMXY[i * col_size + 3 * (j + 1) + 2];
                bMXY[i * col_size + 3 * j + 1] = 0;
                continue;
            }
            bMXY[i * col_size + 3 * j] = p_seq(pwm + 4 * (i + 1), genome[j + 1], P) * PHMM_Tmm * bMXY[(i + 1) * col_size + 3 * (j + 1)]
                + PHMM_q * PHMM_Tmg * bMXY[(i + 1) * col_size + 3 * j + 1] + PHMM_q * PHMM_Tmg * bMXY[i * col_size + 3 * (j + 1) + 2];
            bMXY[i * col_size + 3 * j + 1] = p_seq(pwm + 4 * (i + 1), genome[j + 1], P) * PHMM_Tgm * bMXY[(i + 1) * col_size + 3 * (j + 1)]
                + PHMM_q * PHMM_Tgg * bMXY[(i + 1) * col_size + 3 * j + 1];
            bMXY[i * col_size + 3 * j + 2] = p_seq(pwm + 4 * (i + 1), genome[j + 1], P) * PHMM_Tgm * bMXY[(i + 1) * col_size + 3 * (j + 1)]
                + PHMM_q * PHMM_Tgg * bMXY[i * col_size + 3 * (j + 1) + 2];
        }

    for (i = 1; i < n + 1; i++)
        for (j = 1; j < m + 1; j++) {
            pMXY[(i - 1) * p_col_size + 3 * (j - 1)]     = fMXY[i * col_size + 3 * j]     * bMXY[(i - 1) * col_size + 3 * (j - 1)]     / fE;
            pMXY[(i - 1) * p_col_size + 3 * (j - 1) + 1] = fMXY[i * col_size + 3 * j + 1] * bMXY[(i - 1) * col_size + 3 * (j - 1) + 1] / fE;
            pMXY[(i - 1) * p_col_size + 3 * (j - 1) + 2] = fMXY[i * col_size + 3 * j + 2] * bMXY[(i - 1) * col_size + 3 * (j - 1) + 2] / fE;
        }

    for (i = 0; i < m; i++)
        for (j = 0; j < n; j++) {
            int code = gen_conversion(consensus[j]);
            if (code > 4) continue;     /* cannot happen for a max_char consensus */
            post_out[5 * i + code] += pMXY[j * p_col_size + i * 3 + 2] + pMXY[j * p_col_size + i * 3];
        }
    free(fMXY); free(bMXY); free(pMXY);
}

/* ------------------------------------------------------------------------------------------
 * FM index                       src/bwt.c (vendored BWA), inc/bwt.h
 * ---------------------------------------------------------------------------------------- */
#define OCC_INTV_SHIFT 7
#define OCC_INTERVAL   (1LL << OCC_INTV_SHIFT)
#define OCC_INTV_MASK  (OCC_INTERVAL - 1)

static inline const uint32_t *occ_intv(const gmx_index *ix, uint64_t k) { return ix->bwt + ((k >> 7) << 4); }   /* bwt.h:73 */
static inline int bwt_B0(const gmx_index *ix, uint64_t k)                                                        /* bwt.h:72,78 */
{
    uint32_t w = ix->bwt[((k >> 7) << 4) + sizeof(uint64_t) + ((k & 0x7f) >> 4)];
    return (int)(w >> ((~k & 0xf) << 1) & 3);
}

/* src/bwt.c:98-105 __occ_aux */
static inline int occ_aux(uint64_t y, int c)
{
    y = ((c & 2) ? y : ~y) >> 1 & ((c & 1) ? y : ~y) & 0x5555555555555555ull;
    y = (y & 0x3333333333333333ull) + (y >> 2 & 0x3333333333333333ull);
    return (int)(((y + (y >> 4)) & 0xf0f0f0f0f0f0f0full) * 0x101010101010101ull >> 56);
}

/* src/bwt.c:107-130 bwt_occ */
uint64_t orc_bwt_occ(const gmx_index *ix, uint64_t k, int c)
{
    if (k == ix->seq_len) return ix->L2[c + 1] - ix->L2[c];
    if (k == (uint64_t)(-1)) return 0;
    k -= (k >= ix->primary);
    const uint32_t *p = occ_intv(ix, k);
    uint64_t n; memcpy(&n, (const uint64_t *)p + c, sizeof(n));
    p += sizeof(uint64_t);
    const uint32_t *end = p + (((k >> 5) - ((k & ~OCC_INTV_MASK) >> 5)) << 1);
    for (; p < end; p += 2) n += occ_aux((uint64_t)p[0] << 32 | p[1], c);
    n += occ_aux(((uint64_t)p[0] << 32 | p[1]) & ~((1ull << ((~k & 31) << 1)) - 1), c);
    if (c == 0) n -= ~k & 31;
    return n;
}

/* src/bwt.c:132-163 bwt_2occ -- the fast path computes the same two values as two bwt_occ calls */
void orc_bwt_2occ(const gmx_index *ix, uint64_t k, uint64_t l, int c, uint64_t *ok, uint64_t *ol)
{
    *ok = orc_bwt_occ(ix, k, c);
    *ol = orc_bwt_occ(ix, l, c);
}

/* src/bwt.c:222-239 bwt_match_exact */
int orc_match_exact(const gmx_index *ix, int len, const uint8_t *str, uint64_t *sa_begin, uint64_t *sa_end)
{
    uint64_t k = 0, l = ix->seq_len, ok, ol;
    for (int i = len - 1; i >= 0; --i) {
        uint8_t c = str[i];
        if (c > 3) return 0;
        orc_bwt_2occ(ix, k - 1, l, c, &ok, &ol);
        k = ix->L2[c] + ok + 1;
        l = ix->L2[c] + ol;
        if (k > l) break;
    }
    if (k > l) return 0;
    *sa_begin = k; *sa_end = l;
    return (int)(l - k + 1);
}

/* src/bntseq.c:47-64 nst_nt4_table restricted to what get_sa_int needs */
static inline uint8_t nt4(uint8_t c)
{
    switch (c) {
        case 'A': case 'a': return 0; case 'C': case 'c': return 1;
        case 'G': case 'g': return 2; case 'T': case 't': return 3;
        case '-': return 5; default: return 4;
    }
}

/* src/GenomeBwt.cpp:438-474 get_sa_int */
void orc_get_sa_int(const gmx_index *ix, const uint8_t *seq, int len, uint64_t *in_start, uint64_t *in_end)
{
    uint8_t c_seq[64];
    if (len > 64) { *in_start = *in_end = 0; return; }
    for (int i = 0; i < len; ++i) c_seq[i] = seq[i] < 4 ? seq[i] : nt4(seq[i]);
    uint64_t s = 0, e = 0;
    int r = orc_match_exact(ix, len, c_seq, &s, &e);
    if (r > 0) { *in_start = s; *in_end = e; } else { *in_start = 0; *in_end = 0; }
}

/* src/bwt.c:53-59 bwt_invPsi */
static inline uint64_t inv_psi(const gmx_index *ix, uint64_t k)
{
    uint64_t x = k - (k > ix->primary);
    x = (uint64_t)bwt_B0(ix, x);
    x = ix->L2[x] + orc_bwt_occ(ix, k, (int)x);
    return k == ix->primary ? 0 : x;
}

/* src/bwt.c:86-97 bwt_sa */
uint64_t orc_bwt_sa(const gmx_index *ix, uint64_t k)
{
    uint64_t sa = 0, mask = (uint64_t)ix->sa_intv - 1;
    while (k & mask) { ++sa; k = inv_psi(ix, k); }
    return sa + ix->sa[k / (uint64_t)ix->sa_intv];
}

/* src/bntseq.c:349-363 bns_pos2rid */
static int pos2rid(const gmx_index *ix, int64_t pos_f)
{
    int left, mid, right;
    if (pos_f >= ix->l_pac) return -1;
    left = 0; mid = 0; right = ix->n_seqs;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= ix->seq_offset[mid]) {
            if (mid == ix->n_seqs - 1) break;
            if (pos_f < ix->seq_offset[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

/* src/GenomeBwt.cpp:384-415 GetString (+ bns_intv2rid src/bntseq.c:365-373, bns_get_seq :398-419).
 * Forward-only index: begin < l_pac always.  Returns the string length (size, or 0 = ""). */
int orc_get_string(const gmx_index *ix, uint64_t begin, int size, uint8_t *out)
{
    int64_t rb = (int64_t)begin, re = (int64_t)begin + size;
    if (rb < ix->l_pac && re > ix->l_pac) return 0;                    /* -2: past the end */
    if (rb >= ix->l_pac) return 0;                                     /* never produced by the path */
    int rid_b = pos2rid(ix, rb);
    int rid_e = rb < re ? pos2rid(ix, re - 1) : rid_b;
    if (rid_b != rid_e || rid_b < 0) return 0;
    for (int64_t k = rb; k < re; ++k)
        out[k - rb] = (uint8_t)"acgt"[ix->pac[k >> 2] >> ((~k & 3) << 1) & 3];
    return size;
}

/* ------------------------------------------------------------------------------------------
 * whole path
 * ---------------------------------------------------------------------------------------- */

typedef struct { uint64_t pos; int strand; } orc_spot;

typedef struct {
    uint8_t *key;          /* read-orientation genome string: the std::map key (align_seq2_raw.cpp:125-130) */
    uint8_t *sequence;     /* ScoredSeq::sequence: genome-orientation string of the first hit          */
    double   align_score;  /* ScoredSeq::align_score                                                   */
    double   log_align_score; /* exp(align_score)  (ScoredSeq.h:133; the name is the reference's)      */
    int      first_strand;
    orc_spot *spots; int n_spots, cap_spots;   /* std::set<pair<pos,strand>> kept sorted               */
} orc_group;

typedef struct { orc_group *g; int n, cap; } orc_unique;   /* kept sorted by key: std::map<string,ScoredSeq*> */

typedef struct { uint64_t *keys; int *vals; int cap, n; } orc_locs;   /* map<unsigned long,int> possible_locs */

static void locs_init(orc_locs *m, int cap) { m->cap = cap; m->n = 0; m->keys = (uint64_t *)malloc(sizeof(uint64_t) * cap); m->vals = (int *)malloc(sizeof(int) * cap); for (int i = 0; i < cap; ++i) m->keys[i] = ~0ull; }
static void locs_free(orc_locs *m) { free(m->keys); free(m->vals); }
static int *locs_slot(orc_locs *m, uint64_t key);
static void locs_grow(orc_locs *m)
{
    orc_locs b; locs_init(&b, m->cap * 2);
    for (int i = 0; i < m->cap; ++i) if (m->keys[i] != ~0ull) { *locs_slot(&b, m->keys[i]) = m->vals[i]; }
    locs_free(m); *m = b;
}
static int *locs_slot(orc_locs *m, uint64_t key)     /* operator[]: creates the entry with value 0 */
{
    if (m->n * 2 >= m->cap) locs_grow(m);
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    int i = (int)(h >> 32) & (m->cap - 1);
    while (m->keys[i] != ~0ull && m->keys[i] != key) i = (i + 1) & (m->cap - 1);
    if (m->keys[i] == ~0ull) { m->keys[i] = key; m->vals[i] = 0; m->n++; }
    return &m->vals[i];
}

static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }

static int unique_find(const orc_unique *u, const uint8_t *key, int n, int *insert_at)
{
    int lo = 0, hi = u->n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        int c = memcmp(u->g[mid].key, key, (size_t)n);
        if (c == 0) { *insert_at = mid; return 1; }
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    *insert_at = lo; return 0;
}

static int group_add_spot(orc_group *g, uint64_t pos, int strand)   /* ScoredSeq::add_spot, ScoredSeq.h:237-251 */
{
    int lo = 0, hi = g->n_spots;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        orc_spot *s = &g->spots[mid];
        if (s->pos == pos && s->strand == strand) return 0;
        if (s->pos < pos || (s->pos == pos && s->strand < strand)) lo = mid + 1; else hi = mid;
    }
    if (g->n_spots == g->cap_spots) { g->cap_spots = g->cap_spots ? g->cap_spots * 2 : 4; g->spots = (orc_spot *)realloc(g->spots, sizeof(orc_spot) * g->cap_spots); }
    memmove(g->spots + lo + 1, g->spots + lo, sizeof(orc_spot) * (g->n_spots - lo));
    g->spots[lo].pos = pos; g->spots[lo].strand = strand; g->n_spots++;
    return 1;
}

static void unique_clear(orc_unique *u)
{
    for (int i = 0; i < u->n; ++i) { free(u->g[i].key); free(u->g[i].sequence); free(u->g[i].spots); }
    u->n = 0;
}

typedef struct {
    const gmx_index *ix; const gmx_params *pr;
    int n;                     /* read length */
    double min_align_score, denominator, top_align_score;
    int n_nw;
    orc_unique unique;
} orc_read_state;

/* inc/align_seq2_raw.cpp:22-178 process_hits, restricted to gNW == true.  `pending` holds, in
 * ascending order, exactly the map entries the reference loop would act on in this round: those
 * whose count is >= gMIN_JUMP_MATCHES and not yet -1 (entries whose window is "" are revisited by
 * the reference every round with no effect, :45-51, so visiting them once is equivalent). */
static int process_hits(orc_read_state *st, orc_locs *locs, uint64_t *pending, int n_pending,
                        const float *pwm, int strand)
{
    const gmx_params *pr = st->pr;
    int n = st->n;
    uint8_t *to_match = (uint8_t *)malloc((size_t)n + 1), *key = (uint8_t *)malloc((size_t)n + 1);
    int goon = 1;
    for (int q = 0; q < n_pending && goon; ++q) {
        uint64_t pos = pending[q];
        int *cnt = locs_slot(locs, pos);
        if (*cnt < pr->min_seed_hits) continue;
        if (*cnt == -1) continue;
        if (orc_get_string(st->ix, pos, n, to_match) == 0) continue;
        to_match[n] = 0;
        double align_score = orc_nw_score(pwm, n, to_match, pr->align_scores, pr->gap, pr->max_gap);
        st->n_nw++;
        *cnt = -1;
        if (align_score > st->top_align_score) st->top_align_score = align_score;
        if (align_score >= st->min_align_score) {
            if (strand == GMX_NEG_STRAND) orc_revcomp_str(to_match, n, key); else memcpy(key, to_match, (size_t)n);
            int at;
            if (!unique_find(&st->unique, key, n, &at)) {
                orc_unique *u = &st->unique;
                if (u->n == u->cap) { u->cap = u->cap ? u->cap * 2 : 4; u->g = (orc_group *)realloc(u->g, sizeof(orc_group) * u->cap); }
                memmove(u->g + at + 1, u->g + at, sizeof(orc_group) * (u->n - at));
                orc_group *g = &u->g[at]; memset(g, 0, sizeof(*g));
                g->key = (uint8_t *)malloc((size_t)n); memcpy(g->key, key, (size_t)n);
                g->sequence = (uint8_t *)malloc((size_t)n + 1); memcpy(g->sequence, to_match, (size_t)n + 1);
                g->align_score = align_score; g->log_align_score = exp(align_score);
                g->first_strand = strand;
                group_add_spot(g, pos, strand);
                u->n++;
                st->denominator += exp(align_score);
            } else {
                if (pr->unique_only) { goon = 0; break; }
                if (group_add_spot(&st->unique.g[at], pos, strand)) st->denominator += exp(align_score);
            }
        }
    }
    free(to_match); free(key);
    return goon;
}

/* inc/align_seq2_raw.cpp:180-328 align_sequence, gNW == true */
static int align_sequence(orc_read_state *st, const float *pwm, const uint8_t *consensus, int strand)
{
    const gmx_params *pr = st->pr;
    unsigned i, j, last_kmer_pos = (unsigned)st->n - (unsigned)pr->mer;
    orc_locs locs; locs_init(&locs, 1024);
    uint64_t *pending = NULL; int cap_pending = 0;
    int ret = 1;
    for (i = 0; i < last_kmer_pos; i += (unsigned)pr->jump) {
        uint64_t start = 0, end = 0;
        for (j = 0; j + i < last_kmer_pos; j++) {
            orc_get_sa_int(st->ix, consensus + i + j, pr->mer, &start, &end);
            if (end == 0 && start == 0) continue;
            else if (pr->max_kmer_hits > 0 && end - start + 1 > pr->max_kmer_hits) continue;
            else break;
        }
        i += j;
        if (end == 0 && start == 0) break;
        if (pr->max_kmer_hits > 0 && end - start + 1 > pr->max_kmer_hits) break;

        int n_pending = 0;
        for (unsigned vit = (unsigned)start; vit <= end; vit++) {
            uint64_t sa = orc_bwt_sa(st->ix, vit);
            uint64_t beginning = (sa <= i) ? 0 : (sa - i);
            int *c = locs_slot(&locs, beginning);
            if (*c != -1) {
                (*c)++;
                if (*c == pr->min_seed_hits) {           /* first round in which the loop at :28-35 acts on it */
                    if (n_pending == cap_pending) { cap_pending = cap_pending ? cap_pending * 2 : 64; pending = (uint64_t *)realloc(pending, sizeof(uint64_t) * cap_pending); }
                    pending[n_pending++] = beginning;
                }
            }
            if (vit == 0xffffffffu) break;
        }
        qsort(pending, (size_t)n_pending, sizeof(uint64_t), cmp_u64);
        int goon = process_hits(st, &locs, pending, n_pending, pwm, strand);
        if (!goon) { ret = 0; break; }
        if ((unsigned)st->unique.n > pr->max_matches) { ret = 0; break; }
        if (pr->fast) break;
    }
    free(pending); locs_free(&locs);
    return ret;
}

static void add_score(const gmx_index *ix, const gmx_params *pr, float *amount, uint64_t pos, float amt)
{   /* src/GenomeBwt.cpp:483-490; the bounds check neutralises the reference's overrun */
    if ((int64_t)pos >= ix->l_pac) return;
    amount[pos / pr->gen_size] += amt;
}

/* {Normal,BS,SNP}ScoredSeq::score  src/NormalScoredSeq.cpp:24-76, BSScoredSeq.cpp:24-88, SNPScoredSeq.cpp:25-108 */
static void score_group(const gmx_index *ix, const gmx_params *pr, const orc_group *g, double denom,
                        const float *pwm, int n, float *amount, float *const planes[5])
{
    if (!g->n_spots) return;
    double total_score = g->log_align_score / denom;
    float *opwm = (float *)malloc(sizeof(float) * 4 * (size_t)n);
    if (g->first_strand == GMX_NEG_STRAND) orc_revcomp_pwm(pwm, n, opwm); else memcpy(opwm, pwm, sizeof(float) * 4 * (size_t)n);
    uint8_t *cons = (uint8_t *)malloc((size_t)n + 1);
    for (int i = 0; i < n; ++i) cons[i] = (uint8_t)orc_max_char(opwm + 4 * i);
    cons[n] = 0;

    if (pr->mode == GMX_MODE_SNP) {
        float *hmm = (float *)malloc(sizeof(float) * 5 * (size_t)n), *rev = (float *)malloc(sizeof(float) * 5 * (size_t)n);
        orc_pair_hmm(opwm, n, cons, g->sequence, n, pr->phmm_scores, hmm);
        for (int i = 0; i < n; ++i) {        /* reverse_comp_cpy_phmm, SequenceOperations.h:164-181 */
            float *d = rev + 5 * (n - 1 - i), *s = hmm + 5 * i;
            d[0] = s[3]; d[1] = s[2]; d[2] = s[1]; d[3] = s[0]; d[4] = s[4];
        }
        for (int s = 0; s < g->n_spots; ++s) {
            const float *h = (g->spots[s].strand == g->first_strand) ? hmm : rev;
            for (int i = 0; i < n; ++i) {
                uint64_t p = g->spots[s].pos + (uint64_t)i;
                if ((int64_t)p >= ix->l_pac) continue;
                add_score(ix, pr, amount, p, (float)total_score);
                float scale = (float)total_score;                 /* Genome.cpp:441-502 */
                uint64_t loc = p / pr->gen_size;
                for (int b = 0; b < 5; ++b) planes[b][loc] += h[5 * i + b] * scale;
            }
        }
        free(hmm); free(rev);
    } else {
        char *aligned = (char *)malloc(2 * (size_t)n + 16), *rc = (char *)malloc(2 * (size_t)n + 16);
        char cigar[1024];
        int alen = orc_nw_traceback(opwm, n, cons, g->sequence, n, pr->align_scores, pr->gap, pr->max_gap,
                                    aligned, 2 * n + 16, cigar, sizeof(cigar));
        orc_revcomp_str((const uint8_t *)aligned, alen, (uint8_t *)rc);
        for (int s = 0; s < g->n_spots; ++s) {
            const char *a = aligned;
            if (pr->mode == GMX_MODE_BS && g->spots[s].strand != g->first_strand) a = rc;
            for (int i = 0; i < alen; ++i) {
                uint64_t p = g->spots[s].pos + (uint64_t)i;
                if ((int64_t)p >= ix->l_pac) continue;
                add_score(ix, pr, amount, p, (float)total_score);
                if (pr->mode == GMX_MODE_BS) {                    /* Genome.cpp:507-554 */
                    int which = gen_conversion((uint8_t)a[i]);
                    if (which < 5) planes[which][p / pr->gen_size] += (float)total_score;
                }
            }
        }
        free(aligned); free(rc);
    }
    free(opwm); free(cons);
}

int orc_process_batch(const gmx_index *ix, const gmx_params *pr, const gmx_reads *reads, int do_score,
                      gmx_read_result *results, gmx_hit *hits, int64_t hits_cap, int64_t *n_hits_out,
                      char *cigar_out, int cigar_stride, uint8_t *aligned_out, int aligned_stride,
                      float *amount, float *const planes[5])
{
    if (!pr->use_nw) return GMX_ERR_UNSUPPORTED;
    int64_t n_hits = 0;
    int rc = GMX_OK;
    for (int r = 0; r < reads->n_reads; ++r) {
        gmx_read_result *res = &results[r];
        memset(res, 0, sizeof(*res));
        res->best_group = -1;
        res->hit_begin = res->hit_end = (int32_t)n_hits;
        if (cigar_out) cigar_out[(size_t)r * cigar_stride] = 0;
        if (aligned_out) aligned_out[(size_t)r * aligned_stride] = 0;
        int n = (int)(reads->offsets[r + 1] - reads->offsets[r]);
        const uint8_t *seq = reads->seq + reads->offsets[r];
        const uint8_t *qual = reads->qual ? reads->qual + reads->offsets[r] : NULL;

        /* src/Driver.cpp:446-456 */
        if ((unsigned)n < (unsigned)pr->mer) { res->status = GMX_READ_TOO_SHORT; res->top_score = -2; continue; }

        float *pwm = (float *)malloc(sizeof(float) * 4 * (size_t)n), *rpwm = (float *)malloc(sizeof(float) * 4 * (size_t)n);
        if (reads->pwm) memcpy(pwm, reads->pwm + 4 * reads->offsets[r], sizeof(float) * 4 * (size_t)n);
        else orc_fastq_pwm(seq, qual, n, pr->illumina, pwm);
        uint8_t *cons = (uint8_t *)malloc((size_t)n + 1), *rcons = (uint8_t *)malloc((size_t)n + 1);
        memcpy(cons, seq, (size_t)n); cons[n] = 0;        /* GetConsensus: read.seq, Driver.cpp:352-356 */

        orc_read_state st; memset(&st, 0, sizeof(st));
        st.ix = ix; st.pr = pr; st.n = n;

        double max_align_score = orc_self_score(pwm, cons, n, pr->align_scores);      /* :466 */
        res->max_align_score = (float)max_align_score;
        if (max_align_score < pr->cutoff) {                                              /* :469-480 */
            res->status = GMX_READ_TOO_POOR; res->top_score = -3;
            free(pwm); free(rpwm); free(cons); free(rcons); continue;
        }
        st.min_align_score = pr->perc ? pr->align_score * max_align_score : pr->align_score;   /* :490-497 */

        int too_many = 0;
        if (pr->match_pos) if (!align_sequence(&st, pwm, cons, GMX_POS_STRAND)) too_many = 1;  /* :506-526 */
        if (!too_many && pr->match_neg) {                                                        /* :532-587 */
            orc_revcomp_str(cons, n, rcons); rcons[n] = 0;
            orc_revcomp_pwm(pwm, n, rpwm);
            if (!align_sequence(&st, rpwm, rcons, GMX_NEG_STRAND)) too_many = 1;
        }
        res->n_candidates = st.n_nw;
        if (too_many) {
            res->status = GMX_READ_TOO_MANY; res->top_score = 999999; res->denominator = 0;
            unique_clear(&st.unique); free(st.unique.g);
            free(pwm); free(rpwm); free(cons); free(rcons); continue;
        }
        if (st.unique.n == 0) {                                                          /* :593-602 */
            res->status = GMX_READ_UNMATCHED; res->top_score = 0; res->denominator = 0;
            free(st.unique.g); free(pwm); free(rpwm); free(cons); free(rcons); continue;
        }
        res->status = GMX_READ_MAPPED;                                                   /* :605-610 */
        res->denominator = st.denominator; res->top_score = st.top_align_score;
        res->n_groups = st.unique.n;

        /* hits, in (group key order, pos, strand) order */
        for (int gi = 0; gi < st.unique.n; ++gi) {
            orc_group *g = &st.unique.g[gi];
            for (int s = 0; s < g->n_spots; ++s) {
                if (n_hits < hits_cap && hits) {
                    gmx_hit *h = &hits[n_hits];
                    h->pos = g->spots[s].pos; h->score = (float)g->align_score; h->read = r;
                    h->group = (int16_t)gi; h->strand = (uint8_t)g->spots[s].strand; h->first_strand = (uint8_t)g->first_strand;
                } else if (hits) rc = GMX_ERR_OVERFLOW;
                n_hits++;
            }
        }
        res->hit_end = (int32_t)n_hits;

        /* src/Driver.cpp:614-716 create_match_output */
        double best_log = exp(-1.0);                       /* default ScoredSeq, ScoredSeq.h:115-118 */
        int best = -1;
        for (int gi = 0; gi < st.unique.n; ++gi) {
            orc_group *g = &st.unique.g[gi];
            if (do_score) score_group(ix, pr, g, st.denominator, pwm, n, amount, planes);
            if (g->log_align_score > best_log) { best_log = g->log_align_score; best = gi; }   /* is_greater, strict */
        }
        if (best >= 0) {
            orc_group *g = &st.unique.g[best];
            res->best_group = best;
            res->best_score = (float)g->align_score;
            res->best_posterior = (float)(g->log_align_score / st.denominator);
            res->best_n_positions = g->n_spots;
            res->best_first_strand = g->first_strand;
            res->best_first_pos = g->spots[0].pos;
            if (do_score) {
                /* get_SAM (inc/ScoredSeq.h:314-372) re-runs the traceback for the CIGAR: NEG uses the
                 * max_char consensus of the reverse-complemented PWM, POS the raw read string.  The
                 * CIGAR depends only on the moves, never on the consensus characters, so the gapped
                 * string reported here is the one score() feeds to the accumulators (max_char
                 * consensus of the oriented PWM, src/NormalScoredSeq.cpp:40-62) for both strands. */
                char aligned[2048], cigar[1024];
                int alen;
                const float *opwm = pwm;
                if (g->first_strand == GMX_NEG_STRAND) { orc_revcomp_pwm(pwm, n, rpwm); opwm = rpwm; }
                for (int i = 0; i < n; ++i) rcons[i] = (uint8_t)orc_max_char(opwm + 4 * i);
                rcons[n] = 0;
                alen = orc_nw_traceback(opwm, n, rcons, g->sequence, n, pr->align_scores, pr->gap, pr->max_gap, aligned, sizeof(aligned), cigar, sizeof(cigar));
                res->best_aligned_len = alen;
                if (cigar[0] == 0) strcpy(cigar, "*"); else orc_fix_cigar_for_deletions(cigar);
                if (cigar_out) { strncpy(cigar_out + (size_t)r * cigar_stride, cigar, (size_t)cigar_stride - 1); cigar_out[(size_t)r * cigar_stride + cigar_stride - 1] = 0; }
                if (aligned_out) { int c = alen < aligned_stride ? alen : aligned_stride; memcpy(aligned_out + (size_t)r * aligned_stride, aligned, (size_t)c); if (c < aligned_stride) aligned_out[(size_t)r * aligned_stride + c] = 0; }
            }
        }
        unique_clear(&st.unique); free(st.unique.g);
        free(pwm); free(rpwm); free(cons); free(rcons);
    }
    if (n_hits_out) *n_hits_out = n_hits;
    return rc;
}
