// gnumap_gmx_bridge.cpp -- the reference-side binding: what a GNUMAP maintainer adds next to src/Driver.cpp.
//
// It sees the reference's own headers and globals and talks to libgmx.so only through the C ABI of include/gmx.h.  It
// contains no reference code.  integration/driver_gmx.patch makes src/Driver.cpp call it at four points:
//
//   (1) after gGen.LoadGenome()                  src/Driver.cpp:1428-1429  -> gmx_attach()      contexts (one per GPU) + communicator
//   (2)+(3) the two per-slice loops of parallel_thread_run  :2344-2373     -> gmx_run_slice()   PHASE A + PHASE B of one slice
//   (4) before gGen.PrintFinal                   :1820-1823 (in place of the MPI block :1615-1811) -> gmx_collect()
//   and the slice size of the worker threads     :970, :2307               -> gmx_slice_reads()
//
// `make -C oracle patched` applies the patch to a copy of the reference's Driver.cpp and links the UNMODIFIED other
// objects with this file and libgmx.so into oracle/_ref/gnumap_gmx: the reference program, all of its options, on GPUs.
//
//   g++ -std=c++0x -I<reference>/inc -I<repo>/oracle/gsl_stub -I<repo>/include -DGMX_BRIDGE_TEST_ACCESS -c gnumap_gmx_bridge.cpp
//
// Environment of the patched binary: GMX_DISABLE=1 runs the reference's own CPU loops (A/B runs with one binary);
// GMX_GPUS=n caps the GPUs used (default: min(threads, devices)); GMX_SLICE_READS=n reads per worker-thread slice
// (default 65536; the reference's READS_PER_PROC is 2048); GMX_COMM=peer|nccl picks the reduce backend;
// GMX_NATIVE_PRINT=1 also writes <out>.native.sgr / .gmp with the library's printers.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <pthread.h>

#ifdef GMX_BRIDGE_TEST_ACCESS
// GenomeBwt::index is private (inc/GenomeBwt.h:277); a maintainer adds `bwaidx_t* GetIndex() { return index; }` instead
#define private public
#define protected public
#endif
#include "const_include.h"
#include "GenomeBwt.h"
#include "ScoredSeq.h"
#include "SequenceOperations.h"
#ifdef GMX_BRIDGE_TEST_ACCESS
#undef private
#undef protected
#define GMX_GENOME_INDEX(gen) ((gen).index)
#else
#define GMX_GENOME_INDEX(gen) ((gen).GetIndex())
#endif

#include "gmx.h"

// globals of src/Driver.cpp:135-139 and the tunables of inc/const_define.h that inc/const_include.h does not declare
extern Read **gReadArray;
extern double *gReadDenominator;
extern double *gTopReadScore;
extern bool perc;
extern unsigned int gJUMP_SIZE, gMAX_MATCHES;
extern int gMIN_JUMP_MATCHES;
extern float gCUTOFF_SCORE;
extern bool gFAST;
extern bool gPRINT_ALL_SAM;

#define GMX_BRIDGE_MAX_GPUS 16
static gmx_ctx *gGmx[GMX_BRIDGE_MAX_GPUS];
static pthread_mutex_t gGmxLock[GMX_BRIDGE_MAX_GPUS];        // worker threads that share a GPU take turns on its context
static int gGmxCigar[GMX_BRIDGE_MAX_GPUS];                   // CIGAR slot bytes of each context (GMX_OPT_CIGAR_STRIDE)
static int gGmxN = 0;
static gmx_comm *gGmxComm = 0;
static int gGmxIllumina = 0;                                  // gILLUMINA when the contexts were created

// consensus character of one PWM row, as GetConsensus() picks it (src/Driver.cpp:317-347)
static char gmx_row_char(const float *c)
{
    if (c[0] == c[1] && c[0] == c[2] && c[0] == c[3]) return 'n';
    if (c[0] >= c[1]) return c[0] >= c[2] ? (c[0] >= c[3] ? 'a' : 't') : (c[2] >= c[3] ? 'g' : 't');
    return c[1] >= c[2] ? (c[1] >= c[3] ? 'c' : 't') : (c[2] >= c[3] ? 'g' : 't');
}

static void gmx_die(const char *what, gmx_ctx *ctx)
{
    fprintf(stderr, "gmx: %s: %s\n", what, ctx ? gmx_last_error(ctx) : "no context");
    exit(1);
}

bool gmx_enabled() { const char *e = getenv("GMX_DISABLE"); return !(e && *e && *e != '0'); }

// reads per worker-thread slice: the GPU wants slices far larger than the reference's READS_PER_PROC (2048)
unsigned gmx_slice_reads(unsigned reference_default)
{
    if (!gmx_enabled()) return reference_default;
    const char *e = getenv("GMX_SLICE_READS");
    long v = e ? atol(e) : 65536;
    return (unsigned)(v < 1 ? 1 : v);
}

// patch point 1 ---------------------------------------------------------------------------------------------------
void gmx_attach(GenomeBwt &gen, unsigned n_threads)
{
    if (gPRINT_ALL_SAM) { fprintf(stderr, "gmx: --print_all_sam is not carried by the GPU path (run with GMX_DISABLE=1)\n"); exit(1); }
    bwaidx_t *ix = GMX_GENOME_INDEX(gen);
    std::vector<int64_t> off(ix->bns->n_seqs);
    std::vector<int32_t> len(ix->bns->n_seqs);
    for (int i = 0; i < ix->bns->n_seqs; ++i) { off[i] = ix->bns->anns[i].offset; len[i] = ix->bns->anns[i].len; }
    gmx_index gi;
    memset(&gi, 0, sizeof(gi));
    gi.bwt = ix->bwt->bwt; gi.bwt_words = ix->bwt->bwt_size; gi.primary = ix->bwt->primary;
    for (int i = 0; i < 5; ++i) gi.L2[i] = ix->bwt->L2[i];
    gi.seq_len = ix->bwt->seq_len; gi.sa = (const uint64_t *)ix->bwt->sa; gi.n_sa = ix->bwt->n_sa; gi.sa_intv = ix->bwt->sa_intv;
    gi.n_seqs = ix->bns->n_seqs; gi.pac = ix->pac; gi.l_pac = ix->bns->l_pac;
    gi.seq_offset = &off[0]; gi.seq_len_arr = &len[0];

    gmx_params gp;                       // snapshot of the globals AFTER main() has edited them (src/Driver.cpp:1083-1315)
    memset(&gp, 0, sizeof(gp));
    memcpy(gp.align_scores, gALIGN_SCORES, sizeof(gp.align_scores));
    memcpy(gp.phmm_scores, gPHMM_ALIGN_SCORES, sizeof(gp.phmm_scores));
    gp.gap = gGAP; gp.max_gap = gMAX_GAP; gp.mer = gMER_SIZE; gp.jump = gJUMP_SIZE; gp.min_seed_hits = gMIN_JUMP_MATCHES;
    gp.max_kmer_hits = gMAX_KMER_SIZE; gp.max_matches = gMAX_MATCHES; gp.gen_size = gGEN_SIZE;
    gp.align_score = gALIGN_SCORE; gp.perc = perc; gp.cutoff = gCUTOFF_SCORE;
    gp.match_pos = gMATCH_POS_STRAND; gp.match_neg = gMATCH_NEG_STRAND; gp.unique_only = gUNIQUE; gp.fast = gFAST;
    gp.use_nw = gNW; gp.illumina = gILLUMINA; gp.adjust = gADJUST;
    gp.mode = gSNP ? GMX_MODE_SNP : ((gBISULFITE || gATOG) ? GMX_MODE_BS : GMX_MODE_NORMAL);
    gGmxIllumina = gILLUMINA;

    // one context per GPU; worker thread t drives context t % n (the reference's threads share gGen under a mutex instead)
    int want = (int)n_threads;
    if (const char *e = getenv("GMX_GPUS")) want = atoi(e) < want ? atoi(e) : want;
    if (want < 1) want = 1;
    if (want > GMX_BRIDGE_MAX_GPUS) want = GMX_BRIDGE_MAX_GPUS;
    gGmxN = 0;
    for (int d = 0; d < want; ++d) {
        gmx_ctx *c = 0;
        int rc = gmx_create(&c, &gi, &gp, d);
        if (rc != GMX_OK) {
            if (d == 0 || rc != GMX_ERR_INVALID) {                                  // fewer devices than threads: use what is there
                fprintf(stderr, "gmx: gmx_create on device %d: %s\n", d, c && *gmx_last_error(c) ? gmx_last_error(c) : gmx_strerror(rc));
                exit(1);
            }
            gmx_destroy(c);
            break;
        }
        gmx_set_option(c, GMX_OPT_STAGE_TIMING, 0);                                 // nobody reads gmx_get_stage_stats here
        pthread_mutex_init(&gGmxLock[d], NULL);
        gGmxCigar[d] = 64;                                                          // the library's default
        gGmx[gGmxN++] = c;
    }
    if (gGmxN > 1) {
        int backend = GMX_COMM_AUTO;
        if (const char *e = getenv("GMX_COMM")) backend = !strcmp(e, "nccl") ? GMX_COMM_NCCL : GMX_COMM_PEER;
        if (gmx_comm_create(&gGmxComm, gGmx, gGmxN, backend) != GMX_OK) gmx_die("gmx_comm_create", gGmx[0]);
    }
    fprintf(stderr, "gmx: %d GPU context(s) for %u worker thread(s)\n", gGmxN, n_threads);
}

// patch points 2 + 3 ------------------------------------------------------------------------------------------------
// One call per worker-thread slice; fills what set_top_matches / create_match_output leave behind: gTopReadScore /
// gReadDenominator (incl. the status sentinels), the matched / not-matched counters, the thread's NW count (DEBUG_NW)
// and one TopReadOutput per (position, strand) of the best group for the SAM writer.
void gmx_run_slice(GenomeBwt &gen, unsigned thread_id, unsigned read_begin, unsigned read_end, unsigned &good_seqs, unsigned &bad_seqs,
                   unsigned &n_nw, std::vector<TopReadOutput> &sam_out)
{
    // FASTQ reads: the PWM is a function of (base, quality char) and two bytes per base go to the GPU.  Anything else
    // (PRB / INT / FASTA reads, or a reader that changed its quality offset on the way: src/SeqReader.cpp:1180-1188)
    // goes as the raw PWM rows the reader built, with GetConsensus() as the sequence (src/Driver.cpp:352-364).
    unsigned n = 0;
    bool as_pwm = (int)gILLUMINA != gGmxIllumina;
    for (unsigned k = read_begin; k < read_end && gReadArray[k]; ++k, ++n) {
        const Read *r = gReadArray[k];
        if (r->seq.size() != r->length || r->fq.size() < r->seq.size()) as_pwm = true;
    }
    std::vector<int64_t> off(1, 0);
    std::string seq, qual;
    std::vector<float> pwm;
    for (unsigned k = read_begin; k < read_begin + n; ++k) {
        const Read *r = gReadArray[k];
        if (!as_pwm) { seq += r->seq; qual.append(r->fq, 0, r->seq.size()); }
        else {
            if (r->seq.size() == r->length) seq += r->seq;
            else for (unsigned i = 0; i < r->length; ++i) seq += gmx_row_char(r->pwm[i]);
            for (unsigned i = 0; i < r->length; ++i) pwm.insert(pwm.end(), r->pwm[i], r->pwm[i] + 4);
        }
        off.push_back((int64_t)seq.size());
    }
    gmx_reads in;
    memset(&in, 0, sizeof(in));
    in.n_reads = (int32_t)n; in.offsets = &off[0];
    in.seq = (const uint8_t *)seq.data();
    in.qual = as_pwm ? 0 : (const uint8_t *)qual.data();
    in.pwm = as_pwm && !pwm.empty() ? &pwm[0] : 0;
    std::vector<gmx_read_result> res(n);
    std::vector<gmx_hit> hits;

    const int g = (int)(thread_id % (unsigned)gGmxN);
    gmx_ctx *ctx = gGmx[g];
    pthread_mutex_lock(&gGmxLock[g]);
    // The reference builds CIGARs unbounded; the library writes them into fixed slots and fails the batch rather than cut
    // one.  No alignment of an n-base read needs more than 4 n + 16 bytes ("1I1D" per base), so the slot is sized for the
    // longest read seen so far (a cheap score table makes gap-rich alignments real: -S with gap > mismatch).
    size_t longest = 0;
    for (unsigned i = 0; i < n; ++i) longest = std::max(longest, (size_t)(off[i + 1] - off[i]));
    int want = (int)std::min<size_t>(2048, (4 * longest + 16 + 15) / 16 * 16);
    if (want > gGmxCigar[g]) {
        if (gmx_set_option(ctx, GMX_OPT_CIGAR_STRIDE, want) != GMX_OK) gmx_die("gmx_set_option(CIGAR_STRIDE)", ctx);
        gGmxCigar[g] = want;
    }
    const size_t cs = (size_t)gGmxCigar[g];
    std::vector<char> cigar((size_t)n * cs + cs);
    if (n && gmx_process_batch(ctx, &in, &res[0]) != GMX_OK) gmx_die("gmx_process_batch", ctx);
    int64_t nh = 0;
    if (n && gmx_get_hits(ctx, 0, 0, &nh) != GMX_OK) gmx_die("gmx_get_hits", ctx);
    hits.resize((size_t)nh + 1);
    if (n && gmx_get_hits(ctx, &hits[0], nh + 1, &nh) != GMX_OK) gmx_die("gmx_get_hits", ctx);
    if (n && gmx_get_best_alignments(ctx, &cigar[0], (int)cs, 0, 0) != GMX_OK) gmx_die("gmx_get_best_alignments", ctx);
    pthread_mutex_unlock(&gGmxLock[g]);

    for (unsigned i = 0; i < n; ++i) {
        const unsigned k = read_begin + i;
        gTopReadScore[k] = res[i].top_score;                                       // incl. READ_TOO_SHORT / _POOR / _MANY
        gReadDenominator[k] = res[i].denominator;
        n_nw += (unsigned)res[i].n_candidates;
        // READ_TOO_MANY counts as matched in the reference's statistics (num_matched++, src/Driver.cpp:520,579) and prints nothing
        if (res[i].status == GMX_READ_TOO_MANY) { good_seqs++; continue; }
        if (res[i].status != GMX_READ_MAPPED) { bad_seqs++; continue; }
        good_seqs++;
        if (!GMX_READ_PRINTS_SAM(res[i])) continue;                                // src/Driver.cpp:695: best group below top - SAME_DIFF
        // what ScoredSeq::get_SAM fills (inc/ScoredSeq.h:293-404): one TopReadOutput per (position, strand) of the best group
        const double total = exp((double)res[i].best_score) / res[i].denominator;
        // the denominator was summed from the device's exp(): the host's exp() may differ in the last place, so a
        // sole hit can give a ratio one ulp above 1
        int mapq = (total >= 1) ? 30 : (int)round(-10 * log(1 - total) / log(10.0));
        if (mapq > 30) mapq = 30;
        for (int32_t h = res[i].hit_begin; h < res[i].hit_end; ++h) {
            if (hits[h].group != res[i].best_group) continue;
            std::pair<std::string, unsigned long> seq_pos = gen.GetPosPair((unsigned long)hits[h].pos);
            TopReadOutput out;
            strncpy(out.READ_NAME, gReadArray[k]->name, MAX_NAME_SZ - 1); out.READ_NAME[MAX_NAME_SZ - 1] = '\0';
            strncpy(out.CHR_NAME, seq_pos.first.c_str(), MAX_NAME_SZ - 1); out.CHR_NAME[MAX_NAME_SZ - 1] = '\0';
            out.CHR_POS = seq_pos.second + 1;
            out.strand = hits[h].strand == GMX_NEG_STRAND ? NEG_STRAND : POS_STRAND;
            out.MAPQ = mapq;
            strncpy(out.CIGAR, &cigar[(size_t)i * cs], MAX_CIGAR_SZ - 1); out.CIGAR[MAX_CIGAR_SZ - 1] = '\0';
            out.readIndex = k;
            out.consensus = seq.substr((size_t)off[i], (size_t)(off[i + 1] - off[i]));        // GetConsensus(read)
            out.qual = str2qual(*gReadArray[k]);
            out.A_SCORE = res[i].best_score;
            out.SIM_MATCHES = res[i].best_n_positions;
            out.POST_PROB = res[i].best_posterior;
            sam_out.push_back(out);
        }
    }
}

// optional, before patch point 4: the library's own printers in place of gGen.PrintFinal (src/Driver.cpp:1820-1823).
// GenomeBwt::PrintFinal picks the file by mode (src/GenomeBwt.cpp:911-923); rows are selected on the GPU from the
// accumulators where they are, so the 4-24 B per genome position never cross to the host.
void gmx_print_final(GenomeBwt &gen, const char *fn)
{
    if (gGmxComm && gmx_comm_reduce(gGmxComm, 0) != GMX_OK) gmx_die("gmx_comm_reduce", gGmx[0]);
    const bntseq_t *bns = GMX_GENOME_INDEX(gen)->bns;
    std::vector<const char *> names((size_t)bns->n_seqs);
    for (int i = 0; i < bns->n_seqs; ++i) names[(size_t)i] = bns->anns[i].name;
    const bool gmp = gSNP || gBISULFITE || gATOG;
    int target = -1;                                  // genome base PrintFinalBisulfite reports (src/GenomeBwt.cpp:1136-1160)
    if (!gSNP && gBISULFITE) target = (gMATCH_POS_STRAND && !gBISULFITE2) ? 1 : 2;
    else if (!gSNP && gATOG) target = gMATCH_POS_STRAND ? 0 : 3;
    std::vector<char> text;
    int64_t len = 0;
    for (int pass = 0; pass < 2; ++pass) {            // first pass sizes the buffer
        const int rc = gmp ? gmx_format_gmp(gGmx[0], &names[0], target, 0.001, gSNP_PVAL, gSNP_MONOP ? 1 : 0, text.empty() ? 0 : &text[0], (int64_t)text.size(), &len)
                           : gmx_format_sgr(gGmx[0], &names[0], 0.001, text.empty() ? 0 : &text[0], (int64_t)text.size(), &len);
        if (rc == GMX_OK) break;
        if (rc != GMX_ERR_OVERFLOW || pass) gmx_die(gmp ? "gmx_format_gmp" : "gmx_format_sgr", gGmx[0]);
        text.resize((size_t)len);
    }
    const std::string path = std::string(fn) + (gmp ? ".gmp" : ".sgr");
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { perror(path.c_str()); exit(1); }
    if (len) fwrite(&text[0], 1, (size_t)len, f);
    fclose(f);
}

// patch point 4 ---------------------------------------------------------------------------------------------------
// The sum over the GPUs' accumulators (gmx_finish on the root context reduces first: the library-side counterpart of the
// MPI block, src/Driver.cpp:1615-1811) lands in the arrays GenomeBwt::PrintFinal reads.
void gmx_collect(GenomeBwt &gen, const char *out_prefix)
{
    if (getenv("GMX_NATIVE_PRINT")) gmx_print_final(gen, (std::string(out_prefix) + ".native").c_str());
    float *planes[5] = {gen.GetGenomeAPtr(), gen.GetGenomeCPtr(), gen.GetGenomeGPtr(), gen.GetGenomeTPtr(), gen.GetGenomeNPtr()};
    // The library keeps ceil(l_pac / gGEN_SIZE) bins; the reference allocates l_pac / gGEN_SIZE floats (src/GenomeBwt.cpp:323).
    uint64_t n_amount = 0, n_plane = 0;
    if (gmx_accumulators_device(gGmx[0], 0, &n_amount, 0, &n_plane) != GMX_OK) gmx_die("gmx_accumulators_device", gGmx[0]);
    std::vector<float> amount((size_t)n_amount);
    if (gmx_finish(gGmx[0], n_amount ? &amount[0] : 0, planes) != GMX_OK) gmx_die("gmx_finish", gGmx[0]);
    const uint64_t ref_bins = (uint64_t)GMX_GENOME_INDEX(gen)->bns->l_pac / gGEN_SIZE;
    memcpy(gen.GetGenomeAmtPtr(), amount.data(), sizeof(float) * (size_t)(ref_bins < n_amount ? ref_bins : n_amount));
    if (gGmxComm) { gmx_comm_destroy(gGmxComm); gGmxComm = 0; }
    for (int i = 0; i < gGmxN; ++i) { gmx_destroy(gGmx[i]); gGmx[i] = 0; }
    gGmxN = 0;
}
