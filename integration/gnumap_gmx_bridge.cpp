// gnumap_gmx_bridge.cpp -- the reference-side binding of INTEGRATION.md as a compilable translation unit.
//
// This file is what a GNUMAP maintainer adds next to src/Driver.cpp: it sees the reference's own headers and globals
// and talks to libgmx.so only through the C ABI of include/gmx.h.  It contains no reference code; it is compiled (not
// linked) against the headers under /root/reference/inc by tests/test_abi.py::test_reference_side_binding_compiles
// to prove that the binding matches the reference's real types.
//
//   g++ -std=c++0x -I<reference>/inc -I<repo>/oracle/gsl_stub -I<repo>/include -DGMX_BRIDGE_TEST_ACCESS -c gnumap_gmx_bridge.cpp
//
// Patch points (reference file:line): (1) after gGen.LoadGenome() src/Driver.cpp:1428-1429 -> gmx_attach();
// (2)+(3) the two per-slice loops of parallel_thread_run src/Driver.cpp:2344-2373 -> gmx_run_slice();
// (4) instead of the MPI block src/Driver.cpp:1615-1811, before gGen.PrintFinal :1820 -> gmx_collect().
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

#ifdef GMX_BRIDGE_TEST_ACCESS
// GenomeBwt::index is private (inc/GenomeBwt.h:277); the real patch adds `bwaidx_t* GetIndex() { return index; }`
#define private public
#define protected public
#endif
#include "const_include.h"
#include "GenomeBwt.h"
#include "ScoredSeq.h"
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

static gmx_ctx *gGmx = 0;

static void gmx_die(const char *what)
{
    fprintf(stderr, "gmx: %s: %s\n", what, gGmx ? gmx_last_error(gGmx) : "no context");
    exit(1);
}

// patch point 1 ---------------------------------------------------------------------------------------------------
void gmx_attach(GenomeBwt &gen, int device)
{
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
    if (gmx_create(&gGmx, &gi, &gp, device) != GMX_OK) gmx_die("gmx_create");
}

// patch points 2 + 3 ------------------------------------------------------------------------------------------------
// One call per slice of <= READS_PER_PROC reads; fills what set_top_matches / create_match_output leave behind:
// gTopReadScore / gReadDenominator (incl. the status sentinels), the matched / not-matched counters and one
// TopReadOutput per (position, strand) of the best group for the SAM writer.
void gmx_run_slice(GenomeBwt &gen, unsigned read_begin, unsigned read_end, unsigned &good_seqs, unsigned &bad_seqs, std::vector<TopReadOutput> &sam_out)
{
    std::vector<int64_t> off(1, 0);
    std::string seq, qual;
    unsigned n = 0;
    for (unsigned k = read_begin; k < read_end && gReadArray[k]; ++k, ++n) {       // FASTQ reads: PWM == f(base, quality)
        const Read *r = gReadArray[k];
        seq += r->seq;
        qual += r->fq.substr(0, r->seq.size());
        off.push_back((int64_t)seq.size());
    }                                                                             // PRB / INT reads: fill gmx_reads.pwm instead
    gmx_reads in;
    memset(&in, 0, sizeof(in));
    in.n_reads = (int32_t)n; in.offsets = &off[0];
    in.seq = (const uint8_t *)seq.data(); in.qual = (const uint8_t *)qual.data();
    std::vector<gmx_read_result> res(n);
    if (gmx_process_batch(gGmx, &in, n ? &res[0] : 0) != GMX_OK) gmx_die("gmx_process_batch");
    int64_t nh = 0;
    if (gmx_get_hits(gGmx, 0, 0, &nh) != GMX_OK) gmx_die("gmx_get_hits");
    std::vector<gmx_hit> hits((size_t)nh + 1);
    if (gmx_get_hits(gGmx, &hits[0], nh + 1, &nh) != GMX_OK) gmx_die("gmx_get_hits");
    std::vector<char> cigar((size_t)n * 64 + 64);
    if (n && gmx_get_best_alignments(gGmx, &cigar[0], 64, 0, 0) != GMX_OK) gmx_die("gmx_get_best_alignments");
    for (unsigned i = 0; i < n; ++i) {
        const unsigned k = read_begin + i;
        gTopReadScore[k] = res[i].top_score;                                       // incl. READ_TOO_SHORT / _POOR / _MANY
        gReadDenominator[k] = res[i].denominator;
        if (res[i].status != GMX_READ_MAPPED) { bad_seqs++; continue; }
        good_seqs++;
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
            strncpy(out.CIGAR, &cigar[(size_t)i * 64], MAX_CIGAR_SZ - 1); out.CIGAR[MAX_CIGAR_SZ - 1] = '\0';
            out.readIndex = k;
            out.consensus = gReadArray[k]->seq;
            out.qual = gReadArray[k]->fq;
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
        const int rc = gmp ? gmx_format_gmp(gGmx, &names[0], target, 0.001, gSNP_PVAL, gSNP_MONOP ? 1 : 0, text.empty() ? 0 : &text[0], (int64_t)text.size(), &len)
                           : gmx_format_sgr(gGmx, &names[0], 0.001, text.empty() ? 0 : &text[0], (int64_t)text.size(), &len);
        if (rc == GMX_OK) break;
        if (rc != GMX_ERR_OVERFLOW || pass) gmx_die(gmp ? "gmx_format_gmp" : "gmx_format_sgr");
        text.resize((size_t)len);
    }
    const std::string path = std::string(fn) + (gmp ? ".gmp" : ".sgr");
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { perror(path.c_str()); exit(1); }
    if (len) fwrite(&text[0], 1, (size_t)len, f);
    fclose(f);
}

// patch point 4 ---------------------------------------------------------------------------------------------------
void gmx_collect(GenomeBwt &gen)
{
    float *planes[5] = {gen.GetGenomeAPtr(), gen.GetGenomeCPtr(), gen.GetGenomeGPtr(), gen.GetGenomeTPtr(), gen.GetGenomeNPtr()};
    // several GPUs: one process per GPU, ncclAllReduce(sum, f32) on gmx_accumulators_device() first.
    // The library keeps ceil(l_pac / gGEN_SIZE) bins; the reference allocates l_pac / gGEN_SIZE floats (src/GenomeBwt.cpp:323).
    uint64_t n_amount = 0, n_plane = 0;
    if (gmx_accumulators_device(gGmx, 0, &n_amount, 0, &n_plane) != GMX_OK) gmx_die("gmx_accumulators_device");
    std::vector<float> amount((size_t)n_amount);
    if (gmx_finish(gGmx, n_amount ? &amount[0] : 0, planes) != GMX_OK) gmx_die("gmx_finish");
    const uint64_t ref_bins = (uint64_t)GMX_GENOME_INDEX(gen)->bns->l_pac / gGEN_SIZE;
    memcpy(gen.GetGenomeAmtPtr(), amount.data(), sizeof(float) * (size_t)(ref_bins < n_amount ? ref_bins : n_amount));
    gmx_destroy(gGmx);
    gGmx = 0;
}
