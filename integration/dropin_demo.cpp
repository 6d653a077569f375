// dropin_demo.cpp -- the reference's own host code driving libgmx.so: a whole-program drop-in demonstration.
//
// TEST INFRASTRUCTURE (built by oracle/Makefile `make dropin` into oracle/_ref/gnumap_gmx_demo where /root/reference is
// present; the binary travels to the GPU box).  It links the UNMODIFIED reference objects (everything except Driver.o,
// whose main() and worker loops are what the patch replaces) with integration/gnumap_gmx_bridge.cpp and libgmx.so:
//
//   reference code used as is : globals + tables (const_define.h, a_matrices.c), GenomeBwt::LoadGenome (index files),
//                               SeqReader (FASTQ -> Read*), GenomeBwt::PrintFinal (.sgr / .gmp incl. the SNP LRT calls)
//   replaced by libgmx        : set_top_matches + create_match_output for every read (PHASE A + PHASE B), accumulators
//   restated here (10 lines)  : the SAM record printer of single_write_cond_wait (reference src/Driver.cpp:2166-2205),
//                               with the same ostream operations so that numbers print identically
//
//   gnumap_gmx_demo <genome.fa> <reads.fq> <out_prefix> [normal|snp|snp_monop|bs]
#include <pthread.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "const_include.h"
#include "const_define.h"
#include "GenomeBwt.h"
#include "SeqReader.h"
#include "ScoredSeq.h"
#include "SequenceOperations.h"

const char *pos_matrix = NULL;      /* src/Driver.cpp:72, read by a_matrices.c */
#include "a_matrices.c"

// src/Driver.cpp:135-139
Read **gReadArray;
double *gReadDenominator;
double *gTopReadScore;

void gmx_attach(GenomeBwt &gen, int device);
void gmx_run_slice(GenomeBwt &gen, unsigned read_begin, unsigned read_end, unsigned &good_seqs, unsigned &bad_seqs, std::vector<TopReadOutput> &sam_out);
void gmx_collect(GenomeBwt &gen);
void gmx_print_final(GenomeBwt &gen, const char *fn);

static void write_sam(std::ofstream &of, std::vector<TopReadOutput> &v)
{   // reference src/Driver.cpp:2166-2205
    for (std::vector<TopReadOutput>::iterator vit = v.begin(); vit != v.end(); ++vit) {
        of << (*vit).READ_NAME << "\t";
        if ((*vit).strand == POS_STRAND) of << 0x0000 << "\t"; else of << 0x0010 << "\t";
        of << (*vit).CHR_NAME << "\t" << (*vit).CHR_POS << "\t";
        of << (*vit).MAPQ << "\t";
        if ((*vit).strand == POS_STRAND) of << (*vit).CIGAR << "\t"; else of << reverse_CIGAR((*vit).CIGAR) << "\t";
        of << "*\t0\t0\t";
        if ((*vit).strand == POS_STRAND) { of << (*vit).consensus << "\t"; of << (*vit).qual << "\t"; }
        else { of << reverse_comp((*vit).consensus) << "\t"; of << reverse_qual((*vit).qual) << "\t"; }
        of << "XA:f:" << (float)((*vit).A_SCORE) * (1.0 / gADJUST) << "\t";
        of << "XP:f:" << (float)((*vit).POST_PROB) << "\t";
        of << "X0:i:" << (*vit).SIM_MATCHES << "\n";
    }
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s genome.fa reads.fq out_prefix [normal|snp|snp_monop|bs]\n", argv[0]); return 2; }
    const std::string mode = argc > 4 ? argv[4] : "normal";
    InitProg();                                   // const_define.h:127-164
    gINT2BASE[5] = 'I'; gINT2BASE[6] = 'D';
    setup_alignment_matrices();                   // a_matrices.c:25
    gMER_SIZE = DEF_MER_SIZE;                      // src/Driver.cpp:1162-1207 defaults
    gJUMP_SIZE = gMER_SIZE / 2;
    gVERBOSE = 0;
    if (mode == "snp" || mode == "snp_monop") { gSNP = true; gGEN_SIZE = 1; gSNP_MONOP = mode == "snp_monop"; }                                   // src/Driver.cpp:3203-3211
    if (mode == "bs") { gBISULFITE = true; gGEN_SIZE = 1; gALIGN_SCORES[(int)'c'][3] = gMATCH; }   // :2805-2810, :1260-1268

    GenomeBwt gen;
    gen.use(argv[1]);
    gen.LoadGenome();
    // the reference never zeroes amount_genome (src/GenomeBwt.cpp:323); gmx_finish overwrites it entirely
    gmx_attach(gen, 0);

    const unsigned slice = READS_PER_PROC;
    gReadArray = new Read *[slice + 1];
    gReadDenominator = new double[slice + 1];
    gTopReadScore = new double[slice + 1];
    std::ofstream of((std::string(argv[3]) + ".sam").c_str());
    SeqReader sr;
    sr.use(argv[2]);
    unsigned good = 0, bad = 0;
    bool more = true;
    while (more) {
        unsigned n = 0;
        for (; n < slice; ++n) {
            Read *r = sr.GetNextSequence();
            if (!r) { more = false; break; }
            gReadArray[n] = r;
        }
        gReadArray[n] = 0;
        if (n == 0) break;
        std::vector<TopReadOutput> out;
        gmx_run_slice(gen, 0, n, good, bad, out);
        write_sam(of, out);
        for (unsigned k = 0; k < n; ++k) { delete_read(gReadArray[k]); gReadArray[k] = 0; }
    }
    of.close();
    gmx_print_final(gen, (std::string(argv[3]) + ".native").c_str());   // the library's printers on the same accumulators
    gmx_collect(gen);
    gen.PrintFinal(argv[3]);                      // the reference's own .sgr / .gmp printers on the GPU's accumulators
    fprintf(stdout, "#\tSequences matched: %u\n#\tSequences not matched: %u\n", good, bad);
    return 0;
}
