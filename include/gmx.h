/* gmx.h -- C ABI of the B200-native GNUMAP hot path ("gmx" = GNUMAP mapping accelerator).
 *
 * This is the drop-in boundary described in DESIGN.md / INTEGRATION.md.  The reference
 * (byucsl/gnumap 4.0.0 BETA) has no plugin or FFI layer: its hot path is reached through the
 * abstract class `Genome` (reference inc/Genome.h:88-141), value-type `bin_seq` objects
 * (reference inc/bin_seq.h:47-200) and `virtual ScoredSeq::score` (reference inc/ScoredSeq.h:259),
 * driven by the two per-batch loops of the worker threads (reference src/Driver.cpp:2344-2373).
 * Each entry point below names the reference function(s) it replaces.
 *
 * Rules of the boundary
 *   - plain C: pointers + sizes, no C++/torch types, no exceptions; every call returns GMX_OK (0)
 *     or a negative error code; per-read outcomes are data, not errors;
 *   - the caller owns every host buffer; the library owns all device memory;
 *   - one context drives one GPU (one process per GPU); there is NO CPU fallback: without a usable
 *     CUDA device every call fails with GMX_ERR_NO_DEVICE.
 */
#ifndef GMX_H
#define GMX_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMX_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------------------------ */
#define GMX_OK               0
#define GMX_ERR_INVALID     -1   /* bad argument */
#define GMX_ERR_CUDA        -2   /* CUDA runtime error, see gmx_last_error() */
#define GMX_ERR_NOMEM       -3
#define GMX_ERR_UNSUPPORTED -4   /* option combination not implemented on the device path */
#define GMX_ERR_OVERFLOW    -5   /* an internal device work list overflowed even after regrowth */
#define GMX_ERR_NO_DEVICE   -6   /* no CUDA device: there is no CPU fallback */
#define GMX_ERR_STATE       -7   /* call sequence violated (e.g. score_batch without map_batch) */
#define GMX_ERR_FORMAT      -8   /* malformed FASTQ text (device indexer: use gmx_fastq_scan_host; host: invalid quality) */

/* ---- per-read status (mirrors the sentinels the reference stores in gTopReadScore,
 *      reference inc/const_include.h:184-188 and src/Driver.cpp:446-611) ----------------------- */
#define GMX_READ_MAPPED     0    /* unique.size() > 0                      top = best NW score   */
#define GMX_READ_UNMATCHED  1    /* no accepted hit                         top = 0               */
#define GMX_READ_TOO_SHORT  2    /* length < mer            (READ_TOO_SHORT) top = -2             */
#define GMX_READ_TOO_POOR   3    /* self score < cutoff     (READ_TOO_POOR)  top = -3             */
#define GMX_READ_TOO_MANY   4    /* > max_matches or !unique (READ_TOO_MANY) top = 999999; the reference's
                                    "Sequences matched" line counts these (num_matched++, Driver.cpp:520,579)       */

#define GMX_POS_STRAND 0         /* reference inc/const_include.h:190-191 */
#define GMX_NEG_STRAND 1

#define GMX_MODE_NORMAL 0        /* NormalScoredSeq: amount_genome only                           */
#define GMX_MODE_BS     1        /* BSScoredSeq (-b / --b2 / -d): amount_genome + 5 base planes   */
#define GMX_MODE_SNP    2        /* SNPScoredSeq (--snp): pair-HMM posteriors into 5 planes       */

typedef struct gmx_ctx gmx_ctx;

/* Borrowed host views of a loaded BWA-style index == the reference's bwaidx_t
 * (reference inc/GenomeBwt.h:64-72; bwt_t inc/bwt.h:46-58; bntseq_t inc/bntseq.h:41-66).
 * gmx_create() copies everything it needs to the device; the views may be freed afterwards. */
typedef struct gmx_index {
    const uint32_t *bwt;         /* bwt_t::bwt : occ-interleaved BWT, 64-byte blocks (bwtindex.c:128-150) */
    uint64_t        bwt_words;   /* bwt_t::bwt_size (uint32 words)                                 */
    uint64_t        primary;     /* bwt_t::primary                                                 */
    uint64_t        L2[5];       /* bwt_t::L2                                                      */
    uint64_t        seq_len;     /* bwt_t::seq_len (== l_pac: the index is forward-only)           */
    const uint64_t *sa;          /* bwt_t::sa, n_sa entries, sa[0] == (uint64_t)-1 (bwt.c:83)      */
    uint64_t        n_sa;
    int32_t         sa_intv;     /* 32 in every index GNUMAP writes (bwtindex.c:286)               */
    int32_t         n_seqs;      /* bntseq_t::n_seqs                                               */
    const uint8_t  *pac;         /* 2-bit packed forward genome, MSB first (bntseq.c:224-225)      */
    int64_t         l_pac;       /* bntseq_t::l_pac                                                */
    const int64_t  *seq_offset;  /* bntann1_t::offset per sequence                                 */
    const int32_t  *seq_len_arr; /* bntann1_t::len per sequence                                    */
} gmx_index;

/* Snapshot of the reference globals that reach the hot path, taken AFTER main()
 * has finished editing them (reference src/Driver.cpp:1083-1315). */
typedef struct gmx_params {
    float    align_scores[256][4]; /* gALIGN_SCORES, already gADJUST-scaled (a_matrices.c:59-87)   */
    float    phmm_scores[256][4];  /* gPHMM_ALIGN_SCORES (a_matrices.c:97-126)                     */
    float    gap;                  /* gGAP (scaled)                     const_define.h:69          */
    int32_t  max_gap;              /* gMAX_GAP                          const_define.h:86          */
    int32_t  mer;                  /* gMER_SIZE                         const_define.h:47          */
    int32_t  jump;                 /* gJUMP_SIZE                        const_define.h:102         */
    int32_t  min_seed_hits;        /* gMIN_JUMP_MATCHES                 const_define.h:107         */
    uint32_t max_kmer_hits;        /* gMAX_KMER_SIZE (0 = unlimited)    const_define.h:60          */
    uint32_t max_matches;          /* gMAX_MATCHES                      const_define.h:49          */
    uint32_t gen_size;             /* gGEN_SIZE (accumulator bin width) const_define.h:101         */
    float    align_score;          /* gALIGN_SCORE                      const_define.h:61          */
    int32_t  perc;                 /* perc                              const_define.h:62          */
    float    cutoff;               /* gCUTOFF_SCORE                     const_define.h:63          */
    int32_t  match_pos;            /* gMATCH_POS_STRAND                 const_define.h:94          */
    int32_t  match_neg;            /* gMATCH_NEG_STRAND                 const_define.h:95          */
    int32_t  unique_only;          /* gUNIQUE                           const_define.h:50          */
    int32_t  fast;                 /* gFAST                             const_define.h:92          */
    int32_t  use_nw;               /* gNW (only 1 is implemented on the device)                    */
    int32_t  mode;                 /* GMX_MODE_*  (gSNP / gBISULFITE / gATOG)                      */
    int32_t  illumina;             /* gILLUMINA: Q offset 64 + Q2Prb_ill (SeqReader.cpp:618-622)   */
    float    adjust;               /* gADJUST (only used to print XA, Driver.cpp:2202)             */
} gmx_params;

/* Fill `p` with the reference defaults (const_define.h + setup_alignment_matrices(), Normal mode). */
void gmx_default_params(gmx_params *p);

/* One batch of reads (the reference hands its workers slices of <= 2048 Read* per thread,
 * Driver.cpp:2339; a batch here may be any size).
 *   seq/qual : the FASTQ sequence and quality lines, concatenated, raw ASCII (case preserved);
 *              read r occupies [offsets[r], offsets[r+1]).  The PWM of a FASTQ read is a pure
 *              function of (base, quality char) -- reference src/SeqReader.cpp:1155-1240.
 *   pwm      : optional float[total_len][4] for reads whose PWM is not such a function
 *              (PRB / INT inputs, SeqReader.cpp:541-571,901-978); NULL for FASTQ.
 * A read may be up to 1024 bases long in every mode (GMX_ERR_UNSUPPORTED beyond; the reference has no limit of its own
 * but its float scores leave the range of exp() near 940 bases).  The kernels are tuned for reads of up to 160 bases;
 * longer ones take generic paths (exact vote tables beyond 448, register strips of the pair-HMM up to 256, local-memory
 * strips up to 1024). */
typedef struct gmx_reads {
    int32_t        n_reads;
    const int64_t *offsets;   /* [n_reads + 1] */
    const uint8_t *seq;
    const uint8_t *qual;
    const float   *pwm;
    int32_t        on_device; /* 0: the four arrays are host memory (copied to the GPU inside the call);
                                 1: they already live in the memory of the context's GPU (seq / qual are
                                    read in aligned 8-byte words: the last word may reach up to 7 bytes past the
                                    final base, inside the allocation granule of any cudaMalloc'd buffer)  */
    int32_t        max_len;   /* longest read of the batch; 0 = let the library scan offsets (host only) */
    /* Reads used in place inside a larger text (FASTQ): device-resident batches only.  When `lens` is given, read r
     * is seq[offsets[r] .. +lens[r]) and offsets has n_reads entries; when `qual_offsets` is given its quality string
     * starts at qual[qual_offsets[r]].  NULL for packed batches. */
    const int64_t *qual_offsets;
    const int32_t *lens;
} gmx_reads;

/* Per-read outcome of PHASE A + PHASE B  (== gTopReadScore / gReadDenominator / the best
 * ScoredSeq chosen in create_match_output, reference src/Driver.cpp:606-611,640-716). */
typedef struct gmx_read_result {
    double   top_score;        /* gTopReadScore[k] incl. sentinels                                 */
    double   denominator;      /* gReadDenominator[k]                                              */
    float    max_align_score;  /* self-alignment score (Driver.cpp:466)                            */
    int32_t  status;           /* GMX_READ_*                                                       */
    int32_t  n_groups;         /* unique.size(): distinct genome strings accepted                  */
    int32_t  n_candidates;     /* NW alignments evaluated for this read (== DEBUG_NW counter)      */
    /* best ScoredSeq (first in key order with the largest exp(score), strict >):                 */
    float    best_score;       /* A_SCORE                                                          */
    float    best_posterior;   /* POST_PROB = exp(score)/denominator                               */
    int32_t  best_n_positions; /* SIM_MATCHES (X0)                                                 */
    int32_t  best_first_strand;
    uint64_t best_first_pos;   /* smallest (pos,strand) of the best group, absolute 0-based        */
    int32_t  hit_begin;        /* [hit_begin, hit_end) into the gmx_hit array of the batch         */
    int32_t  hit_end;
    int32_t  best_group;       /* index (key order) of the best group within this read, or -1      */
    int32_t  best_aligned_len; /* length of the gapped `aligned` string of the best group          */
} gmx_read_result;

/* create_match_output hands a read's best group to the SAM writer only when that group's score (the NW score of its
 * FIRST hit, ScoredSeq::align_score) lies within SAME_DIFF of the read's top NW score (reference src/Driver.cpp:695,
 * inc/const_include.h:188); a mapped read whose best group misses that test -- e.g. a group whose later member, on the
 * other strand, scored a few ulps higher -- counts as matched and scores into the accumulators but prints nothing. */
#define GMX_SAME_DIFF 0.00001
#define GMX_READ_PRINTS_SAM(res) ((res).status == GMX_READ_MAPPED && (double)(res).best_score > (res).top_score - GMX_SAME_DIFF)

/* One accepted (position, strand) of one group (== one element of ScoredSeq::positions). */
typedef struct gmx_hit {
    uint64_t pos;              /* absolute 0-based genome position                                 */
    float    score;            /* NW score of the group's first hit (ScoredSeq::align_score)       */
    int32_t  read;             /* read index in the batch                                          */
    int16_t  group;            /* group index within the read, key (lexicographic) order           */
    uint8_t  strand;           /* GMX_POS_STRAND / GMX_NEG_STRAND                                  */
    uint8_t  first_strand;     /* ScoredSeq::firstStrand of the group                              */
} gmx_hit;

/* ---- context ------------------------------------------------------------------------------ */

/* After gGen.LoadGenome() (Driver.cpp:1428-1429): upload index + tables to GPU `device`,
 * de-sample the suffix array, allocate zeroed accumulators. */
int  gmx_create(gmx_ctx **ctx, const gmx_index *index, const gmx_params *params, int device);
void gmx_destroy(gmx_ctx *ctx);
const char *gmx_strerror(int code);
const char *gmx_last_error(const gmx_ctx *ctx);
int  gmx_abi_version(void);
/* Run all subsequent work of `ctx` on this cudaStream_t (default: a private stream). */
int  gmx_set_stream(gmx_ctx *ctx, void *cuda_stream);
int  gmx_synchronize(gmx_ctx *ctx);

/* ---- kernel-level entry points (each is also a stage of gmx_map_batch) -------------------- */

/* K1: GenomeBwt::get_sa_int (GenomeBwt.cpp:438-474) -> bwt_match_exact (bwt.c:222-239).
 * kmers: n x len ASCII.  Writes the inclusive SA interval [k,l], or (0,0) when absent / non-ACGT. */
int gmx_fm_search(gmx_ctx *ctx, const uint8_t *kmers, int32_t len, int64_t n,
                  uint64_t *k_out, uint64_t *l_out);

/* K1b: GenomeBwt::get_sa_coord (GenomeBwt.cpp:431-436) -> bwt_sa (bwt.c:86-97).
 * mode 0: one read of the de-sampled suffix array; mode 1: LF-walk over the sampled SA exactly as
 * bwt_sa does (validation path). */
int gmx_sa_locate(gmx_ctx *ctx, const uint64_t *ranks, int64_t n, int32_t mode, uint64_t *pos_out);

/* GenomeBwt::GetString (GenomeBwt.cpp:384-415): n windows of `size` bases as "acgt" chars;
 * len_out[i] = size, or 0 when the window crosses a sequence boundary / the genome end. */
int gmx_get_windows(gmx_ctx *ctx, const uint64_t *begin, int64_t n, int32_t size,
                    uint8_t *chars_out, int32_t *len_out);

/* a3: bin_seq::get_align_score(read, consensus, 0, n-1) (bin_seq.cpp:739-759,860-893). */
int gmx_self_score(gmx_ctx *ctx, const gmx_reads *reads, float *score_out);

/* K2a: bin_seq::get_align_score(read, gen) (bin_seq.cpp:761-850): banded PWM NW, score only.
 * Task t aligns read read_idx[t] (strand[t]: the PWM is reverse-complemented for NEG) against the
 * explicit window `windows + t*win_stride` (ASCII, length == read length). */
int gmx_nw_score(gmx_ctx *ctx, const gmx_reads *reads, int64_t n_tasks, const int32_t *read_idx,
                 const uint8_t *strand, const uint8_t *windows, int32_t win_stride, float *score_out);

/* K2b: bin_seq::get_align_score_w_traceback (bin_seq.cpp:445-718).  Consensus is the lower-case
 * max_char() consensus of the (strand-oriented) PWM (ScoredSeq.h:57-103) unless `consensus` is
 * given (n_tasks x win_stride ASCII).  aligned_out: n_tasks x aligned_stride bytes (NUL padded);
 * cigar_out: n_tasks x cigar_stride bytes (NUL terminated, forward order as the reference builds). */
int gmx_nw_traceback(gmx_ctx *ctx, const gmx_reads *reads, int64_t n_tasks, const int32_t *read_idx,
                     const uint8_t *strand, const uint8_t *windows, int32_t win_stride,
                     const uint8_t *consensus,
                     uint8_t *aligned_out, int32_t aligned_stride, int32_t *aligned_len_out,
                     char *cigar_out, int32_t cigar_stride);

/* K2c: bin_seq::pairHMM (bin_seq.cpp:60-244).  post_out: n_tasks x win_len x 5 floats (A,C,G,T,N). */
int gmx_pair_hmm(gmx_ctx *ctx, const gmx_reads *reads, int64_t n_tasks, const int32_t *read_idx,
                 const uint8_t *strand, const uint8_t *windows, int32_t win_stride, float *post_out);

/* ---- batch pipeline ----------------------------------------------------------------------- */

/* PHASE A for a whole batch: set_top_matches (Driver.cpp:432-612) = self score, k-mer walk +
 * backward search + locate + diagonal vote (align_seq2_raw.cpp:180-328), window fetch + banded NW
 * + acceptance + grouping + denominator (align_seq2_raw.cpp:22-178).  Results stay on the device
 * for gmx_score_batch; `results` (n_reads entries, may be NULL) receives the PHASE A fields. */
int gmx_map_batch(gmx_ctx *ctx, const gmx_reads *reads, gmx_read_result *results);

/* PHASE B for the batch last mapped: create_match_output (Driver.cpp:614-753) =
 * {Normal,BS,SNP}ScoredSeq::score -- traceback / pair-HMM, posterior, atomic scatter into the
 * genome accumulators -- and selection of the best group.  Fills the best_* fields. */
int gmx_score_batch(gmx_ctx *ctx, gmx_read_result *results);

/* gmx_map_batch + gmx_score_batch in one call with one D2H at the end. */
int gmx_process_batch(gmx_ctx *ctx, const gmx_reads *reads, gmx_read_result *results);

/* Accepted hits of the last batch, sorted by (read, group, pos, strand).  Call with hits == NULL
 * to obtain the count. */
int gmx_get_hits(gmx_ctx *ctx, gmx_hit *hits, int64_t capacity, int64_t *n_hits);

/* CIGAR (as get_SAM builds it, ScoredSeq.h:314-372, before strand reversal) and gapped aligned
 * string of the best group of each read of the last scored batch; stride bytes per read. */
int gmx_get_best_alignments(gmx_ctx *ctx, char *cigar_out, int32_t cigar_stride,
                            uint8_t *aligned_out, int32_t aligned_stride);

/* Device accumulators (float): amount_genome[ceil(l_pac/gen_size)] and, in BS/SNP modes, five
 * planes [l_pac] (A,C,G,T,N).  Exposed so the caller can run the final NCCL sum-reduce
 * (replacing the MPI block, Driver.cpp:1615-1811) directly on them. */
int gmx_accumulators_device(gmx_ctx *ctx, void **amount, uint64_t *n_amount,
                            void *planes[5], uint64_t *n_plane);
int gmx_reset_accumulators(gmx_ctx *ctx);

/* Before gGen.PrintFinal (Driver.cpp:1820-1823): synchronise and download the accumulators into
 * the host arrays GetGenomeAmtPtr() / GetGenome{A,C,G,T,N}Ptr() (GenomeBwt.h:199-209).
 * planes may be NULL in Normal mode.  On the root context of a gmx_comm the accumulators of all its contexts are
 * summed into the root first (replacing the MPI block, Driver.cpp:1615-1811). */
int gmx_finish(gmx_ctx *ctx, float *amount_genome, float *const planes[5]);

/* ---- several GPUs behind one process ------------------------------------------------
 * The reference's worker threads (`-c N`, src/Driver.cpp:1527-1554) share one set of accumulators under a mutex; its
 * MPI nodes sum theirs at the end (MPI_Allreduce of amount_genome src/Driver.cpp:1660-1672, MPI_Reduce of the five read
 * planes to rank 0 :1719-1767).  Here every GPU has its own context (worker thread t drives context t % n) and the
 * contexts' accumulators are terms of one sum.  gmx_comm_create ties n contexts (same genome, mode and gen_size; at
 * most 16) together; ctxs[0] is the root.  gmx_finish on the root first reduces into it, so the patched driver's call
 * sequence stays create / process ... / finish.
 *   GMX_COMM_PEER  the library's own reduce kernels over peer memory: GPU g sums slice g of every context's arrays with
 *                  direct NVLink loads, in context order (deterministic), and stores the sums into the root's memory.
 *                  Also works for contexts that share a device.
 *   GMX_COMM_NCCL  ncclCommInitAll over the contexts' devices + ncclReduce (ncclAllReduce when all != 0); libnccl.so.2
 *                  is opened at run time.  Needs distinct devices.
 *   GMX_COMM_AUTO  = GMX_COMM_PEER.
 * gmx_comm_reduce(comm, 0) leaves the sum in the root and zeroes the other contexts' accumulators (a later reduce adds
 * only what is new); gmx_comm_reduce(comm, 1) leaves the sum in every context (terminal: do not accumulate further).
 * One process per GPU (torchrun / MPI) does not need a gmx_comm: run the collective of your launcher on
 * gmx_accumulators_device(). */
typedef struct gmx_comm gmx_comm;
#define GMX_COMM_AUTO 0
#define GMX_COMM_PEER 1
#define GMX_COMM_NCCL 2
int  gmx_comm_create(gmx_comm **comm, gmx_ctx *const *ctxs, int n, int backend);
int  gmx_comm_reduce(gmx_comm *comm, int all);
/* CUDA-event time and bytes per context of the last reduce, and the backend in use. */
int  gmx_comm_stats(const gmx_comm *comm, float *ms, uint64_t *bytes, int *backend);
void gmx_comm_destroy(gmx_comm *comm);      /* before gmx_destroy of its contexts */

/* ---- next row: FASTQ text -> reads -------------------------------------------
 * SeqReader::get_more_fastq (reference src/SeqReader.cpp:1023-1292).  One record per read, offsets into the text. */
typedef struct gmx_fastq_rec {
    int64_t name_off;          /* read name without the '@' (Read::name, SeqReader.cpp:1263)          */
    int64_t seq_off;           /* Read::seq                                                           */
    int64_t qual_off;          /* Read::fq (may be longer than the sequence; the PWM uses seq_len of it) */
    int32_t name_len, seq_len, qual_len, pad;
} gmx_fastq_rec;

/* Host restatement with the reference's recovery from malformed records.  Returns GMX_ERR_FORMAT when a quality
 * character maps to Q < 0 (the reference throws "Invalid Fastq Character"), GMX_ERR_OVERFLOW when capacity is short
 * (*n_recs then holds the count).  Needs no GPU. */
int gmx_fastq_scan_host(const char *text, int64_t len, int illumina, gmx_fastq_rec *recs, int64_t capacity, int64_t *n_recs);

/* The same index on the GPU for well-formed text (4 lines per record, '@' / '+' in place, len(qual) >= len(seq)):
 * newline compaction + one thread per record.  text may be host (copied in) or device memory.  Any record that would
 * need the reference's recovery path gives GMX_ERR_FORMAT: fall back to gmx_fastq_scan_host.  recs may be NULL. */
int gmx_fastq_scan(gmx_ctx *ctx, const char *text, int64_t len, int text_on_device, gmx_fastq_rec *recs, int64_t capacity, int64_t *n_recs);

/* gmx_fastq_scan + gmx_process_batch with the reads used in place inside the (device copy of the) text.  A large host text
 * is pipelined piece by piece (GMX_OPT_FASTQ_PIECE): should a LATER piece turn out malformed, the call returns GMX_ERR_FORMAT
 * with *n_reads = the reads before that piece -- they have been mapped and scored, results / recs hold them -- and the
 * caller continues with gmx_fastq_scan_host from the end of record *n_reads - 1 (*n_reads == 0: nothing was done). */
int gmx_process_fastq(gmx_ctx *ctx, const char *text, int64_t len, int text_on_device, gmx_read_result *results, int64_t capacity,
                      int64_t *n_reads, gmx_fastq_rec *recs);

/* ---- next row: SAM emission ------------------------------------------------------
 * ScoredSeq::get_SAM (reference inc/ScoredSeq.h:293-404) + the SAM writer (src/Driver.cpp:2146-2217): the body
 * lines of the last scored batch, in read order, one per (position, strand) of each read's best group, byte for
 * byte as the reference prints them.  Works with GMX_OPT_COLLECT_HITS 0 or 1.  After gmx_process_fastq with
 * GMX_OPT_COLLECT_HITS = 0 the records are formatted on the GPU (text, record index, results and CIGARs are resident
 * there) and only the finished text crosses to `out`; any other batch is formatted by host threads.  names / seq / qual come from
 * `recs` into `text` (gmx_fastq_rec: name_off/len, seq_off/len, qual_off/qual_len); chrom_names[i] is the name of
 * sequence i of the index.  Returns GMX_ERR_OVERFLOW with *len = bytes needed when `cap` is short. */
int gmx_format_sam(gmx_ctx *ctx, const char *text, const gmx_fastq_rec *recs, const gmx_read_result *results, int64_t n_reads,
                   const char *const *chrom_names, char *out, int64_t cap, int64_t *len);

/* "%g" of a double exactly as printf prints it -- the writer the device SAM formatter uses for XA:f / XP:f (the reference
 * streams floats with the default ostream format), exported so that it can be checked on the host.  Returns the length,
 * GMX_ERR_UNSUPPORTED for values outside ~1e-17 .. 1e21 (not finite included), GMX_ERR_OVERFLOW when cap is short. */
int gmx_format_g(double v, char *out, int cap);

/* ---- next row: .sgr output (Normal-mode part) ----------------------------------
 * GenomeBwt::PrintFinalSGR (reference src/GenomeBwt.cpp:1212-1273): one line "chrom\tpos\t%.5f" per accumulator bin
 * whose value exceeds min_print (the reference's MIN_PRINT is the double 0.001; the float is compared in double as
 * there), in genome order, from the accumulators as they
 * stand on the device (after the caller's NCCL reduce, if any).  The printable bins are selected on the device.
 * Returns GMX_ERR_OVERFLOW with *len = bytes needed when `cap` is short. */
int gmx_format_sgr(gmx_ctx *ctx, const char *const *chrom_names, double min_print, char *out, int64_t cap, int64_t *len);

/* ---- next row: .gmp output with the SNP call (SNP / bisulfite / A->G part) -------
 * GenomeBwt::PrintFinalSNP (reference src/GenomeBwt.cpp:930-1005) in GMX_MODE_SNP: one row
 * "chrom\tpos\t%.5f" + the five read planes as "\t%.5f" + the call column of PrintSNPCall (:1011-1092) for every
 * position whose amount exceeds min_print (MIN_PRINT 0.001 there).  The call is the reference's likelihood-ratio
 * test -- dipLRT (:760-873), or LRT (:739-755) when snp_monoploid (--snp_monop) -- against snp_pval (--snp_pval,
 * gSNP_PVAL, default 0.001); the chi-square CDF the reference takes from GSL is computed by the library.
 * GenomeBwt::PrintFinalBisulfite (:1094-1205) in GMX_MODE_BS: "chrom\tpos\t%f" + five "\t%.5f" for every position
 * with amount > 0 whose genome base is target_base (0..3 = a,c,g,t; the reference reports 'c' for -b on the + strand,
 * 'g' for --b2 / - strand only, 'a' / 't' for A->G on the + / - strand, :1136-1160); min_print, snp_pval and
 * snp_monoploid are ignored there and target_base is ignored in SNP mode.
 * Rows are selected and gathered on the device from the accumulators as they stand (after the caller's NCCL reduce, if
 * any); the host formats them on many threads.  GMX_ERR_STATE in Normal mode; GMX_ERR_OVERFLOW with *len = bytes
 * needed when `cap` is short. */
int gmx_format_gmp(gmx_ctx *ctx, const char *const *chrom_names, int target_base, double min_print, float snp_pval, int snp_monoploid,
                   char *out, int64_t cap, int64_t *len);

/* GenomeBwt::is_snp (reference src/GenomeBwt.cpp:874-898: LRT :739-755 when snp_monoploid, else dipLRT :760-873) for
 * the read counts (A,C,G,T,N) of one position, and the call column PrintSNPCall (:1011-1092) prints for it when the
 * genome holds base `genome_base` (0..4 = a,c,g,t,n) there: "\tN", "\t[YN]:g->x p_val=%.2e" or
 * "\t[YN]:g->x/y p_val=%.2e".  Pure host arithmetic (no context, no device): it is what gmx_format_gmp runs per row.
 * first / second = most and second most supported base (second = -1 when the test was monoploid-only), any output
 * pointer may be NULL; text needs 48 bytes. */
int gmx_snp_call(const float counts[5], int genome_base, int snp_monoploid, float snp_pval, int *first, int *second, int *diploid,
                 double *pval, char *text, int text_cap);

/* ---- next row: index construction ------------------------------------------------
 * bwa_index (reference src/bwtindex.c:187-293) on the GPU, from the base codes (0..3, one byte per base, the sequences of
 * the FASTA concatenated; the synthetic genomes hold no N) to what the reference keeps in `.gnumap.bwt` / `.sa` / `.pac`:
 * the occ-interleaved BWT (bwt_bwtupdate_core :128-150), primary, L2, the 1/32-sampled suffix array (bwt_cal_sa,
 * src/bwt.c:62-84; sa[0] = (uint64_t)-1) and the 2-bit pac.  The suffixes are sorted by prefix doubling over radix sorts
 * (csrc/index_build.cuh); the BWT of a text is unique, so the arrays are the reference's bit for bit.  Needs no context;
 * output arrays are host memory sized by gmx_index_sizes.  *rounds (may be NULL) = sorting rounds used. */
int gmx_index_sizes(int64_t l_pac, uint64_t *bwt_words, uint64_t *n_sa, uint64_t *pac_bytes);
int gmx_index_build(const uint8_t *codes, int64_t l_pac, int device, uint32_t *bwt, uint64_t *primary, uint64_t L2[5],
                    uint64_t *sa, uint8_t *pac, int32_t *rounds, char *err, int err_cap);

/* ---- options ----------------------------------------------------------------------------- */
#define GMX_OPT_COLLECT_HITS 1   /* 1 (default): keep every accepted (pos,strand) for gmx_get_hits; 0: only the
                                    per-read results and the best group's CIGAR leave the device           */
#define GMX_OPT_CHUNK_READS  2   /* reads processed per internal chunk (default 524288)                     */
#define GMX_OPT_VOTE_FILTER  3   /* 1 (default): counting-filter + exact-verification vote kernel with the exact
                                    hash-table kernels as its overflow path; 0: exact hash tables for every task  */
#define GMX_OPT_FILTER_SHIFT 4   /* tuning: log2 scale of the kmin == 2 vote filter (default 0 = 4 bytes per SA hit)     */
#define GMX_OPT_CIGAR_STRIDE  5   /* bytes per CIGAR slot of the batch pipeline (default 64, a multiple of 16 up to 2048).  The
                                    reference builds CIGARs unbounded (TopReadOutput::CIGAR holds MAX_CIGAR_SZ = 1024); a batch in
                                    which some alignment needs more text than the slot fails with GMX_ERR_OVERFLOW instead of
                                    returning a cut string -- after the batch has reached the accumulators, so a caller that
                                    cannot rule long CIGARs out sizes the slot up front: 4 * read length + 16 bytes hold any
                                    alignment (integration/gnumap_gmx_bridge.cpp does that per slice) */
#define GMX_OPT_VOTE_SLOTS   6   /* tuning: 32-hit slots per step of the vote kernel, 4 or 6 (default: from seq_len / 4^mer)  */
#define GMX_OPT_FASTQ_PIECE  9   /* bytes per piece of a host FASTQ text in gmx_process_fastq (default 96 MiB; 0 = upload and index the
                                    whole text first).  Texts of two pieces or more are cut at record boundaries and piece p + 1
                                    crosses PCIe and is indexed while piece p is mapped; the first piece is a quarter of the
                                    regular size (its upload is the exposed one)                                                */
#define GMX_OPT_SAM_DEVICE   8   /* 1 (default): gmx_format_sam formats the batch of the last gmx_process_fastq on the GPU (the text, the
                                    record index, the results and the CIGARs are resident there); 0: always the host formatter     */
#define GMX_OPT_VOTE_COMPACT 7   /* tuning: occupancy variants of the vote kernel for tasks of <= 32 k-mers: 0 off, 1 two bits per
                                    diagonal (24 warps / SM), 2 (default) three bits (32 or 24 warps / SM by hits per task)         */
#define GMX_OPT_OPTIMISTIC   10  /* 1 (default): chunks of gmx_process_batch / gmx_process_fastq are issued without a host wait, over
                                    candidate / group-leader bounds predicted from the previous chunk; the counters are looked at one
                                    chunk later and a chunk that did not fit its bounds is run again the synchronous way (results are
                                    the same either way).  0: every chunk waits for its counts.  2: testing -- bounds that are too
                                    small on purpose, so that every optimistic chunk is run again                              */
#define GMX_OPT_STAGE_TIMING 11  /* 1 (default): a CUDA-event pair around every stage of every chunk feeds gmx_get_stage_stats; 0: no
                                    events are recorded (units / bytes / launches are still counted, ms stay 0)                 */
int gmx_set_option(gmx_ctx *ctx, int option, int64_t value);
/* chunks issued optimistically / of those, run again -- since the context was created */
int gmx_chunk_stats(gmx_ctx *ctx, uint64_t *optimistic, uint64_t *rerun);

/* ---- instrumentation ---------------------------------------------------------------------- */
#define GMX_N_STAGES 12
typedef struct gmx_stage_stats {
    const char *name[GMX_N_STAGES];
    float    ms[GMX_N_STAGES];        /* CUDA-event time of each stage, last batch              */
    uint64_t units[GMX_N_STAGES];     /* work units of each stage (lookups, hits, cells, ...)   */
    uint64_t bytes[GMX_N_STAGES];     /* algorithmic bytes of each stage  */
    int32_t  launches[GMX_N_STAGES];  /* kernel launches of each stage                          */
    int32_t  n_stages;
} gmx_stage_stats;
int gmx_get_stage_stats(gmx_ctx *ctx, gmx_stage_stats *out);

/* Measured ALU ceiling of the context's GPU for the roofline of the NW kernels (a mul/add micro-kernel, best of five):
 * kind 0 = FP32 multiply + add WITHOUT fused multiply-add (what the bit-exact K2a / K2b issue), 1 = FP32 add + max,
 * 2 = FP64 multiply + add without FMA (K2c's pipe).  *tera_ops = 1e-12 x lane-operations per second. */
int gmx_measure_alu_peak(gmx_ctx *ctx, int kind, double *tera_ops);

#ifdef __cplusplus
}
#endif
#endif /* GMX_H */
