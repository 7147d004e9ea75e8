/*
 * pf_types.h -- plain-C data layouts shared by the C-ABI (include/pf_gpu.h), the CPU oracle
 * (oracle/) and the reference shim (oracle/ref_shim.cpp).  No torch / CUDA types appear here.
 *
 * Flat batch conventions
 * ----------------------
 *  - A *sequence batch* is one char array `bases` plus `seq_off[n_seq+1]` (byte offsets, ascending);
 *    sequence s is bases[seq_off[s] .. seq_off[s+1]).  Characters are the ones the reference accepts:
 *    ACGT/acgt are symbols, anything else is an "N" for the KMC path (kmer_api.h:264-275) and an
 *    ordinary character for SeqAlign (which only compares chars; '-' is the gap, SeqAlign.cpp:498-506).
 *  - A *bubble batch* adds `bubble_off[n_bubbles+1]`, a CSR from bubbles to consecutive sequences,
 *    already in the order the caller wants them aligned (the reference's callers sort first,
 *    CDBG.cpp:2035, :2263).
 */
#ifndef PF_TYPES_H
#define PF_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirror of CKMCFileInfo (KMC/kmc_api/kmc_file.h:19-30) plus the on-disk dialect. */
typedef struct pf_kmc_info {
    uint32_t kmer_length;
    uint32_t mode;
    uint32_t counter_size;
    uint32_t lut_prefix_length;
    uint32_t signature_len;
    uint32_t min_count;
    uint64_t max_count;
    uint64_t total_kmers;
    uint32_t both_strands; /* 1 = canonical DB (kmc run without -b)                      */
    uint32_t kmc_version;  /* 0 = KMC1 layout, 0x200 = KMC2 layout (kmc_file.cpp:190)   */
    uint32_t n_bins;       /* KMC2: number of bins (LUT slices); KMC1: 1                 */
    uint32_t reserved;
} pf_kmc_info_t;

/* How a k-mer window is turned into the key that is searched. */
enum {
    PF_LOOKUP_CANONICAL = 0,   /* min(kmer, revcomp) -- GetCountersForRead on a both-strands DB (kmc_file.cpp:1060,1290) */
    PF_LOOKUP_FWD_THEN_RC = 1, /* as written, else reverse complement -- PloidyFrost readCov (CDBG.cpp:38-43)            */
    PF_LOOKUP_FWD = 2          /* as written only -- CheckKmer / strand-specific DB (kmc_file.cpp:330, CDBG.cpp:99-117)   */
};

/* Per-sequence coverage reduction (the readCov semantics, CDBG.cpp:29-120). */
typedef struct pf_cov {
    uint64_t sum;          /* sum of counters of all found k-mers                                         */
    uint32_t min;          /* min counter, initialised to 10000 like CDBG.cpp:71                           */
    uint32_t n_kmers;      /* number of k-mer windows = max(len-k+1, 0)                                     */
    int32_t first_missing; /* index of first window whose k-mer is absent (or non-ACGT); -1 if none        */
    int32_t first_outside; /* index of first window whose count is not low < c < up (strict); -1 if none   */
} pf_cov_t;

/*
 * Result of SeqAlign::SequenceAlignment (SeqAlign.cpp:550) for a batch of bubbles.  All arrays are
 * owned by whoever produced the batch and stay valid until that producer's next call / free.
 *
 * partition (vector<vector<unsigned short>>, one entry per column) is returned sparsely: only columns
 * with a non-zero class vector are listed (these are exactly the columns with partition[c].back() > 0,
 * because compareStrPair numbers either every row or none, SeqAlign.cpp:75-92,121-145).
 *   snp_pos   = var_col where var_kind == 0
 *   indel_pos = var_col where var_kind == 1
 *   var_kind == 2 : numbered column that is in neither list (a continued indel column that shows
 *                   more than two symbols, SeqAlign.cpp:121)
 */
typedef struct pf_msa_batch {
    uint32_t n_bubbles;
    uint32_t reserved;
    const int32_t *status;    /* [n] 0 = ok; >0 = PF_BUBBLE_* capacity/limit code (no result for that bubble)   */
    const uint32_t *n_rows;   /* [n] rows of the chosen MSA; 0 = empty alignment (reference returns str empty)  */
    const uint32_t *aln_len;  /* [n] columns                                                                     */
    const uint64_t *rows_off; /* [n+1] char offsets into rows; bubble b holds n_rows*aln_len chars, row-major   */
    const char *rows;
    const uint64_t *var_off;  /* [n+1] offsets into var_col / var_kind                                           */
    const uint32_t *var_col;
    const uint8_t *var_kind;
    const uint64_t *cls_off;  /* [n+1] offsets into cls; bubble b holds n_var*n_rows entries, [var][row]         */
    const uint16_t *cls;      /* 1-based class ids in order of first appearance                                   */
    const uint64_t *ilen_off; /* [n+1] offsets into ilen                                                          */
    const uint32_t *ilen;     /* indel_len_vec (may be shorter than the indel list: a trailing gap never closes) */
} pf_msa_batch_t;

/* per-bubble status codes (status[] above) */
enum {
    PF_BUBBLE_OK = 0,
    PF_BUBBLE_TOO_MANY_ROWS = 1,  /* more sequences than the device path supports                     */
    PF_BUBBLE_TOO_LONG = 2,       /* DP matrix does not fit the device work area                      */
    PF_BUBBLE_CAND_OVERFLOW = 3,  /* more co-optimal alignments than the device arena can hold        */
    PF_BUBBLE_STEP_LIMIT = 4,     /* traceback exceeded the step budget (combinatorial explosion)     */
    PF_BUBBLE_BAD_INPUT = 5       /* fewer than 2 sequences, or a sequence contains a literal '-'     */
};

/*
 * Lookup phase B (pf_site_cov): the allele-class coverages of the variable columns of aligned bubbles, from the site k-mers of
 * CDBG.cpp:2295-2509.  Sites are the variable columns of the MSA batch in its order (site_off == var_off of that batch);
 * cov holds, per site, n_rows(bubble) entries of which the first n_class are the class coverages (cov_off == cls_off).
 */
typedef struct pf_site_batch {
    uint32_t n_bubbles;
    uint32_t reserved;
    const uint64_t *site_off; /* [n+1]                                                                      */
    const uint8_t *status;    /* per site: PF_SITE_*                                                        */
    const uint8_t *n_class;   /* per site: number of allele classes (max class id of the column)            */
    const uint64_t *cov_off;  /* [n+1]                                                                      */
    const uint64_t *cov;      /* class coverage = sum of the counters of the class's distinct site k-mers   */
} pf_site_batch_t;

/* Site k-mers without lookups (pf_site_kmers): for every variable column of every bubble of the last alignment and every
 * row, the k-mer the reference forms at that column (CCDBG.cpp:1057-1241 / CDBG.cpp:2295-2509) as a right-aligned 2-bit
 * value (A=0 C=1 G=2 T=3, first base in the highest bits).  keys of bubble b start at key_off[b]; column j (site_off[b] + j)
 * owns n_rows[b] consecutive keys.  status: PF_SITE_OK, PF_SITE_UNDEFINED or PF_SITE_SKIPPED per column. */
typedef struct pf_site_kmers {
    uint32_t n_bubbles;
    uint32_t reserved;
    const uint64_t *site_off; /* [n+1]  == var_off of the alignment                                         */
    const uint64_t *key_off;  /* [n+1]  == cls_off of the alignment                                         */
    const uint64_t *keys;     /* one per (variable column, row)                                             */
    const uint8_t *status;    /* per variable column                                                        */
} pf_site_kmers_t;

enum {
    PF_SITE_OK = 0,
    PF_SITE_DROPPED = 1,   /* a counter is not inside (low, up): the reference skips the site (CDBG.cpp:2415-2418)            */
    PF_SITE_MISSING = 2,   /* a site k-mer is not in the database: the reference prints it and exits (CDBG.cpp:52-56)         */
    PF_SITE_UNDEFINED = 3, /* the reference itself would read outside a row here (substr throws / never returns)             */
    PF_SITE_SKIPPED = 4    /* the caller asked not to compute this bubble (strict bubbles use the branch means instead)       */
};

#ifdef __cplusplus
}
#endif
#endif /* PF_TYPES_H */
