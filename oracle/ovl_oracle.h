/*
 *  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 *  CPU restatement (plain C) of marbl/canu's overlapInCore ("ovl") hot path,
 *  used ONLY as the parity checker by tests/, __graft_entry__.smoke() and
 *  bench.py's cpu_baseline / --impl reference legs.  Nothing under canu_b200/
 *  may include, link or call this.
 *
 *  Parity status: PINNED.  The reference ships no golden vectors for this path
 *  (SURVEY.md 8c), so the restatement is pinned against outputs of the
 *  reference binary itself, built unmodified by oracle/build_ref.sh into
 *  oracle/_ref/ and run in this container; the resulting fixtures are committed
 *  under tests/golden/ together with the generating script
 *  (tests/golden/make_golden.py).  tests/test_oracle_golden.py checks this file
 *  against every one of them.
 *
 *  Reference files restated (paths under /root/reference/src/overlapInCore):
 *    overlapInCore.C                         constants, hash shifts        (:26-35, :454-459)
 *    overlapInCore-Build_Hash_Index.C        table build, skip k-mers      (:28-404, :415-631)
 *    overlapInCore-Find_Overlaps.C           lookup, Add_Ref, Add_Match    (:26-336)
 *    overlapInCore-Process_String_Overlaps.C per-pair control, merge rules (:22-690)
 *    overlapInCore-Process_Overlaps.C        per-read driver               (:25-122)
 *    overlapInCore-Output.C                  ovOverlap record building     (:27-264)
 *    liboverlap/prefixEditDistance*.C        banded extension + traceback
 *    liboverlap/Binomial_Bound.C             Edit_Match_Limit table
 *    ../stores/ovOverlap.H                   24-byte record bit layout     (:34-78)
 */
#ifndef OVL_ORACLE_H
#define OVL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OVO_MAX_READLEN_BITS 21
#define OVO_MAX_READLEN      ((1u << OVO_MAX_READLEN_BITS) - 1)

typedef struct {
  uint32_t kmer_len;          /* -k N                                             */
  double   max_erate;         /* --maxerate, ALREADY rounded through float (strtof) */
  double   align_noise;       /* --alignnoise, likewise; default 1.0              */
  int32_t  partial;           /* -partial                                         */
  int32_t  unique_per_pair;   /* -u (1, default) / -m (0)                         */
  int32_t  min_olap_len;      /* --minlength                                      */
  int32_t  no_hopeless;       /* -z                                               */
  int32_t  min_kmers;         /* --minkmers flag                                  */
  uint32_t hash_bits;         /* --hashbits (default 22)                          */
  double   hash_load;         /* --hashload (default 0.6)                         */
  uint64_t hash_data_len;     /* --hashdatalen (default 100000000)                */
} ovo_params;

/* One overlap record, in the reference's in-memory ovOverlap layout
   (a_iid, b_iid, two 64-bit words; stores/ovOverlap.H:49-78). */
typedef struct {
  uint32_t a_iid, b_iid;
  uint64_t w0, w1;
} ovo_record;

typedef struct {
  uint64_t kmer_hits_without_olap;
  uint64_t kmer_hits_with_olap;
  uint64_t kmer_hits_skipped;
  uint64_t multi_overlap;
  uint64_t total_overlaps;
  uint64_t contained;
  uint64_t dovetail;
  /* instrumentation (SURVEY.md 8d definitions) */
  uint64_t extend_calls;      /* forward()+reverse() invocations that ran a DP     */
  uint64_t dp_cells;          /* inner `for d` bodies evaluated                    */
  uint64_t char_compares;     /* iterations of the slide `while`                   */
  uint64_t hash_inserts;      /* k-mers put in the table                           */
  uint64_t ref_lookups;       /* ref windows tested                                */
  uint64_t seed_hits;         /* Add_Ref calls                                     */
} ovo_stats;

/* Per candidate pair trace (kernel-granularity goldens). */
typedef struct {
  uint32_t ref_id, hash_id;
  int32_t  dir;               /* 0 forward, 1 reverse                              */
  int32_t  consistent;
  int32_t  diag_ct, diag_bgn, diag_end;
  int32_t  n_seeds;           /* seeds follow in seed arrays, list order           */
  int64_t  seed_begin;        /* index of first seed in the seed trace             */
} ovo_pair_trace;

typedef struct { int32_t start, offset, len; } ovo_seed;

/* Per Extend_Alignment call trace. */
typedef struct {
  uint32_t ref_id, hash_id;
  int32_t  dir;
  int32_t  seed_start, seed_offset, seed_len;
  int32_t  s_lo, s_hi, t_lo, t_hi, errors, kind, delta_ct;
} ovo_ext_trace;

typedef struct ovo_ctx ovo_ctx;

ovo_ctx *ovo_create(const ovo_params *p);
void     ovo_destroy(ovo_ctx *c);

/* Reads 1..n (ID = index+1) as ASCII in one buffer; any case; len 0 = absent read.
   lib_ids may be NULL. The buffer is copied. */
int      ovo_set_reads(ovo_ctx *c, uint32_t n, const char *bases, const uint64_t *offsets, const uint32_t *lens);

/* Skip k-mers: n strings of kmer_len chars, concatenated (no separators). */
int      ovo_set_skip_kmers(ovo_ctx *c, uint32_t n, const char *kmers);

/* Enable traces (0/1). */
void     ovo_enable_trace(ovo_ctx *c, int pairs, int exts);

/* Run -h hb-he -r rb-re (inclusive, 1-based, clamped to the read count) with
   `threads` OpenMP threads (output order unspecified, like the reference). */
int      ovo_run(ovo_ctx *c, uint32_t hb, uint32_t he, uint32_t rb, uint32_t re, int threads);

uint64_t          ovo_num_records(const ovo_ctx *c);
const ovo_record *ovo_records(const ovo_ctx *c);
void              ovo_get_stats(const ovo_ctx *c, ovo_stats *s);

uint64_t              ovo_num_pair_traces(const ovo_ctx *c);
const ovo_pair_trace *ovo_pair_traces(const ovo_ctx *c);
const ovo_seed       *ovo_seed_traces(const ovo_ctx *c);
uint64_t              ovo_num_ext_traces(const ovo_ctx *c);
const ovo_ext_trace  *ovo_ext_traces(const ovo_ctx *c);

/* Host tables (liboverlap/prefixEditDistance.C:23-107, Binomial_Bound.C:104-188). */
uint32_t       ovo_max_errors(const ovo_ctx *c);
const int32_t *ovo_edit_match_limit(const ovo_ctx *c);       /* ovo_max_errors() entries */
int32_t        ovo_error_bound(const ovo_ctx *c, int32_t len);
double         ovo_branch_match_value(const ovo_ctx *c);

/* Stand-alone extension of one seed (strings are lower-cased copies made inside). */
int      ovo_extend_one(ovo_ctx *c, const char *S, int32_t s_len, const char *T, int32_t t_len,
                        int32_t seed_start, int32_t seed_offset, int32_t seed_len,
                        ovo_ext_trace *out, int32_t *delta_out, int32_t delta_cap);

#ifdef __cplusplus
}
#endif
#endif
