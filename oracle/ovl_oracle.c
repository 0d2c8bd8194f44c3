/*
 *  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See ovl_oracle.h for the header
 *  comment (scope, parity status "pinned", and the reference files restated).
 *
 *  This is a from-scratch CPU restatement of the reference algorithm in our own
 *  structure; every function cites the reference file:line it follows
 *  (paths relative to /root/reference/src/overlapInCore unless noted).
 *
 *  Build:  gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC ovl_oracle.c -o libovl_oracle.so -lm
 *  (-ffp-contract=off: the reference is built without -march, i.e. without FMA;
 *   the branch-point score `Row * Branch_Match_Value - e` must round twice.)
 */
#include "ovl_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <ctype.h>
#include <limits.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- constants (overlapInCore.H:58-160, overlapInCore.C:26-35) ---- */
#define ENTRIES_PER_BUCKET   21
#define HASH_CHECK_MASK      0x1fu
#define CHECK_MASK           0xffu
#define PROBE_MASK           0x3eu
#define HOPELESS_MATCH       90
#define MAX_DISTINCT_OLAPS   3
#define MIN_INTERSECTION     10
#define SHIFT_SLACK          1
#define STRING_OLAP_SHIFT    8
#define STRING_OLAP_MODULUS  (1 << STRING_OLAP_SHIFT)
#define STRING_OLAP_MASK     (STRING_OLAP_MODULUS - 1)
#define MIN_BRANCH_END_DIST  20

#define STRING_NUM_BITS      31
#define OFFSET_BITS          31
#define STRING_NUM_MASK      ((1ull << STRING_NUM_BITS) - 1)
#define OFFSET_MASK          ((1ull << OFFSET_BITS) - 1)
#define BIT_EMPT             62
#define BIT_LAST             63

typedef uint64_t sref_t;                         /* String_Ref_t, overlapInCore.H:283-313 */
#define SR_NUM(x)    ((x) & STRING_NUM_MASK)
#define SR_OFF(x)    (((x) >> STRING_NUM_BITS) & OFFSET_MASK)
#define SR_EMPTY(x)  (((x) >> BIT_EMPT) & 1ull)
#define SR_LAST(x)   (((x) >> BIT_LAST) & 1ull)
#define SR_MAKE(num, off)  (((uint64_t)(num) & STRING_NUM_MASK) | (((uint64_t)(off) & OFFSET_MASK) << STRING_NUM_BITS))
#define SR_SET_EMPTY(x)    ((x) |  (1ull << BIT_EMPT))
#define SR_SET_LAST(x)     ((x) |  (1ull << BIT_LAST))
#define SR_CLR_LAST(x)     ((x) & ~(1ull << BIT_LAST))

typedef struct {                                 /* Hash_Bucket_t, overlapInCore.H:316-321 */
  sref_t   entry[ENTRIES_PER_BUCKET];
  uint8_t  check[ENTRIES_PER_BUCKET];
  uint8_t  hits[ENTRIES_PER_BUCKET];
  int16_t  count;
} bucket_t;

typedef struct {                                 /* Hash_Frag_Info_t, overlapInCore.H:323-327 */
  uint32_t length;
  uint8_t  lscreen, rscreen;
} finfo_t;

typedef struct { int32_t offset, len, start, next; } mnode_t;   /* Match_Node_t, prefixEditDistance.H:57-62 */

typedef struct {                                 /* String_Olap_t, overlapInCore.H:194-204 */
  uint32_t string_num;
  int32_t  match_list;
  double   diag_sum;
  int32_t  diag_ct, diag_bgn, diag_end;
  int32_t  next;
  uint8_t  full, consistent;
} solap_t;

typedef struct {                                 /* Olap_Info_t, overlapInCore.H:207-216 (delta dropped: only delta_ct is consumed) */
  int s_lo, s_hi, t_lo, t_hi;
  double quality;
  int delta_ct;
  int s_left_boundary, s_right_boundary, t_left_boundary, t_right_boundary;
  int min_diag, max_diag;
} oinfo_t;

/* Per-thread extension state (class prefixEditDistance, prefixEditDistance.H:67-326) */
typedef struct {
  int32_t **row;            /* row[e] points at diagonal 0 of error row e; valid d in [-e-2, e+2] */
  int32_t   rows_alloc;
  int32_t  *left_delta, *right_delta, *stack;
  int32_t   left_delta_len, right_delta_len;
  uint64_t  calls, cells, compares;
} ped_t;

typedef struct {                                 /* Work_Area_t, overlapInCore.H:222-274 */
  solap_t  *olap;   int32_t olap_size, olap_next;
  mnode_t  *node;   int32_t node_size, node_next;
  int       left_end_screened, right_end_screened;
  ped_t     ped;
  oinfo_t   distinct[MAX_DISTINCT_OLAPS];
  ovo_stats st;
  ovo_record *rec;  uint64_t rec_len, rec_max;
  ovo_pair_trace *pt; uint64_t pt_len, pt_max;
  ovo_seed       *sd; uint64_t sd_len, sd_max;
  ovo_ext_trace  *et; uint64_t et_len, et_max;
  char     *fwd, *rev;      /* lower-cased ref read, both orientations */
} work_t;

struct ovo_ctx {
  ovo_params P;
  uint64_t HSF1, HSF2, SV1, SV2, SV3, hash_mask, table_size;
  int      use_hopeless;
  uint64_t filter_by_kmer_count;
  uint64_t max_hash_data_len;

  uint32_t max_errors;
  int32_t *eml;
  double   branch_match_value, min_branch_tail_slope;

  uint32_t  n_reads;
  char     *bases;  uint64_t *off;  uint32_t *len;

  uint32_t  n_skip;  char *skip;

  /* the reference's process globals (overlapInCore.C:46-99) */
  bucket_t *table;   uint32_t *check_array;
  char     *data;    uint64_t data_len, used_data_len, extra_data_len;
  sref_t   *next_ref;
  sref_t   *extra_ref; uint64_t extra_ref_ct, extra_ref_max;
  uint64_t  string_ct, extra_string_ct, extra_string_subcount, hash_string_num_offset;
  int64_t  *string_start; uint64_t string_start_size;
  finfo_t  *string_info;  uint64_t string_info_size;
  uint64_t  hash_entries;

  int trace_pairs, trace_exts;

  ovo_record *rec;  uint64_t rec_len, rec_max;
  ovo_stats   st;
  ovo_pair_trace *pt; uint64_t pt_len, pt_max;
  ovo_seed       *sd; uint64_t sd_len, sd_max;
  ovo_ext_trace  *et; uint64_t et_len, et_max;
};

static int bit_equiv(int ch) {                   /* Bit_Equivalent, overlapInCore.C:500-503 */
  switch (ch) { case 'a': case 'A': return 0; case 'c': case 'C': return 1;
                case 'g': case 'G': return 2; case 't': case 'T': return 3; }
  return 0;
}
static int char_is_bad(int ch) {                 /* Char_Is_Bad, overlapInCore.C:505-512 */
  ch = tolower(ch);
  return !(ch == 'a' || ch == 'c' || ch == 'g' || ch == 't');
}

#define HASH_FN(c,k)    ((((k) ^ ((k) >> (c)->HSF1) ^ ((k) >> (c)->HSF2))) & (c)->hash_mask)     /* overlapInCore.H:174 */
#define HCHECK_FN(c,k)  ((((k) ^ ((k) >> (c)->SV1) ^ ((k) >> (c)->SV2))) & HASH_CHECK_MASK)      /* :177 */
#define KCHECK_FN(c,k)  ((((k) ^ ((k) >> (c)->SV1) ^ ((k) >> (c)->SV3))) & CHECK_MASK)           /* :180 */
#define PROBE_FN(c,k)   (((((k) ^ ((k) >> (c)->SV2) ^ ((k) >> (c)->SV3))) & PROBE_MASK) | 1)     /* :183 */

/* ======================================================================= */
/*  Edit_Match_Limit  (liboverlap/Binomial_Bound.C:36-188)                 */
/* ======================================================================= */

#define NORMAL_DISTRIB_THOLD  3.62
#define EDIT_DIST_PROB_BOUND  1e-4

static int binomial_bound(int e, double p, int start) {       /* Binomial_Bound.C:36-99 */
  double q = 1.0 - p;
  if (start < e) start = e;
  for (int n = start; n < (int)OVO_MAX_READLEN; n++) {
    if (n <= 35) {
      double sum = 0.0, p_power = 1.0, q_power = pow(q, n);
      int bin_coeff = 1, ct = 0;
      for (int k = 0; k < e && 1.0 - sum > EDIT_DIST_PROB_BOUND; k++) {
        double x = bin_coeff * p_power * q_power;
        sum += x;
        bin_coeff *= n - ct;
        bin_coeff /= ++ct;
        p_power *= p;
        q_power /= q;
      }
      if (1.0 - sum > EDIT_DIST_PROB_BOUND) return n;
    } else {
      double normal_z = (e - 0.5 - n * p) / sqrt(n * p * q);
      if (normal_z <= NORMAL_DISTRIB_THOLD) return n;
      double sum = 0.0, mu_power = 1.0, factorial = 1.0, poisson_coeff = exp(-n * p);
      for (int k = 0; k < e; k++) {
        sum += mu_power * poisson_coeff / factorial;
        mu_power *= n * p;
        factorial *= k + 1;
      }
      if (1.0 - sum > EDIT_DIST_PROB_BOUND) return n;
    }
  }
  return (int)OVO_MAX_READLEN;
}

static void init_match_limit(int32_t *ml, double erate, int32_t max_errors) {   /* Binomial_Bound.C:104-188 */
  int32_t e = 0, s = 1;
  int32_t l = max_errors < 2000 ? max_errors : 2000;
  while (e <= 1) ml[e++] = 0;                                 /* ERRORS_FOR_FREE = 1 */
  while (e < l) {
    s = binomial_bound(e - 1, erate, s);
    ml[e] = s - 1;
    e++;
  }
  double sl = 0.982064188397525 / erate + 0.067835741959926;  /* AS_MAX_READLEN_BITS == 21, :159-161 */
  double vl = ml[e - 1] + sl;
  while (e < max_errors) {
    ml[e] = (int32_t)ceil(vl);
    vl += sl;
    e++;
  }
}

static int32_t error_bound(const ovo_ctx *c, int32_t len) {   /* prefixEditDistance.C:58-61 */
  return (int32_t)ceil(len * c->P.max_erate);
}

/* ======================================================================= */
/*  prefixEditDistance  (liboverlap/prefixEditDistance-*.C)                */
/* ======================================================================= */

static void ped_init(ped_t *p, uint32_t max_errors) {
  memset(p, 0, sizeof(*p));
  p->rows_alloc  = 0;
  p->row         = (int32_t **)calloc(max_errors + 1, sizeof(int32_t *));
  p->left_delta  = (int32_t *)malloc(sizeof(int32_t) * (max_errors + 1));
  p->right_delta = (int32_t *)malloc(sizeof(int32_t) * (max_errors + 1));
  p->stack       = (int32_t *)malloc(sizeof(int32_t) * (max_errors + 1));
}
static void ped_free(ped_t *p, uint32_t max_errors) {
  for (uint32_t i = 0; i <= max_errors; i++)
    if (p->row[i]) free(p->row[i] - (i + 2));
  free(p->row); free(p->left_delta); free(p->right_delta); free(p->stack);
}
/* Row e can be indexed from -2-e to 2+e (prefixEditDistance-allocateMoreSpace.C:52-58). */
static inline int32_t *ped_row(ped_t *p, int e) {
  if (p->row[e] == NULL) {
    int32_t *m = (int32_t *)malloc(sizeof(int32_t) * (2 * e + 5));
    p->row[e] = m + (e + 2);
  }
  return p->row[e];
}

static int sign_of(int a) { return (a > 0) - (a < 0); }

/* prefixEditDistance-forward.C:32-71 */
static void set_right_delta(ped_t *p, int e, int d) {
  int last = p->row[e][d];
  p->right_delta_len = 0;
  for (int k = e; k > 0; k--) {
    int from = d, j;
    int max = 1 + p->row[k - 1][d];
    if ((j = p->row[k - 1][d - 1]) > max)     { from = d - 1; max = j; }
    if ((j = 1 + p->row[k - 1][d + 1]) > max) { from = d + 1; max = j; }
    if (from == d - 1) {
      p->stack[p->right_delta_len++] = max - last - 1;
      d--;
      last = p->row[k - 1][from];
    } else if (from == d + 1) {
      p->stack[p->right_delta_len++] = last - (max - 1);
      d++;
      last = p->row[k - 1][from];
    }
  }
  p->stack[p->right_delta_len++] = last + 1;
  int k = 0;
  for (int i = p->right_delta_len - 1; i > 0; i--)
    p->right_delta[k++] = abs(p->stack[i]) * sign_of(p->stack[i - 1]);
  p->right_delta_len--;
}

/* prefixEditDistance-reverse.C:36-93 */
static void set_left_delta(ped_t *p, int e, int d, int *leftover, int *t_end, int t_len) {
  int last = p->row[e][d];
  p->left_delta_len = 0;
  for (int k = e; k > 0; k--) {
    int from = d, j;
    int max = 1 + p->row[k - 1][d];
    if ((j = p->row[k - 1][d - 1]) > max)     { from = d - 1; max = j; }
    if ((j = 1 + p->row[k - 1][d + 1]) > max) { from = d + 1; max = j; }
    if (from == d - 1) {
      p->left_delta[p->left_delta_len++] = max - last - 1;
      d--;
      last = p->row[k - 1][from];
    } else if (from == d + 1) {
      p->left_delta[p->left_delta_len++] = last - (max - 1);
      d++;
      last = p->row[k - 1][from];
    }
  }
  *leftover = last;
  if (p->left_delta_len > 1 && p->left_delta[0] == 1 && *t_end + t_len > 0) {
    if (p->left_delta[1] > 0) p->left_delta[0] = p->left_delta[1] + 1;
    else                      p->left_delta[0] = p->left_delta[1] - 1;
    for (int i = 2; i < p->left_delta_len; i++)
      p->left_delta[i - 1] = p->left_delta[i];
    p->left_delta_len--;
    (*t_end)--;
    if (p->left_delta_len == 0)
      (*leftover)++;
  }
}

#define PRUNE(e, d, extra)  (cur[d] + (extra) < c->eml[e])     /* prefixEditDistance.H:248-258 */

/*  forward (dirn=+1, prefixEditDistance-forward.C:94-314) and reverse (dirn=-1,
 *  prefixEditDistance-reverse.C:114-330) share everything except the string
 *  direction, the "force last error to be mismatch" fix-up (forward only,
 *  :215-221) and the traceback routine.  */
static int32_t ped_extend(const ovo_ctx *c, ped_t *p, int dirn,
                          const char *A, int32_t m, const char *T, int32_t n,
                          int32_t error_limit, int32_t *a_end, int32_t *t_end,
                          int32_t *leftover, int *match_to_end) {
  int max_score_len = 0, max_score_best_d = 0, max_score_best_e = 0;
  int best_d = 0, best_e = 0, longest = 0;
  int row, d, e;
  double max_score = 0.0;
  const double bmv = c->branch_match_value;

  if (dirn > 0) p->right_delta_len = 0; else p->left_delta_len = 0;

#define AT(s, i)  ((s)[dirn > 0 ? (i) : -(i)])
#define MATCH(i, j) (AT(A, i) == AT(T, j) || AT(A, i) == 'n' || AT(T, j) == 'n')

  for (row = 0; row < m && MATCH(row, row); row++)
    p->compares++;
  if (row < m) p->compares++;

  ped_row(p, 0)[0] = row;

  if (row == m) {
    if (dirn > 0) { *a_end = *t_end = m; }
    else          { *a_end = *t_end = -m; *leftover = m; }
    *match_to_end = 1;
    return 0;
  }
  p->calls++;                          /* instrumentation: extensions that actually ran the DP */

  int left = 0, right = 0;

  for (e = 1; e <= error_limit; e++) {
    int32_t *cur  = ped_row(p, e);
    int32_t *prev = p->row[e - 1];

    left  = (left - 1  > -e) ? left - 1  : -e;
    right = (right + 1 <  e) ? right + 1 :  e;

    prev[left] = -2;  prev[left - 1] = -2;  prev[right] = -2;  prev[right + 1] = -2;

    for (d = left; d <= right; d++) {
      int j;
      p->cells++;
      row = 1 + prev[d];
      if ((j = prev[d - 1]) > row)     row = j;
      if ((j = 1 + prev[d + 1]) > row) row = j;
      while (row < m && row + d < n) {
        p->compares++;
        if (!MATCH(row, row + d)) break;
        row++;
      }
      cur[d] = row;

      if (row == m || row + d == n) {
        double score    = row * bmv - e;
        int    tail_len = row - max_score_len;
        int    abort_   = 0;
        double slope    = (double)(max_score - score) / tail_len;

        if (c->P.partial && score < max_score) abort_ = 1;
        if (e > MIN_BRANCH_END_DIST / 2 && tail_len >= MIN_BRANCH_END_DIST && slope >= c->min_branch_tail_slope)
          abort_ = 1;

        if (abort_) {
          if (dirn > 0) {
            *a_end = max_score_len;
            *t_end = max_score_len + max_score_best_d;
            set_right_delta(p, max_score_best_e, max_score_best_d);
          } else {
            *a_end = -max_score_len;
            *t_end = -max_score_len - max_score_best_d;
            set_left_delta(p, max_score_best_e, max_score_best_d, leftover, t_end, n);
          }
          *match_to_end = 0;
          return max_score_best_e;
        }

        if (dirn > 0) {
          if (row == m && 1 + prev[d + 1] == cur[d] && d < right) {   /* forward.C:215-221 */
            d++;
            cur[d] = cur[d - 1];
          }
          *a_end = row;
          *t_end = row + d;
          set_right_delta(p, e, d);
        } else {
          *a_end = -row;
          *t_end = -row - d;
          set_left_delta(p, e, d, leftover, t_end, n);
        }
        *match_to_end = 1;
        return e;
      }
    }

    while (left <= right && left < 0 && PRUNE(e, left, 0)) left++;
    if (left >= 0)
      while (left <= right && PRUNE(e, left, left)) left++;
    if (left > right)
      break;
    while (right > 0 && PRUNE(e, right, right)) right--;
    if (right <= 0)
      while (PRUNE(e, right, 0)) right--;

    for (d = left; d <= right; d++)
      if (cur[d] > longest) { best_d = d; best_e = e; longest = cur[d]; }

    double score = longest * bmv - e;
    if (score > max_score) {
      max_score = score;
      max_score_len = longest;
      max_score_best_d = best_d;
      max_score_best_e = best_e;
    }
  }

  if (dirn > 0) {
    *a_end = max_score_len;
    *t_end = max_score_len + max_score_best_d;
    set_right_delta(p, max_score_best_e, max_score_best_d);
  } else {
    *a_end = -max_score_len;
    *t_end = -max_score_len - max_score_best_d;
    set_left_delta(p, max_score_best_e, max_score_best_d, leftover, t_end, n);
  }
  *match_to_end = 0;
  return max_score_best_e;
#undef AT
#undef MATCH
}

enum { K_NONE = 0, K_LEFT_BRANCH = 1, K_RIGHT_BRANCH = 2, K_DOVETAIL = 3 };   /* Overlap_t, prefixEditDistance.H:35-40 */

/* prefixEditDistance-extend.C:36-183 */
static int extend_alignment(const ovo_ctx *c, ped_t *p, const mnode_t *match,
                            const char *S, int32_t s_len, const char *T, int32_t t_len,
                            int32_t *s_lo, int32_t *s_hi, int32_t *t_lo, int32_t *t_hi, int32_t *errors) {
  int32_t right_errors = 0, left_errors = 0, leftover = 0;
  int r_to_end = 1, l_to_end = 1;

  int32_t s_left_begin  = match->start - 1;
  int32_t s_right_begin = match->start + match->len;
  int32_t s_right_len   = s_len - s_right_begin;
  int32_t t_left_begin  = match->offset - 1;
  int32_t t_right_begin = match->offset + match->len;
  int32_t t_right_len   = t_len - t_right_begin;

  int32_t total_olap = (match->start < match->offset ? match->start : match->offset) + match->len +
                       (s_right_len < t_right_len ? s_right_len : t_right_len);
  int32_t error_limit = error_bound(c, total_olap);

  p->left_delta_len = 0;
  p->right_delta_len = 0;

  if (s_right_len == 0 || t_right_len == 0) {
    *s_hi = 0; *t_hi = 0; r_to_end = 1;
  } else if (s_right_len <= t_right_len) {
    right_errors = ped_extend(c, p, +1, S + s_right_begin, s_right_len, T + t_right_begin, t_right_len,
                              error_limit, s_hi, t_hi, NULL, &r_to_end);
    for (int i = 0; i < p->right_delta_len; i++) p->right_delta[i] *= -1;
  } else {
    right_errors = ped_extend(c, p, +1, T + t_right_begin, t_right_len, S + s_right_begin, s_right_len,
                              error_limit, t_hi, s_hi, NULL, &r_to_end);
  }
  *s_hi += s_right_begin - 1;
  *t_hi += t_right_begin - 1;

  if (s_left_begin < 0 || t_left_begin < 0) {
    *s_lo = 0; *t_lo = 0; l_to_end = 1;
  } else if (s_right_begin <= t_right_begin) {
    left_errors = ped_extend(c, p, -1, S + s_left_begin, s_left_begin + 1, T + t_left_begin, t_left_begin + 1,
                             error_limit - right_errors, s_lo, t_lo, &leftover, &l_to_end);
  } else {
    left_errors = ped_extend(c, p, -1, T + t_left_begin, t_left_begin + 1, S + s_left_begin, s_left_begin + 1,
                             error_limit - right_errors, t_lo, s_lo, &leftover, &l_to_end);
    for (int i = 0; i < p->left_delta_len; i++) p->left_delta[i] *= -1;
  }
  *s_lo += s_left_begin + 1;
  *t_lo += t_left_begin + 1;

  *errors = left_errors + right_errors;

  int kind = !r_to_end ? (!l_to_end ? K_NONE : K_RIGHT_BRANCH) : (!l_to_end ? K_LEFT_BRANCH : K_DOVETAIL);

  if (p->right_delta_len > 0) {
    if (p->right_delta[0] > 0)
      p->left_delta[p->left_delta_len++] = -(p->right_delta[0] + leftover + match->len);
    else
      p->left_delta[p->left_delta_len++] = -(p->right_delta[0] - leftover - match->len);
  }
  for (int i = 1; i < p->right_delta_len; i++)
    p->left_delta[p->left_delta_len++] = -p->right_delta[i];
  p->right_delta_len = 0;

  return kind;
}

/* ======================================================================= */
/*  Hash table build  (overlapInCore-Build_Hash_Index.C)                   */
/* ======================================================================= */

static void fail(const char *msg) { fprintf(stderr, "ovl_oracle: %s\n", msg); abort(); }

/* Build_Hash_Index.C:267-321 */
static void hash_insert(ovo_ctx *c, sref_t ref, uint64_t key, const char *s) {
  int64_t sub   = (int64_t)HASH_FN(c, key);
  int     shift = (int)HCHECK_FN(c, key);
  c->check_array[sub] |= (1u << shift);
  uint8_t kc    = (uint8_t)KCHECK_FN(c, key);
  int64_t probe = (int64_t)PROBE_FN(c, key);
  uint64_t ct = 0;
  const uint32_t K = c->P.kmer_len;
  do {
    bucket_t *b = &c->table[sub];
    int i;
    for (i = 0; i < b->count; i++)
      if (b->check[i] == kc) {
        sref_t h = b->entry[i];
        const char *t = c->data + c->string_start[SR_NUM(h)] + SR_OFF(h);
        if (strncmp(s, t, K) == 0) {
          if (SR_LAST(h)) c->extra_ref_ct++;
          c->next_ref[c->string_start[SR_NUM(ref)] + SR_OFF(ref)] = h;
          c->extra_ref_ct++;
          ref = SR_CLR_LAST(ref);
          b->entry[i] = ref;
          if (b->hits[i] < 255) b->hits[i]++;
          return;
        }
      }
    if (b->count < ENTRIES_PER_BUCKET) {
      ref = SR_SET_LAST(ref);
      b->entry[i] = ref;
      b->check[i] = kc;
      b->count++;
      c->hash_entries++;
      b->hits[i] = 1;
      return;
    }
    sub = (sub + probe) % (int64_t)c->table_size;
  } while (++ct < c->table_size);
  fail("hash table full");
}

/* Build_Hash_Index.C:331-404 (HASH_KMER_SKIP == 0) */
static void put_string_in_hash(ovo_ctx *c, uint32_t i) {
  const uint32_t K = c->P.kmer_len;
  const char *p = c->data + c->string_start[i];
  const char *window = p;
  uint64_t key = 0, bad = 0;
  for (uint32_t j = 0; j < K; j++) {
    bad |= (uint64_t)char_is_bad(*p) << j;
    key |= (uint64_t)bit_equiv(*(p++)) << (2 * j);
  }
  sref_t ref = SR_MAKE(i, 0);
  if (!bad) { hash_insert(c, ref, key, window); c->st.hash_inserts++; }
  while (*p != 0) {
    window++;
    ref = SR_MAKE(i, SR_OFF(ref) + 1);
    bad >>= 1;
    bad |= (uint64_t)char_is_bad(*p) << (K - 1);
    key >>= 2;
    key |= (uint64_t)bit_equiv(*(p++)) << (2 * (K - 1));
    if (bad) continue;
    hash_insert(c, ref, key, window);
    c->st.hash_inserts++;
  }
}

/* Build_Hash_Index.C:28-86 */
static sref_t add_extra_hash_string(ovo_ctx *c, const char *s) {
  const uint32_t K = c->P.kmer_len;
  uint64_t max_extra_subcount = OVO_MAX_READLEN / K;
  uint64_t sub;
  uint64_t new_len = c->used_data_len + K;

  if (c->extra_string_subcount < max_extra_subcount) {
    sub = c->string_ct + c->extra_string_ct - 1;
  } else {
    sub = c->string_ct + c->extra_string_ct;
    if (sub >= c->string_start_size) {
      uint64_t n = (uint64_t)fmax(sub * 1.1, c->string_start_size * 1.5);
      c->string_start = (int64_t *)realloc(c->string_start, sizeof(int64_t) * n);
      c->string_start_size = n;
    }
    c->string_start[sub] = (int64_t)c->used_data_len;
    c->extra_string_ct++;
    c->extra_string_subcount = 0;
    new_len++;
  }
  if (new_len >= c->extra_data_len) {
    uint64_t n = (uint64_t)fmax(new_len * 1.1, c->extra_data_len * 1.5);
    c->data = (char *)realloc(c->data, n);
    c->extra_data_len = n;
  }
  strncpy(c->data + c->string_start[sub] + K * c->extra_string_subcount, s, K + 1);
  c->used_data_len = new_len;

  sref_t ref = SR_MAKE(sub, c->extra_string_subcount * (uint64_t)K);
  ref = SR_SET_LAST(ref);
  ref = SR_SET_EMPTY(ref);
  c->extra_string_subcount++;
  return ref;
}

/* Build_Hash_Index.C:98-121 */
static void mark_screened_ends_single(ovo_ctx *c, sref_t ref) {
  int32_t s_num = (int32_t)SR_NUM(ref);
  int32_t len = (int32_t)c->string_info[s_num].length;
  if (SR_OFF(ref) < HOPELESS_MATCH) c->string_info[s_num].lscreen = 1;
  /* the reference evaluates `len - offset - Kmer_Len + 1 < HOPELESS_MATCH` in uint64 arithmetic */
  if ((uint64_t)len - SR_OFF(ref) - (uint64_t)c->P.kmer_len + 1 < (uint64_t)HOPELESS_MATCH)
    c->string_info[s_num].rscreen = 1;
}
static void mark_screened_ends_chain(ovo_ctx *c, sref_t ref) {
  mark_screened_ends_single(c, ref);
  while (!SR_LAST(ref)) {
    ref = c->next_ref[c->string_start[SR_NUM(ref)] + SR_OFF(ref)];
    mark_screened_ends_single(c, ref);
  }
}

/* Build_Hash_Index.C:132-177 */
static void hash_mark_empty(ovo_ctx *c, uint64_t key, const char *s) {
  int64_t sub   = (int64_t)HASH_FN(c, key);
  uint8_t kc    = (uint8_t)KCHECK_FN(c, key);
  int64_t probe = (int64_t)PROBE_FN(c, key);
  uint64_t ct = 0;
  const uint32_t K = c->P.kmer_len;
  do {
    bucket_t *b = &c->table[sub];
    int i;
    for (i = 0; i < b->count; i++)
      if (b->check[i] == kc) {
        sref_t h = b->entry[i];
        const char *t = c->data + c->string_start[SR_NUM(h)] + SR_OFF(h);
        if (strncmp(s, t, K) == 0) {
          if (!SR_EMPTY(b->entry[i])) mark_screened_ends_chain(c, b->entry[i]);
          b->entry[i] = SR_SET_EMPTY(b->entry[i]);
          return;
        }
      }
    if (b->count < ENTRIES_PER_BUCKET) {
      if (c->use_hopeless) {
        b = NULL;   /* add_extra_hash_string may realloc c->data but never c->table */
        sref_t r = add_extra_hash_string(c, s);
        b = &c->table[sub];
        b->entry[i] = SR_SET_EMPTY(r);
        b->check[i] = kc;
        b->count++;
        b->hits[i] = 0;
        c->hash_entries++;
        int shift = (int)HCHECK_FN(c, key);
        c->check_array[sub] |= (1u << shift);    /* on the PROBED bucket: SURVEY.md 7.11 */
      }
      return;
    }
    sub = (sub + probe) % (int64_t)c->table_size;
  } while (++ct < c->table_size);
  fail("hash table full");
}

static char comp_lower(char ch) {                /* utility/src/sequence/sequence-v1.C:27-60 restricted to lower case */
  switch (ch) { case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'n': return 'n'; }
  return 0;
}
static void revcomp_inplace(char *s, int len) {  /* sequence-v1.C:111-129 */
  int i = 0, j = len - 1;
  while (i < j) { char a = s[i]; s[i++] = comp_lower(s[j]); s[j--] = comp_lower(a); }
  if (i == j) s[i] = comp_lower(s[i]);
}

/* Build_Hash_Index.C:186-257 */
static void mark_skip_kmers(ovo_ctx *c) {
  const uint32_t K = c->P.kmer_len;
  char line[64];
  for (uint32_t n = 0; n < c->n_skip; n++) {
    for (uint32_t i = 0; i < K; i++) line[i] = (char)tolower(c->skip[(uint64_t)n * K + i]);
    line[K] = 0;
    uint64_t key = 0;
    for (uint32_t i = 0; i < K; i++) key |= (uint64_t)bit_equiv(line[i]) << (2 * i);
    hash_mark_empty(c, key, line);
    revcomp_inplace(line, (int)K);
    key = 0;
    for (uint32_t i = 0; i < K; i++) key |= (uint64_t)bit_equiv(line[i]) << (2 * i);
    hash_mark_empty(c, key, line);
  }
}

/* Build_Hash_Index.C:415-631; returns the last read ID loaded */
static uint32_t build_hash_index(ovo_ctx *c, uint32_t bgn, uint32_t end) {
  c->hash_string_num_offset = bgn;
  c->string_ct = 0;
  c->extra_string_ct = 0;
  c->extra_string_subcount = OVO_MAX_READLEN / c->P.kmer_len;
  uint64_t total_len = 0;

  memset(c->table, 0, sizeof(bucket_t) * c->table_size);
  memset(c->check_array, 0, sizeof(uint32_t) * c->table_size);

  c->extra_ref_ct = 0;
  c->hash_entries = 0;
  uint64_t hash_entry_limit = (uint64_t)(c->P.hash_load * c->table_size * ENTRIES_PER_BUCKET);

  uint64_t max_alloc = 0;
  uint32_t cur;
  for (cur = bgn; max_alloc < c->max_hash_data_len && cur <= end; cur++) {
    uint32_t rl = c->len[cur];
    if (rl < (uint32_t)c->P.min_olap_len) continue;
    max_alloc += rl + 1;
  }

  free(c->data);     c->data = (char *)malloc(max_alloc + 1);
  free(c->next_ref); c->next_ref = (sref_t *)malloc(sizeof(sref_t) * (max_alloc + 1));
  memset(c->next_ref, 0xff, sizeof(sref_t) * (max_alloc + 1));
  c->data_len = c->extra_data_len = max_alloc;

  uint64_t need = (uint64_t)end - bgn + 2;
  if (c->string_start_size < need) {
    c->string_start = (int64_t *)realloc(c->string_start, sizeof(int64_t) * need);
    c->string_start_size = need;
  }
  if (c->string_info_size < need) {
    c->string_info = (finfo_t *)realloc(c->string_info, sizeof(finfo_t) * need);
    c->string_info_size = need;
  }

  for (cur = bgn; total_len < c->max_hash_data_len && c->hash_entries < hash_entry_limit && cur <= end; cur++, c->string_ct++) {
    uint64_t sc = c->string_ct;
    c->string_start[sc] = -1;
    c->string_info[sc].length = 0;
    c->string_info[sc].lscreen = 1;
    c->string_info[sc].rscreen = 1;

    uint32_t len = c->len[cur];
    if (len < (uint32_t)c->P.min_olap_len) continue;

    c->string_start[sc] = (int64_t)total_len;
    c->string_info[sc].length = len;
    c->string_info[sc].lscreen = 0;
    c->string_info[sc].rscreen = 0;

    const char *src = c->bases + c->off[cur];
    for (uint32_t i = 0; i < len; i++, total_len++)
      c->data[total_len] = (char)tolower(src[i]);
    c->data[total_len] = 0;
    total_len++;

    put_string_in_hash(c, (uint32_t)sc);
  }

  if (c->string_ct == 0)
    return end;

  c->used_data_len = total_len;

  if (c->extra_ref_ct > c->extra_ref_max) {
    free(c->extra_ref);
    c->extra_ref_max = c->extra_ref_ct;
    c->extra_ref = (sref_t *)malloc(sizeof(sref_t) * c->extra_ref_max);
  }

  mark_skip_kmers(c);

  /* coalesce chains, Build_Hash_Index.C:613-628 */
  c->extra_ref_ct = 0;
  for (uint64_t i = 0; i < c->table_size; i++)
    for (int j = 0; j < c->table[i].count; j++) {
      sref_t ref = c->table[i].entry[j];
      if (!SR_LAST(ref) && !SR_EMPTY(ref)) {
        c->extra_ref[c->extra_ref_ct] = ref;
        sref_t e = c->table[i].entry[j];
        e = (e & ~STRING_NUM_MASK) | (c->extra_ref_ct >> OFFSET_BITS);
        e = (e & ~(OFFSET_MASK << STRING_NUM_BITS)) | ((c->extra_ref_ct & OFFSET_MASK) << STRING_NUM_BITS);
        c->table[i].entry[j] = e;
        c->extra_ref_ct++;
        do {
          ref = c->next_ref[c->string_start[SR_NUM(ref)] + SR_OFF(ref)];
          c->extra_ref[c->extra_ref_ct++] = ref;
        } while (!SR_LAST(ref));
      }
    }

  return cur - 1;
}

/* ======================================================================= */
/*  Seed collection  (overlapInCore-Find_Overlaps.C)                       */
/* ======================================================================= */

/* Find_Overlaps.C:26-96 */
static void add_match(const ovo_ctx *c, work_t *w, sref_t ref, int32_t *start, int offset, int *consistent) {
  const int K = (int)c->P.kmer_len;
  int32_t *p;
  int diag = 0, expected_start = 0, num_checked = 0, move_to_front = 0;
  int new_diag = (int)SR_OFF(ref) - offset;

  for (p = start; *p != 0; p = &w->node[*p].next) {
    mnode_t *nd = &w->node[*p];
    expected_start = nd->start + nd->len - K + 1;
    diag = nd->offset - nd->start;
    if (expected_start < offset) break;
    if (expected_start == offset) {
      if (new_diag == diag) {
        nd->len += 1;
        if (move_to_front) {
          int32_t save = *p;
          *p = w->node[*p].next;
          w->node[save].next = *start;
          *start = save;
        }
        return;
      } else
        move_to_front = 1;
    }
    num_checked++;
  }

  if (w->node_next == w->node_size) {
    /* start may point into w->node (a .next field) or into w->olap; only the first can move */
    int in_nodes = ((char *)start >= (char *)w->node && (char *)start < (char *)(w->node + w->node_size));
    size_t delta = in_nodes ? (size_t)((char *)start - (char *)w->node) : 0;
    w->node_size *= 2;
    w->node = (mnode_t *)realloc(w->node, sizeof(mnode_t) * w->node_size);
    if (in_nodes) start = (int32_t *)((char *)w->node + delta);
  }

  if (*start != 0 && (num_checked > 0 || abs(diag - new_diag) > 3 || offset < expected_start + K - 2))
    *consistent = 0;

  int32_t save = *start;
  *start = w->node_next++;
  mnode_t *nn = &w->node[*start];
  nn->offset = (int32_t)SR_OFF(ref);
  nn->len = K;
  nn->start = offset;
  nn->next = save;
}

/* Find_Overlaps.C:105-163 */
static void add_ref(const ovo_ctx *c, work_t *w, sref_t ref, int offset) {
  uint32_t str_num = (uint32_t)SR_NUM(ref);
  uint32_t sub = (str_num ^ (str_num >> STRING_OLAP_SHIFT)) & STRING_OLAP_MASK;
  uint32_t prev;

  while (w->olap[sub].full && w->olap[sub].string_num != str_num) {
    prev = sub;
    sub = (uint32_t)w->olap[sub].next;
    if (sub == 0) {
      if (w->olap_next == w->olap_size) {
        w->olap_size *= 2;
        w->olap = (solap_t *)realloc(w->olap, sizeof(solap_t) * w->olap_size);
      }
      sub = (uint32_t)w->olap_next++;
      w->olap[prev].next = (int32_t)sub;
      w->olap[sub].full = 0;
      break;
    }
  }
  solap_t *o = &w->olap[sub];
  if (!o->full) {
    o->string_num = str_num;
    o->match_list = 0;
    o->diag_sum = 0.0;
    o->diag_ct = 0;
    o->diag_bgn = (int32_t)OVO_MAX_READLEN;
    o->diag_end = 0;
    o->next = 0;
    o->full = 1;
    o->consistent = 1;
  }
  int consistent = o->consistent;
  o->diag_sum += (double)SR_OFF(ref) - offset;
  o->diag_ct++;
  if (o->diag_bgn > offset) o->diag_bgn = offset;
  if (o->diag_end < offset) o->diag_end = offset;
  add_match(c, w, ref, &o->match_list, offset, &consistent);
  w->olap[sub].consistent = (uint8_t)consistent;
  w->st.seed_hits++;
}

/* Find_Overlaps.C:177-222 */
static sref_t hash_find(const ovo_ctx *c, uint64_t key, int64_t sub, const char *s, int64_t *where, int *hi_hits) {
  sref_t h = 0;
  uint8_t kc    = (uint8_t)KCHECK_FN(c, key);
  int64_t probe = (int64_t)PROBE_FN(c, key);
  uint64_t ct = 0;
  const uint32_t K = c->P.kmer_len;
  *hi_hits = 0;
  do {
    const bucket_t *b = &c->table[sub];
    for (int i = 0; i < b->count; i++)
      if (b->check[i] == kc) {
        h = b->entry[i];
        int is_empty = (int)SR_EMPTY(h);
        if (!SR_LAST(h) && !is_empty) {
          *where = (int64_t)((SR_NUM(h) << OFFSET_BITS) + SR_OFF(h));
          h = c->extra_ref[*where];
        }
        const char *t = c->data + c->string_start[SR_NUM(h)] + SR_OFF(h);
        if (strncmp(s, t, K) == 0) {
          if (is_empty) { h = SR_SET_EMPTY(h); *hi_hits = 1; }
          return h;
        }
      }
    if (b->count < ENTRIES_PER_BUCKET)
      return SR_SET_EMPTY(h);
    sub = (sub + probe) % (int64_t)c->table_size;
  } while (++ct < c->table_size);
  return SR_SET_EMPTY(h);
}

static void process_string_olaps(ovo_ctx *c, work_t *w, const char *S, int len, uint32_t id, int dir);

/* Find_Overlaps.C:235-336.  The reference pipelines Next_Key/Next_Check one
 * window ahead; the sequence of windows tested is 0 .. len-K, restated directly. */
static void find_overlaps(ovo_ctx *c, work_t *w, const char *frag, int frag_len, uint32_t frag_num, int dir) {
  const int K = (int)c->P.kmer_len;

  memset(w->olap, 0, STRING_OLAP_MODULUS * sizeof(solap_t));
  w->olap_next = STRING_OLAP_MODULUS;
  w->node_next = 1;
  w->left_end_screened = 0;
  w->right_end_screened = 0;

  uint64_t key = 0;
  for (int j = 0; j < K; j++)
    key |= (uint64_t)bit_equiv(frag[j]) << (2 * j);

  for (int offset = 0; offset + K <= frag_len; offset++) {
    if (offset > 0) {
      key >>= 2;
      key |= (uint64_t)bit_equiv(frag[offset + K - 1]) << (2 * (K - 1));
    }
    w->st.ref_lookups++;
    int64_t sub = (int64_t)HASH_FN(c, key);
    int shift = (int)HCHECK_FN(c, key);
    if ((c->check_array[sub] & (1u << shift)) == 0)
      continue;
    int64_t where = 0; int hi_hits = 0;
    sref_t ref = hash_find(c, key, sub, frag + offset, &where, &hi_hits);
    if (hi_hits) {
      if (offset == 0) {
        w->left_end_screened = 1;                                   /* :274-276 */
      } else {
        if (offset < HOPELESS_MATCH) w->left_end_screened = 1;      /* :310-316 */
        /* mixed int/uint64 arithmetic in the reference: evaluated in uint64 */
        if ((uint64_t)frag_len - (uint64_t)offset - (uint64_t)K + 1 < (uint64_t)HOPELESS_MATCH) w->right_end_screened = 1;
      }
    }
    if (!SR_EMPTY(ref)) {
      for (;;) {
        if (frag_num < SR_NUM(ref) + c->hash_string_num_offset)
          add_ref(c, w, ref, offset);
        if (SR_LAST(ref)) break;
        ref = c->extra_ref[++where];
      }
    }
  }
  process_string_olaps(c, w, frag, frag_len, frag_num, dir);
}

/* ======================================================================= */
/*  Output  (overlapInCore-Output.C, ../stores/ovOverlap.H)                */
/* ======================================================================= */

typedef struct { uint32_t ahg5, ahg3, bhg5, bhg3, span, evalue, flipped, obt, dup, utg; } ovl_fields;

static void pack_record(ovo_record *r, uint32_t a, uint32_t b, const ovl_fields *f) {   /* ovOverlap.H:49-67 */
  const uint64_t M = (1ull << OVO_MAX_READLEN_BITS) - 1;
  r->a_iid = a; r->b_iid = b;
  r->w0 = ((uint64_t)f->ahg5 & M) | (((uint64_t)f->ahg3 & M) << 21) | (((uint64_t)f->evalue & 0xffff) << 42) |
          ((uint64_t)(f->flipped & 1) << 58) | ((uint64_t)(f->obt & 1) << 59) | ((uint64_t)(f->dup & 1) << 60) | ((uint64_t)(f->utg & 1) << 61);
  r->w1 = ((uint64_t)f->bhg5 & M) | (((uint64_t)f->bhg3 & M) << 21) | (((uint64_t)f->span & M) << 42);
}

static uint32_t encode_evalue(double q) {                     /* ovOverlap.H:31-35 */
  return (q < 65535 / 100000.0) ? (uint32_t)(100000.0 * q + 0.5) : 65535u;
}

static void emit(work_t *w, uint32_t a, uint32_t b, const ovl_fields *f) {
  if (w->rec_len == w->rec_max) {
    w->rec_max = w->rec_max ? w->rec_max * 2 : 1024;
    w->rec = (ovo_record *)realloc(w->rec, sizeof(ovo_record) * w->rec_max);
  }
  pack_record(&w->rec[w->rec_len++], a, b, f);
}

static void set_a_hang(ovl_fields *f, int32_t a) { f->ahg5 = (a < 0) ? 0 : (uint32_t)a;  f->bhg5 = (a < 0) ? (uint32_t)-a : 0; }   /* ovOverlap.H:166 */
static void set_b_hang(ovl_fields *f, int32_t b) { f->bhg3 = (b < 0) ? 0 : (uint32_t)b;  f->ahg3 = (b < 0) ? (uint32_t)-b : 0; }   /* :167 */

/* Output.C:27-191 */
static void output_overlap(work_t *w, uint32_t s_id, int s_len, int dir, uint32_t t_id, int t_len, const oinfo_t *o) {
  ovl_fields f; memset(&f, 0, sizeof(f));
  f.utg = 1;
  f.span = (uint32_t)(((o->s_hi - o->s_lo) + (o->t_hi - o->t_lo) + o->delta_ct) / 2);

  int32_t s_right_hang = s_len - o->s_hi - 1;
  int32_t t_right_hang = t_len - o->t_hi - 1;
  int sleft = (o->s_lo > o->t_lo) || (o->s_lo == o->t_lo && s_right_hang > t_right_hang);
  uint32_t a = sleft ? s_id : t_id, b = sleft ? t_id : s_id;
  char orient; int32_t ahg, bhg;
  if (sleft) { orient = (dir == 0) ? 'N' : 'O'; ahg = o->s_lo; bhg = t_right_hang - s_right_hang; }
  else       { orient = (dir == 0) ? 'N' : 'I'; ahg = o->t_lo; bhg = s_right_hang - t_right_hang; }
  if (orient == 'O' && s_right_hang >= t_right_hang) {
    orient = 'I';
    ahg = -(t_right_hang - s_right_hang);
    bhg = -(o->s_lo);
  }
  f.evalue = encode_evalue(o->quality);
  switch (orient) {
    case 'N': set_a_hang(&f, ahg);  set_b_hang(&f, bhg);  f.flipped = 0; break;
    case 'I': set_a_hang(&f, ahg);  set_b_hang(&f, bhg);  f.flipped = 1; break;
    case 'O': set_a_hang(&f, -bhg); set_b_hang(&f, -ahg); f.flipped = 1; break;
  }
  emit(w, a, b, &f);
  w->st.total_overlaps++;
  if (bhg <= 0) w->st.contained++; else w->st.dovetail++;
}

/* Output.C:195-264 */
static void output_partial_overlap(work_t *w, uint32_t s_id, uint32_t t_id, int dir, const oinfo_t *o, int s_len, int t_len) {
  ovl_fields f; memset(&f, 0, sizeof(f));
  w->st.total_overlaps++;
  f.obt = 1; f.dup = 1;
  f.span = (uint32_t)(((o->s_hi - o->s_lo) + (o->t_hi - o->t_lo) + o->delta_ct) / 2);
  if (dir == 0) {
    f.ahg5 = (uint32_t)o->s_lo;               f.ahg3 = (uint32_t)(s_len - (o->s_hi + 1));
    f.bhg5 = (uint32_t)o->t_lo;               f.bhg3 = (uint32_t)(t_len - (o->t_hi + 1));
    f.flipped = 0;
  } else {
    f.ahg5 = (uint32_t)(s_len - (o->s_hi + 1)); f.ahg3 = (uint32_t)o->s_lo;
    f.bhg5 = (uint32_t)(t_len - (o->t_hi + 1)); f.bhg3 = (uint32_t)o->t_lo;
    f.flipped = 1;
  }
  f.evalue = encode_evalue(o->quality);
  emit(w, s_id, t_id, &f);
}

/* ======================================================================= */
/*  Per-pair control  (overlapInCore-Process_String_Overlaps.C)            */
/* ======================================================================= */

/* :42-96 */
static void combine_into_one_olap(oinfo_t *o, int ct, int *deleted) {
  int best = 0;
  int min_diag = o[0].min_diag, max_diag = o[0].max_diag;
  int slb = o[0].s_left_boundary, srb = o[0].s_right_boundary, tlb = o[0].t_left_boundary, trb = o[0].t_right_boundary;
  for (int i = 1; i < ct; i++) {
    int leni = 1 + ((o[i].s_hi - o[i].s_lo < o[i].t_hi - o[i].t_lo) ? o[i].s_hi - o[i].s_lo : o[i].t_hi - o[i].t_lo);
    int lenb = 1 + ((o[best].s_hi - o[best].s_lo < o[best].t_hi - o[best].t_lo) ? o[best].s_hi - o[best].s_lo : o[best].t_hi - o[best].t_lo);
    if (o[i].quality < o[best].quality || (o[i].quality == o[best].quality && leni > lenb)) best = i;
    if (o[i].min_diag < min_diag) min_diag = o[i].min_diag;
    if (o[i].max_diag > max_diag) max_diag = o[i].max_diag;
    if (o[i].s_left_boundary  < slb) slb = o[i].s_left_boundary;
    if (o[i].s_right_boundary > srb) srb = o[i].s_right_boundary;
    if (o[i].t_left_boundary  < tlb) tlb = o[i].t_left_boundary;
    if (o[i].t_right_boundary > trb) trb = o[i].t_right_boundary;
  }
  o[best].min_diag = min_diag; o[best].max_diag = max_diag;
  o[best].s_left_boundary = slb; o[best].s_right_boundary = srb;
  o[best].t_left_boundary = tlb; o[best].t_right_boundary = trb;
  for (int i = 0; i < ct; i++) deleted[i] = (i != best);
}

/* :108-162 */
static void merge_intersecting_olaps(oinfo_t *p, int ct, int *deleted) {
  for (int i = 0; i < ct - 1; i++)
    for (int j = i + 1; j < ct; j++) {
      if (deleted[i] || deleted[j]) continue;
      int lo_diag = p[i].min_diag, hi_diag = p[i].max_diag;
      if ((lo_diag <= 0 && p[j].min_diag > 0) || (lo_diag > 0 && p[j].min_diag <= 0)) continue;
      if ((lo_diag >= 0 && p[j].t_right_boundary - lo_diag - p[j].s_left_boundary >= MIN_INTERSECTION) ||
          (lo_diag <= 0 && p[j].s_right_boundary + lo_diag - p[j].t_left_boundary >= MIN_INTERSECTION) ||
          (hi_diag >= 0 && p[j].t_right_boundary - hi_diag - p[j].s_left_boundary >= MIN_INTERSECTION) ||
          (hi_diag <= 0 && p[j].s_right_boundary + hi_diag - p[j].t_left_boundary >= MIN_INTERSECTION)) {
        oinfo_t *discard, *keep;
        if (p[i].quality < p[j].quality) { keep = p + i; discard = p + j; deleted[j] = 1; }
        else                             { keep = p + j; discard = p + i; deleted[i] = 1; }
        if (discard->min_diag < keep->min_diag) keep->min_diag = discard->min_diag;
        if (discard->max_diag > keep->max_diag) keep->max_diag = discard->max_diag;
        if (discard->s_left_boundary  < keep->s_left_boundary)  keep->s_left_boundary  = discard->s_left_boundary;
        if (discard->s_right_boundary > keep->s_right_boundary) keep->s_right_boundary = discard->s_right_boundary;
        if (discard->t_left_boundary  < keep->t_left_boundary)  keep->t_left_boundary  = discard->t_left_boundary;
        if (discard->t_right_boundary > keep->t_right_boundary) keep->t_right_boundary = discard->t_right_boundary;
      }
    }
}

/* :291-311 */
static void choose_best_partial(oinfo_t *o, int ct, int *deleted) {
  int best = 0;
  double matching_bases = (1.0 - o[0].quality) * (2 + o[0].s_hi - o[0].s_lo + o[0].t_hi - o[0].t_lo);
  for (int i = 1; i < ct; i++) {
    double mb = (1.0 - o[i].quality) * (2 + o[i].s_hi - o[i].s_lo + o[i].t_hi - o[i].t_lo);
    if (matching_bases < mb || (matching_bases == mb && o[i].quality < o[best].quality))
      best = i;
  }
  for (int i = 0; i < ct; i++) deleted[i] = (i != best);
}

/* :177-244 */
static void add_overlap(const ovo_ctx *c, work_t *w, int s_lo, int s_hi, int t_lo, int t_hi, double qual, oinfo_t *o, int *ct) {
  if (!c->P.partial) {
    int new_diag = t_lo - s_lo;
    for (int i = 0; i < *ct; i++) {
      int old_diag = o[i].t_lo - o[i].s_lo;
      if ((new_diag >  0 && old_diag >  0 && o[i].t_right_boundary - new_diag - o[i].s_left_boundary >= MIN_INTERSECTION) ||
          (new_diag <= 0 && old_diag <= 0 && o[i].s_right_boundary + new_diag - o[i].t_left_boundary >= MIN_INTERSECTION)) {
        if (new_diag < o[i].min_diag) o[i].min_diag = new_diag;
        if (new_diag > o[i].max_diag) o[i].max_diag = new_diag;
        if (s_lo < o[i].s_left_boundary)  o[i].s_left_boundary  = s_lo;
        if (s_hi > o[i].s_right_boundary) o[i].s_right_boundary = s_hi;
        if (t_lo < o[i].t_left_boundary)  o[i].t_left_boundary  = t_lo;
        if (t_hi > o[i].t_right_boundary) o[i].t_right_boundary = t_hi;
        if (qual < o[i].quality) {
          o[i].s_lo = s_lo; o[i].s_hi = s_hi; o[i].t_lo = t_lo; o[i].t_hi = t_hi;
          o[i].quality = qual;
          o[i].delta_ct = w->ped.left_delta_len;
        }
        return;
      }
    }
  }
  if (*ct >= MAX_DISTINCT_OLAPS) return;
  oinfo_t *n = &o[*ct];
  n->s_lo = n->s_left_boundary  = s_lo;
  n->s_hi = n->s_right_boundary = s_hi;
  n->t_lo = n->t_left_boundary  = t_lo;
  n->t_hi = n->t_right_boundary = t_hi;
  n->quality = qual;
  n->delta_ct = w->ped.left_delta_len;
  n->min_diag = n->max_diag = t_lo - s_lo;
  (*ct)++;
}

/* :262-281 */
static int lies_on_alignment(const ped_t *p, int start, int offset, int s_lo, int t_lo) {
  int diag = t_lo - s_lo, new_diag = offset - start;
  for (int i = 0; i < p->left_delta_len; i++) {
    s_lo += abs(p->left_delta[i]);
    if (start < s_lo) return abs(new_diag - diag) <= SHIFT_SLACK;
    if (p->left_delta[i] < 0) diag++;
    else { s_lo++; diag--; }
  }
  return abs(new_diag - diag) <= SHIFT_SLACK;
}

static void trace_ext(ovo_ctx *c, work_t *w, uint32_t s_id, uint32_t t_id, int dir, const mnode_t *m,
                      int s_lo, int s_hi, int t_lo, int t_hi, int errors, int kind) {
  if (!c->trace_exts) return;
  if (w->et_len == w->et_max) { w->et_max = w->et_max ? w->et_max * 2 : 1024; w->et = (ovo_ext_trace *)realloc(w->et, sizeof(ovo_ext_trace) * w->et_max); }
  ovo_ext_trace *t = &w->et[w->et_len++];
  t->ref_id = s_id; t->hash_id = t_id; t->dir = dir;
  t->seed_start = m->start; t->seed_offset = m->offset; t->seed_len = m->len;
  t->s_lo = s_lo; t->s_hi = s_hi; t->t_lo = t_lo; t->t_hi = t_hi; t->errors = errors; t->kind = kind; t->delta_ct = w->ped.left_delta_len;
}

/* :355-547 (Frag_Olap_Limit is never reached: Canu never passes -l; hit_limit == false) */
static void process_matches(ovo_ctx *c, work_t *w, int32_t *start, const char *S, int s_len, uint32_t s_id, int dir,
                            const char *T, finfo_t t_info, uint32_t t_id, int consistent) {
  int t_len = (int)t_info.length;
  int kind = K_NONE;
  int s_lo = 0, s_hi = 0, t_lo = 0, t_hi = 0, errors = 0;
  int overlaps_output = 0;

  if (c->use_hopeless && w->node[*start].next == 0 && !c->P.partial) {
    int is_hopeless = 0;
    int s_head = w->node[*start].start, t_head = w->node[*start].offset;
    if (s_head <= t_head) { if (s_head > HOPELESS_MATCH && !w->left_end_screened) is_hopeless = 1; }
    else                  { if (t_head > HOPELESS_MATCH && !t_info.lscreen)       is_hopeless = 1; }
    int s_tail = s_len - s_head - w->node[*start].len + 1;
    int t_tail = t_len - t_head - w->node[*start].len + 1;
    if (s_tail <= t_tail) { if (s_tail > HOPELESS_MATCH && !w->right_end_screened) is_hopeless = 1; }
    else                  { if (t_tail > HOPELESS_MATCH && !t_info.rscreen)        is_hopeless = 1; }
    if (is_hopeless) {
      *start = 0;
      w->st.kmer_hits_without_olap++;
      return;
    }
  }

  oinfo_t *distinct = w->distinct;
  int distinct_ct = 0;

  while (*start != 0) {
    int max_len = w->node[*start].len;
    mnode_t *longest = &w->node[*start];
    for (int p = w->node[*start].next; p != 0; p = w->node[p].next)
      if (w->node[p].len > max_len) { max_len = w->node[p].len; longest = &w->node[p]; }

    kind = extend_alignment(c, &w->ped, longest, S, s_len, T, t_len, &s_lo, &s_hi, &t_lo, &t_hi, &errors);
    trace_ext(c, w, s_id, t_id, dir, longest, s_lo, s_hi, t_lo, t_hi, errors, kind);

    if (kind == K_DOVETAIL || c->P.partial) {
      if (1 + s_hi - s_lo >= c->P.min_olap_len && 1 + t_hi - t_lo >= c->P.min_olap_len) {
        int olap_len = 1 + ((s_hi - s_lo < t_hi - t_lo) ? s_hi - s_lo : t_hi - t_lo);
        double quality = (double)errors / olap_len;
        if (errors <= error_bound(c, olap_len))
          add_overlap(c, w, s_lo, s_hi, t_lo, t_hi, quality, distinct, &distinct_ct);
      }
    }

    if (consistent)
      *start = 0;

    for (int32_t *ref = start; *ref != 0; ) {
      mnode_t *ptr = &w->node[*ref];
      if (ptr == longest ||
          ((kind == K_DOVETAIL || c->P.partial) &&
           s_lo - SHIFT_SLACK <= ptr->start &&
           ptr->start + ptr->len <= (s_hi + 1) + SHIFT_SLACK - 1 &&
           lies_on_alignment(&w->ped, ptr->start, ptr->offset, s_lo, t_lo)))
        *ref = ptr->next;
      else
        ref = &ptr->next;
    }
  }

  if (distinct_ct > 0) {
    int deleted[MAX_DISTINCT_OLAPS] = {0};
    if (c->P.partial) {
      if (c->P.unique_per_pair) choose_best_partial(distinct, distinct_ct, deleted);
    } else {
      if (c->P.unique_per_pair) combine_into_one_olap(distinct, distinct_ct, deleted);
      else                      merge_intersecting_olaps(distinct, distinct_ct, deleted);
    }
    for (int i = 0; i < distinct_ct; i++)
      if (!deleted[i]) {
        if (c->P.partial) output_partial_overlap(w, s_id, t_id, dir, &distinct[i], s_len, t_len);
        else              output_overlap(w, s_id, s_len, dir, t_id, t_len, &distinct[i]);
        overlaps_output++;
      }
  }

  if (overlaps_output == 0) w->st.kmer_hits_without_olap++;
  else { w->st.kmer_hits_with_olap++; if (overlaps_output > 1) w->st.multi_overlap++; }
}

/* :22-33 */
static uint64_t compute_minimum_kmers(const ovo_ctx *c, double ovl_len) {
  if (c->filter_by_kmer_count == 0) return 0;
  ovl_len = (ovl_len < 0 ? ovl_len * -1.0 : ovl_len);
  uint64_t expected = 0;
  if (!(ovl_len < c->P.kmer_len))
    expected = (uint64_t)(int)floor(exp(-1.0 * (double)c->P.kmer_len * c->P.max_erate) * (ovl_len - c->P.kmer_len + 1));
  return c->filter_by_kmer_count > expected ? c->filter_by_kmer_count : expected;
}

static void trace_pair(ovo_ctx *c, work_t *w, uint32_t ref_id, uint32_t hash_id, int dir, const solap_t *o) {
  if (!c->trace_pairs) return;
  if (w->pt_len == w->pt_max) { w->pt_max = w->pt_max ? w->pt_max * 2 : 1024; w->pt = (ovo_pair_trace *)realloc(w->pt, sizeof(ovo_pair_trace) * w->pt_max); }
  ovo_pair_trace *t = &w->pt[w->pt_len++];
  t->ref_id = ref_id; t->hash_id = hash_id; t->dir = dir; t->consistent = o->consistent;
  t->diag_ct = o->diag_ct; t->diag_bgn = o->diag_bgn; t->diag_end = o->diag_end;
  t->seed_begin = (int64_t)w->sd_len; t->n_seeds = 0;
  for (int p = o->match_list; p != 0; p = w->node[p].next) {
    if (w->sd_len == w->sd_max) { w->sd_max = w->sd_max ? w->sd_max * 2 : 4096; w->sd = (ovo_seed *)realloc(w->sd, sizeof(ovo_seed) * w->sd_max); }
    w->sd[w->sd_len].start = w->node[p].start; w->sd[w->sd_len].offset = w->node[p].offset; w->sd[w->sd_len].len = w->node[p].len;
    w->sd_len++; t->n_seeds++;
  }
}

/* :581-690 with ct <= Frag_Olap_Limit always true (default UINT64_MAX; -l unsupported) */
static void process_string_olaps(ovo_ctx *c, work_t *w, const char *S, int len, uint32_t id, int dir) {
  int ct = 0;
  for (int i = 0; i < w->olap_next; i++)
    if (w->olap[i].full) {
      uint32_t root = w->olap[i].string_num;
      if (root + c->hash_string_num_offset > id) {
        if (i != ct) w->olap[ct] = w->olap[i];
        w->olap[ct].diag_sum /= w->olap[ct].diag_ct;
        ct++;
      }
    }
  for (int i = 0; i < ct; i++) {
    uint32_t root = w->olap[i].string_num;
    if (compute_minimum_kmers(c, w->olap[i].diag_end - w->olap[i].diag_bgn) > (uint64_t)w->olap[i].diag_ct) {
      w->st.kmer_hits_skipped++;
      continue;
    }
    trace_pair(c, w, id, (uint32_t)(root + c->hash_string_num_offset), dir, &w->olap[i]);
    process_matches(c, w, &w->olap[i].match_list, S, len, id, dir,
                    c->data + c->string_start[root], c->string_info[root],
                    (uint32_t)(root + c->hash_string_num_offset), w->olap[i].consistent);
  }
}

/* ======================================================================= */
/*  Driver  (overlapInCore.C:162-277, overlapInCore-Process_Overlaps.C)    */
/* ======================================================================= */

static void work_init(const ovo_ctx *c, work_t *w) {
  memset(w, 0, sizeof(*w));
  w->olap_size = 5000;  w->olap = (solap_t *)malloc(sizeof(solap_t) * w->olap_size);
  w->node_size = 10000; w->node = (mnode_t *)malloc(sizeof(mnode_t) * w->node_size);
  ped_init(&w->ped, c->max_errors);
}
static void work_free(const ovo_ctx *c, work_t *w) {
  free(w->olap); free(w->node); ped_free(&w->ped, c->max_errors);
  free(w->rec); free(w->pt); free(w->sd); free(w->et); free(w->fwd); free(w->rev);
}

#define APPEND(dst, dst_len, dst_max, src, src_len, T) do { \
    if ((dst_len) + (src_len) > (dst_max)) { (dst_max) = ((dst_len) + (src_len)) * 2 + 16; (dst) = (T *)realloc((dst), sizeof(T) * (dst_max)); } \
    if (src_len) memcpy((dst) + (dst_len), (src), sizeof(T) * (src_len)); (dst_len) += (src_len); } while (0)

static void work_merge(ovo_ctx *c, work_t *w) {
  uint64_t seed_base = c->sd_len;
  for (uint64_t i = 0; i < w->pt_len; i++) w->pt[i].seed_begin += (int64_t)seed_base;
  APPEND(c->rec, c->rec_len, c->rec_max, w->rec, w->rec_len, ovo_record);
  APPEND(c->pt,  c->pt_len,  c->pt_max,  w->pt,  w->pt_len,  ovo_pair_trace);
  APPEND(c->sd,  c->sd_len,  c->sd_max,  w->sd,  w->sd_len,  ovo_seed);
  APPEND(c->et,  c->et_len,  c->et_max,  w->et,  w->et_len,  ovo_ext_trace);
  w->rec_len = w->pt_len = w->sd_len = w->et_len = 0;
  c->st.kmer_hits_without_olap += w->st.kmer_hits_without_olap;
  c->st.kmer_hits_with_olap    += w->st.kmer_hits_with_olap;
  c->st.kmer_hits_skipped      += w->st.kmer_hits_skipped;
  c->st.multi_overlap          += w->st.multi_overlap;
  c->st.total_overlaps         += w->st.total_overlaps;
  c->st.contained              += w->st.contained;
  c->st.dovetail               += w->st.dovetail;
  c->st.ref_lookups            += w->st.ref_lookups;
  c->st.seed_hits              += w->st.seed_hits;
  c->st.extend_calls           += w->ped.calls;
  c->st.dp_cells               += w->ped.cells;
  c->st.char_compares          += w->ped.compares;
  memset(&w->st, 0, sizeof(w->st));
  w->ped.calls = w->ped.cells = w->ped.compares = 0;
}

/* Process_Overlaps.C:48-80 for one ref read */
static void process_ref_read(ovo_ctx *c, work_t *w, uint32_t fi) {
  uint32_t len = c->len[fi];
  if (len < (uint32_t)c->P.min_olap_len || len < c->P.kmer_len) return;
  w->fwd = (char *)realloc(w->fwd, len + 1);
  const char *src = c->bases + c->off[fi];
  for (uint32_t i = 0; i < len; i++) w->fwd[i] = (char)tolower(src[i]);
  w->fwd[len] = 0;
  find_overlaps(c, w, w->fwd, (int)len, fi, 0);
  revcomp_inplace(w->fwd, (int)len);
  find_overlaps(c, w, w->fwd, (int)len, fi, 1);
}

int ovo_run(ovo_ctx *c, uint32_t hb, uint32_t he, uint32_t rb, uint32_t re, int threads) {
  if (hb < 1) hb = 1;
  if (he > c->n_reads) he = c->n_reads;
  if (rb < 1) rb = 1;
  if (re > c->n_reads) re = c->n_reads;
  if (threads < 1) threads = 1;

  c->rec_len = c->pt_len = c->sd_len = c->et_len = 0;
  memset(&c->st, 0, sizeof(c->st));

  if (!c->table) {
    c->table = (bucket_t *)malloc(sizeof(bucket_t) * c->table_size);
    c->check_array = (uint32_t *)malloc(sizeof(uint32_t) * c->table_size);
  }

  work_t *wa = (work_t *)malloc(sizeof(work_t) * threads);
  for (int t = 0; t < threads; t++) work_init(c, &wa[t]);

  /* overlapInCore.C:204-263.  We implement the INTENDED semantics: every read of
     the -h range is hashed and every read of the -r range is searched (the
     reference's `<` loop bounds drop a trailing read in some -t/-h settings,
     SURVEY.md 7.5; the golden fixtures are minted with settings where that
     quirk does not fire). */
  uint32_t bgn = hb;
  while (bgn <= he) {
    uint32_t end = build_hash_index(c, bgn, he);
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
#endif
    for (int64_t fi = rb; fi <= (int64_t)re; fi++) {
#ifdef _OPENMP
      work_t *w = &wa[omp_get_thread_num()];
#else
      work_t *w = &wa[0];
#endif
      process_ref_read(c, w, (uint32_t)fi);
    }
    bgn = end + 1;
  }

  for (int t = 0; t < threads; t++) { work_merge(c, &wa[t]); work_free(c, &wa[t]); }
  free(wa);
  return 0;
}

/* ======================================================================= */
/*  Context                                                                */
/* ======================================================================= */

ovo_ctx *ovo_create(const ovo_params *p) {
  ovo_ctx *c = (ovo_ctx *)calloc(1, sizeof(ovo_ctx));
  c->P = *p;
  if (c->P.hash_bits == 0) c->P.hash_bits = 22;
  if (c->P.hash_load == 0) c->P.hash_load = 0.6;
  if (c->P.hash_data_len == 0) c->P.hash_data_len = 100000000ull;
  if (c->P.align_noise == 0) c->P.align_noise = 1.0;

  /* overlapInCore.C:400-411, :454-459 */
  c->use_hopeless = !c->P.no_hopeless;
  if (c->P.max_erate > 0.06) c->use_hopeless = 0;
  c->filter_by_kmer_count = 0;
  if (c->P.min_kmers)
    c->filter_by_kmer_count = (uint64_t)(int)floor(exp(-1.0 * (double)c->P.kmer_len * c->P.max_erate) * (c->P.min_olap_len - (double)c->P.kmer_len + 1));
  c->max_hash_data_len = c->P.hash_data_len + OVO_MAX_READLEN;
  c->HSF1 = c->P.kmer_len - (c->P.hash_bits / 2);
  c->HSF2 = 2 * (uint64_t)c->P.kmer_len - c->P.hash_bits;
  c->SV1  = c->HSF1 + 2;
  c->SV2  = (c->HSF1 + c->HSF2) / 2;
  c->SV3  = c->HSF2 - 2;
  c->hash_mask  = (1ull << c->P.hash_bits) - 1;
  c->table_size = 1ull << c->P.hash_bits;

  /* prefixEditDistance.C:23-107 */
  c->max_errors = 1 + (uint32_t)(int)ceil(c->P.max_erate * OVO_MAX_READLEN);
  c->min_branch_tail_slope = (c->P.max_erate > 0.06) ? 1.0 : 0.20;
  c->eml = (int32_t *)calloc(c->max_errors + 1, sizeof(int32_t));
  init_match_limit(c->eml, c->P.max_erate * c->P.align_noise, (int32_t)c->max_errors);
  c->branch_match_value = c->P.max_erate / (1 + c->P.max_erate);
  return c;
}

void ovo_destroy(ovo_ctx *c) {
  if (!c) return;
  free(c->eml); free(c->bases); free(c->off); free(c->len); free(c->skip);
  free(c->table); free(c->check_array); free(c->data); free(c->next_ref); free(c->extra_ref);
  free(c->string_start); free(c->string_info);
  free(c->rec); free(c->pt); free(c->sd); free(c->et);
  free(c);
}

int ovo_set_reads(ovo_ctx *c, uint32_t n, const char *bases, const uint64_t *offsets, const uint32_t *lens) {
  free(c->bases); free(c->off); free(c->len);
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; i++) total += lens[i];
  c->n_reads = n;
  c->bases = (char *)malloc(total + 1);
  c->off = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
  c->len = (uint32_t *)calloc(n + 1, sizeof(uint32_t));
  uint64_t p = 0;
  for (uint32_t i = 0; i < n; i++) {
    c->off[i + 1] = p;
    c->len[i + 1] = lens[i];
    memcpy(c->bases + p, bases + offsets[i], lens[i]);
    p += lens[i];
  }
  return 0;
}

int ovo_set_skip_kmers(ovo_ctx *c, uint32_t n, const char *kmers) {
  free(c->skip);
  c->n_skip = n;
  c->skip = (char *)malloc((uint64_t)n * c->P.kmer_len + 1);
  memcpy(c->skip, kmers, (uint64_t)n * c->P.kmer_len);
  return 0;
}

void ovo_enable_trace(ovo_ctx *c, int pairs, int exts) { c->trace_pairs = pairs; c->trace_exts = exts; }

uint64_t          ovo_num_records(const ovo_ctx *c) { return c->rec_len; }
const ovo_record *ovo_records(const ovo_ctx *c)     { return c->rec; }
void              ovo_get_stats(const ovo_ctx *c, ovo_stats *s) { *s = c->st; }
uint64_t              ovo_num_pair_traces(const ovo_ctx *c) { return c->pt_len; }
const ovo_pair_trace *ovo_pair_traces(const ovo_ctx *c)     { return c->pt; }
const ovo_seed       *ovo_seed_traces(const ovo_ctx *c)     { return c->sd; }
uint64_t              ovo_num_ext_traces(const ovo_ctx *c)  { return c->et_len; }
const ovo_ext_trace  *ovo_ext_traces(const ovo_ctx *c)      { return c->et; }

uint32_t       ovo_max_errors(const ovo_ctx *c)        { return c->max_errors; }
const int32_t *ovo_edit_match_limit(const ovo_ctx *c)  { return c->eml; }
int32_t        ovo_error_bound(const ovo_ctx *c, int32_t len) { return error_bound(c, len); }
double         ovo_branch_match_value(const ovo_ctx *c) { return c->branch_match_value; }

int ovo_extend_one(ovo_ctx *c, const char *S, int32_t s_len, const char *T, int32_t t_len,
                   int32_t seed_start, int32_t seed_offset, int32_t seed_len,
                   ovo_ext_trace *out, int32_t *delta_out, int32_t delta_cap) {
  ped_t p; ped_init(&p, c->max_errors);
  char *s = (char *)malloc(s_len + 2), *t = (char *)malloc(t_len + 2);
  s[0] = t[0] = 0;   /* guard byte so reverse() may read index -1 safely: strings start at +1 */
  for (int i = 0; i < s_len; i++) s[i + 1] = (char)tolower(S[i]);
  for (int i = 0; i < t_len; i++) t[i + 1] = (char)tolower(T[i]);
  mnode_t m = { seed_offset, seed_len, seed_start, 0 };
  memset(out, 0, sizeof(*out));
  out->seed_start = seed_start; out->seed_offset = seed_offset; out->seed_len = seed_len;
  out->kind = extend_alignment(c, &p, &m, s + 1, s_len, t + 1, t_len, &out->s_lo, &out->s_hi, &out->t_lo, &out->t_hi, &out->errors);
  out->delta_ct = p.left_delta_len;
  for (int i = 0; i < p.left_delta_len && i < delta_cap; i++) delta_out[i] = p.left_delta[i];
  free(s); free(t);
  ped_free(&p, c->max_errors);
  return 0;
}
