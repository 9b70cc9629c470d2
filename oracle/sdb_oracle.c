/*
 * sdb_oracle.c -- CPU ORACLE (test infrastructure only; see sdb_oracle.h).
 * Plain-C restatement of the reference's KNN / HNSW / graph-expansion algorithms.
 * Paths cited are relative to /root/reference/surrealdb/core/src.
 */
#include "sdb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * val::Number (Int | Float)                                            val/number.rs
 * ---------------------------------------------------------------------------------------- */
static inline orc_num N_f(double f) {
  orc_num n;
  n.tag = 0;
  n.v.f = f;
  return n;
}
static inline orc_num N_i(int64_t i) {
  orc_num n;
  n.tag = 1;
  n.v.i = i;
  return n;
}
static inline double N_to_float(orc_num a) { return a.tag ? (double)a.v.i : a.v.f; }
static inline int64_t wrap_add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int64_t wrap_sub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
static inline int64_t wrap_mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }

/* ops::Add / Sub / Mul / Div for Number -- val/number.rs:926-1044 */
static orc_num N_add(orc_num a, orc_num b) {
  if (a.tag && b.tag) return N_i(wrap_add(a.v.i, b.v.i));
  if (!a.tag && !b.tag) return N_f(a.v.f + b.v.f);
  if (a.tag) return N_f((double)a.v.i + b.v.f);
  return N_f(a.v.f + (double)b.v.i);
}
static orc_num N_sub(orc_num a, orc_num b) {
  if (a.tag && b.tag) return N_i(wrap_sub(a.v.i, b.v.i));
  if (!a.tag && !b.tag) return N_f(a.v.f - b.v.f);
  if (a.tag) return N_f((double)a.v.i - b.v.f);
  return N_f(a.v.f - (double)b.v.i);
}
static orc_num N_mul(orc_num a, orc_num b) {
  if (a.tag && b.tag) return N_i(wrap_mul(a.v.i, b.v.i));
  if (!a.tag && !b.tag) return N_f(a.v.f * b.v.f);
  if (a.tag) return N_f((double)a.v.i * b.v.f);
  return N_f(a.v.f * (double)b.v.i);
}
static orc_num N_div(orc_num a, orc_num b) {
  if (a.tag && b.tag) return b.v.i == 0 ? N_f(NAN) /* Rust panics; unreachable for metrics */ : N_i(a.v.i / b.v.i);
  if (!a.tag && !b.tag) return N_f(a.v.f / b.v.f);
  if (a.tag) return N_f((double)a.v.i / b.v.f);
  return N_f(a.v.f / (double)b.v.i);
}
static orc_num N_abs(orc_num a) { return a.tag ? N_i(a.v.i < 0 ? -a.v.i : a.v.i) : N_f(fabs(a.v.f)); }

/* f64::total_cmp */
static inline int64_t f64_total_key(double d) {
  int64_t b;
  memcpy(&b, &d, 8);
  b ^= (int64_t)(((uint64_t)(b >> 63)) >> 1);
  return b;
}
static inline int f64_total_cmp(double a, double b) {
  int64_t x = f64_total_key(a), y = f64_total_key(b);
  return (x > y) - (x < y);
}
/* total_cmp_f64 inside Number::cmp: -0.0 == 0.0, otherwise total_cmp  val/number.rs:621-629 */
static inline int num_total_cmp_f64(double a, double b) {
  if (a == 0.0 && b == 0.0) return 0;
  return f64_total_cmp(a, b);
}
static int cmp_int_float(int64_t v, double w) { /* val/number.rs:647-661 */
  if (!isfinite(w)) return signbit(w) ? 1 : -1; /* greater!(w).reverse() */
  __int128 l = (__int128)v;
  __int128 r = (__int128)w; /* truncation like `as i128` (finite) */
  if (l != r) return (l > r) - (l < r);
  double ip;
  double fr = modf(w, &ip);
  return num_total_cmp_f64(0.0, fr);
}
int orc_num_cmp(const orc_num* a, const orc_num* b) {
  if (a->tag && b->tag) return (a->v.i > b->v.i) - (a->v.i < b->v.i);
  if (!a->tag && !b->tag) return num_total_cmp_f64(a->v.f, b->v.f);
  if (a->tag) return cmp_int_float(a->v.i, b->v.f);
  return -cmp_int_float(b->v.i, a->v.f);
}
/* PartialEq for Number  val/number.rs:788-808 */
static int N_eq(orc_num a, orc_num b) {
  if (!a.tag && !b.tag) {
    uint64_t x, y;
    memcpy(&x, &a.v.f, 8);
    memcpy(&y, &b.v.f, 8);
    return x == y || (a.v.f == 0.0 && b.v.f == 0.0);
  }
  return orc_num_cmp(&a, &b) == 0;
}

/* ------------------------------------------------------------------------------------------
 * a1-a3: Vec<Number> metrics                                 fnc/util/math/vector.rs
 * ---------------------------------------------------------------------------------------- */
#define CHECK_DIM(na, nb) \
  if ((na) != (nb)) return ORC_EDIM /* check_same_dimension  vector.rs:23-32 */

/* fn dot: a.iter().zip(b).map(|(a,b)| a*b).sum()  with Sum = fold(Number::Int(0), +)
 * vector.rs:279-281, val/number.rs:1052-1068 */
static orc_num num_dot(const orc_num* a, const orc_num* b, size_t n) {
  orc_num acc = N_i(0);
  for (size_t i = 0; i < n; i++) acc = N_add(acc, N_mul(a[i], b[i]));
  return acc;
}
/* magnitude_squared: v.iter().map(|a| a.to_float().powi(2)).sum::<f64>()  vector.rs:301-303 */
static double num_mag2(const orc_num* a, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; i++) {
    double x = N_to_float(a[i]);
    s += x * x;
  }
  return s;
}
int orc_num_dot(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  CHECK_DIM(na, nb);
  *out = num_dot(a, b, na);
  return ORC_OK;
}
int orc_num_magnitude(const orc_num* a, size_t na, orc_num* out) { /* vector.rs:310-314 */
  *out = N_f(sqrt(num_mag2(a, na)));
  return ORC_OK;
}
int orc_num_cosine_similarity(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  CHECK_DIM(na, nb); /* vector.rs:77-83 */
  orc_num d = num_dot(a, b, na);
  orc_num m = N_mul(N_f(sqrt(num_mag2(a, na))), N_f(sqrt(num_mag2(b, nb))));
  *out = N_div(d, m);
  return ORC_OK;
}
int orc_num_cosine_distance(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  CHECK_DIM(na, nb); /* vector.rs:65-71 */
  orc_num d = num_dot(a, b, na);
  orc_num m = N_mul(N_f(sqrt(num_mag2(a, na))), N_f(sqrt(num_mag2(b, nb))));
  *out = N_sub(N_i(1), N_div(d, m));
  return ORC_OK;
}
int orc_num_euclidean(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  CHECK_DIM(na, nb); /* vector.rs:288-299 */
  double s = 0.0;
  for (size_t i = 0; i < na; i++) {
    double x = N_to_float(N_sub(a[i], b[i]));
    s += x * x;
  }
  *out = N_f(sqrt(s));
  return ORC_OK;
}
int orc_num_manhattan(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  CHECK_DIM(na, nb); /* vector.rs:152-157 */
  orc_num acc = N_i(0);
  for (size_t i = 0; i < na; i++) acc = N_add(acc, N_abs(N_sub(a[i], b[i])));
  *out = acc;
  return ORC_OK;
}
/* Rust f64::max: returns the non-NaN operand */
static inline double rust_fmax(double a, double b) {
  if (isnan(a)) return b;
  if (isnan(b)) return a;
  return a > b ? a : b;
}
int orc_num_chebyshev(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  CHECK_DIM(na, nb); /* vector.rs:215-225: fold(f64::MIN, f64::max) */
  double m = -1.7976931348623157e308;
  for (size_t i = 0; i < na; i++) m = rust_fmax(m, fabs(N_to_float(a[i]) - N_to_float(b[i])));
  *out = N_f(m);
  return ORC_OK;
}
int orc_num_hamming(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  CHECK_DIM(na, nb); /* vector.rs:111-116 */
  int64_t c = 0;
  for (size_t i = 0; i < na; i++) c += !N_eq(a[i], b[i]);
  *out = N_i(c);
  return ORC_OK;
}
int orc_num_minkowski(const orc_num* a, size_t na, const orc_num* b, size_t nb, double p, orc_num* out) {
  CHECK_DIM(na, nb); /* vector.rs:163-174 */
  double s = 0.0;
  for (size_t i = 0; i < na; i++) s += pow(fabs(N_to_float(a[i]) - N_to_float(b[i])), p);
  *out = N_f(pow(s, 1.0 / p));
  return ORC_OK;
}
static double num_mean(const orc_num* a, size_t n) { /* fnc/util/math/mod.rs:54-69 */
  if (n == 0) return NAN;
  double s = 0.0;
  for (size_t i = 0; i < n; i++) s = s + N_to_float(a[i]);
  return s / (double)n;
}
static double num_deviation(const orc_num* a, size_t n, double mean) { /* vector.rs:9-21 (sample=false) */
  if (n == 0) return NAN;
  if (n == 1) return 0.0;
  double s = 0.0;
  for (size_t i = 0; i < n; i++) {
    double x = N_to_float(a[i]) - mean;
    s += x * x;
  }
  return sqrt(s / (double)n);
}
int orc_num_pearson(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  CHECK_DIM(na, nb); /* vector.rs:133-146 */
  double m1 = num_mean(a, na), m2 = num_mean(b, nb);
  double covar = 0.0;
  for (size_t i = 0; i < na; i++) covar += (N_to_float(a[i]) - m1) * (N_to_float(b[i]) - m2);
  covar = covar / (double)na;
  *out = N_f(covar / (num_deviation(a, na, m1) * num_deviation(b, nb, m2)));
  return ORC_OK;
}
int orc_num_jaccard(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out) {
  /* vector.rs:121-127: union = set(a); inter = |{x in b : !union.insert(x)}|; inter/|union| */
  orc_num* set = (orc_num*)malloc(sizeof(orc_num) * (na + nb + 1));
  size_t ns = 0, inter = 0;
  for (size_t i = 0; i < na; i++) {
    int found = 0;
    for (size_t j = 0; j < ns && !found; j++) found = N_eq(set[j], a[i]);
    if (!found) set[ns++] = a[i];
  }
  for (size_t i = 0; i < nb; i++) {
    int found = 0;
    for (size_t j = 0; j < ns && !found; j++) found = N_eq(set[j], b[i]);
    if (found)
      inter++;
    else
      set[ns++] = b[i];
  }
  free(set);
  *out = N_f((double)inter / (double)ns);
  return ORC_OK;
}
int orc_num_distance(int metric, double p, const orc_num* a, size_t na, const orc_num* b, size_t nb,
                     orc_num* out) { /* Distance::compute catalog/schema/index.rs:287-303 */
  switch (metric) {
    case ORC_COSINE: return orc_num_cosine_distance(a, na, b, nb, out);
    case ORC_CHEBYSHEV: return orc_num_chebyshev(a, na, b, nb, out);
    case ORC_EUCLIDEAN: return orc_num_euclidean(a, na, b, nb, out);
    case ORC_HAMMING: return orc_num_hamming(a, na, b, nb, out);
    case ORC_JACCARD: return orc_num_jaccard(a, na, b, nb, out);
    case ORC_MANHATTAN: return orc_num_manhattan(a, na, b, nb, out);
    case ORC_MINKOWSKI: return orc_num_minkowski(a, na, b, nb, p, out);
    case ORC_PEARSON: return orc_num_pearson(a, na, b, nb, out);
  }
  return ORC_EINVAL;
}

/* ---- all-Float fast path (same op sequence as the Number path when every element is Float:
 * dot acc starts as Int(0); Int(0)+Float(p0) = 0 as f64 + p0; then plain left-to-right f64). */
#define DEF_FAST(NAME_COS, NAME_EUC, NAME_MAG, T)                                   \
  double NAME_MAG(const T* a, size_t n) {                                           \
    double s = 0.0;                                                                 \
    for (size_t i = 0; i < n; i++) {                                                \
      double x = (double)a[i];                                                      \
      s += x * x;                                                                   \
    }                                                                               \
    return sqrt(s);                                                                 \
  }                                                                                 \
  double NAME_COS(const T* a, const double* b, size_t n) {                          \
    double dot = 0.0, ma2 = 0.0, mb2 = 0.0;                                         \
    for (size_t i = 0; i < n; i++) dot = dot + (double)a[i] * b[i];                 \
    for (size_t i = 0; i < n; i++) {                                                \
      double x = (double)a[i];                                                      \
      ma2 += x * x;                                                                 \
    }                                                                               \
    for (size_t i = 0; i < n; i++) mb2 += b[i] * b[i];                              \
    return 1.0 - dot / (sqrt(ma2) * sqrt(mb2));                                     \
  }                                                                                 \
  double NAME_EUC(const T* a, const double* b, size_t n) {                          \
    double s = 0.0;                                                                 \
    for (size_t i = 0; i < n; i++) {                                                \
      double x = (double)a[i] - b[i];                                               \
      s += x * x;                                                                   \
    }                                                                               \
    return sqrt(s);                                                                 \
  }
DEF_FAST(orc_f64_cosine_distance, orc_f64_euclidean, orc_f64_magnitude, double)
DEF_FAST(orc_f32row_cosine_distance, orc_f32row_euclidean, orc_f32row_magnitude, float)

/* all-Float fast paths of the remaining Distance::compute metrics (same op order as orc_num_* on Floats);
 * `row` is f32 or f64, `q` f64.  Used by orc_knn_topk for the metrics the GPU serves through its exact kernel. */
static double g_minkowski_p = 3.0;
void orc_set_minkowski_order(double p) { g_minkowski_p = p; }
#define DEF_MORE(SUFFIX, T)                                                                      \
  static double fast_manhattan_##SUFFIX(const T* a, const double* b, size_t n) {                 \
    double s = 0.0;                                                                              \
    for (size_t i = 0; i < n; i++) s = s + fabs((double)a[i] - b[i]);                            \
    return s;                                                                                    \
  }                                                                                              \
  static double fast_chebyshev_##SUFFIX(const T* a, const double* b, size_t n) {                 \
    double m = -1.7976931348623157e308;                                                          \
    for (size_t i = 0; i < n; i++) m = rust_fmax(m, fabs((double)a[i] - b[i]));                  \
    return m;                                                                                    \
  }                                                                                              \
  static double fast_hamming_##SUFFIX(const T* a, const double* b, size_t n) {                   \
    int64_t c = 0;                                                                               \
    for (size_t i = 0; i < n; i++) c += !N_eq(N_f((double)a[i]), N_f(b[i]));                     \
    return (double)c;                                                                            \
  }                                                                                              \
  static double fast_pearson_##SUFFIX(const T* a, const double* b, size_t n) {                   \
    double s1 = 0.0, s2 = 0.0;                                                                   \
    for (size_t i = 0; i < n; i++) s1 = s1 + (double)a[i];                                       \
    for (size_t i = 0; i < n; i++) s2 = s2 + b[i];                                               \
    const double m1 = n ? s1 / (double)n : NAN, m2 = n ? s2 / (double)n : NAN;                   \
    double covar = 0.0, d1 = 0.0, d2 = 0.0;                                                      \
    for (size_t i = 0; i < n; i++) covar += ((double)a[i] - m1) * (b[i] - m2);                   \
    covar = covar / (double)n;                                                                   \
    for (size_t i = 0; i < n; i++) { double x = (double)a[i] - m1; d1 += x * x; }                \
    for (size_t i = 0; i < n; i++) { double x = b[i] - m2; d2 += x * x; }                        \
    const double sd1 = n == 0 ? NAN : (n == 1 ? 0.0 : sqrt(d1 / (double)n));                     \
    const double sd2 = n == 0 ? NAN : (n == 1 ? 0.0 : sqrt(d2 / (double)n));                     \
    return covar / (sd1 * sd2);                                                                  \
  }                                                                                              \
  /* vector.rs:163-174: sum |a-b|^p, then ^(1/p); p = the Distance::Minkowski(order) of the index / operator */        \
  static double fast_minkowski_##SUFFIX(const T* a, const double* b, size_t n) {                 \
    double s = 0.0;                                                                              \
    for (size_t i = 0; i < n; i++) s += pow(fabs((double)a[i] - b[i]), g_minkowski_p);           \
    return pow(s, 1.0 / g_minkowski_p);                                                          \
  }                                                                                              \
  /* vector.rs:121-127 on Floats: union = set(a); every x of b already in the (growing) union counts */               \
  static double fast_jaccard_##SUFFIX(const T* a, const double* b, size_t n) {                   \
    double* set = (double*)malloc(sizeof(double) * (2 * n + 1));                                 \
    size_t ns = 0, inter = 0;                                                                    \
    for (size_t i = 0; i < n; i++) {                                                             \
      int found = 0;                                                                             \
      for (size_t j = 0; j < ns && !found; j++) found = N_eq(N_f(set[j]), N_f((double)a[i]));    \
      if (!found) set[ns++] = (double)a[i];                                                      \
    }                                                                                            \
    for (size_t i = 0; i < n; i++) {                                                             \
      int found = 0;                                                                             \
      for (size_t j = 0; j < ns && !found; j++) found = N_eq(N_f(set[j]), N_f(b[i]));            \
      if (found) inter++;                                                                        \
      else set[ns++] = b[i];                                                                     \
    }                                                                                            \
    free(set);                                                                                   \
    return (double)inter / (double)ns;                                                           \
  }
DEF_MORE(f64, double)
DEF_MORE(f32, float)
double orc_f64_metric(int metric, const double* a, const double* b, size_t n) {
  switch (metric) {
    case ORC_COSINE: return orc_f64_cosine_distance(a, b, n);
    case ORC_EUCLIDEAN: return orc_f64_euclidean(a, b, n);
    case ORC_MANHATTAN: return fast_manhattan_f64(a, b, n);
    case ORC_CHEBYSHEV: return fast_chebyshev_f64(a, b, n);
    case ORC_HAMMING: return fast_hamming_f64(a, b, n);
    case ORC_PEARSON: return fast_pearson_f64(a, b, n);
    case ORC_MINKOWSKI: return fast_minkowski_f64(a, b, n);
    case ORC_JACCARD: return fast_jaccard_f64(a, b, n);
  }
  return NAN;
}

/* ------------------------------------------------------------------------------------------
 * a4: KnnTopK                                         exec/operators/knn_topk.rs:166-267
 * Bounded selection with the DistanceEntry order: worst = max by (distance, seq); a new entry
 * replaces the worst only when strictly closer (Number::cmp).  Output (distance asc, seq asc).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double d;
  uint64_t seq;
  uint64_t row;
} topk_ent;
static inline int topk_worse(const topk_ent* a, const topk_ent* b) { /* a is farther/later than b */
  int c = num_total_cmp_f64(a->d, b->d);
  if (c) return c > 0;
  return a->seq > b->seq;
}
static double row_distance(const void* corpus, int is_f64, size_t r, size_t dim, const double* q, int metric) {
  if (is_f64) {
    const double* row = (const double*)corpus + r * dim;
    return orc_f64_metric(metric, row, q, dim);
  }
  const float* row = (const float*)corpus + r * dim;
  switch (metric) {
    case ORC_COSINE: return orc_f32row_cosine_distance(row, q, dim);
    case ORC_EUCLIDEAN: return orc_f32row_euclidean(row, q, dim);
    case ORC_MANHATTAN: return fast_manhattan_f32(row, q, dim);
    case ORC_CHEBYSHEV: return fast_chebyshev_f32(row, q, dim);
    case ORC_HAMMING: return fast_hamming_f32(row, q, dim);
    case ORC_PEARSON: return fast_pearson_f32(row, q, dim);
    case ORC_MINKOWSKI: return fast_minkowski_f32(row, q, dim);
    case ORC_JACCARD: return fast_jaccard_f32(row, q, dim);
  }
  return NAN;
}
static int topk_cmp_sort(const void* x, const void* y) {
  const topk_ent *a = (const topk_ent*)x, *b = (const topk_ent*)y;
  int c = num_total_cmp_f64(a->d, b->d);
  if (c) return c;
  return (a->seq > b->seq) - (a->seq < b->seq);
}
size_t orc_knn_topk(const void* corpus, int is_f64, size_t n_rows, size_t dim, const uint8_t* skip,
                    const double* q, int metric, size_t k, uint64_t* out_rows, double* out_dist) {
  if (k == 0) return 0;
  topk_ent* heap = (topk_ent*)malloc(sizeof(topk_ent) * k);
  size_t n = 0, worst = 0;
  uint64_t seq = 0;
  for (size_t r = 0; r < n_rows; r++) {
    if (skip && skip[r]) continue; /* extract_vector -> None: knn_topk.rs:199-203 */
    topk_ent e;
    e.d = row_distance(corpus, is_f64, r, dim, q, metric);
    e.seq = seq++;
    e.row = r;
    if (n < k) {
      heap[n] = e;
      if (n == 0 || topk_worse(&heap[n], &heap[worst])) worst = n;
      n++;
    } else if (num_total_cmp_f64(e.d, heap[worst].d) < 0) { /* entry.distance < worst.distance */
      heap[worst] = e;
      worst = 0;
      for (size_t i = 1; i < n; i++)
        if (topk_worse(&heap[i], &heap[worst])) worst = i;
    }
  }
  qsort(heap, n, sizeof(topk_ent), topk_cmp_sort);
  for (size_t i = 0; i < n; i++) {
    out_rows[i] = heap[i].row;
    out_dist[i] = heap[i].d;
  }
  free(heap);
  return n;
}

typedef struct {
  const void* corpus;
  int is_f64;
  size_t n_rows, dim;
  const double* queries;
  size_t q0, q1;
  int metric;
  size_t k;
  uint64_t* out_rows;
  double* out_dist;
} batch_job;
static void* batch_worker(void* p) {
  batch_job* j = (batch_job*)p;
  for (size_t q = j->q0; q < j->q1; q++)
    orc_knn_topk(j->corpus, j->is_f64, j->n_rows, j->dim, NULL, j->queries + q * j->dim, j->metric, j->k,
                 j->out_rows + q * j->k, j->out_dist + q * j->k);
  return NULL;
}
void orc_knn_topk_batch(const void* corpus, int is_f64, size_t n_rows, size_t dim, const double* queries,
                        size_t nq, int metric, size_t k, uint64_t* out_rows, double* out_dist, int nt) {
  if (nt < 1) nt = 1;
  if ((size_t)nt > nq) nt = (int)(nq ? nq : 1);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
  batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * nt);
  for (int t = 0; t < nt; t++) {
    batch_job j = {corpus, is_f64, n_rows, dim, queries, nq * t / nt, nq * (t + 1) / nt, metric, k, out_rows, out_dist};
    jobs[t] = j;
    pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}

/* ------------------------------------------------------------------------------------------
 * a6: typed ndarray metrics                                   idx/trees/vector.rs:235-289
 * ndarray 0.17.2 numeric_util::unrolled_dot / unrolled_fold (crate not vendored: restated from
 * its published source; PARITY UNPINNED -- isolated here so it can be corrected in one place):
 *   8 partial sums p0..p7 over the floor(n/8)*8 prefix (lane j takes elements = j mod 8),
 *   sum = 0; sum += p0+p4; sum += p1+p5; sum += p2+p6; sum += p3+p7; then the <=7 tail in order.
 * ---------------------------------------------------------------------------------------- */
float orc_nd_dot_f32(const float* a, const float* b, size_t n) {
  float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; j++) p[j] = p[j] + a[i + j] * b[i + j];
  float sum = 0.0f;
  sum = sum + (p[0] + p[4]);
  sum = sum + (p[1] + p[5]);
  sum = sum + (p[2] + p[6]);
  sum = sum + (p[3] + p[7]);
  for (; i < n; i++) sum = sum + a[i] * b[i];
  return sum;
}
float orc_nd_sumsq_f32(const float* a, size_t n) { /* (a*a).sum(): temp array of f32 products, unrolled_fold */
  float p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; j++) {
      float t = a[i + j] * a[i + j];
      p[j] = p[j] + t;
    }
  float sum = 0.0f;
  sum = sum + (p[0] + p[4]);
  sum = sum + (p[1] + p[5]);
  sum = sum + (p[2] + p[6]);
  sum = sum + (p[3] + p[7]);
  for (; i < n; i++) {
    float t = a[i] * a[i];
    sum = sum + t;
  }
  return sum;
}
static double nd_dot_f64(const double* a, const double* b, size_t n) {
  double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; j++) p[j] = p[j] + a[i + j] * b[i + j];
  double sum = 0.0;
  sum = sum + (p[0] + p[4]);
  sum = sum + (p[1] + p[5]);
  sum = sum + (p[2] + p[6]);
  sum = sum + (p[3] + p[7]);
  for (; i < n; i++) sum = sum + a[i] * b[i];
  return sum;
}
double orc_vec_cosine_f32(const float* a, const float* b, size_t n) { /* vector.rs:243-249 */
  double dot = (double)orc_nd_dot_f32(a, b, n);
  double na = sqrt((double)orc_nd_sumsq_f32(a, n));
  double nb = sqrt((double)orc_nd_sumsq_f32(b, n));
  return 1.0 - dot / (na * nb);
}
double orc_vec_cosine_f64(const double* a, const double* b, size_t n) { /* vector.rs:235-241 */
  double dot = nd_dot_f64(a, b, n);
  double na = sqrt(nd_dot_f64(a, a, n));
  double nb = sqrt(nd_dot_f64(b, b, n));
  return 1.0 - dot / (na * nb);
}
/* ndarray-stats 0.7.0 sq_l2_dist: Zip fold in element type, sequential; l2_dist = sqrt(to_f64) */
double orc_vec_l2_f32(const float* a, const float* b, size_t n) {
  float s = 0.0f;
  for (size_t i = 0; i < n; i++) {
    float d = a[i] - b[i];
    s = s + d * d;
  }
  return sqrt((double)s);
}
double orc_vec_l2_f64(const double* a, const double* b, size_t n) {
  double s = 0.0;
  for (size_t i = 0; i < n; i++) {
    double d = a[i] - b[i];
    s = s + d * d;
  }
  return sqrt(s);
}
double orc_vec_distance_f32(int metric, const float* a, const float* b, size_t n) { /* vector.rs:659-672 */
  switch (metric) {
    case ORC_COSINE: return orc_vec_cosine_f32(a, b, n);
    case ORC_EUCLIDEAN: return orc_vec_l2_f32(a, b, n);
    case ORC_MANHATTAN: { /* l1_dist, f32 accumulate */
      float s = 0.0f;
      for (size_t i = 0; i < n; i++) s = s + fabsf(a[i] - b[i]);
      return (double)s;
    }
    case ORC_CHEBYSHEV: { /* linf_dist */
      float m = 0.0f;
      for (size_t i = 0; i < n; i++) {
        float d = fabsf(a[i] - b[i]);
        if (d > m) m = d;
      }
      return (double)m;
    }
    case ORC_HAMMING: {
      size_t c = 0;
      for (size_t i = 0; i < n; i++) c += a[i] != b[i];
      return (double)c;
    }
  }
  return NAN;
}

/* ------------------------------------------------------------------------------------------
 * A4: DoublePriorityQueue                                        idx/trees/knn.rs:15-123
 * BTreeMap<FloatKey, VecDeque<id>>: stored here as one array sorted by (total_cmp key, arrival
 * seq).  pop_first = smallest key, oldest id; pop_last = largest key, NEWEST id.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double d;
  uint64_t seq, id;
} dpq_ent;
struct orc_dpq {
  dpq_ent* e;
  size_t n, cap;
  uint64_t next_seq;
};
static void dpq_init(orc_dpq* q) {
  q->e = NULL;
  q->n = q->cap = 0;
  q->next_seq = 0;
}
static void dpq_destroy(orc_dpq* q) {
  free(q->e);
  q->e = NULL;
  q->n = q->cap = 0;
}
static void dpq_copy(orc_dpq* dst, const orc_dpq* src) {
  dst->n = src->n;
  dst->cap = src->n + 8;
  dst->next_seq = src->next_seq;
  dst->e = (dpq_ent*)malloc(sizeof(dpq_ent) * dst->cap);
  if (src->n) memcpy(dst->e, src->e, sizeof(dpq_ent) * src->n);
}
static void dpq_push(orc_dpq* q, double d, uint64_t id) {
  if (q->n == q->cap) {
    q->cap = q->cap ? q->cap * 2 : 16;
    q->e = (dpq_ent*)realloc(q->e, sizeof(dpq_ent) * q->cap);
  }
  /* upper bound on key: new entry goes after all entries with key <= d (FIFO inside a key) */
  size_t lo = 0, hi = q->n;
  int64_t k = f64_total_key(d);
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    if (f64_total_key(q->e[mid].d) <= k)
      lo = mid + 1;
    else
      hi = mid;
  }
  memmove(q->e + lo + 1, q->e + lo, sizeof(dpq_ent) * (q->n - lo));
  q->e[lo].d = d;
  q->e[lo].id = id;
  q->e[lo].seq = q->next_seq++;
  q->n++;
}
static int dpq_pop_first(orc_dpq* q, double* d, uint64_t* id) {
  if (!q->n) return 0;
  *d = q->e[0].d;
  *id = q->e[0].id;
  memmove(q->e, q->e + 1, sizeof(dpq_ent) * (q->n - 1));
  q->n--;
  return 1;
}
static int dpq_pop_last(orc_dpq* q, double* d, uint64_t* id) {
  if (!q->n) return 0;
  q->n--;
  *d = q->e[q->n].d;
  *id = q->e[q->n].id;
  return 1;
}
orc_dpq* orc_dpq_new(void) {
  orc_dpq* q = (orc_dpq*)malloc(sizeof(orc_dpq));
  dpq_init(q);
  return q;
}
void orc_dpq_free(orc_dpq* q) {
  if (!q) return;
  dpq_destroy(q);
  free(q);
}
void orc_dpq_push(orc_dpq* q, double d, uint64_t id) { dpq_push(q, d, id); }
size_t orc_dpq_len(const orc_dpq* q) { return q->n; }
int orc_dpq_pop_first(orc_dpq* q, double* d, uint64_t* id) { return dpq_pop_first(q, d, id); }
int orc_dpq_pop_last(orc_dpq* q, double* d, uint64_t* id) { return dpq_pop_last(q, d, id); }
int orc_dpq_peek_first(const orc_dpq* q, double* d, uint64_t* id) {
  if (!q->n) return 0;
  *d = q->e[0].d;
  *id = q->e[0].id;
  return 1;
}
int orc_dpq_peek_last_dist(const orc_dpq* q, double* d) {
  if (!q->n) return 0;
  *d = q->e[q->n - 1].d;
  return 1;
}

/* ------------------------------------------------------------------------------------------
 * A6: KnnResultBuilder                                          idx/trees/knn.rs:363-437
 * BTreeSet<(FloatKey, VectorId::DocId)>: sorted unique array by (total_cmp dist, doc).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double d;
  uint64_t doc;
} krb_ent;
struct orc_krb {
  size_t knn, n, cap;
  krb_ent* e;
};
orc_krb* orc_krb_new(size_t knn) {
  orc_krb* b = (orc_krb*)malloc(sizeof(orc_krb));
  b->knn = knn;
  b->n = 0;
  b->cap = knn + 2;
  b->e = (krb_ent*)malloc(sizeof(krb_ent) * b->cap);
  return b;
}
void orc_krb_free(orc_krb* b) {
  if (!b) return;
  free(b->e);
  free(b);
}
int orc_krb_check_add(const orc_krb* b, double dist) { /* knn.rs:386-394: plain f64 `>` */
  if (b->n >= b->knn && b->n > 0 && dist > b->e[b->n - 1].d) return 0;
  return 1;
}
static void krb_add_one(orc_krb* b, double d, uint64_t doc) { /* knn.rs:409-431 */
  size_t pos = 0;
  while (pos < b->n) {
    int c = f64_total_cmp(b->e[pos].d, d);
    if (c > 0 || (c == 0 && b->e[pos].doc >= doc)) break;
    pos++;
  }
  if (pos < b->n && f64_total_cmp(b->e[pos].d, d) == 0 && b->e[pos].doc == doc) return; /* set: dup */
  if (b->n == b->cap) {
    b->cap *= 2;
    b->e = (krb_ent*)realloc(b->e, sizeof(krb_ent) * b->cap);
  }
  memmove(b->e + pos + 1, b->e + pos, sizeof(krb_ent) * (b->n - pos));
  b->e[pos].d = d;
  b->e[pos].doc = doc;
  b->n++;
  if (b->n > b->knn) b->n--; /* pop_last */
}
void orc_krb_add(orc_krb* b, double dist, const uint64_t* docs, size_t n_docs) {
  for (size_t i = 0; i < n_docs; i++) krb_add_one(b, dist, docs[i]);
}
size_t orc_krb_collect(const orc_krb* b, double* out_dist, uint64_t* out_doc) {
  for (size_t i = 0; i < b->n; i++) {
    out_dist[i] = b->e[i].d;
    out_doc[i] = b->e[i].doc;
  }
  return b->n;
}
size_t orc_vec_knn_f32(const float* corpus, size_t n, size_t dim, int metric, const float* q, size_t k,
                       uint64_t* out_ids, double* out_dist) { /* hnsw/mod.rs:1186-1197 */
  orc_krb* b = orc_krb_new(k);
  for (size_t r = 0; r < n; r++) {
    double d = orc_vec_distance_f32(metric, corpus + r * dim, q, dim);
    if (orc_krb_check_add(b, d)) {
      uint64_t doc = r;
      orc_krb_add(b, d, &doc, 1);
    }
  }
  size_t c = orc_krb_collect(b, out_dist, out_ids);
  orc_krb_free(b);
  return c;
}

/* ------------------------------------------------------------------------------------------
 * a7-a9: HNSW                                       idx/trees/hnsw/{mod,layer,heuristic}.rs
 * Edge sets keep INSERTION order (the reference's ArraySet does; its AHashSet iterates in an
 * irreproducible hash order -- SURVEY appendix A5; ties have measure zero on continuous data).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t* v;
  uint32_t n, cap;
  uint8_t present;
} eset;
typedef struct {
  eset* nodes; /* indexed by element id */
  size_t n_alloc;
  size_t m_max;
} hlayer;
struct orc_hnsw {
  size_t dim, m, m0, efc;
  int metric, heur;
  double ml;
  uint64_t rng[4];
  float* vec;
  size_t n, cap;
  int64_t enter_point;
  hlayer* layers; /* layers[0] = layer0, layers[i] = upper layer i */
  size_t n_layers;
  uint64_t cnt_visited, cnt_expanded;
  struct {
    uint32_t* stamp;
    size_t n;
    uint32_t cur;
  } vis; /* scratch visited set (same layout as vset); makes a handle single-threaded */
};
static void eset_insert(eset* s, uint32_t v) {
  for (uint32_t i = 0; i < s->n; i++)
    if (s->v[i] == v) return;
  if (s->n == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 8;
    s->v = (uint32_t*)realloc(s->v, sizeof(uint32_t) * s->cap);
  }
  s->v[s->n++] = v;
}
static int eset_contains(const eset* s, uint32_t v) {
  for (uint32_t i = 0; i < s->n; i++)
    if (s->v[i] == v) return 1;
  return 0;
}
static void layer_ensure(hlayer* L, size_t id) {
  if (id < L->n_alloc) return;
  size_t na = L->n_alloc ? L->n_alloc : 64;
  while (na <= id) na *= 2;
  L->nodes = (eset*)realloc(L->nodes, sizeof(eset) * na);
  memset(L->nodes + L->n_alloc, 0, sizeof(eset) * (na - L->n_alloc));
  L->n_alloc = na;
}
static eset* layer_get(const hlayer* L, size_t id) {
  if (id >= L->n_alloc || !L->nodes[id].present) return NULL;
  return &L->nodes[id];
}
static int layer_add_empty(hlayer* L, size_t id) { /* graph.rs:43-50 */
  layer_ensure(L, id);
  if (L->nodes[id].present) return 0;
  L->nodes[id].present = 1;
  return 1;
}
/* xoshiro256++ (rand 0.8 SmallRng on 64-bit) seeded by SplitMix64 (seed_from_u64).  The reference
 * seeds from thread_rng (hnsw/mod.rs:176) so graphs are irreproducible anyway. */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(uint64_t s[4]) {
  uint64_t r = rotl64(s[0] + s[3], 23) + s[0];
  uint64_t t = s[1] << 17;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl64(s[3], 45);
  return r;
}
static void rng_seed(uint64_t s[4], uint64_t seed) {
  for (int i = 0; i < 4; i++) {
    seed += 0x9E3779B97F4A7C15ULL;
    uint64_t z = seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    s[i] = z ^ (z >> 31);
  }
}
orc_hnsw* orc_hnsw_new(size_t dim, int metric, size_t m, size_t m0, size_t efc, double ml, int heur,
                       uint64_t seed) {
  orc_hnsw* h = (orc_hnsw*)calloc(1, sizeof(orc_hnsw));
  h->dim = dim;
  h->metric = metric;
  h->m = m;
  h->m0 = m0;
  h->efc = efc;
  h->ml = ml;
  h->heur = heur;
  rng_seed(h->rng, seed);
  h->enter_point = -1;
  h->layers = (hlayer*)calloc(1, sizeof(hlayer));
  h->layers[0].m_max = m0;
  h->n_layers = 1;
  return h;
}
void orc_hnsw_free(orc_hnsw* h) {
  if (!h) return;
  for (size_t l = 0; l < h->n_layers; l++) {
    for (size_t i = 0; i < h->layers[l].n_alloc; i++) free(h->layers[l].nodes[i].v);
    free(h->layers[l].nodes);
  }
  free(h->vis.stamp);
  free(h->layers);
  free(h->vec);
  free(h);
}
size_t orc_hnsw_len(const orc_hnsw* h) { return h->n; }
const float* orc_hnsw_vectors(const orc_hnsw* h) { return h->vec; }
static inline double h_dist(const orc_hnsw* h, const float* a, const float* b) {
  return orc_vec_distance_f32(h->metric, a, b, h->dim);
}
static inline const float* h_vec(const orc_hnsw* h, uint64_t id) { return h->vec + id * h->dim; }

/* visited set: stamp array (ahash HashSet<u64> in the reference; membership semantics only) */
typedef struct {
  uint32_t* stamp;
  size_t n;
  uint32_t cur;
} vset;
static void vset_reset(vset* v, size_t n) { /* O(1) amortised clear; grows on demand */
  if (v->n < n || !v->stamp) {
    size_t nn = v->n ? v->n : 1024;
    while (nn < n) nn *= 2;
    free(v->stamp);
    v->stamp = (uint32_t*)calloc(nn, sizeof(uint32_t));
    v->n = nn;
    v->cur = 0;
  }
  if (++v->cur == 0) {
    memset(v->stamp, 0, sizeof(uint32_t) * v->n);
    v->cur = 1;
  }
}
static inline int vset_insert(vset* v, uint64_t id) {
  if (v->stamp[id] == v->cur) return 0;
  v->stamp[id] = v->cur;
  return 1;
}

/* generic adjacency accessor so the same walk serves the builder's layers and imported CSR */
typedef struct {
  const hlayer* L;
  const uint64_t* row_ptr;
  const uint32_t* col_idx;
} adj_src;
static inline size_t adj_get(const adj_src* a, uint64_t id, const uint32_t** out) {
  if (a->L) {
    const eset* s = layer_get(a->L, id);
    if (!s) {
      *out = NULL;
      return (size_t)-1;
    }
    *out = s->v;
    return s->n;
  }
  *out = a->col_idx + a->row_ptr[id];
  return (size_t)(a->row_ptr[id + 1] - a->row_ptr[id]);
}

/* HnswLayer::search  layer.rs:184-223 (pending_docs = None => are_all_docs_in_pending is false) */
/* pending_docs of HnswLayer::search (layer.rs:184-223) evaluated ahead of time per element: all_docs_pending[e] != 0
 * iff every document of element e is in the pending bitmap (are_all_docs_in_pending, layer.rs:320-339).  Set only for
 * the duration of orc_hnsw_search_csr_pending (the oracle is single-threaded test infrastructure). */
static const uint8_t* g_all_docs_pending = NULL;
static void layer_search(const float* vectors, size_t dim, int metric, const adj_src* adj, const float* q,
                         orc_dpq* candidates, vset* visited, orc_dpq* w, size_t ef, uint64_t* counters) {
  double fq = 1.7976931348623157e308; /* f64::MAX */
  if (w->n) fq = w->e[w->n - 1].d;
  double cd;
  uint64_t c;
  while (dpq_pop_first(candidates, &cd, &c)) {
    if (cd > fq) break;
    const uint32_t* nb;
    size_t deg = adj_get(adj, c, &nb);
    if (deg == (size_t)-1) continue;
    if (counters) counters[1]++;
    for (size_t i = 0; i < deg; i++) {
      uint64_t e = nb[i];
      if (!vset_insert(visited, e)) continue;
      double ed = orc_vec_distance_f32(metric, vectors + e * dim, q, dim);
      if (counters) counters[0]++;
      if (ed < fq || w->n < ef) {
        /* layer.rs:209: an element whose documents ALL have pending updates still enters w, but is not expanded */
        if (!g_all_docs_pending || !g_all_docs_pending[e]) dpq_push(candidates, ed, e);
        dpq_push(w, ed, e);
        if (w->n > ef) {
          double dd;
          uint64_t ii;
          dpq_pop_last(w, &dd, &ii);
        }
        fq = w->n ? w->e[w->n - 1].d : 1.7976931348623157e308;
      }
    }
  }
}
/* HnswLayer::search_with_filter + add_if_truthy  layer.rs:226-306 (pending_docs = None).  `truthy[e]` restates
 * HnswTruthyDocumentFilter::check_any_doc_truthy (hnsw/filter.rs:52-136) evaluated ahead of time per element. */
static void layer_search_with_filter(const float* vectors, size_t dim, int metric, const adj_src* adj, const float* q,
                                     orc_dpq* candidates, vset* visited, orc_dpq* w, size_t ef,
                                     const uint8_t* truthy, uint64_t* counters) {
  double fq = 1.7976931348623157e308;
  if (w->n) fq = w->e[w->n - 1].d;
  double cd;
  uint64_t c;
  while (dpq_pop_first(candidates, &cd, &c)) {
    if (cd > fq) break;
    const uint32_t* nb;
    size_t deg = adj_get(adj, c, &nb);
    if (deg == (size_t)-1) continue;
    if (counters) counters[1]++;
    for (size_t i = 0; i < deg; i++) {
      uint64_t e = nb[i];
      if (!vset_insert(visited, e)) continue;
      double ed = orc_vec_distance_f32(metric, vectors + e * dim, q, dim);
      if (counters) counters[0]++;
      if (ed < fq || w->n < ef) {
        dpq_push(candidates, ed, e);
        if (truthy[e]) { /* add_if_truthy */
          dpq_push(w, ed, e);
          if (w->n > ef) {
            double dd;
            uint64_t ii;
            dpq_pop_last(w, &dd, &ii);
          }
          fq = w->e[w->n - 1].d;
        }
      }
    }
  }
}
/* search_single_with_filter  layer.rs:111-149 */
static void layer_search_single_with_filter(const float* vectors, size_t n, size_t dim, int metric, const adj_src* adj,
                                            const float* q, double ep_dist, uint64_t ep, size_t ef,
                                            const uint8_t* truthy, orc_dpq* w_out, uint64_t* counters, vset* visp) {
  vset_reset(visp, n);
  vset_insert(visp, ep);
  orc_dpq cand;
  dpq_init(&cand);
  dpq_push(&cand, ep_dist, ep);
  dpq_init(w_out);
  if (truthy[ep]) dpq_push(w_out, ep_dist, ep);
  layer_search_with_filter(vectors, dim, metric, adj, q, &cand, visp, w_out, ef, truthy, counters);
  dpq_destroy(&cand);
}
/* search_single  layer.rs:76-90 : returns w (caller destroys) */
static void layer_search_single(const float* vectors, size_t n, size_t dim, int metric, const adj_src* adj,
                                const float* q, double ep_dist, uint64_t ep, size_t ef, orc_dpq* w_out,
                                uint64_t* counters, vset* visp) {
  vset_reset(visp, n);
#define vis (*visp)
  vset_insert(&vis, ep);
  orc_dpq cand;
  dpq_init(&cand);
  dpq_push(&cand, ep_dist, ep);
  dpq_copy(w_out, &cand);
  layer_search(vectors, dim, metric, adj, q, &cand, &vis, w_out, ef, counters);
  dpq_destroy(&cand);
#undef vis
}
/* search_multi  layer.rs:151-162 */
static void layer_search_multi(const orc_hnsw* h, const hlayer* L, const float* q, orc_dpq* cand /*consumed*/,
                               size_t ef, orc_dpq* w_out) {
  dpq_copy(w_out, cand);
  vset* vis = (vset*)&h->vis;
  vset_reset(vis, h->n + 1);
  for (size_t i = 0; i < w_out->n; i++) vset_insert(vis, w_out->e[i].id);
  adj_src a = {L, NULL, NULL};
  layer_search(h->vec, h->dim, h->metric, &a, q, cand, vis, w_out, ef, NULL);
}

/* Heuristic::is_closer  heuristic.rs:201-216 */
static int heur_is_closer(const orc_hnsw* h, double e_dist, uint64_t e_id, eset* r) {
  const float* cv = h_vec(h, e_id);
  for (uint32_t i = 0; i < r->n; i++) {
    double rd = h_dist(h, h_vec(h, r->v[i]), cv); /* get_distance(current_vec, r_id) = calculate(stored r, cur) */
    if (e_dist > rd) return 0;
  }
  eset_insert(r, (uint32_t)e_id);
  return 1;
}
/* extend_candidates  heuristic.rs:112-152 */
static void heur_extend(const orc_hnsw* h, const hlayer* L, uint64_t q_id, const float* q_pt, orc_dpq* c,
                        int64_t ignore) {
  size_t n0 = c->n;
  vset ex;
  memset(&ex, 0, sizeof(ex));
  vset_reset(&ex, h->n + 1);
  for (size_t i = 0; i < n0; i++) vset_insert(&ex, c->e[i].id);
  if (ignore >= 0) vset_insert(&ex, (uint64_t)ignore);
  dpq_ent* ext = NULL;
  size_t n_ext = 0, cap_ext = 0;
  for (size_t i = 0; i < n0; i++) { /* c.to_vec(): ascending */
    const eset* conn = layer_get(L, c->e[i].id);
    if (!conn) continue;
    for (uint32_t j = 0; j < conn->n; j++) {
      uint64_t adj = conn->v[j];
      if (adj != q_id && vset_insert(&ex, adj)) {
        if (n_ext == cap_ext) {
          cap_ext = cap_ext ? cap_ext * 2 : 32;
          ext = (dpq_ent*)realloc(ext, sizeof(dpq_ent) * cap_ext);
        }
        ext[n_ext].d = h_dist(h, h_vec(h, adj), q_pt);
        ext[n_ext].id = adj;
        n_ext++;
      }
    }
  }
  for (size_t i = 0; i < n_ext; i++) dpq_push(c, ext[i].d, ext[i].id);
  free(ext);
  free(ex.stamp);
}
/* Heuristic::select  heuristic.rs:36-58 ; c is consumed */
static void heur_select(const orc_hnsw* h, const hlayer* L, uint64_t q_id, const float* q_pt, orc_dpq* c,
                        int64_t ignore, eset* res) {
  size_t m_max = L->m_max;
  if (h->heur & 1) heur_extend(h, L, q_id, q_pt, c, ignore);
  int keep = (h->heur & 2) != 0;
  if (c->n <= m_max) { /* c.to_dynamic_set(res) */
    for (size_t i = 0; i < c->n; i++) eset_insert(res, (uint32_t)c->e[i].id);
    return;
  }
  uint32_t* pruned = keep ? (uint32_t*)malloc(sizeof(uint32_t) * c->n) : NULL;
  size_t n_pruned = 0;
  double ed;
  uint64_t eid;
  while (dpq_pop_first(c, &ed, &eid)) {
    if (heur_is_closer(h, ed, eid, res)) {
      if (res->n == m_max) break;
    } else if (keep) {
      pruned[n_pruned++] = (uint32_t)eid;
    }
  }
  if (keep) { /* heuristic_keep  heuristic.rs:83-110: pruned.drain(0..n) */
    size_t nfill = m_max - res->n;
    for (size_t i = 0; i < nfill && i < n_pruned; i++) eset_insert(res, pruned[i]);
    free(pruned);
  }
}
/* HnswLayer::insert  layer.rs:342-387 ; eps consumed, returns new eps in *eps */
static void layer_insert(orc_hnsw* h, hlayer* L, uint64_t q_id, const float* q_pt, orc_dpq* eps) {
  orc_dpq w, wsel;
  layer_search_multi(h, L, q_pt, eps, h->efc, &w);
  dpq_destroy(eps);
  dpq_copy(eps, &w); /* eps = w.clone() */
  dpq_copy(&wsel, &w);
  dpq_destroy(&w);
  eset neighbors;
  memset(&neighbors, 0, sizeof(neighbors));
  heur_select(h, L, q_id, q_pt, &wsel, -1, &neighbors);
  dpq_destroy(&wsel);
  /* add_node_and_bidirectional_edges  graph.rs:52-64 */
  layer_ensure(L, q_id);
  for (uint32_t i = 0; i < neighbors.n; i++) {
    uint32_t e = neighbors.v[i];
    layer_ensure(L, e);
    L->nodes[e].present = 1;
    eset_insert(&L->nodes[e], (uint32_t)q_id);
  }
  free(L->nodes[q_id].v);
  L->nodes[q_id].v = NULL;
  L->nodes[q_id].n = L->nodes[q_id].cap = 0;
  L->nodes[q_id].present = 1;
  for (uint32_t i = 0; i < neighbors.n; i++) eset_insert(&L->nodes[q_id], neighbors.v[i]);
  /* shrink over-full neighbours  layer.rs:362-378 */
  for (uint32_t i = 0; i < neighbors.n; i++) {
    uint32_t e_id = neighbors.v[i];
    eset* conn = &L->nodes[e_id];
    if (conn->n > L->m_max) {
      const float* e_pt = h_vec(h, e_id);
      orc_dpq ec; /* build_priority_list  layer.rs:389-405 */
      dpq_init(&ec);
      for (uint32_t j = 0; j < conn->n; j++) dpq_push(&ec, h_dist(h, e_pt, h_vec(h, conn->v[j])), conn->v[j]);
      eset nc;
      memset(&nc, 0, sizeof(nc));
      heur_select(h, L, e_id, e_pt, &ec, -1, &nc);
      dpq_destroy(&ec);
      free(conn->v);
      conn->v = nc.v;
      conn->n = nc.n;
      conn->cap = nc.cap;
    }
  }
  free(neighbors.v);
}
static size_t h_random_level(orc_hnsw* h) { /* hnsw/mod.rs:263-266 */
  double unif = (double)(rng_next(h->rng) >> 11) * (1.0 / 9007199254740992.0);
  double lv = floor(-log(unif) * h->ml);
  if (!(lv < 64.0)) lv = 64.0; /* unif == 0 saturates in Rust; clamp so the oracle stays finite */
  return (size_t)lv;
}
uint64_t orc_hnsw_insert_level(orc_hnsw* h, const float* v, size_t q_level) { /* hnsw/mod.rs:230-260 */
  uint64_t q_id = h->n;
  size_t top_up_layers = h->n_layers - 1;
  if (q_level > top_up_layers) {
    h->layers = (hlayer*)realloc(h->layers, sizeof(hlayer) * (q_level + 1));
    for (size_t i = top_up_layers + 1; i <= q_level; i++) {
      memset(&h->layers[i], 0, sizeof(hlayer));
      h->layers[i].m_max = h->m;
    }
    h->n_layers = q_level + 1;
  }
  if (h->n == h->cap) {
    h->cap = h->cap ? h->cap * 2 : 1024;
    h->vec = (float*)realloc(h->vec, sizeof(float) * h->cap * h->dim);
  }
  memcpy(h->vec + q_id * h->dim, v, sizeof(float) * h->dim);
  h->n++;
  const float* q_pt = h_vec(h, q_id);
  if (h->enter_point < 0) { /* insert_first_element  hnsw/mod.rs:268-289 */
    for (size_t l = 1; l <= q_level; l++) layer_add_empty(&h->layers[l], q_id);
    layer_add_empty(&h->layers[0], q_id);
    h->enter_point = (int64_t)q_id;
    return q_id;
  }
  /* insert_element  hnsw/mod.rs:297-377 */
  uint64_t ep_id = (uint64_t)h->enter_point;
  double ep_dist = h_dist(h, h_vec(h, ep_id), q_pt);
  if (q_level < top_up_layers) {
    for (size_t l = top_up_layers; l > q_level; l--) { /* layers[q_level..top_up_layers].rev() */
      adj_src a = {&h->layers[l], NULL, NULL};
      orc_dpq w;
      layer_search_single(h->vec, h->n, h->dim, h->metric, &a, q_pt, ep_dist, ep_id, 1, &w, NULL, (vset*)&h->vis);
      if (w.n) {
        ep_dist = w.e[0].d;
        ep_id = w.e[0].id;
      }
      dpq_destroy(&w);
    }
  }
  orc_dpq eps;
  dpq_init(&eps);
  dpq_push(&eps, ep_dist, ep_id);
  size_t insert_to_up = q_level < top_up_layers ? q_level : top_up_layers;
  for (size_t l = insert_to_up; l >= 1; l--) layer_insert(h, &h->layers[l], q_id, q_pt, &eps);
  layer_insert(h, &h->layers[0], q_id, q_pt, &eps);
  dpq_destroy(&eps);
  for (size_t l = top_up_layers + 1; l <= q_level; l++) layer_add_empty(&h->layers[l], q_id);
  if (q_level > top_up_layers) h->enter_point = (int64_t)q_id;
  return q_id;
}
uint64_t orc_hnsw_insert(orc_hnsw* h, const float* v) { return orc_hnsw_insert_level(h, v, h_random_level(h)); }

static size_t hnsw_search_csr_vs(const float*, size_t, size_t, int, size_t, const uint64_t* const*,
                                 const uint32_t* const*, int64_t, const float*, size_t, size_t, uint64_t*,
                                 double*, uint64_t*, vset*);
/* Hnsw::knn_search_with_filter  hnsw/mod.rs:488-515 over an imported graph; truthy: one byte per element */
size_t orc_hnsw_search_csr_filtered(const float* vectors, size_t n, size_t dim, int metric, size_t n_layers,
                                    const uint64_t* const* row_ptr, const uint32_t* const* col_idx,
                                    int64_t entry_point, const float* q, size_t k, size_t ef, const uint8_t* truthy,
                                    uint64_t* out_ids, double* out_dist, uint64_t* counters) {
  if (entry_point < 0 || k == 0) return 0;
  vset lv;
  memset(&lv, 0, sizeof(lv));
  uint64_t ep = (uint64_t)entry_point;
  double ep_dist = orc_vec_distance_f32(metric, vectors + ep * dim, q, dim);
  if (counters) counters[0]++;
  for (size_t l = n_layers - 1; l >= 1; l--) { /* search_ep is NOT filtered  hnsw/mod.rs:521-548 */
    adj_src a = {NULL, row_ptr[l], col_idx[l]};
    orc_dpq w;
    layer_search_single(vectors, n, dim, metric, &a, q, ep_dist, ep, 1, &w, counters, &lv);
    if (w.n) {
      ep_dist = w.e[0].d;
      ep = w.e[0].id;
    }
    dpq_destroy(&w);
  }
  adj_src a0 = {NULL, row_ptr[0], col_idx[0]};
  orc_dpq w;
  layer_search_single_with_filter(vectors, n, dim, metric, &a0, q, ep_dist, ep, ef, truthy, &w, counters, &lv);
  size_t c = w.n < k ? w.n : k;
  for (size_t i = 0; i < c; i++) {
    out_ids[i] = w.e[i].id;
    out_dist[i] = w.e[i].d;
  }
  dpq_destroy(&w);
  free(lv.stamp);
  return c;
}
size_t orc_hnsw_search_csr(const float* vectors, size_t n, size_t dim, int metric, size_t n_layers,
                           const uint64_t* const* row_ptr, const uint32_t* const* col_idx, int64_t entry_point,
                           const float* q, size_t k, size_t ef, uint64_t* out_ids, double* out_dist,
                           uint64_t* counters) {
  if (entry_point < 0 || k == 0) return 0;
  vset lv;
  memset(&lv, 0, sizeof(lv));
  size_t r = hnsw_search_csr_vs(vectors, n, dim, metric, n_layers, row_ptr, col_idx, entry_point, q, k, ef,
                                out_ids, out_dist, counters, &lv);
  free(lv.stamp);
  return r;
}
/* Hnsw::knn_search with pending_docs = Some(..)  hnsw/mod.rs:459-482 (search_ep passes the bitmap down too) */
size_t orc_hnsw_search_csr_pending(const float* vectors, size_t n, size_t dim, int metric, size_t n_layers,
                                   const uint64_t* const* row_ptr, const uint32_t* const* col_idx, int64_t entry_point,
                                   const float* q, size_t k, size_t ef, const uint8_t* all_docs_pending,
                                   uint64_t* out_ids, double* out_dist, uint64_t* counters) {
  g_all_docs_pending = all_docs_pending;
  size_t r = orc_hnsw_search_csr(vectors, n, dim, metric, n_layers, row_ptr, col_idx, entry_point, q, k, ef, out_ids,
                                 out_dist, counters);
  g_all_docs_pending = NULL;
  return r;
}
static size_t hnsw_search_csr_vs(const float* vectors, size_t n, size_t dim, int metric, size_t n_layers,
                                 const uint64_t* const* row_ptr, const uint32_t* const* col_idx,
                                 int64_t entry_point, const float* q, size_t k, size_t ef, uint64_t* out_ids,
                                 double* out_dist, uint64_t* counters, vset* vs) {
  uint64_t ep = (uint64_t)entry_point;
  double ep_dist = orc_vec_distance_f32(metric, vectors + ep * dim, q, dim);
  if (counters) counters[0]++;
  for (size_t l = n_layers - 1; l >= 1; l--) { /* search_ep  hnsw/mod.rs:521-548 */
    adj_src a = {NULL, row_ptr[l], col_idx[l]};
    orc_dpq w;
    layer_search_single(vectors, n, dim, metric, &a, q, ep_dist, ep, 1, &w, counters, vs);
    if (w.n) {
      ep_dist = w.e[0].d;
      ep = w.e[0].id;
    }
    dpq_destroy(&w);
  }
  adj_src a0 = {NULL, row_ptr[0], col_idx[0]};
  orc_dpq w;
  layer_search_single(vectors, n, dim, metric, &a0, q, ep_dist, ep, ef, &w, counters, vs);
  size_t c = w.n < k ? w.n : k; /* to_vec_limit(k) */
  for (size_t i = 0; i < c; i++) {
    out_ids[i] = w.e[i].id;
    out_dist[i] = w.e[i].d;
  }
  dpq_destroy(&w);
  return c;
}
size_t orc_hnsw_search(const orc_hnsw* hc, const float* q, size_t k, size_t ef, uint64_t* out_ids,
                       double* out_dist) { /* Hnsw::knn_search  hnsw/mod.rs:459-482 */
  orc_hnsw* h = (orc_hnsw*)hc;
  uint64_t counters[2] = {0, 0};
  h->cnt_visited = h->cnt_expanded = 0;
  if (h->enter_point < 0 || k == 0) return 0;
  uint64_t ep = (uint64_t)h->enter_point;
  double ep_dist = h_dist(h, h_vec(h, ep), q);
  counters[0]++;
  for (size_t l = h->n_layers - 1; l >= 1; l--) {
    adj_src a = {&h->layers[l], NULL, NULL};
    orc_dpq w;
    layer_search_single(h->vec, h->n, h->dim, h->metric, &a, q, ep_dist, ep, 1, &w, counters, (vset*)&h->vis);
    if (w.n) {
      ep_dist = w.e[0].d;
      ep = w.e[0].id;
    }
    dpq_destroy(&w);
  }
  adj_src a0 = {&h->layers[0], NULL, NULL};
  orc_dpq w;
  layer_search_single(h->vec, h->n, h->dim, h->metric, &a0, q, ep_dist, ep, ef, &w, counters, (vset*)&h->vis);
  size_t c = w.n < k ? w.n : k;
  for (size_t i = 0; i < c; i++) {
    out_ids[i] = w.e[i].id;
    out_dist[i] = w.e[i].d;
  }
  dpq_destroy(&w);
  h->cnt_visited = counters[0];
  h->cnt_expanded = counters[1];
  return c;
}
void orc_hnsw_last_counters(const orc_hnsw* h, uint64_t* visited, uint64_t* expanded) {
  *visited = h->cnt_visited;
  *expanded = h->cnt_expanded;
}
int orc_hnsw_check_props(const orc_hnsw* h) { /* check_hnsw_props  layer.rs:571-587 */
  for (size_t l = 0; l < h->n_layers; l++) {
    const hlayer* L = &h->layers[l];
    for (size_t i = 0; i < L->n_alloc; i++) {
      const eset* s = &L->nodes[i];
      if (!s->present) continue;
      if (i >= h->n) return 0;
      if (s->n > L->m_max) return 0;
      if (eset_contains(s, (uint32_t)i)) return 0;
      for (uint32_t j = 0; j < s->n; j++)
        if (s->v[j] >= h->n) return 0;
    }
  }
  for (size_t i = 0; i < h->n; i++) /* every element is in layer 0 */
    if (!layer_get(&h->layers[0], i)) return 0;
  return 1;
}
size_t orc_hnsw_n_layers(const orc_hnsw* h) { return h->n_layers; }
int64_t orc_hnsw_entry_point(const orc_hnsw* h) { return h->enter_point; }
size_t orc_hnsw_layer_edges(const orc_hnsw* h, size_t layer) {
  size_t e = 0;
  const hlayer* L = &h->layers[layer];
  for (size_t i = 0; i < L->n_alloc && i < h->n; i++)
    if (L->nodes[i].present) e += L->nodes[i].n;
  return e;
}
void orc_hnsw_export_layer(const orc_hnsw* h, size_t layer, uint64_t* row_ptr, uint32_t* col_idx,
                           uint8_t* present) {
  const hlayer* L = &h->layers[layer];
  uint64_t e = 0;
  for (size_t i = 0; i < h->n; i++) {
    row_ptr[i] = e;
    const eset* s = (i < L->n_alloc && L->nodes[i].present) ? &L->nodes[i] : NULL;
    if (present) present[i] = s != NULL;
    if (s) {
      memcpy(col_idx + e, s->v, sizeof(uint32_t) * s->n);
      e += s->n;
    }
  }
  row_ptr[h->n] = e;
}

/* ------------------------------------------------------------------------------------------
 * a11-a13: graph expansion                    exec/operators/scan/graph.rs:203-279,
 *                                             exec/parts/lookup.rs:144-169  
 * CSR rows hold, per source vertex, its `->edge->node` targets in the KV key order of the
 * connecting edge records (two fused GraphEdgeScans collapse to one CSR row when each edge has
 * exactly one target).  Multiset semantics: duplicates kept, frontier order preserved.
 * ---------------------------------------------------------------------------------------- */
uint64_t orc_graph_hop(const uint64_t* row_ptr, const uint32_t* col_idx, const uint32_t* frontier,
                       uint64_t n_frontier, uint32_t limit, uint32_t* out) {
  uint64_t w = 0;
  for (uint64_t i = 0; i < n_frontier; i++) {
    uint64_t b = row_ptr[frontier[i]], e = row_ptr[frontier[i] + 1];
    if (limit && e - b > limit) e = b + limit; /* GraphEdgeScan.limit  graph.rs:83,238,261 */
    for (uint64_t j = b; j < e; j++) {
      if (out) out[w] = col_idx[j];
      w++;
    }
  }
  return w;
}
uint64_t orc_graph_collect(const uint64_t* row_ptr, const uint32_t* col_idx, uint64_t n_nodes,
                           const uint32_t* start, uint64_t n_start, uint32_t min_depth, uint32_t max_depth,
                           int inclusive, uint32_t* out, uint64_t out_cap) {
  /* recursion/collect.rs:74-143: seen = hashes of start values; per level expand every frontier
   * node (frontier order), dedup first-seen against `seen`, emit when depth+1 >= min_depth.
   * NOTE the start value is NOT in `seen` unless `inclusive`, so a cycle re-emits it. */
  uint8_t* seen = (uint8_t*)calloc(n_nodes ? n_nodes : 1, 1);
  uint32_t* frontier = (uint32_t*)malloc(sizeof(uint32_t) * (n_start ? n_start : 1));
  uint64_t nf = 0, n_out = 0;
  for (uint64_t i = 0; i < n_start; i++) { /* the reference starts from ONE value; >1 = that many roots */
    frontier[nf++] = start[i];
    if (inclusive) { /* collect.rs:83-86: start is only marked seen when inclusive */
      if (n_out < out_cap) out[n_out++] = start[i];
      seen[start[i]] = 1;
    }
  }
  uint32_t depth = 0;
  while (nf && (max_depth == 0 || depth < max_depth)) {
    uint64_t cnt = orc_graph_hop(row_ptr, col_idx, frontier, nf, 0, NULL);
    uint32_t* next = (uint32_t*)malloc(sizeof(uint32_t) * (cnt ? cnt : 1));
    uint64_t nn = 0;
    for (uint64_t i = 0; i < nf; i++)
      for (uint64_t j = row_ptr[frontier[i]]; j < row_ptr[frontier[i] + 1]; j++) {
        uint32_t t = col_idx[j];
        if (seen[t]) continue;
        seen[t] = 1;
        next[nn++] = t;
        if (depth + 1 >= min_depth && n_out < out_cap) out[n_out++] = t;
      }
    free(frontier);
    frontier = next;
    nf = nn;
    depth++;
  }
  free(frontier);
  free(seen);
  return n_out;
}

/* ------------------------------------------------------------------------------------------
 * synthetic data generator (bench/test inputs; mirrored bit-for-bit by csrc/gen.cuh)
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
float orc_gen_f32(uint64_t seed, uint64_t index) {
  uint64_t h = mix64(seed * 0x9E3779B97F4A7C15ULL + index + 0x632BE59BD9B4E019ULL);
  uint32_t m = (uint32_t)(h >> 40); /* 24 random bits */
  return (float)m * (1.0f / 8388608.0f) - 1.0f; /* exact: [-1, 1) on a 2^-23 grid */
}
void orc_gen_fill_f32(uint64_t seed, uint64_t first, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; i++) out[i] = orc_gen_f32(seed, first + i);
}
