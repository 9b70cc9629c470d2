/*
 * sdb_oracle.h -- CPU ORACLE for the SurrealDB KNN / graph-scan hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * path (surrealdb_b200/, include/sdbgpu.h) never links or calls anything here.
 *
 * It is a plain-C restatement of the reference's algorithms (the reference is Rust and
 * cannot be compiled in this image: no cargo/rustc).  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/surrealdb/core/src).
 *
 * Parity pinning: the arithmetic below is checked against every known-answer vector the
 * reference's own tests hold for this path (tests/test_oracle_*.py cite them).  ONE piece is
 * "parity unpinned": the f32 8-lane summation order of ndarray 0.17.2 (`orc_nd_dot_f32`,
 * `orc_nd_sum_f32`) -- the crate is not vendored under /root/reference and the reference's
 * tests pin f32 metrics only through order-insensitive assertions.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  -ffp-contract=off
 * matters: Rust never fuses a*b+c.
 */
#ifndef SDB_ORACLE_H
#define SDB_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- val::Number (Int | Float); Decimal is out of scope (never produced by vector literals of
 *      ints/floats).  core/val/number.rs:49-56 ---- */
typedef struct {
  int64_t tag; /* 0 = Float, 1 = Int */
  union {
    double f;
    int64_t i;
  } v;
} orc_num;

enum { ORC_OK = 0, ORC_EDIM = 1, ORC_EINVAL = 2 };

/* metric ids follow catalog::Distance (catalog/schema/index.rs:247-284) */
enum {
  ORC_CHEBYSHEV = 0,
  ORC_COSINE = 1,
  ORC_EUCLIDEAN = 2,
  ORC_HAMMING = 3,
  ORC_JACCARD = 4,
  ORC_MANHATTAN = 5,
  ORC_MINKOWSKI = 6,
  ORC_PEARSON = 7
};

/* ---- a1-a3: Vec<Number> metrics.  fnc/util/math/vector.rs ---- */
int orc_num_dot(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
int orc_num_magnitude(const orc_num* a, size_t na, orc_num* out);
int orc_num_cosine_similarity(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
int orc_num_cosine_distance(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
int orc_num_euclidean(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
int orc_num_manhattan(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
int orc_num_chebyshev(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
int orc_num_hamming(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
int orc_num_minkowski(const orc_num* a, size_t na, const orc_num* b, size_t nb, double p, orc_num* out);
int orc_num_pearson(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
int orc_num_jaccard(const orc_num* a, size_t na, const orc_num* b, size_t nb, orc_num* out);
/* Distance::compute  catalog/schema/index.rs:287-303 */
/* order of Distance::Minkowski used by the all-Float fast paths (orc_f64_metric / orc_knn_topk); default 3 */
void orc_set_minkowski_order(double p);
int orc_num_distance(int metric, double minkowski_p, const orc_num* a, size_t na, const orc_num* b,
                     size_t nb, orc_num* out);
/* Number::cmp (Int/Float only)  val/number.rs:620-680 ; returns -1/0/1 */
int orc_num_cmp(const orc_num* a, const orc_num* b);

/* All-Float fast path: identical arithmetic to the orc_num_* functions when every element is
 * Number::Float.  `row` may be f32 (promoted element-wise, exactly) or f64. */
double orc_f64_cosine_distance(const double* a, const double* b, size_t n);
double orc_f64_euclidean(const double* a, const double* b, size_t n);
double orc_f32row_cosine_distance(const float* row, const double* q, size_t n);
double orc_f32row_euclidean(const float* row, const double* q, size_t n);
double orc_f64_magnitude(const double* a, size_t n);
/* all-Float fast path of Distance::compute for COSINE, EUCLIDEAN, MANHATTAN, CHEBYSHEV, HAMMING, PEARSON */
double orc_f64_metric(int metric, const double* a, const double* b, size_t n);
double orc_f32row_magnitude(const float* a, size_t n);

/* ---- a4: KnnTopK selection.  exec/operators/knn_topk.rs:166-267
 * rows arrive in scan order 0..n_rows-1; `skip` (nullable) marks rows whose field is missing /
 * non-numeric / dimension-mismatched (they do not consume a seq number -- knn_topk.rs:199-217).
 * elem_is_f64: corpus element type (0 = f32 rows, 1 = f64 rows).  Output nearest-first,
 * ties by scan order.  Returns count written (<= k). */
size_t orc_knn_topk(const void* corpus, int elem_is_f64, size_t n_rows, size_t dim, const uint8_t* skip,
                    const double* query, int metric, size_t k, uint64_t* out_rows, double* out_dist);
/* multi-threaded batch driver over independent queries (the reference runs one task per query);
 * used by bench.py cpu_baseline only. */
void orc_knn_topk_batch(const void* corpus, int elem_is_f64, size_t n_rows, size_t dim,
                        const double* queries, size_t n_queries, int metric, size_t k,
                        uint64_t* out_rows, double* out_dist, int n_threads);

/* ---- a6: typed ndarray metrics.  idx/trees/vector.rs:235-289 ---- */
float orc_nd_dot_f32(const float* a, const float* b, size_t n);  /* ndarray unrolled_dot  (UNPINNED) */
float orc_nd_sumsq_f32(const float* a, size_t n);                /* (a*a).sum()           (UNPINNED) */
double orc_vec_cosine_f32(const float* a, const float* b, size_t n);  /* vector.rs:243-249 */
double orc_vec_l2_f32(const float* a, const float* b, size_t n);      /* vector.rs:273-289 via ndarray-stats */
double orc_vec_cosine_f64(const double* a, const double* b, size_t n); /* vector.rs:235-241 */
double orc_vec_l2_f64(const double* a, const double* b, size_t n);
double orc_vec_distance_f32(int metric, const float* a, const float* b, size_t n);

/* ---- A4: DoublePriorityQueue  idx/trees/knn.rs:15-123 (exposed for the reference's unit KATs) */
typedef struct orc_dpq orc_dpq;
orc_dpq* orc_dpq_new(void);
void orc_dpq_free(orc_dpq*);
void orc_dpq_push(orc_dpq*, double d, uint64_t id);
size_t orc_dpq_len(const orc_dpq*);
int orc_dpq_pop_first(orc_dpq*, double* d, uint64_t* id);
int orc_dpq_pop_last(orc_dpq*, double* d, uint64_t* id);
int orc_dpq_peek_first(const orc_dpq*, double* d, uint64_t* id);
int orc_dpq_peek_last_dist(const orc_dpq*, double* d);

/* ---- A6: KnnResultBuilder  idx/trees/knn.rs:363-437 ; (dist, doc) ordered set, bounded to knn */
typedef struct orc_krb orc_krb;
orc_krb* orc_krb_new(size_t knn);
void orc_krb_free(orc_krb*);
int orc_krb_check_add(const orc_krb*, double dist);
void orc_krb_add(orc_krb*, double dist, const uint64_t* docs, size_t n_docs);
size_t orc_krb_collect(const orc_krb*, double* out_dist, uint64_t* out_doc);

/* ---- a7-a9, A5, A7: HNSW (F32 vectors).  idx/trees/hnsw/{mod,layer,heuristic}.rs ---- */
typedef struct orc_hnsw orc_hnsw;
/* heuristic flags: bit0 = extend_candidates, bit1 = keep_pruned_connections */
orc_hnsw* orc_hnsw_new(size_t dim, int metric, size_t m, size_t m0, size_t efc, double ml,
                       int heuristic_flags, uint64_t seed);
void orc_hnsw_free(orc_hnsw*);
uint64_t orc_hnsw_insert(orc_hnsw*, const float* v);                       /* hnsw/mod.rs:389-394 */
uint64_t orc_hnsw_insert_level(orc_hnsw*, const float* v, size_t level);   /* hnsw/mod.rs:230-260 */
size_t orc_hnsw_len(const orc_hnsw*);
/* Hnsw::knn_search hnsw/mod.rs:459-482 -> (dist, element) ascending, <= k */
size_t orc_hnsw_search(const orc_hnsw*, const float* q, size_t k, size_t ef, uint64_t* out_ids,
                       double* out_dist);
/* counters of the last orc_hnsw_search on this handle (distance evaluations, expanded nodes) */
void orc_hnsw_last_counters(const orc_hnsw*, uint64_t* visited, uint64_t* expanded);
int orc_hnsw_check_props(const orc_hnsw*);                                  /* layer.rs:571-587 */
/* export for the device loader: n_layers (incl. layer 0), entry point; per layer CSR in insertion
 * order of each node's edge set.  Two-call protocol: pass NULL arrays to get sizes. */
size_t orc_hnsw_n_layers(const orc_hnsw*);
int64_t orc_hnsw_entry_point(const orc_hnsw*);
size_t orc_hnsw_layer_edges(const orc_hnsw*, size_t layer);
void orc_hnsw_export_layer(const orc_hnsw*, size_t layer, uint64_t* row_ptr /* n+1 */,
                           uint32_t* col_idx, uint8_t* present /* n */);
const float* orc_hnsw_vectors(const orc_hnsw*);
/* Same search over an explicit (imported) graph, so the walk can be checked independently of the
 * builder: layers given as CSR over element ids 0..n-1. */
size_t orc_hnsw_search_csr(const float* vectors, size_t n, size_t dim, int metric, size_t n_layers,
                           const uint64_t* const* row_ptr, const uint32_t* const* col_idx,
                           int64_t entry_point, const float* q, size_t k, size_t ef,
                           uint64_t* out_ids, double* out_dist, uint64_t* counters /* [2] nullable */);
/* Hnsw::knn_search_with_filter (hnsw/mod.rs:488-515, layer.rs:111-149,226-306) over an imported graph;
 * truthy[e] != 0 iff any document of element e passes the WHERE condition (hnsw/filter.rs:52-136). */
size_t orc_hnsw_search_csr_filtered(const float* vectors, size_t n, size_t dim, int metric, size_t n_layers,
                                    const uint64_t* const* row_ptr, const uint32_t* const* col_idx,
                                    int64_t entry_point, const float* q, size_t k, size_t ef, const uint8_t* truthy,
                                    uint64_t* out_ids, double* out_dist, uint64_t* counters /* [2] nullable */);
/* Hnsw::knn_search with a pending-docs bitmap (hnsw/mod.rs:459-482, layer.rs:209,320-339): all_docs_pending[e] != 0
 * iff every document of element e has a pending update -- such an element enters w but is never expanded */
size_t orc_hnsw_search_csr_pending(const float* vectors, size_t n, size_t dim, int metric, size_t n_layers,
                                   const uint64_t* const* row_ptr, const uint32_t* const* col_idx, int64_t entry_point,
                                   const float* q, size_t k, size_t ef, const uint8_t* all_docs_pending,
                                   uint64_t* out_ids, double* out_dist, uint64_t* counters /* [2] nullable */);
/* TestCollection::knn  hnsw/mod.rs:1186-1197: brute force through KnnResultBuilder, docs = row ids */
size_t orc_vec_knn_f32(const float* corpus, size_t n, size_t dim, int metric, const float* q, size_t k,
                       uint64_t* out_ids, double* out_dist);

/* ---- a11-a13, A8: graph expansion over CSR adjacency (targets pre-sorted in KV key order) ---- */
/* one `->edge->node` step for every frontier element, in frontier order, duplicates kept,
 * optional per-source limit (0 = none).  Returns number written; pass out=NULL to count. */
uint64_t orc_graph_hop(const uint64_t* row_ptr, const uint32_t* col_idx, const uint32_t* frontier,
                       uint64_t n_frontier, uint32_t per_source_limit, uint32_t* out);
/* evaluate_recurse_collect  exec/operators/recursion/collect.rs:74-143: level-synchronous BFS,
 * first-seen dedup, emits from min_depth; `inclusive` emits the start nodes first. Returns count. */
uint64_t orc_graph_collect(const uint64_t* row_ptr, const uint32_t* col_idx, uint64_t n_nodes,
                           const uint32_t* start, uint64_t n_start, uint32_t min_depth,
                           uint32_t max_depth /* 0 = unbounded */, int inclusive, uint32_t* out,
                           uint64_t out_cap);

/* ---- synthetic data: counter-based generator shared bit-for-bit with the CUDA side
 *      (surrealdb_b200/csrc/gen.cuh).  value(seed, i) in [-1,1). ---- */
float orc_gen_f32(uint64_t seed, uint64_t index);
void orc_gen_fill_f32(uint64_t seed, uint64_t first_index, uint64_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
