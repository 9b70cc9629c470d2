// internal.cuh -- shared declarations of the sdbgpu library (not part of the ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/sdbgpu.h"

namespace sdb {

void set_error(const char* fmt, ...);

#define SDB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      ::sdb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return SDB_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)
#define SDB_TRY(call)                  \
  do {                                 \
    sdb_status s__ = (call);           \
    if (s__ != SDB_OK) return s__;     \
  } while (0)

constexpr int TILE_ROWS = 256;    // screening tile = 256 corpus rows (one tcgen05 N=256 MMA tile)
constexpr int PASS_RATIO = 8;     // default geometric threshold-refinement ratio (api.cu:pass_ratio picks per batch size)
constexpr int SPECIAL_CAP = 1024; // rows with zero / non-finite norm handled by exact ranking

// ---- ordered keys -------------------------------------------------------------------------------
// Number::cmp on Floats (val/number.rs:620-633): -0.0 == 0.0, otherwise f64::total_cmp.
// All bit manipulation is done on integers obtained through an opaque move: nvcc otherwise rewrites
// `bits(d) | signbit` into fneg(fabs(d)), implements it with a DADD, and the DADD canonicalises NaNs --
// which silently destroyed the sign/payload of NaN distances (found on B200, round 1).
__host__ __device__ inline uint64_t f64_bits(double d) {
  uint64_t b;
#ifdef __CUDA_ARCH__
  asm volatile("mov.b64 %0, %1;" : "=l"(b) : "d"(d));
#else
  memcpy(&b, &d, 8);
#endif
  return b;
}
__host__ __device__ inline uint64_t dist_key(double d) {
  uint64_t b = f64_bits(d);
  if ((b << 1) == 0) b = 0;  // canonicalise -0.0
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline uint32_t f32_key(float f) {  // ascending key of a float (total order)
  uint32_t b;
#ifdef __CUDA_ARCH__
  asm volatile("mov.b32 %0, %1;" : "=r"(b) : "f"(f));
#else
  memcpy(&b, &f, 4);
#endif
  return (b >> 31) ? ~b : (b | 0x80000000u);
}

// ---- streaming threshold refinement (screen_tc.cu, candidates.cu) -------------------------------------------------
// Per query a 256-bin histogram of the scores appended so far, log-linear above the query's floor `lo`:
//   t = (score - lo) / w0 + 1 (>= 1),  bin = (bits(t) - bits(1.0f)) >> 19  -- 16 bins per octave of t, clamped to 255.
// bin() is monotone and edge(bin(s)) <= s up to float rounding (the refiner subtracts a guard), so
// "suffix count from the top reaches k at bin b" proves that the k-th best score seen so far is >= edge(b).
constexpr uint32_t HIST_BINS = 256;
constexpr uint32_t PROBE_TILES_MAX = 64;                   // tiles scored by the probe launch
constexpr uint32_t PROBE_STRIDE = PROBE_TILES_MAX * 8;     // chunk maxima per query (8 chunks of 32 rows per tile)
struct HistParam {
  float lo;      // floor: tau at the time the histogram was seeded
  float inv_w0;  // 1 / w0
  float w0;      // width scale (a quarter of the query's error margin, never 0)
  float margin;  // 2.1 x the screen's error bound in score units: tau = (k-th best score) - margin
};
__host__ __device__ inline uint32_t hist_bin(const HistParam& p, float score) {
  float t = (score - p.lo) * p.inv_w0 + 1.0f;
  t = t >= 1.0f ? t : 1.0f;  // also maps NaN to bin 0
  uint32_t b;
#ifdef __CUDA_ARCH__
  b = (__float_as_uint(t) - 0x3F800000u) >> 19;
#else
  uint32_t u;
  memcpy(&u, &t, 4);
  b = (u - 0x3F800000u) >> 19;
#endif
  return b < HIST_BINS - 1 ? b : HIST_BINS - 1;
}
__host__ __device__ inline double hist_edge(const HistParam& p, uint32_t b) {  // lower edge of bin b
  const uint32_t u = 0x3F800000u + (b << 19);
  float t;
#ifdef __CUDA_ARCH__
  t = __uint_as_float(u);
#else
  memcpy(&t, &u, 4);
#endif
  return (double)p.lo + (double)p.w0 * ((double)t - 1.0);
}

struct PassDesc {
  uint32_t stride;  // tiles t = i * stride
  uint32_t excl;    // 0: every i; else = the schedule ratio R: skip i % R == 0 (already done by an earlier pass)
  uint32_t count;   // number of tiles in this pass
  uint32_t perm;    // 0: visit in order; else an odd multiplier coprime with `count`: the j-th tile visited is
                    // (j * perm) % count, so that every stretch of the streaming pass samples the whole corpus
};
__host__ __device__ inline uint32_t pass_tile(const PassDesc& p, uint32_t w) {
  if (p.perm) w = (uint32_t)(((uint64_t)w * p.perm) % p.count);
  uint32_t i = p.excl ? (w / (p.excl - 1)) * p.excl + (w % (p.excl - 1)) + 1 : w;
  return i * p.stride;
}

struct Comm;  // comm.cu: NCCL communicator attached to a context (nullptr = single shard)

struct Ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;      // odd ticket slots: batch i+1 runs here, so its screen overlaps batch i's tail
  cudaStream_t copy_stream = nullptr;  // host<->device copies of the asynchronous entry points
  uint64_t launches = 0;
  std::mutex mu;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled (driver entry point), resolved lazily
  void* h_stage = nullptr;       // pinned staging buffer for large device->host results (grow-only)
  size_t h_stage_bytes = 0;
  Comm* comm = nullptr;
  // cancellation (sdb_ctx_cancel): one int in pinned memory that the host side polls between kernel phases, mirrored
  // into a word in DEVICE memory (copied on its own stream by sdb_ctx_cancel) that long-running kernels (the HNSW walk)
  // poll per query -- polling the pinned word itself from thousands of warps is a PCIe read each (ncu r2: 8 % of the
  // walk's stall samples)
  volatile int* h_cancel = nullptr;
  int* d_cancel = nullptr;
  cudaStream_t cancel_stream = nullptr;
  int prio_high = 0;  // greatest launch priority of the device (cudaDeviceGetStreamPriorityRange)
  cudaEvent_t trace_epoch = nullptr;  // SDB_TRACE
  double trace_host0 = 0.0;           // host clock (s) at the epoch
};
inline bool ctx_cancelled(const Ctx* ctx) { return ctx->h_cancel && *ctx->h_cancel != 0; }

// per-device kernel attributes (dynamic shared-memory limits).  cudaFuncSetAttribute is per DEVICE, so these run in
// sdb_ctx_create after cudaSetDevice -- never behind a process-wide flag (a second context on another GPU of the same
// process would otherwise launch with the 48 KB default and fail).
sdb_status screen_tc_init_device();
sdb_status candidates_init_device();
sdb_status exact_init_device();
void comm_destroy(Ctx* ctx);  // comm.cu
void comm_corpus_released(struct Corpus* c);  // comm.cu: drops the corpus's sharded-search state (slot buffers, arena)
int comm_size(const Ctx* ctx);
int comm_rank(const Ctx* ctx);
sdb_status comm_allreduce_sum(Ctx* ctx, void* d_buf, size_t count, int elem_bytes, cudaStream_t st);

struct Cand {  // one screened candidate
  float score; // larger = closer
  uint32_t row;
};

// one asynchronous batch (sdb_knn_submit* ... sdb_knn_wait): everything wait() needs to finish it on the host side
struct Ticket {
  bool busy = false;
  uint32_t id = 0;
  const double* d_queries = nullptr;  // caller's device queries (must stay valid until wait)
  uint32_t nq = 0, k = 0;
  uint64_t row_base = 0;
  uint64_t *d_out_rows = nullptr, *h_out_rows = nullptr;  // device result (caller's or the slot's) / optional host copy
  double *d_out_dist = nullptr, *h_out_dist = nullptr;
  uint32_t *d_out_count = nullptr, *h_out_count = nullptr;
  const volatile int* cancel = nullptr;
  cudaStream_t stream = nullptr;  // every kernel / copy of this batch (slot parity picks the context's stream)
  int set = 0;                    // scratch set of this batch
  int screen = 0;       // sdb_screen this batch ran
  uint32_t rung = 0, n_rungs = 0, n_batch_rungs = 0, n_repaired = 0;
  uint32_t n_passes = 0;
  uint64_t launches0 = 0;
  cudaEvent_t ev_begin = nullptr, ev_screen0 = nullptr, ev_screen1 = nullptr, ev_end = nullptr;
  uint32_t* h_flags = nullptr;   // pinned: per query, bit0 overflow, bit1 proof failed
  uint32_t* h_qflags = nullptr;  // pinned: per query, bit0 needs the exact path, bit1 NaN input
  uint32_t* h_stat = nullptr;    // pinned: [0] queries flagged, [1] candidates re-ranked, [2] max per query, [3] survivors gathered
  uint32_t h_cap = 0;
  // host-buffer entry points: per-slot device staging
  double* d_in_q = nullptr;
  uint64_t* d_res_rows = nullptr;
  double* d_res_dist = nullptr;
  uint32_t* d_res_count = nullptr;
  size_t in_cap = 0, res_cap = 0, res_cap_q = 0;
  cudaEvent_t ev_h2d = nullptr, ev_out = nullptr;
  cudaEvent_t ev_main = nullptr;  // recorded after this batch's last screen launch (the next batch's screen waits for it)
  bool wait_h2d = false;  // the batch's stream still has to wait for ev_h2d (queries travelling on the copy stream)
  // SDB_TRACE=1: named timestamps of this batch on its stream, printed at wait time relative to the context's epoch
  std::vector<std::pair<const char*, cudaEvent_t>> trace;
};
constexpr int N_TICKETS = 4;
bool trace_enabled();
void trace_mark(Ctx* ctx, Ticket& t, const char* name, cudaStream_t st);  // api.cu
void trace_dump(Ctx* ctx, Ticket& t);
void trace_host(Ctx* ctx, uint32_t ticket, const char* name);  // host-side timestamp on the same time base

// Per-batch search scratch.  Two sets exist per corpus: consecutive batches alternate between them (and between the
// context's two streams), so the screen of batch i+1 can run while the tail of batch i (candidate selection, f32
// re-score, exact re-rank, final ordering, all-gather + merge) is still in flight.  The Corpus object itself carries
// the ACTIVE set's fields (it derives from Scratch): enqueue_batch swaps the ticket's set in before it launches
// anything, and every launch captures the pointers by value.
struct Scratch {
  uint32_t sc_nq = 0, sc_cap = 0;
  double* d_q64 = nullptr;
  float* d_q32 = nullptr;
  __nv_bfloat16* d_qbf16 = nullptr;
  double* d_qmag = nullptr;
  uint32_t* d_qflags = nullptr;  // bit0: query needs the exact path; bit1: query has NaN input
  float* d_qbferr = nullptr;     // |q - bf16(q)| / |q| per query
  int8_t* d_q8 = nullptr;        // int8 queries nq_pad x dim_pad8
  float* d_q8scale = nullptr;    // max|q|/127 per query
  float* d_q8err = nullptr;      // |q - dequant(q)| / |q| per query
  Cand* d_sub = nullptr;         // thread-private candidate sub-lists of the tensor-core screens
  uint32_t* d_sub_cnt = nullptr; // [nq][sub_slots]
  uint32_t sub_slots = 0, sub_cap = 0, last_slots = 0;
  float* d_bscale = nullptr;     // per query: factor that turns tau into similarity*|q| units (1 or q8scale * i8_scale)
  float* d_beps = nullptr;       // per query: rigorous screen error bound (cosine units / relative dot error)
  float* d_margin = nullptr;     // per query: 2.1 x that bound in score units (0 in approximate mode)
  float* d_margin2 = nullptr;    // stage B (f32 re-score of the candidates): margin, error bound, threshold
  float* d_beps2 = nullptr;
  float* d_tau2 = nullptr;
  float* d_qlow = nullptr;       // per query: lower / upper bound of any score (histogram geometry)
  float* d_qcap = nullptr;
  HistParam* d_hparam = nullptr; // per query histogram geometry of the streaming screen
  uint32_t* d_hist = nullptr;    // [nq][HIST_BINS]
  float* d_probe = nullptr;      // [nq][PROBE_STRIDE] chunk maxima of the probe launch
  float* d_tau = nullptr;
  Cand* d_cand = nullptr;
  uint32_t* d_cand_cnt = nullptr;
  uint32_t* d_flags = nullptr;   // per query: bit0 overflow, bit1 verification failed
  uint32_t* d_stat = nullptr;    // [0] queries flagged by cand_final, [1] candidates re-ranked, [2] max per query, [3] gathered
  uint64_t* d_rr_key = nullptr;  // re-rank results: nq x rr_stride
  double* d_rr_dist = nullptr;
  uint32_t* d_rr_row = nullptr;
  uint32_t rr_stride = 0;
  uint32_t sc_gen = 0;           // bumped whenever this set is reallocated
};

struct Corpus : Scratch {
  Ctx* ctx = nullptr;
  uint32_t dim = 0, dim_pad = 0;  // dim_pad: bf16 screen copy row length (multiple of 64)
  sdb_dtype dtype = SDB_F32;
  sdb_metric metric = SDB_COSINE;
  sdb_screen screen = SDB_SCREEN_AUTO;
  bool exact = true;   // false: skip the proof / exact fallback (approximate mode)
  bool stream_refine = true;  // tensor-core screens: one streaming launch with in-kernel threshold refinement
  double minkowski_p = 3.0;   // order of SDB_MINKOWSKI
  sdb_screen rung_scr = SDB_SCREEN_AUTO;  // the first-choice screen the remembered rung belongs to
  uint32_t rung_k = 0;                    // ... and the k it was learnt for
  uint32_t rung = 0;                      // rung of the precision ladder the last batch settled on (api.cu)
  uint64_t cap = 0, n = 0;
  uint64_t row_base = 0;            // global id of row 0 (row-sharded corpora)
  bool finalized = false;
  void* d_rows = nullptr;           // master copy, cap x dim (f32 or f64)
  double* d_mag = nullptr;          // exact f64 magnitude per row (reference arithmetic)
  float* d_snorm = nullptr;         // cosine: 1/|x| ; euclid: |x|^2 ; NaN = never a screen candidate
  __nv_bfloat16* d_bf16 = nullptr;  // screen copy cap_pad x dim_pad (rows padded to TILE_ROWS)
  float bf16_rel_err = 0.00390625f; // max over rows of |x - bf16(x)| / |x| (measured at finalize, rounded up)
  int8_t* d_i8 = nullptr;           // int8 screen copy cap_pad x dim_pad8 of the normalised rows (one global scale), cosine only
  uint32_t dim_pad8 = 0;            // multiple of 128
  float max_rel_qerr = 0.f;         // max over rows of |x/|x| - s * x8|
  float i8_scale = 1.f;             // global scale s of the int8 copy
  uint8_t* d_skip = nullptr;        // optional skip mask
  uint8_t* d_removed = nullptr;     // tombstones (sdb_corpus_remove); OR-ed with the skip mask at finalize
  uint64_t n_removed = 0;
  uint32_t* d_special = nullptr;    // rows ranked exactly on every query
  uint32_t n_special = 0;
  uint32_t n_outliers = 0;          // of those: rows made special because one component dominates (int8 scale)
  bool special_overflow = false;
  float max_norm = 0.f;
  // exact path scratch
  uint64_t* d_ex_key = nullptr;  // N keys
  uint32_t* d_sel = nullptr;     // radix-select state
  uint64_t ex_cap = 0;
  double* d_fb_q = nullptr;      // fallback query scratch (one query: f64 copy, |q|, flags)
  double* d_rp_q = nullptr;      // repair sub-batch: the failed queries of a batch, gathered, and their results
  uint64_t* d_rp_rows = nullptr;
  double* d_rp_dist = nullptr;
  uint32_t* d_rp_cnt = nullptr;
  size_t rp_cap_q = 0, rp_cap_o = 0, rp_cap_n = 0;
  double* d_fb_qmag = nullptr;
  uint32_t* d_fb_qflags = nullptr;
  Scratch sets[2];  // the inactive set's fields are parked here (see Scratch)
  int active_set = 0;
  cudaEvent_t last_main = nullptr;  // ev_main of the batch whose screen was enqueued last
  // asynchronous batches
  Ticket tickets[N_TICKETS];
  uint32_t next_ticket = 1;
  // sharded search (comm.cu): this rank's result block + the all-gathered blocks
  uint8_t* d_block = nullptr;
  uint8_t* d_gather = nullptr;
  size_t block_cap = 0, gather_cap = 0;
  sdb_knn_stats stats{};
  std::mutex mu;
};

// ---- launch wrappers (defined in the .cu files) ---------------------------------------------------
// corpus.cu
sdb_status corpus_finalize_device(Corpus* c);
sdb_status corpus_remove_device(Corpus* c, const uint64_t* h_ids, uint64_t n);
sdb_status corpus_reapply_tombstones(Corpus* c, cudaStream_t st);
// screen_simt.cu
sdb_status screen_simt_pass(Corpus* c, uint32_t nq, const PassDesc& p, cudaStream_t st);
// screen_tc.cu
// mode 0: pass 0 (every score of the pass's tiles written to fixed slots), 1: threshold pass, 2: streaming pass with
// in-kernel threshold refinement (histogram + refiner warp), 3: probe (chunk maxima of a few tiles, no candidates)
sdb_status screen_tc_pass(Corpus* c, uint32_t nq, uint32_t k, const PassDesc& p, bool int8, int mode, cudaStream_t st);
bool screen_tc_available();
// candidates.cu
sdb_status scratch_for(Corpus* c, uint32_t nq, uint32_t cap);
sdb_status prep_queries(Corpus* c, const double* d_queries, uint32_t nq, cudaStream_t st);
// one query prepared into the fallback scratch (d_fb_*), independent of the batch scratch
sdb_status prep_fallback_query(Corpus* c, const double* d_query, cudaStream_t st);
// resets tau / counts / flags and derives, per query, the screen's error bound, the selection margin and the score range
sdb_status cand_begin(Corpus* c, uint32_t nq, int screen, cudaStream_t st);
sdb_status cand_set_count(Corpus* c, uint32_t nq, uint32_t value, cudaStream_t st);
// per query: gather the main list + the private sub-lists, find the k-th best score s_k, keep every candidate with
// score >= tau = s_k - margin (all of them while fewer than k exist), publish tau.  seed_hist: also (re)build the
// query's histogram (geometry + counts of the kept candidates) for the streaming pass that follows.
sdb_status cand_select(Corpus* c, uint32_t nq, uint32_t k, bool drop_invalid, uint32_t n_slots, bool seed_hist,
                       cudaStream_t st, int stage = 0);
// stage B: re-score every kept candidate in f32 (master rows x f32 query) so that cand_select(stage 1) can shrink the
// set before the FP64-bound exact re-rank
sdb_status cand_refine(Corpus* c, uint32_t nq, cudaStream_t st);
// after a probe launch over n_tiles tiles: tau = (k-th largest chunk maximum) - margin, histogram geometry, empty lists
sdb_status cand_seed_from_probe(Corpus* c, uint32_t nq, uint32_t k, uint32_t n_tiles, cudaStream_t st);
sdb_status cand_rerank(Corpus* c, uint32_t nq, cudaStream_t st, bool small_sets = false);
sdb_status cand_final(Corpus* c, uint32_t nq, uint32_t k, uint64_t row_base, uint64_t* d_out_rows, double* d_out_dist,
                      uint32_t* d_out_count, cudaStream_t st);
// exact.cu: query vector / |q| / flags are passed explicitly (batch scratch row or the fallback scratch)
sdb_status exact_query(Corpus* c, const double* d_q64, const double* d_qmag, const uint32_t* d_qflags, uint32_t k,
                       uint64_t row_base, uint64_t* d_out_rows, double* d_out_dist, uint32_t* d_out_count,
                       cudaStream_t st);
// gen.cu
sdb_status exact_project(Corpus* c, int fn, double* d_vals, cudaStream_t st);
sdb_status gen_fill_f32(Ctx* ctx, float* d_out, uint64_t seed, uint64_t first, uint64_t n, cudaStream_t st);

// hnsw.cu: range / monotonicity check of a device CSR handed over the ABI (SDB_EINVAL with a message on violation)
sdb_status csr_check(Ctx* ctx, const uint64_t* d_rp, const uint32_t* d_ci, uint64_t n_rows, uint64_t n_edges,
                     uint64_t id_limit, unsigned long long counts[2], const char* what, cudaStream_t st);
sdb_status csr_validate(Ctx* ctx, const uint64_t* d_rp, const uint32_t* d_ci, uint64_t n_rows, uint64_t n_edges,
                        uint64_t id_limit, const char* what, cudaStream_t st);
// graph.cu: out[0..n) = exclusive scan of in[0..n); *d_total = sum (in/out may alias)
sdb_status exclusive_scan(Ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_total, cudaStream_t st);
// stage.cu: He / Hn value decoders (host blobs in, device arrays out)
sdb_status stage_decode_vectors(Ctx* ctx, const uint8_t* blob, const uint64_t* off, const uint64_t* ids, uint64_t n,
                                uint32_t dim, sdb_dtype out_dtype, uint64_t n_rows, void* d_out, uint8_t* d_present,
                                uint64_t* n_bad, cudaStream_t st);
sdb_status stage_decode_nodes(Ctx* ctx, const uint8_t* blob, const uint64_t* off, const uint64_t* node_ids, uint64_t n,
                              uint64_t n_elems, uint64_t** d_row_ptr_out, uint32_t** d_col_idx_out, uint64_t* n_edges,
                              uint64_t* n_bad, cudaStream_t st);

inline void count_launch(Ctx* ctx, uint64_t n = 1) { ctx->launches += n; }

}  // namespace sdb

struct sdb_ctx : sdb::Ctx {};
struct sdb_corpus : sdb::Corpus {};
