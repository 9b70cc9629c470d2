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

struct PassDesc {
  uint32_t stride;  // tiles t = i * stride
  uint32_t excl;    // 0: every i; else = the schedule ratio R: skip i % R == 0 (already done by an earlier pass)
  uint32_t count;   // number of tiles in this pass
};
__host__ __device__ inline uint32_t pass_tile(const PassDesc& p, uint32_t w) {
  uint32_t i = p.excl ? (w / (p.excl - 1)) * p.excl + (w % (p.excl - 1)) + 1 : w;
  return i * p.stride;
}

struct Ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  std::mutex mu;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled (driver entry point), resolved lazily
  void* h_stage = nullptr;       // pinned staging buffer for large device->host results (grow-only)
  size_t h_stage_bytes = 0;
};

struct Cand {  // one screened candidate
  float score; // larger = closer
  uint32_t row;
};

struct Corpus {
  Ctx* ctx = nullptr;
  uint32_t dim = 0, dim_pad = 0;  // dim_pad: bf16 screen copy row length (multiple of 64)
  sdb_dtype dtype = SDB_F32;
  sdb_metric metric = SDB_COSINE;
  sdb_screen screen = SDB_SCREEN_AUTO;
  bool exact = true;  // false: skip the proof / exact fallback (approximate mode)
  sdb_screen ladder_scr = SDB_SCREEN_AUTO;  // the screen the remembered rung belongs to
  uint32_t sc_gen = 0;                      // bumped whenever the per-query scratch is reallocated
  uint32_t ladder_k = 0;                    // ... and the k it was learnt for
  uint32_t ladder_rung = 0;                 // rung of the (screen, slack) ladder the last batch settled on (api.cu)
  uint64_t cap = 0, n = 0;
  bool finalized = false;
  void* d_rows = nullptr;           // master copy, cap x dim (f32 or f64)
  double* d_mag = nullptr;          // exact f64 magnitude per row (reference arithmetic)
  float* d_snorm = nullptr;         // cosine: 1/|x| ; euclid: |x|^2 ; NaN = never a screen candidate
  __nv_bfloat16* d_bf16 = nullptr;  // screen copy cap_pad x dim_pad (rows padded to TILE_ROWS)
  int8_t* d_i8 = nullptr;           // int8 screen copy cap_pad x dim_pad8 (per-row scale max|x|/127), cosine only
  uint32_t dim_pad8 = 0;            // multiple of 128
  float max_rel_qerr = 0.f;         // max over rows of |x/|x| - s * x8|
  float i8_scale = 1.f;             // global scale s of the int8 copy
  uint8_t* d_skip = nullptr;        // optional skip mask
  uint32_t* d_special = nullptr;    // rows ranked exactly on every query
  uint32_t n_special = 0;
  bool special_overflow = false;
  float max_norm = 0.f;
  // ---- search scratch (grown on demand) ----
  uint32_t sc_nq = 0, sc_cap = 0, sc_kp = 0;
  double* d_q64 = nullptr;
  float* d_q32 = nullptr;
  __nv_bfloat16* d_qbf16 = nullptr;
  double* d_qmag = nullptr;
  uint32_t* d_qflags = nullptr;  // bit0: query needs the exact path; bit1: query has NaN input
  int8_t* d_q8 = nullptr;        // int8 queries nq_pad x dim_pad8
  float* d_q8scale = nullptr;    // max|q|/127 per query
  float* d_q8err = nullptr;      // |q - dequant(q)| / |q| per query
  Cand* d_sub = nullptr;         // thread-private candidate sub-lists of the tensor-core screens
  uint32_t* d_sub_cnt = nullptr; // [nq][sub_slots]
  uint32_t sub_slots = 0, sub_cap = 0, last_slots = 0;
  float* d_bscale = nullptr;     // per query: factor that turns tau into similarity*|q| units (1 or q8scale)
  float* d_beps = nullptr;       // per query: rigorous screen error bound in cosine units
  float* d_tau = nullptr;
  Cand* d_cand = nullptr;
  uint32_t* d_cand_cnt = nullptr;
  uint32_t* d_flags = nullptr;   // per query: bit0 overflow, bit1 verification failed
  uint64_t* d_rr_key = nullptr;  // re-rank results: nq x rr_stride
  double* d_rr_dist = nullptr;
  uint32_t* d_rr_row = nullptr;
  uint32_t rr_stride = 0;
  // exact path scratch
  uint64_t* d_ex_key = nullptr;  // N keys
  uint32_t* d_sel = nullptr;     // radix-select state
  uint64_t ex_cap = 0;
  // host-entry staging
  uint64_t* d_out_rows = nullptr;
  double* d_out_dist = nullptr;
  uint32_t* d_out_count = nullptr;
  double* d_in_q = nullptr;
  size_t out_cap = 0, out_cap_q = 0;
  sdb_knn_stats stats{};
  std::mutex mu;
};

// ---- launch wrappers (defined in the .cu files) ---------------------------------------------------
// corpus.cu
sdb_status corpus_finalize_device(Corpus* c);
// screen_simt.cu
sdb_status screen_simt_pass(Corpus* c, uint32_t nq, const PassDesc& p, cudaStream_t st);
// screen_tc.cu
sdb_status screen_tc_pass(Corpus* c, uint32_t nq, const PassDesc& p, bool int8, cudaStream_t st);
bool screen_tc_available();
// candidates.cu
sdb_status scratch_for(Corpus* c, uint32_t nq, uint32_t cap, uint32_t kp);
sdb_status prep_queries(Corpus* c, const double* d_queries, uint32_t nq, cudaStream_t st);
sdb_status cand_reset(Corpus* c, uint32_t nq, cudaStream_t st);
sdb_status cand_set_count(Corpus* c, uint32_t nq, uint32_t value, cudaStream_t st);
sdb_status cand_compact(Corpus* c, uint32_t nq, uint32_t kp, bool drop_invalid, uint32_t n_slots, cudaStream_t st);
sdb_status cand_rerank(Corpus* c, uint32_t nq, cudaStream_t st);
sdb_status set_bounds(Corpus* c, uint32_t nq, int screen, float eps_rel, cudaStream_t st);
sdb_status cand_final(Corpus* c, uint32_t nq, uint32_t k, uint32_t kp, float eps_rel, uint64_t row_base,
                      uint64_t* d_out_rows, double* d_out_dist, uint32_t* d_out_count, cudaStream_t st);
// exact.cu
sdb_status exact_query(Corpus* c, uint32_t q, uint32_t k, uint64_t row_base, uint64_t* d_out_rows,
                       double* d_out_dist, uint32_t* d_out_count, cudaStream_t st);
// gen.cu
sdb_status exact_project(Corpus* c, int fn, double* d_vals, cudaStream_t st);
sdb_status gen_fill_f32(Ctx* ctx, float* d_out, uint64_t seed, uint64_t first, uint64_t n, cudaStream_t st);

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
