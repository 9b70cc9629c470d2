// api.cu -- the extern "C" boundary (include/sdbgpu.h): contexts, corpus lifecycle, brute-force KNN driver.
#include <cmath>

#include "internal.cuh"

namespace sdb {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}


// schedule ratio: every pass looks at (R-1) x the rows seen so far; SDB_PASS_RATIO overrides it (tuning only)
// Cost model (measured on B200): a pass costs a fixed ~0.1 ms (launch, pipeline fill, compaction)
// plus the survivor appends, ~(R-1) * k' per query.  Large batches are append-dominated -> small R; small batches are
// launch-dominated -> fewer, larger passes.
static uint32_t pass_ratio(uint32_t nq) {
  static int env = -1;
  if (env < 0) {
    env = 0;
    if (const char* e = getenv("SDB_PASS_RATIO")) {
      const int v = atoi(e);
      if (v >= 2 && v <= 64) env = v;
    }
  }
  if (env) return (uint32_t)env;
  return nq >= 512 ? 4u : PASS_RATIO;
}

static std::vector<PassDesc> build_passes(uint64_t n_rows, uint32_t cand_cap, uint32_t nq) {
  std::vector<PassDesc> v;
  const uint64_t PASS_RATIO = pass_ratio(nq);  // shadows the compile-time default
  const uint64_t T = (n_rows + TILE_ROWS - 1) / TILE_ROWS;
  if (T == 0) return v;
  const uint64_t max0 = cand_cap / TILE_ROWS;  // pass 0 appends every row it sees
  uint64_t stride = 1;
  while ((T + stride - 1) / stride > max0) stride *= PASS_RATIO;
  PassDesc p0{(uint32_t)stride, 0u, (uint32_t)((T + stride - 1) / stride)};
  v.push_back(p0);
  for (uint64_t s = stride / PASS_RATIO; s >= 1; s /= PASS_RATIO) {
    const uint64_t M = (T + s - 1) / s;
    PassDesc p{(uint32_t)s, (uint32_t)PASS_RATIO, (uint32_t)(M - (M + PASS_RATIO - 1) / PASS_RATIO)};
    v.push_back(p);
    if (s == 1) break;
  }
  return v;
}

static sdb_status knn_device_locked(Corpus* c, const double* d_queries, uint32_t nq, uint32_t k, uint64_t row_base,
                                    uint64_t* d_out_rows, double* d_out_dist, uint32_t* d_out_count,
                                    const volatile int* cancel) {
  Ctx* ctx = c->ctx;
  cudaStream_t st = ctx->stream;
  if (!c->finalized) {
    set_error("corpus not finalized (call sdb_corpus_finalize after the last append)");
    return SDB_EINVAL;
  }
  if (nq == 0) return SDB_OK;
  if (k == 0) {
    SDB_CUDA(cudaMemsetAsync(d_out_count, 0, sizeof(uint32_t) * nq, st));
    SDB_CUDA(cudaStreamSynchronize(st));
    return SDB_OK;
  }
  const uint64_t launches0 = ctx->launches;
  sdb_knn_stats stt{};
  struct Events {  // destroyed on every return path
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Events() {
      for (auto& x : e)
        if (x) cudaEventDestroy(x);
    }
  } evs;
  cudaEvent_t* ev = evs.e;
  for (auto& e : evs.e) SDB_CUDA(cudaEventCreate(&e));
  SDB_CUDA(cudaEventRecord(ev[0], st));

  // ---- choose the screen ----
  // AUTO: the tensor-core screens read 1/2 (bf16) or 1/4 (int8) of the bytes of the f32 stream and are HBM-bound for
  // small batches, so they win at every batch size; int8 is used when the corpus quantises well enough for its
  // error bound to be provable (cosine only); the f32 SIMT stream stays selectable (SDB_SCREEN_SIMT_F32).
  const bool int8_ok = c->d_i8 && c->metric == SDB_COSINE && screen_tc_available();
  sdb_screen scr = c->screen;
  if (scr == SDB_SCREEN_AUTO)
    scr = !screen_tc_available() ? SDB_SCREEN_SIMT_F32 : (int8_ok && c->max_rel_qerr <= 0.006f ? SDB_SCREEN_TC_INT8 : SDB_SCREEN_TC_BF16);
  if (scr == SDB_SCREEN_TC_INT8 && !int8_ok) scr = SDB_SCREEN_TC_BF16;
  if (scr == SDB_SCREEN_TC_BF16 && !screen_tc_available()) scr = SDB_SCREEN_SIMT_F32;
  const bool screenable = c->metric == SDB_COSINE || c->metric == SDB_EUCLIDEAN;
  if (c->dtype == SDB_F64 || c->special_overflow || k > 256 || !screenable) scr = SDB_SCREEN_NONE_EXACT;
  // ---- the ladder: (screen, slack multiplier) rungs, cheapest first.  A rung is abandoned when the exactness proof
  // fails for more than a handful of queries (each failure would otherwise cost a full f64 pass over the corpus in the
  // exact kernel); the rung that worked is remembered per corpus so later batches start there.  More slack (k' x 4)
  // lowers tau relative to the k-th distance -- what high-dimensional / large-k workloads with tightly packed
  // similarities need (BASELINE config 4: 1536 dims, k = 100) -- at the price of 4x the survivor appends.
  struct Rung { sdb_screen scr; uint32_t mult; };
  std::vector<Rung> rungs;
  if (scr == SDB_SCREEN_TC_INT8) rungs = {{SDB_SCREEN_TC_INT8, 1}, {SDB_SCREEN_TC_INT8, 4}, {SDB_SCREEN_TC_BF16, 4}};
  else if (scr == SDB_SCREEN_TC_BF16) rungs = {{SDB_SCREEN_TC_BF16, 1}, {SDB_SCREEN_TC_BF16, 4}};
  else if (scr == SDB_SCREEN_SIMT_F32) rungs = {{SDB_SCREEN_SIMT_F32, 1}};
  if (rungs.empty()) {  // exact-only: the exact kernel still needs the prepared queries (f64 copy, |q|, flags)
    SDB_TRY(scratch_for(c, nq, 4096, k + (k > 54 ? k : 54)));
    SDB_TRY(prep_queries(c, d_queries, nq, st));
  }
  uint32_t rung = 0;
  if (c->ladder_scr == scr && c->ladder_k == k && c->ladder_rung < rungs.size()) rung = c->ladder_rung;
  std::vector<uint32_t> h_flags(nq, 2u), h_qflags(nq, 0u);
  SDB_CUDA(cudaEventRecord(ev[1], st));
  bool first = true;
  for (; rung < rungs.size(); rung++) {
    const sdb_screen rs = rungs[rung].scr;
    uint32_t kp = k + (k > 54 ? k : 54) + (rs == SDB_SCREEN_TC_INT8 ? 64 : 0);  // looser screen => more slack
    kp = kp * rungs[rung].mult;
    if (kp > 1024u) kp = 1024u > 2 * k ? 1024u : 2 * k;
    uint32_t cap = 4096;
    while (cap < 2u * pass_ratio(nq) * kp && cap < 16384u) cap <<= 1;  // a pass appends ~(R-1)*kp survivors per query
    const uint32_t gen_before = c->sc_gen;
    SDB_TRY(scratch_for(c, nq, cap, kp));
    cap = c->sc_cap;
    if (first || gen_before != c->sc_gen) SDB_TRY(prep_queries(c, d_queries, nq, st));  // (re)allocation drops the prepared queries
    if (first) SDB_CUDA(cudaEventRecord(ev[1], st));
    first = false;
    const float eps_rel = rs == SDB_SCREEN_SIMT_F32
                              ? (float)((c->dim / 16.0 + 16.0) * 1.1920929e-7)
                              : (float)(0.00390625 * 1.01 + c->dim * 4.76837158e-7 + 1e-5);
    std::vector<PassDesc> passes = build_passes(c->n, cap, nq);
    SDB_TRY(cand_reset(c, nq, st));
    SDB_TRY(set_bounds(c, nq, (int)rs, eps_rel, st));
    for (const PassDesc& p : passes) {
      if (cancel && *cancel) {
        cudaStreamSynchronize(st);
        set_error("query cancelled");
        return SDB_ECANCELLED;
      }
      if (rs == SDB_SCREEN_SIMT_F32) SDB_TRY(screen_simt_pass(c, nq, p, st));
      else SDB_TRY(screen_tc_pass(c, nq, p, rs == SDB_SCREEN_TC_INT8, st));
      SDB_TRY(cand_compact(c, nq, kp, rs == SDB_SCREEN_TC_INT8, rs == SDB_SCREEN_SIMT_F32 ? 0u : c->last_slots, st));
    }
    SDB_CUDA(cudaEventRecord(ev[2], st));
    SDB_TRY(cand_rerank(c, nq, st));
    SDB_TRY(cand_final(c, nq, k, kp, eps_rel, row_base, d_out_rows, d_out_dist, d_out_count, st));
    SDB_CUDA(cudaMemcpyAsync(h_flags.data(), c->d_flags, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, st));
    stt.n_passes += (uint32_t)passes.size();
    stt.n_reranked += (uint64_t)nq * (kp + c->n_special);
    scr = rs;
    if (!c->exact || rung + 1 == rungs.size()) break;
    SDB_CUDA(cudaStreamSynchronize(st));
    uint32_t n_fail = 0;
    for (uint32_t q = 0; q < nq; q++) n_fail += (h_flags[q] & 2u) ? 1u : 0u;
    if (n_fail <= 2 + nq / 64) break;
    stt.n_candidates += n_fail;  // (diagnostic: queries handed up the ladder)
  }
  if (!rungs.empty()) {
    c->ladder_scr = rungs[0].scr;
    c->ladder_k = k;
    c->ladder_rung = rung < rungs.size() ? rung : (uint32_t)rungs.size() - 1;
  }
  if (scr == SDB_SCREEN_NONE_EXACT) SDB_CUDA(cudaEventRecord(ev[2], st));
  SDB_CUDA(cudaMemcpyAsync(h_qflags.data(), c->d_qflags, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaStreamSynchronize(st));
  // ---- exact path for everything the screens could not prove ----
  for (uint32_t q = 0; q < nq; q++) {
    if (((h_flags[q] & 2u) && (c->exact || scr == SDB_SCREEN_NONE_EXACT)) || (h_qflags[q] & 1u)) {
      if (cancel && *cancel) {
        set_error("query cancelled");
        return SDB_ECANCELLED;
      }
      SDB_TRY(exact_query(c, q, k, row_base, d_out_rows, d_out_dist, d_out_count, st));
      stt.n_fallback++;
    }
  }
  SDB_CUDA(cudaEventRecord(ev[3], st));
  SDB_CUDA(cudaStreamSynchronize(st));
  SDB_CUDA(cudaEventElapsedTime(&stt.screen_ms, ev[1], ev[2]));
  SDB_CUDA(cudaEventElapsedTime(&stt.total_ms, ev[0], ev[3]));
  stt.screen_used = (uint32_t)scr;
  stt.n_special_rows = c->n_special;
  stt.kernel_launches = ctx->launches - launches0;
  c->stats = stt;
  return SDB_OK;
}

// ---- global top-k merge of per-shard lists (after the NCCL all-gather) -----------------------------
__global__ void __launch_bounds__(1024) topk_merge_kernel(uint32_t n_lists, uint32_t nq, uint32_t k,
                                                          const uint64_t* __restrict__ rows,
                                                          const double* __restrict__ dist,
                                                          const uint32_t* __restrict__ counts, uint64_t st_rows,
                                                          uint64_t st_dist, uint64_t st_cnt,
                                                          uint64_t* __restrict__ out_rows, double* __restrict__ out_dist,
                                                          uint32_t* __restrict__ out_count) {
  extern __shared__ uint64_t s_mem[];
  const uint32_t q = blockIdx.x;
  const uint32_t total = n_lists * k;
  uint32_t p2 = 1;
  while (p2 < total) p2 <<= 1;
  uint64_t* s_key = s_mem;
  uint64_t* s_row = s_mem + p2;
  double* s_d = reinterpret_cast<double*>(s_mem + 2 * p2);
  __shared__ uint32_t s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) {
    uint64_t key = ~0ull, row = ~0ull;
    double d = 0.0;
    if (i < total) {
      const uint32_t l = i / k, j = i % k;
      if (j < counts[(size_t)l * st_cnt + q]) {
        const size_t o = (size_t)q * k + j;
        d = dist[(size_t)l * st_dist + o];
        key = dist_key(d);
        row = rows[(size_t)l * st_rows + o];
        atomicAdd(&s_n, 1u);
      }
    }
    s_key[i] = key;
    s_row[i] = row;
    s_d[i] = d;
  }
  __syncthreads();
  for (uint32_t kk = 2; kk <= p2; kk <<= 1)
    for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
          const uint64_t ka = s_key[i], kb = s_key[ixj], ra = s_row[i], rb = s_row[ixj];
          const bool a_gt_b = ka > kb || (ka == kb && ra > rb);
          const bool up = ((i & kk) == 0);
          if (up ? a_gt_b : !a_gt_b) {
            s_key[i] = kb; s_key[ixj] = ka;
            s_row[i] = rb; s_row[ixj] = ra;
            const double da = s_d[i];
            s_d[i] = s_d[ixj];
            s_d[ixj] = da;
          }
        }
      }
      __syncthreads();
    }
  const uint32_t n_out = s_n < k ? s_n : k;
  for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
    out_rows[(size_t)q * k + i] = s_row[i];
    out_dist[(size_t)q * k + i] = s_d[i];
  }
  if (threadIdx.x == 0) out_count[q] = n_out;
}

}  // namespace sdb

using namespace sdb;

extern "C" {

const char* sdb_last_error(void) { return g_err; }
const char* sdb_version(void) { return "sdbgpu 0.1.0 (sm_100a)"; }

sdb_status sdb_ctx_create(int device, sdb_ctx** out) {
  if (!out) return SDB_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return SDB_ECUDA;
  }
  if (device < 0 || device >= n) {
    set_error("device %d out of range (0..%d)", device, n - 1);
    return SDB_EINVAL;
  }
  cudaDeviceProp prop;
  SDB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library ships sm_100a code only", device, prop.major, prop.minor);
    return SDB_ECUDA;
  }
  SDB_CUDA(cudaSetDevice(device));
  sdb_ctx* c = new sdb_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  SDB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {  // keep stream-ordered allocations cached in the pool instead of returning them to the OS at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  *out = c;
  return SDB_OK;
}
void sdb_ctx_destroy(sdb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  delete c;
}
uint64_t sdb_ctx_kernel_launches(const sdb_ctx* c) { return c ? c->launches : 0; }
void* sdb_ctx_stream(const sdb_ctx* c) { return c ? (void*)c->stream : nullptr; }
void* sdb_pinned_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    set_error("cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
void sdb_pinned_free(void* p) {
  if (p) cudaFreeHost(p);
}
void sdb_free(void* p) { free(p); }

sdb_status sdb_corpus_create(sdb_ctx* ctx, uint32_t dim, sdb_dtype dt, sdb_metric m, uint64_t cap, sdb_corpus** out) {
  if (!ctx || !out || dim == 0 || dim > 65535 || cap == 0 || cap >= 0xFFFFFFF0ull) {
    set_error("sdb_corpus_create: bad argument (dim 1..65535, 0 < capacity < 2^32)");
    return SDB_EINVAL;
  }
  // COSINE / EUCLIDEAN: screened (K1/K2) + exact re-rank.  MANHATTAN / CHEBYSHEV / HAMMING / PEARSON: served by the
  // exact kernel alone (sequential f64, bit-identical to Distance::compute).  MINKOWSKI (powf is not bit-reproducible
  // across libm implementations) and JACCARD (set semantics, not a vector metric) stay on the reference's CPU path.
  const bool screenable = m == SDB_COSINE || m == SDB_EUCLIDEAN;
  if (!screenable && m != SDB_MANHATTAN && m != SDB_CHEBYSHEV && m != SDB_HAMMING && m != SDB_PEARSON) {
    set_error("metric %d not implemented on the GPU path (MINKOWSKI and JACCARD stay on the CPU)", (int)m);
    return SDB_EUNSUPPORTED;
  }
  if (dt != SDB_F32 && dt != SDB_F64) return SDB_EINVAL;
  SDB_CUDA(cudaSetDevice(ctx->device));
  sdb_corpus* c = new sdb_corpus();
  c->ctx = ctx;
  c->dim = dim;
  c->dim_pad = (dim + 63) / 64 * 64;
  c->dim_pad8 = (dim + 127) / 128 * 128;
  c->dtype = dt;
  c->metric = m;
  c->cap = cap;
  const size_t esz = dt == SDB_F32 ? 4 : 8;
  const uint64_t cap_pad = (cap + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
  cudaError_t e = cudaMalloc(&c->d_rows, esz * cap * dim);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_mag, sizeof(double) * cap);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_snorm, sizeof(float) * cap_pad);
  if (e == cudaSuccess && dt == SDB_F32 && screenable) e = cudaMalloc(&c->d_bf16, sizeof(__nv_bfloat16) * cap_pad * c->dim_pad);
  if (e == cudaSuccess && dt == SDB_F32 && m == SDB_COSINE) e = cudaMalloc(&c->d_i8, (size_t)cap_pad * c->dim_pad8);
  if (e != cudaSuccess) {
    set_error("corpus allocation failed: %s", cudaGetErrorString(e));
    sdb_corpus_destroy(c);
    return SDB_ENOMEM;
  }
  *out = c;
  return SDB_OK;
}
void sdb_corpus_destroy(sdb_corpus* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  void* ptrs[] = {c->d_i8, c->d_q8, c->d_q8scale, c->d_q8err, c->d_bscale, c->d_beps,
                  c->d_sub, c->d_sub_cnt, c->d_rows, c->d_mag, c->d_snorm, c->d_bf16, c->d_skip, c->d_special, c->d_q64, c->d_q32,
                  c->d_qbf16, c->d_qmag, c->d_qflags, c->d_tau, c->d_cand, c->d_cand_cnt, c->d_flags,
                  c->d_rr_key, c->d_rr_dist, c->d_rr_row, c->d_ex_key, c->d_sel, c->d_out_rows, c->d_out_dist,
                  c->d_out_count, c->d_in_q};
  for (void* p : ptrs) cudaFree(p);
  delete c;
}
uint64_t sdb_corpus_rows(const sdb_corpus* c) { return c ? c->n : 0; }

static sdb_status append_common(sdb_corpus* c, const void* src, uint64_t n, cudaMemcpyKind kind) {
  if (!c || (!src && n)) return SDB_EINVAL;
  std::lock_guard<std::mutex> g(c->mu);
  if (c->n + n > c->cap) {
    set_error("append of %llu rows exceeds capacity %llu", (unsigned long long)n, (unsigned long long)c->cap);
    return SDB_EOVERFLOW;
  }
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  const size_t esz = c->dtype == SDB_F32 ? 4 : 8;
  SDB_CUDA(cudaMemcpyAsync((char*)c->d_rows + esz * c->n * c->dim, src, esz * n * c->dim, kind, c->ctx->stream));
  SDB_CUDA(cudaStreamSynchronize(c->ctx->stream));
  c->n += n;
  c->finalized = false;
  return SDB_OK;
}
sdb_status sdb_corpus_append(sdb_corpus* c, const void* rows, uint64_t n) {
  return append_common(c, rows, n, cudaMemcpyHostToDevice);
}
sdb_status sdb_corpus_append_device(sdb_corpus* c, const void* d_rows, uint64_t n) {
  return append_common(c, d_rows, n, cudaMemcpyDeviceToDevice);
}
sdb_status sdb_corpus_append_synthetic(sdb_corpus* c, uint64_t seed, uint64_t first_row, uint64_t n) {
  if (!c) return SDB_EINVAL;
  if (c->dtype != SDB_F32) {
    set_error("synthetic rows are f32");
    return SDB_EUNSUPPORTED;
  }
  std::lock_guard<std::mutex> g(c->mu);
  if (c->n + n > c->cap) return SDB_EOVERFLOW;
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  SDB_TRY(gen_fill_f32(c->ctx, (float*)c->d_rows + c->n * c->dim, seed, first_row * c->dim, n * c->dim, c->ctx->stream));
  SDB_CUDA(cudaStreamSynchronize(c->ctx->stream));
  c->n += n;
  c->finalized = false;
  return SDB_OK;
}
sdb_status sdb_corpus_set_skip(sdb_corpus* c, const uint8_t* skip, uint64_t n) {
  if (!c || n > c->cap) return SDB_EINVAL;
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  if (!skip) {
    cudaFree(c->d_skip);
    c->d_skip = nullptr;
  } else {
    if (!c->d_skip) SDB_CUDA(cudaMalloc(&c->d_skip, c->cap));
    SDB_CUDA(cudaMemsetAsync(c->d_skip, 0, c->cap, c->ctx->stream));
    SDB_CUDA(cudaMemcpyAsync(c->d_skip, skip, n, cudaMemcpyHostToDevice, c->ctx->stream));
    SDB_CUDA(cudaStreamSynchronize(c->ctx->stream));
  }
  c->finalized = false;
  return SDB_OK;
}
sdb_status sdb_corpus_finalize(sdb_corpus* c) {
  if (!c) return SDB_EINVAL;
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  return corpus_finalize_device(c);
}
sdb_status sdb_corpus_set_screen(sdb_corpus* c, sdb_screen s) {
  if (!c || (int)s < 0 || (int)s > 4) return SDB_EINVAL;
  c->screen = s;
  return SDB_OK;
}
sdb_status sdb_corpus_set_exact(sdb_corpus* c, int exact) {
  if (!c) return SDB_EINVAL;
  c->exact = exact != 0;
  return SDB_OK;
}
sdb_status sdb_knn_last_stats(const sdb_corpus* c, sdb_knn_stats* out) {
  if (!c || !out) return SDB_EINVAL;
  *out = c->stats;
  return SDB_OK;
}

sdb_status sdb_knn_bruteforce_device(sdb_corpus* c, const double* d_queries, uint32_t nq, uint32_t k,
                                     uint64_t row_base, uint64_t* d_out_rows, double* d_out_dist,
                                     uint32_t* d_out_count) {
  if (!c || (nq && (!d_queries || !d_out_count || (k && (!d_out_rows || !d_out_dist))))) return SDB_EINVAL;
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  return knn_device_locked(c, d_queries, nq, k, row_base, d_out_rows, d_out_dist, d_out_count, nullptr);
}

sdb_status sdb_knn_bruteforce(sdb_corpus* c, const double* queries, uint32_t nq, uint32_t k, uint64_t* out_rows,
                              double* out_dist, uint32_t* out_count, const volatile int* cancel_flag) {
  if (!c || (nq && (!queries || !out_count || (k && (!out_rows || !out_dist))))) return SDB_EINVAL;
  if (nq == 0) return SDB_OK;
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  cudaStream_t st = c->ctx->stream;
  const size_t need = (size_t)nq * (k ? k : 1);
  if (c->out_cap < need || c->out_cap_q < nq) {
    cudaFree(c->d_out_rows);
    cudaFree(c->d_out_dist);
    cudaFree(c->d_out_count);
    cudaFree(c->d_in_q);
    c->d_out_rows = nullptr; c->d_out_dist = nullptr; c->d_out_count = nullptr; c->d_in_q = nullptr;
    SDB_CUDA(cudaMalloc(&c->d_out_rows, sizeof(uint64_t) * need));
    SDB_CUDA(cudaMalloc(&c->d_out_dist, sizeof(double) * need));
    SDB_CUDA(cudaMalloc(&c->d_out_count, sizeof(uint32_t) * nq));
    SDB_CUDA(cudaMalloc(&c->d_in_q, sizeof(double) * (size_t)nq * c->dim));
    c->out_cap = need;
    c->out_cap_q = nq;
  }
  SDB_CUDA(cudaMemcpyAsync(c->d_in_q, queries, sizeof(double) * (size_t)nq * c->dim, cudaMemcpyHostToDevice, st));
  SDB_TRY(knn_device_locked(c, c->d_in_q, nq, k, 0, c->d_out_rows, c->d_out_dist, c->d_out_count, cancel_flag));
  if (k) {
    SDB_CUDA(cudaMemcpyAsync(out_rows, c->d_out_rows, sizeof(uint64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaMemcpyAsync(out_dist, c->d_out_dist, sizeof(double) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
  }
  SDB_CUDA(cudaMemcpyAsync(out_count, c->d_out_count, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaStreamSynchronize(st));
  return SDB_OK;
}

sdb_status sdb_corpus_project(sdb_corpus* c, const double* query, int fn, double* out) {
  const bool is_metric = fn == SDB_COSINE || fn == SDB_EUCLIDEAN || fn == SDB_MANHATTAN || fn == SDB_CHEBYSHEV ||
                         fn == SDB_HAMMING || fn == SDB_PEARSON;
  if (!c || !out || (!query && fn != SDB_FN_MAGNITUDE)) return SDB_EINVAL;
  if (!is_metric && fn != SDB_FN_SIMILARITY_COSINE && fn != SDB_FN_DOT && fn != SDB_FN_MAGNITUDE) {
    set_error("vector function %d not implemented on the GPU path", fn);
    return SDB_EUNSUPPORTED;
  }
  std::lock_guard<std::mutex> g(c->mu);
  if (!c->finalized) {
    set_error("corpus not finalized (call sdb_corpus_finalize after the last append)");
    return SDB_EINVAL;
  }
  if (c->n == 0) return SDB_OK;
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  cudaStream_t st = c->ctx->stream;
  double *d_q = nullptr, *d_vals = nullptr;
  auto run = [&]() -> sdb_status {
    SDB_CUDA(cudaMallocAsync(&d_q, sizeof(double) * c->dim, st));
    SDB_CUDA(cudaMallocAsync(&d_vals, sizeof(double) * c->n, st));
    if (query) SDB_CUDA(cudaMemcpyAsync(d_q, query, sizeof(double) * c->dim, cudaMemcpyHostToDevice, st));
    else SDB_CUDA(cudaMemsetAsync(d_q, 0, sizeof(double) * c->dim, st));
    SDB_TRY(scratch_for(c, 1, 4096, 64));
    SDB_TRY(prep_queries(c, d_q, 1, st));
    SDB_TRY(exact_project(c, fn, d_vals, st));
    SDB_CUDA(cudaMemcpyAsync(out, d_vals, sizeof(double) * c->n, cudaMemcpyDeviceToHost, st));
    return SDB_OK;
  };
  const sdb_status rc = run();
  if (d_q) cudaFreeAsync(d_q, st);  // also on the error paths
  if (d_vals) cudaFreeAsync(d_vals, st);
  if (cudaStreamSynchronize(st) != cudaSuccess && rc == SDB_OK) {
    set_error("sdb_corpus_project: %s", cudaGetErrorString(cudaGetLastError()));
    return SDB_ECUDA;
  }
  return rc;
}

sdb_status sdb_topk_merge_device(sdb_ctx* ctx, uint32_t n_lists, uint32_t nq, uint32_t k, const uint64_t* d_rows,
                                 const double* d_dist, const uint32_t* d_counts, uint64_t stride_rows,
                                 uint64_t stride_dist, uint64_t stride_counts, uint64_t* d_out_rows,
                                 double* d_out_dist, uint32_t* d_out_count) {
  if (!stride_rows) stride_rows = (uint64_t)nq * k;
  if (!stride_dist) stride_dist = (uint64_t)nq * k;
  if (!stride_counts) stride_counts = nq;
  if (!ctx || !n_lists || !k || !d_rows || !d_dist || !d_counts || !d_out_rows || !d_out_dist || !d_out_count)
    return SDB_EINVAL;
  if (nq == 0) return SDB_OK;
  uint32_t p2 = 1;
  while (p2 < n_lists * k) p2 <<= 1;
  const size_t smem = (size_t)p2 * 24;
  if (smem > 200 * 1024) {
    set_error("merge of %u lists x k=%u exceeds the shared-memory sorter", n_lists, k);
    return SDB_EUNSUPPORTED;
  }
  std::lock_guard<std::mutex> g(ctx->mu);
  SDB_CUDA(cudaSetDevice(ctx->device));
  SDB_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  topk_merge_kernel<<<nq, 1024, smem, ctx->stream>>>(n_lists, nq, k, d_rows, d_dist, d_counts, stride_rows, stride_dist,
                                                     stride_counts, d_out_rows, d_out_dist, d_out_count);
  count_launch(ctx);
  SDB_CUDA(cudaGetLastError());
  SDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SDB_OK;
}

}  // extern "C"
