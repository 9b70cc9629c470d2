// candidates.cu -- everything between the screening kernels and the answer:
//   prep_queries : f64 queries -> f32 / bf16 screen copies, exact |q|, special-query flags
//   cand_select  : per query keep the screened candidates whose score reaches tau = (k-th best score) - error margin
//   cand_rerank  : exact f64 distances (the reference's arithmetic, op for op) of the survivors
//   cand_final   : order by (distance, scan position), emit top-k, PROVE that no unscreened row could
//                  belong to it (error-bound check) or flag the query for the exact kernel.
#include "exactmath.cuh"
#include "internal.cuh"
#include "rowwalk.cuh"

namespace sdb {

// ------------------------------------------------------------------------------------------------
__global__ void prep_queries_kernel(const double* __restrict__ q64, uint32_t dim, uint32_t dim_pad, int metric,
                                    float* __restrict__ q32, __nv_bfloat16* __restrict__ qbf, double* __restrict__ qmag,
                                    uint32_t* __restrict__ qflags, float* __restrict__ qbferr, uint32_t nq) {
  const uint32_t q = blockIdx.x;
  __shared__ uint32_t s_flags;
  __shared__ float s_err2;
  if (threadIdx.x == 0) {
    s_flags = 0;
    s_err2 = 0.f;
  }
  __syncthreads();
  uint32_t fl = 0;
  float err2 = 0.f;
  if (q < nq) {
    for (uint32_t c = threadIdx.x; c < dim_pad; c += blockDim.x) {
      const double v = c < dim ? q64[(size_t)q * dim + c] : 0.0;
      const float f = (float)v;
      if (q32 && c < dim) q32[(size_t)q * dim + c] = f;
      const __nv_bfloat16 h = __float2bfloat16_rn(f);
      if (qbf) qbf[(size_t)q * dim_pad + c] = h;
      const float d = f - __bfloat162float(h);
      err2 = fmaf(d, d, err2);
      if (v != v) fl |= 3u;               // NaN input: exact path, positive-NaN propagation
      else if (!isfinite(f)) fl |= 1u;    // inf or beyond f32 range: the screen cannot bound its error
    }
  } else {  // padding queries of the bf16 operand tile
    for (uint32_t c = threadIdx.x; c < dim_pad; c += blockDim.x)
      if (qbf) qbf[(size_t)q * dim_pad + c] = __float2bfloat16_rn(0.f);
  }
  if (fl) atomicOr(&s_flags, fl);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, o);
  if ((threadIdx.x & 31) == 0 && err2 > 0.f) atomicAdd(&s_err2, err2);
  __syncthreads();
  if (threadIdx.x == 0 && q < nq) {
    double s = 0.0;  // magnitude(): sequential f64, fnc/util/math/vector.rs:301-314
    for (uint32_t c = 0; c < dim; c++) {
      const double v = q64[(size_t)q * dim + c];
      s = __dadd_rn(s, __dmul_rn(v, v));
    }
    const double m = __dsqrt_rn(s);
    qmag[q] = m;
    uint32_t f = s_flags;
    if (metric == SDB_COSINE && (!(m > 0.0) || !isfinite(m))) f |= 1u;
    if (metric != SDB_COSINE && !isfinite(m)) f |= 1u;
    qflags[q] = f;
    // |q - bf16(q)| / |q|, rounded up; + 2^-23 for the f64 -> f32 rounding of the query itself
    if (qbferr) qbferr[q] = (m > 0.0 && isfinite(m)) ? (sqrtf(s_err2) / (float)m) * 1.0001f + 2.4e-7f : 1.f;
  }
}

// symmetric int8 quantisation of the queries (same scheme as the corpus rows), one block per query
__global__ void __launch_bounds__(128) prep_queries_i8_kernel(const float* __restrict__ q32, const double* __restrict__ qmag,
                                                              uint32_t dim, uint32_t dim_pad8, uint32_t nq,
                                                              int8_t* __restrict__ q8, float* __restrict__ q8scale,
                                                              float* __restrict__ q8err) {
  __shared__ float s_red[4];
  const uint32_t q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int8_t* o = q8 + (size_t)q * dim_pad8;
  if (q >= nq) {
    for (uint32_t c = threadIdx.x; c < dim_pad8; c += blockDim.x) o[c] = 0;
    return;
  }
  const float* x = q32 + (size_t)q * dim;
  float mx = 0.f;
  for (uint32_t c = threadIdx.x; c < dim; c += blockDim.x) mx = fmaxf(mx, fabsf(x[c]));
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  const bool ok = mx > 0.f && isfinite(mx);
  const float s = ok ? mx / 127.f : 1.f, inv = 1.f / s;
  float err2 = 0.f;
  for (uint32_t c = threadIdx.x; c < dim_pad8; c += blockDim.x) {
    int v = 0;
    if (ok && c < dim) {
      v = __float2int_rn(x[c] * inv);
      v = v > 127 ? 127 : (v < -127 ? -127 : v);
      const float d = x[c] - (float)v * s;
      err2 = fmaf(d, d, err2);
    }
    o[c] = (int8_t)v;
  }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, o2);
  if (lane == 0) s_red[warp] = err2;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float e = s_red[0] + s_red[1] + s_red[2] + s_red[3];
    q8scale[q] = s;
    // relative to the f64 norm; +2^-23 covers the f64 -> f32 rounding of the query itself; rounded up
    q8err[q] = ok ? (sqrtf(e) / (float)qmag[q]) * 1.0001f + 2.4e-7f : 1.f;
  }
}

// ------------------------------------------------------------------------------------------------
// start of a screen: tau = -inf, empty lists, and per query
//   beps   rigorous error bound of the screen that is about to run: cosine -> |sim~ - sim| <= beps;
//          euclid -> |score~ - score| <= beps with score = 2 q.x - |x|^2 (absolute)
//   bscale factor that turns a score into similarity * |q| units (1, or s_q * s for the int8 screen)
//   margin 2.1 x beps in SCORE units (0 in approximate mode).  The k rows with the best screened scores have exact
//          scores >= s_k - beps, so the exact k-th best is >= s_k - beps, and every row whose exact score can reach
//          that has a screened score >= s_k - 2 beps: filtering at tau = s_k - margin keeps all of them.
//   qlow / qcap  bounds of any score of this query (histogram geometry)
// Error bounds (both operands are rounded): bf16  e_x + e_q + e_x e_q  with the MEASURED residual norms
// e_x = max_rows |x - bf16(x)|/|x| (finalize) and e_q = |q - bf16(q)|/|q| (prep) -- at most 2^-8 each -- plus fp32
// accumulation D * 2^-21 and 1e-5 for the f32 screening norm; int8  (1 + e_q) e_x + e_q  (integer accumulation is
// exact); f32 SIMT (D/16 + 16) * 2^-23.
__global__ void cand_begin_kernel(float* __restrict__ tau, uint32_t* __restrict__ cnt, uint32_t* __restrict__ flags,
                                  uint32_t* __restrict__ stat, float* __restrict__ bscale, float* __restrict__ beps,
                                  float* __restrict__ margin, float* __restrict__ margin2, float* __restrict__ beps2,
                                  float* __restrict__ tau2, float* __restrict__ qlow, float* __restrict__ qcap,
                                  const double* __restrict__ qmag, const float* __restrict__ q8scale,
                                  const float* __restrict__ q8err, const float* __restrict__ qbferr, uint32_t nq,
                                  int screen, int metric, uint32_t dim, float max_rel_qerr, float i8_scale,
                                  float bf16_rel_err, float max_norm, int exact) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) {
    stat[0] = 0;
    stat[1] = 0;
    stat[2] = 0;
    stat[3] = 0;
  }
  if (q >= nq) return;
  tau[q] = __int_as_float(0xff800000);  // -inf
  cnt[q] = 0;
  flags[q] = 0;
  const double qm = qmag[q];
  double eps_rel, bs = 1.0;
  if (screen == SDB_SCREEN_TC_INT8) {
    const double eq = q8err[q];
    eps_rel = (1.0 + eq) * (double)max_rel_qerr + eq + 2e-6;
    bs = (double)q8scale[q] * (double)i8_scale * 1.000001;
  } else if (screen == SDB_SCREEN_TC_BF16) {
    const double eq = qbferr[q], ex = bf16_rel_err;
    eps_rel = ex + eq + ex * eq + dim * 4.76837158e-7 + 1e-5;
  } else {
    eps_rel = (dim / 16.0 + 16.0) * 1.1920929e-7;
  }
  double eps, mg, lo, hi;
  if (metric == SDB_COSINE) {
    eps = eps_rel;
    mg = 2.1 * eps * qm / bs;
    hi = qm / bs * (1.01 + eps);
    lo = -hi;
  } else {
    const double mn = (double)max_norm;
    eps = 2.0 * eps_rel * qm * mn + 4.8e-7 * (mn * mn + 2.0 * qm * mn) + 1e-30;
    mg = 2.1 * eps;
    hi = qm * qm * 1.01 + eps + 1e-30;
    lo = -(mn * mn + 2.0 * qm * mn) * 1.01 - eps - 1e-30;
  }
  if (!exact) mg = 0.0;
  if (!(qm > 0.0) || !isfinite(qm) || !isfinite(mg) || !isfinite(hi) || !isfinite(lo)) {  // exact path anyway (qflags)
    mg = 0.0;
    hi = 1.0;
    lo = -1.0;
  }
  // stage B (cand_refine: f32 re-score of the candidates with the f32 rows and the f32 query): any summation order of
  // D products in f32 with FMA stays within (D + 16) * 2^-24 of sum |q_i x_i| <= |q||x| (the +16 covers the f64 -> f32
  // rounding of the query, the f32 screening norm and the final scaling)
  const double e2_rel = (dim + 16.0) * 5.9604645e-8;
  double e2, mg2;
  if (metric == SDB_COSINE) {
    e2 = e2_rel;
    mg2 = 2.1 * e2 * qm;
  } else {
    const double mn = (double)max_norm;
    e2 = 2.0 * e2_rel * qm * mn + 2.4e-7 * mn * mn + 1e-30;
    mg2 = 2.1 * e2;
  }
  if (!exact || !(qm > 0.0) || !isfinite(qm) || !isfinite(mg2)) mg2 = 0.0;
  bscale[q] = (float)bs;
  beps[q] = __double2float_ru(eps);
  margin[q] = __double2float_ru(mg);
  margin2[q] = __double2float_ru(mg2);
  beps2[q] = __double2float_ru(e2);
  tau2[q] = __int_as_float(0xff800000);
  qlow[q] = __double2float_rd(lo);
  qcap[q] = __double2float_ru(hi);
}
sdb_status cand_begin(Corpus* c, uint32_t nq, int screen, cudaStream_t st) {
  cand_begin_kernel<<<(nq + 255) / 256, 256, 0, st>>>(c->d_tau, c->d_cand_cnt, c->d_flags, c->d_stat, c->d_bscale, c->d_beps,
                                                      c->d_margin, c->d_margin2, c->d_beps2, c->d_tau2, c->d_qlow, c->d_qcap,
                                                      c->d_qmag, c->d_q8scale,
                                                      c->d_q8err, c->d_qbferr, nq, screen, (int)c->metric, c->dim,
                                                      c->max_rel_qerr, c->i8_scale, c->bf16_rel_err, c->max_norm,
                                                      c->exact ? 1 : 0);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

static sdb_status ensure_scratch(Corpus* c, uint32_t nq, uint32_t cap) {
  const uint32_t nq_pad = (nq + 127) / 128 * 128;
  if (c->sc_nq >= nq_pad && c->sc_cap >= cap) return SDB_OK;
  void* old[] = {c->d_q64, c->d_q32, c->d_qbf16, c->d_qmag, c->d_qflags, c->d_qbferr, c->d_tau, c->d_cand, c->d_cand_cnt,
                 c->d_flags, c->d_stat, c->d_rr_key, c->d_rr_dist, c->d_rr_row, c->d_sub, c->d_sub_cnt, c->d_q8,
                 c->d_q8scale, c->d_q8err, c->d_bscale, c->d_beps, c->d_margin, c->d_qlow, c->d_qcap, c->d_hparam,
                 c->d_hist, c->d_probe, c->d_margin2, c->d_beps2, c->d_tau2};
  for (void* p : old) cudaFree(p);
  const uint32_t nqa = nq_pad > c->sc_nq ? nq_pad : c->sc_nq;
  const uint32_t capa = cap > c->sc_cap ? cap : c->sc_cap;
  c->sc_nq = c->sc_cap = 0;
  c->sc_gen++;  // everything below is reallocated: prepared queries, candidate lists ... are gone
  c->rr_stride = capa + SPECIAL_CAP;
  SDB_CUDA(cudaMalloc(&c->d_q64, sizeof(double) * (size_t)nqa * c->dim));
  SDB_CUDA(cudaMalloc(&c->d_q32, sizeof(float) * (size_t)nqa * c->dim));
  SDB_CUDA(cudaMalloc(&c->d_qbf16, sizeof(__nv_bfloat16) * (size_t)nqa * c->dim_pad));
  SDB_CUDA(cudaMalloc(&c->d_qmag, sizeof(double) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_qflags, sizeof(uint32_t) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_qbferr, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_tau, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_cand, sizeof(Cand) * (size_t)nqa * capa));
  SDB_CUDA(cudaMalloc(&c->d_cand_cnt, sizeof(uint32_t) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_flags, sizeof(uint32_t) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_stat, sizeof(uint32_t) * 4));
  SDB_CUDA(cudaMalloc(&c->d_rr_key, sizeof(uint64_t) * (size_t)nqa * c->rr_stride));
  SDB_CUDA(cudaMalloc(&c->d_rr_dist, sizeof(double) * (size_t)nqa * c->rr_stride));
  SDB_CUDA(cudaMalloc(&c->d_rr_row, sizeof(uint32_t) * (size_t)nqa * c->rr_stride));
  SDB_CUDA(cudaMalloc(&c->d_q8, (size_t)nqa * (c->dim_pad8 ? c->dim_pad8 : 128)));
  SDB_CUDA(cudaMalloc(&c->d_q8scale, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_q8err, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_bscale, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_beps, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_margin, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_margin2, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_beps2, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_tau2, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_qlow, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_qcap, sizeof(float) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_hparam, sizeof(HistParam) * nqa));
  SDB_CUDA(cudaMalloc(&c->d_hist, sizeof(uint32_t) * (size_t)nqa * HIST_BINS));
  SDB_CUDA(cudaMalloc(&c->d_probe, sizeof(float) * (size_t)nqa * PROBE_STRIDE));
  c->sub_slots = 2 * (uint32_t)c->ctx->sm_count;
  c->sub_cap = 16;
  SDB_CUDA(cudaMalloc(&c->d_sub, sizeof(Cand) * (size_t)nqa * c->sub_slots * c->sub_cap));
  SDB_CUDA(cudaMalloc(&c->d_sub_cnt, sizeof(uint32_t) * (size_t)nqa * c->sub_slots));
  c->sc_nq = nqa;
  c->sc_cap = capa;
  return SDB_OK;
}

sdb_status scratch_for(Corpus* c, uint32_t nq, uint32_t cap) { return ensure_scratch(c, nq, cap); }

sdb_status prep_queries(Corpus* c, const double* d_queries, uint32_t nq, cudaStream_t st) {
  // d_queries may alias c->d_q64
  if (d_queries != c->d_q64)
    SDB_CUDA(cudaMemcpyAsync(c->d_q64, d_queries, sizeof(double) * (size_t)nq * c->dim, cudaMemcpyDeviceToDevice, st));
  const uint32_t nq_pad = (nq + 127) / 128 * 128;
  prep_queries_kernel<<<nq_pad, 128, 0, st>>>(c->d_q64, c->dim, c->dim_pad, (int)c->metric, c->d_q32, c->d_qbf16,
                                              c->d_qmag, c->d_qflags, c->d_qbferr, nq);
  count_launch(c->ctx);
  if (c->d_i8) {
    prep_queries_i8_kernel<<<nq_pad, 128, 0, st>>>(c->d_q32, c->d_qmag, c->dim, c->dim_pad8, nq, c->d_q8, c->d_q8scale,
                                                   c->d_q8err);
    count_launch(c->ctx);
  }
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

sdb_status prep_fallback_query(Corpus* c, const double* d_query, cudaStream_t st) {
  if (!c->d_fb_q) {
    SDB_CUDA(cudaMalloc(&c->d_fb_q, sizeof(double) * c->dim));
    SDB_CUDA(cudaMalloc(&c->d_fb_qmag, sizeof(double)));
    SDB_CUDA(cudaMalloc(&c->d_fb_qflags, sizeof(uint32_t)));
  }
  if (d_query != c->d_fb_q)
    SDB_CUDA(cudaMemcpyAsync(c->d_fb_q, d_query, sizeof(double) * c->dim, cudaMemcpyDeviceToDevice, st));
  // the f32 / bf16 copies are not needed by the exact kernel: q32 goes to a throw-away row of the f64 buffer's size
  prep_queries_kernel<<<1, 128, 0, st>>>(c->d_fb_q, c->dim, c->dim, (int)c->metric, nullptr, nullptr, c->d_fb_qmag,
                                         c->d_fb_qflags, nullptr, 1);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void cand_set_count_kernel(uint32_t* cnt, uint32_t nq, uint32_t value) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) cnt[i] = value;
}
sdb_status cand_set_count(Corpus* c, uint32_t nq, uint32_t value, cudaStream_t st) {
  cand_set_count_kernel<<<(nq + 255) / 256, 256, 0, st>>>(c->d_cand_cnt, nq, value);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

constexpr uint32_t SEL_SMEM_KEYS = 1408;  // 11 KB: cand_select fits beside a resident screen CTA (15 KB are free)

__device__ __forceinline__ Cand key_to_cand(uint64_t key) {
  uint32_t fk = (uint32_t)(key >> 32);
  fk = (fk >> 31) ? (fk & 0x7fffffffu) : ~fk;
  Cand cd;
  cd.score = __uint_as_float(fk);
  cd.row = 0xFFFFFFFFu - (uint32_t)key;
  return cd;
}

// Seed of the streaming pass from a probe launch: the probe wrote, per query, the maximum score of every 32-row chunk
// of a few tiles.  The chunks are disjoint row sets, so the k-th largest chunk maximum s is reached by at least k
// different rows: the k-th best score of the corpus is >= s, and tau = s - margin is a valid first threshold.  Also
// sets up the query's histogram (geometry, zero counts) and empties its lists.  One block of 128 threads per query.
__global__ void __launch_bounds__(128) cand_seed_probe_kernel(const float* __restrict__ probe, uint32_t n_vals, uint32_t k,
                                                              const float* __restrict__ margin,
                                                              const float* __restrict__ qlow, const float* __restrict__ qcap,
                                                              float* __restrict__ tau, uint32_t* __restrict__ cnt,
                                                              HistParam* __restrict__ hparam, uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_key[PROBE_STRIDE];
  const uint32_t q = blockIdx.x;
  uint32_t p2 = 1;
  while (p2 < n_vals) p2 <<= 1;
  for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) {
    float v = __int_as_float(0xff800000);
    if (i < n_vals) v = probe[(size_t)q * PROBE_STRIDE + i];
    s_key[i] = (v == v) ? f32_key(v) : f32_key(__int_as_float(0xff800000));
  }
  __syncthreads();
  for (uint32_t kk = 2; kk <= p2; kk <<= 1)  // descending bitonic sort of <= 512 keys
    for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = s_key[i], b = s_key[ixj];
          const bool up = ((i & kk) == 0);
          if (up ? a < b : a > b) {
            s_key[i] = b;
            s_key[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  for (uint32_t i = threadIdx.x; i < HIST_BINS; i += blockDim.x) hist[(size_t)q * HIST_BINS + i] = 0;
  if (threadIdx.x == 0) {
    float t = __int_as_float(0xff800000);
    const float mg = margin[q];
    if (k != 0 && k <= n_vals) {
      uint32_t fk = s_key[k - 1];
      fk = (fk >> 31) ? (fk & 0x7fffffffu) : ~fk;
      const float s_k = __uint_as_float(fk);
      if (s_k > __int_as_float(0xff800000)) t = mg > 0.f ? __fsub_rd(s_k, mg) : s_k;
    }
    HistParam hp;
    hp.lo = t > __int_as_float(0xff800000) ? t : qlow[q];
    float w0 = fmaxf(mg * 0.25f, (qcap[q] - qlow[q]) * 6.1035156e-5f);
    if (!(w0 > 1e-30f) || !isfinite(w0)) w0 = 1e-30f;
    hp.w0 = w0;
    hp.inv_w0 = 1.f / w0;
    hp.margin = mg;
    hparam[q] = hp;
    tau[q] = t;
    cnt[q] = 0;
  }
}
sdb_status cand_seed_from_probe(Corpus* c, uint32_t nq, uint32_t k, uint32_t n_tiles, cudaStream_t st) {
  cand_seed_probe_kernel<<<nq, 128, 0, st>>>(c->d_probe, n_tiles * 8, k, c->d_margin, c->d_qlow, c->d_qcap, c->d_tau,
                                             c->d_cand_cnt, c->d_hparam, c->d_hist);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

// Per query: gather the main list and the thread-private sub-lists of the tensor-core screens, find the k-th best
// screened score s_k, and keep every candidate with score >= tau = s_k - margin (see cand_begin_kernel).  While fewer
// than k candidates exist everything is kept and tau stays where it is.  tau only ever rises: all rows with a score
// >= an earlier tau were appended under thresholds <= that tau, so the kept set always contains every row seen so far
// whose score reaches the current tau.
// seed: (re)build the query's histogram for the streaming pass -- geometry from (tau | the score range) and the margin,
// counts from the kept candidates.
__global__ void __launch_bounds__(256) cand_select_kernel(Cand* __restrict__ cand, uint32_t* __restrict__ cnt,
                                                           float* __restrict__ tau, uint32_t* __restrict__ flags,
                                                           uint32_t cap, uint32_t k, const float* __restrict__ margin,
                                                           const float* __restrict__ snorm, const Cand* __restrict__ sub,
                                                           const uint32_t* __restrict__ sub_cnt, uint32_t n_slots,
                                                           uint32_t subcap, HistParam* __restrict__ hparam,
                                                           uint32_t* __restrict__ hist, const float* __restrict__ qlow,
                                                           const float* __restrict__ qcap, uint32_t* __restrict__ stat,
                                                           uint64_t* __restrict__ g_keys, uint32_t g_stride) {
  // keys live in a small shared-memory window (the kernel must fit next to a resident screen CTA: 12 KB); a query that
  // gathered more -- a tight cluster -- sorts in its row of the (still unused) re-rank key buffer instead
  __shared__ uint64_t s_keys_win[SEL_SMEM_KEYS];
  __shared__ uint32_t s_upper;
  __shared__ uint32_t s_n, s_over;
  __shared__ uint32_t s_hist[256];
  __shared__ uint64_t s_prefix;
  __shared__ uint32_t s_remaining, s_out;
  const uint32_t q = blockIdx.x;
  uint32_t n_main = cnt[q];
  if (threadIdx.x == 0) {
    s_over = n_main > cap ? 1u : 0u;
    s_n = 0;
    s_prefix = 0;
    s_remaining = k;
    s_out = 0;
    s_upper = n_main > cap ? cap : n_main;
  }
  if (n_main > cap) n_main = cap;
  __syncthreads();
  {  // upper bound of what the gather can produce decides where the keys go (uniform per block)
    uint32_t part = 0;
    for (uint32_t s = threadIdx.x; s < n_slots; s += blockDim.x) {
      const uint32_t c = sub_cnt[(size_t)q * n_slots + s];
      part += c > subcap ? subcap : c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31u) == 0 && part) atomicAdd(&s_upper, part);
  }
  __syncthreads();
  uint64_t* s_keys = s_upper > SEL_SMEM_KEYS ? g_keys + (size_t)q * g_stride : s_keys_win;
  auto push = [&](Cand cd) {
    float sc = cd.score;
    if (snorm) {  // integer screens score invalid rows 0: the row's NaN screening norm marks them
      const float sn = snorm[cd.row];
      if (!(sn == sn)) sc = sn;
    }
    if (sc == sc) {  // NaN scores (skipped / special / padding rows) are dropped
      const uint32_t i = atomicAdd(&s_n, 1u);
      if (i < cap) s_keys[i] = ((uint64_t)f32_key(sc) << 32) | (uint64_t)(0xFFFFFFFFu - cd.row);
      else s_over = 1u;
    }
  };
  Cand* cq = cand + (size_t)q * cap;
  for (uint32_t i = threadIdx.x; i < n_main; i += blockDim.x) push(cq[i]);
  for (uint32_t s = threadIdx.x; s < n_slots; s += blockDim.x) {
    uint32_t c = sub_cnt[(size_t)q * n_slots + s];
    if (c > subcap) c = subcap;  // the rest went to the main list
    const Cand* sp = sub + ((size_t)q * n_slots + s) * subcap;
    for (uint32_t e = 0; e < c; e++) push(sp[e]);
  }
  __syncthreads();
  const uint32_t n = s_n < cap ? s_n : cap;
  if (threadIdx.x == 0 && s_over) flags[q] |= 1u;  // candidates were dropped: this query must be re-run exactly
  if (threadIdx.x == 0 && stat) atomicAdd(stat + 3, s_n);  // survivors gathered (diagnostics)
  const float tau_old = tau[q];
  float tau_new = tau_old;
  uint64_t kth = 0;
  const float mg = margin[q];
  if (k != 0 && n >= k) {
    // ---- exact selection of the k-th largest 64-bit key (keys are unique: the row is part of the key) by an MSB-first
    //      radix select, 8 bits per round.  Warp-aggregated histogram updates: the scores of one query share their
    //      leading bytes, so plain shared-memory atomics would serialise on one bin.
    const uint32_t lane = threadIdx.x & 31u;
    // with a margin only the k-th SCORE matters (the keep rule is score >= s_k - margin): 4 rounds over the score half
    // of the key; approximate mode keeps exactly k entries and needs the full (score, row) key: 8 rounds
    const uint32_t n_rounds = mg > 0.f ? 4u : 8u;
    for (uint32_t round = 0; round < n_rounds; round++) {
      const uint32_t shift = 56 - 8 * round;
      s_hist[threadIdx.x] = 0;
      __syncthreads();
      const uint64_t prefix = s_prefix;
      for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {  // uniform trip count: every lane reaches the match below
        const uint32_t i = i0 + threadIdx.x;
        uint32_t digit = 0xFFFFFFFFu;
        if (i < n) {
          const uint64_t key = s_keys[i];
          if (round == 0 || (key >> (shift + 8)) == prefix) digit = (uint32_t)(key >> shift) & 255u;
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (digit != 0xFFFFFFFFu && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&s_hist[digit], (uint32_t)__popc(peers));
      }
      __syncthreads();
      if (threadIdx.x < 32) {  // one warp: find the digit d with  #(digits > d) < remaining <= #(digits >= d)
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          c[j] = s_hist[255 - (lane * 8 + j)];  // lane 0 holds the 8 largest digits
          sum += c[j];
        }
        uint32_t incl = sum;  // inclusive prefix over lanes (descending digits)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= (uint32_t)o) incl += t;
        }
        const uint32_t rem = s_remaining;
        __syncwarp();  // every lane has read s_remaining before the selected lane overwrites it
        const uint32_t before = incl - sum;  // keys with a digit above this lane's range
        if (before < rem && rem <= incl) {
          uint32_t acc = before;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (acc < rem && rem <= acc + c[j]) {
              s_prefix = (prefix << 8) | (uint64_t)(255 - (lane * 8 + j));
              s_remaining = rem - acc;
            }
            acc += c[j];
          }
        }
      }
      __syncthreads();
    }
    kth = n_rounds == 8 ? s_prefix : (s_prefix << 32);  // the k-th largest key itself (score half only with a margin)
    const float s_k = key_to_cand(kth).score;
    const float thr = mg > 0.f ? __fsub_rd(s_k, mg) : s_k;
    tau_new = thr > tau_old ? thr : tau_old;  // (tau_old = -inf the first time)
  }
  // ---- keep: everything while no threshold exists; else score >= tau (approximate mode, margin 0: key >= k-th key) ----
  const bool by_key = (k != 0 && n >= k) && !(mg > 0.f) && tau_new == key_to_cand(kth).score;
  HistParam hp;
  if (hparam) {
    const float lo = tau_new > __int_as_float(0xff800000) ? tau_new : qlow[q];
    float w0 = fmaxf(mg * 0.25f, (qcap[q] - qlow[q]) * 6.1035156e-5f);
    if (!(w0 > 1e-30f) || !isfinite(w0)) w0 = 1e-30f;
    hp.lo = lo;
    hp.w0 = w0;
    hp.inv_w0 = 1.f / w0;
    hp.margin = mg;
    s_hist[threadIdx.x] = 0;
  }
  __syncthreads();
  // the kept entries overwrite the head of the main list: positions < n_main were all read during the gather above
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t key = s_keys[i];
    const Cand cd = key_to_cand(key);
    const bool keep = by_key ? key >= kth : cd.score >= tau_new;  // tau_new = -inf keeps everything
    if (keep) {
      cq[atomicAdd(&s_out, 1u)] = cd;
      if (hparam) atomicAdd(&s_hist[hist_bin(hp, cd.score)], 1u);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tau[q] = tau_new;
    cnt[q] = s_out;
    if (hparam) hparam[q] = hp;
  }
  if (hparam) hist[(size_t)q * HIST_BINS + threadIdx.x] = s_hist[threadIdx.x];
}

sdb_status cand_select(Corpus* c, uint32_t nq, uint32_t k, bool drop_invalid, uint32_t n_slots, bool seed_hist,
                       cudaStream_t st, int stage) {
  if (stage == 1) {  // stage B: the lists hold f32 scores now; own threshold / margin, nothing else to gather
    cand_select_kernel<<<nq, 256, 0, st>>>(c->d_cand, c->d_cand_cnt, c->d_tau2, c->d_flags, c->sc_cap, k, c->d_margin2,
                                           nullptr, c->d_sub, c->d_sub_cnt, 0u, c->sub_cap, nullptr, c->d_hist, c->d_qlow,
                                           c->d_qcap, nullptr, c->d_rr_key, c->rr_stride);
    count_launch(c->ctx);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
  }
  cand_select_kernel<<<nq, 256, 0, st>>>(  // 256 threads: several blocks per SM, the whole batch is one wave
      c->d_cand, c->d_cand_cnt, c->d_tau, c->d_flags, c->sc_cap, k, c->d_margin, drop_invalid ? c->d_snorm : nullptr,
      c->d_sub, c->d_sub_cnt, n_slots, c->sub_cap, seed_hist ? c->d_hparam : nullptr, c->d_hist, c->d_qlow, c->d_qcap,
      c->d_stat, c->d_rr_key, c->rr_stride);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

// ------------------------------------------------------------------------------------------------
// stage B: f32 re-score of the candidates.  FP64 is the scarce resource of this part (the sequential-f64 re-rank is
// bound by the FP64 pipe, not by memory), so the candidates of the coarse screen -- everything within the int8 / bf16
// error margin of the k-th best, ~100 rows per query on spread-out data, thousands inside a tight cluster -- are first
// re-scored with the f32 master rows and the f32 query on the FP32 pipe: one warp per row, row-contiguous LDG.128,
// shuffle reduction.  The error of that score is bounded rigorously (cand_begin_kernel: beps2), so the same rule
// "keep everything within 2.1 x the bound of the k-th best" (cand_select, stage 1) shrinks the set to k plus a few
// near-ties, and only those reach the f64 kernel.
template <bool COSINE>
__global__ void __launch_bounds__(128) cand_refine_f32_kernel(const float* __restrict__ rows, uint32_t dim,
                                                               const float* __restrict__ snorm,
                                                               const float* __restrict__ q32, Cand* __restrict__ cand,
                                                               const uint32_t* __restrict__ cnt, uint32_t cap) {
  const uint32_t q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_c = cnt[q] < cap ? cnt[q] : cap;
  const float* qv = q32 + (size_t)q * dim;
  Cand* cq = cand + (size_t)q * cap;
  const uint32_t warps_total = gridDim.y * 4;
  const bool vec4 = (dim & 3u) == 0;
  for (uint32_t e0 = (blockIdx.y * 4 + warp) * 2; e0 < n_c; e0 += warps_total * 2) {  // two rows per warp in flight
    const uint32_t r0 = cq[e0].row;
    const bool has1 = e0 + 1 < n_c;
    const uint32_t r1 = has1 ? cq[e0 + 1].row : r0;
    const float* x0 = rows + (size_t)r0 * dim;
    const float* x1 = rows + (size_t)r1 * dim;
    float a0 = 0.f, a1 = 0.f;
    if (vec4) {
      for (uint32_t c = lane * 4; c < dim; c += 128) {
        const float4 u = __ldg(reinterpret_cast<const float4*>(x0 + c));
        const float4 v = __ldg(reinterpret_cast<const float4*>(x1 + c));
        const float4 w = __ldg(reinterpret_cast<const float4*>(qv + c));
        a0 = fmaf(u.x, w.x, a0); a0 = fmaf(u.y, w.y, a0); a0 = fmaf(u.z, w.z, a0); a0 = fmaf(u.w, w.w, a0);
        a1 = fmaf(v.x, w.x, a1); a1 = fmaf(v.y, w.y, a1); a1 = fmaf(v.z, w.z, a1); a1 = fmaf(v.w, w.w, a1);
      }
    } else {
      for (uint32_t c = lane; c < dim; c += 32) {
        const float w = __ldg(qv + c);
        a0 = fmaf(__ldg(x0 + c), w, a0);
        a1 = fmaf(__ldg(x1 + c), w, a1);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (lane == 0) {
      const float s0 = __ldg(snorm + r0);
      cq[e0].score = COSINE ? a0 * s0 : fmaf(2.f, a0, -s0);
      if (has1) {
        const float s1 = __ldg(snorm + r1);
        cq[e0 + 1].score = COSINE ? a1 * s1 : fmaf(2.f, a1, -s1);
      }
    }
  }
}
sdb_status cand_refine(Corpus* c, uint32_t nq, cudaStream_t st) {
  if (c->dtype != SDB_F32) return SDB_OK;
  const dim3 grid(nq, 8);  // 128-thread blocks (register budget beside a resident screen CTA)
  if (c->metric == SDB_COSINE)
    cand_refine_f32_kernel<true><<<grid, 128, 0, st>>>((const float*)c->d_rows, c->dim, c->d_snorm, c->d_q32, c->d_cand,
                                                       c->d_cand_cnt, c->sc_cap);
  else
    cand_refine_f32_kernel<false><<<grid, 128, 0, st>>>((const float*)c->d_rows, c->dim, c->d_snorm, c->d_q32, c->d_cand,
                                                        c->d_cand_cnt, c->sc_cap);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

// ------------------------------------------------------------------------------------------------
// exact re-rank.  Work unit = (query, group of 128 entries): blocks (x = query, y = 0..3) stride over the groups of
// their query; each warp takes 32 entries (candidates first, then the special rows) and walks their rows 64 columns at
// a time: 64 coalesced 128-byte row segments are requested back to back (all in flight before the first is consumed --
// the kernel is bound by the latency of these gathers, not by the f64 arithmetic), transposed through shared memory,
// and every lane then accumulates ITS row strictly left to right in the reference's arithmetic.  Candidate counts
// vary per query by orders of magnitude (a handful ... a whole cluster).
constexpr uint32_t QCHUNK = 1024;  // query columns staged in shared memory per step

constexpr uint32_t RR_GROUPS_Y = 4;  // blocks per query; block y takes the groups y, y + 4, ... of its query

template <typename T, int WARPS, int COLS, int QC = QCHUNK>
__global__ void __launch_bounds__(WARPS * 32) cand_rerank_kernel(
    const T* __restrict__ rows, uint32_t dim, int metric, const double* __restrict__ mag,
    const double* __restrict__ q64, const double* __restrict__ qmag, const uint32_t* __restrict__ qflags,
    const Cand* __restrict__ cand, const uint32_t* __restrict__ cnt, uint32_t cap, const uint32_t* __restrict__ special,
    uint32_t n_special, uint64_t* __restrict__ rr_key, double* __restrict__ rr_dist, uint32_t* __restrict__ rr_row,
    uint32_t rr_stride) {
  constexpr int NL = COLS / 32;  // 128-byte segments per row and step
  __shared__ T tile[WARPS][32][COLS + 1];
  __shared__ double s_q[QC];
  const uint32_t q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_c = cnt[q] < cap ? cnt[q] : cap;
  const uint32_t n_e = n_c + n_special;
  const bool q_nan = (qflags[q] & 2u) != 0;
  const double qm = qmag[q];
  for (uint32_t e_first = blockIdx.y * (WARPS * 32); e_first < n_e; e_first += gridDim.y * (WARPS * 32)) {  // uniform per block
    const uint32_t e = e_first + warp * 32 + lane;
    uint32_t my_row = NO_ROW;
    if (e < n_c) my_row = cand[(size_t)q * cap + e].row;
    else if (e < n_e) my_row = special[e - n_c];
    const bool warp_active = e_first + warp * 32 < n_e;
    ExactAcc acc;
    for (uint32_t cb = 0; cb < dim; cb += QC) {
      const uint32_t cw = dim - cb < (uint32_t)QC ? dim - cb : (uint32_t)QC;
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < cw; i += blockDim.x) s_q[i] = q64[(size_t)q * dim + cb + i];
      __syncthreads();
      if (!warp_active) continue;
      const T* base = rows + cb;
      for (uint32_t c0 = 0; c0 < cw; c0 += COLS) {
        T vals[32 * NL];
#pragma unroll
        for (int r = 0; r < 32; r++) {  // 32 x NL independent 128-byte gathers in flight
          const uint32_t row = __shfl_sync(0xffffffffu, my_row, r);
#pragma unroll
          for (int h = 0; h < NL; h++) {
            const uint32_t c = c0 + h * 32 + lane;
            vals[r * NL + h] = (row != NO_ROW && c < cw) ? __ldg(base + (size_t)row * dim + c) : T(0);
          }
        }
#pragma unroll
        for (int r = 0; r < 32; r++)
#pragma unroll
          for (int h = 0; h < NL; h++) tile[warp][r][h * 32 + lane] = vals[r * NL + h];
        __syncwarp();
        if (my_row != NO_ROW) {
          const uint32_t lim = cw - c0 < (uint32_t)COLS ? cw - c0 : (uint32_t)COLS;
          if (metric == SDB_COSINE) {
            for (uint32_t j = 0; j < lim; j++) acc.cosine_step((double)tile[warp][lane][j], s_q[c0 + j]);
          } else {
            for (uint32_t j = 0; j < lim; j++) acc.euclid_step((double)tile[warp][lane][j], s_q[c0 + j]);
          }
        }
        __syncwarp();
      }
    }
    if (my_row != NO_ROW) {
      const double d = metric == SDB_COSINE ? cosine_finish(acc, mag[my_row], qm, q_nan) : euclid_finish(acc, q_nan);
      const size_t o = (size_t)q * rr_stride + e;
      rr_key[o] = dist_key(d);
      rr_dist[o] = d;
      rr_row[o] = my_row;
    }
  }
}

// f32 rows whose length is a multiple of 4 (16-byte aligned rows): the gathers are issued ROW-contiguously -- one
// LDG.128 per lane covers 512 consecutive bytes of ONE row, so DRAM sees 512-byte bursts per candidate row instead of
// 128-byte pieces of 32 different rows (random 128-byte gathers reach ~2 TB/s on HBM3e, page-friendly ones several
// times that).  The 32 x 128 tile is then walked per lane with conflict-free LDS.128 (row stride 132 floats).
constexpr uint32_t RRV_COLS = 128, RRV_STRIDE = RRV_COLS + 4;
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) cand_rerank_v4_kernel(
    const float* __restrict__ rows, uint32_t dim, int metric, const double* __restrict__ mag,
    const double* __restrict__ q64, const double* __restrict__ qmag, const uint32_t* __restrict__ qflags,
    const Cand* __restrict__ cand, const uint32_t* __restrict__ cnt, uint32_t cap, const uint32_t* __restrict__ special,
    uint32_t n_special, uint64_t* __restrict__ rr_key, double* __restrict__ rr_dist, uint32_t* __restrict__ rr_row,
    uint32_t rr_stride) {
  extern __shared__ __align__(16) uint8_t rr_smem[];
  double* s_q = reinterpret_cast<double*>(rr_smem);                                   // [QCHUNK]
  float* tiles = reinterpret_cast<float*>(rr_smem + sizeof(double) * QCHUNK);          // [WARPS][32][RRV_STRIDE]
  const uint32_t q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = tiles + (size_t)warp * 32 * RRV_STRIDE;
  const uint32_t n_c = cnt[q] < cap ? cnt[q] : cap;
  const uint32_t n_e = n_c + n_special;
  const bool q_nan = (qflags[q] & 2u) != 0;
  const double qm = qmag[q];
  for (uint32_t e_first = blockIdx.y * (WARPS * 32); e_first < n_e; e_first += gridDim.y * (WARPS * 32)) {  // uniform per block
    const uint32_t e = e_first + warp * 32 + lane;
    uint32_t my_row = NO_ROW;
    if (e < n_c) my_row = cand[(size_t)q * cap + e].row;
    else if (e < n_e) my_row = special[e - n_c];
    const bool warp_active = e_first + warp * 32 < n_e;
    ExactAcc acc;
    for (uint32_t cb = 0; cb < dim; cb += QCHUNK) {
      const uint32_t cw = dim - cb < QCHUNK ? dim - cb : QCHUNK;  // multiple of 4
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < cw; i += blockDim.x) s_q[i] = q64[(size_t)q * dim + cb + i];
      __syncthreads();
      if (!warp_active) continue;
      for (uint32_t c0 = 0; c0 < cw; c0 += RRV_COLS) {
        const uint32_t c = c0 + lane * 4;  // this lane's 4 columns of every row
#pragma unroll
        for (int half = 0; half < 2; half++) {
          float4 vals[16];
#pragma unroll
          for (int r = 0; r < 16; r++) {  // 16 independent 512-byte row bursts in flight
            const uint32_t row = __shfl_sync(0xffffffffu, my_row, half * 16 + r);
            vals[r] = (row != NO_ROW && c < cw) ? __ldg(reinterpret_cast<const float4*>(rows + (size_t)row * dim + cb + c))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int r = 0; r < 16; r++)
            *reinterpret_cast<float4*>(tile + (size_t)(half * 16 + r) * RRV_STRIDE + lane * 4) = vals[r];
        }
        __syncwarp();
        if (my_row != NO_ROW) {
          const uint32_t lim = cw - c0 < RRV_COLS ? cw - c0 : RRV_COLS;  // multiple of 4
          const float* mine = tile + (size_t)lane * RRV_STRIDE;
          if (metric == SDB_COSINE) {
            for (uint32_t j = 0; j < lim; j += 4) {
              const float4 v = *reinterpret_cast<const float4*>(mine + j);
              acc.cosine_step((double)v.x, s_q[c0 + j]);
              acc.cosine_step((double)v.y, s_q[c0 + j + 1]);
              acc.cosine_step((double)v.z, s_q[c0 + j + 2]);
              acc.cosine_step((double)v.w, s_q[c0 + j + 3]);
            }
          } else {
            for (uint32_t j = 0; j < lim; j += 4) {
              const float4 v = *reinterpret_cast<const float4*>(mine + j);
              acc.euclid_step((double)v.x, s_q[c0 + j]);
              acc.euclid_step((double)v.y, s_q[c0 + j + 1]);
              acc.euclid_step((double)v.z, s_q[c0 + j + 2]);
              acc.euclid_step((double)v.w, s_q[c0 + j + 3]);
            }
          }
        }
        __syncwarp();
      }
    }
    if (my_row != NO_ROW) {
      const double d = metric == SDB_COSINE ? cosine_finish(acc, mag[my_row], qm, q_nan) : euclid_finish(acc, q_nan);
      const size_t o = (size_t)q * rr_stride + e;
      rr_key[o] = dist_key(d);
      rr_dist[o] = d;
      rr_row[o] = my_row;
    }
  }
}
constexpr size_t RRV_SMEM = sizeof(double) * QCHUNK + sizeof(float) * 4 * 32 * RRV_STRIDE;  // 4 warps: 75.8 KB


// After the f32 stage a query keeps k candidates plus a few near-ties (about 10 at k = 10), and the exact re-rank is one
// strictly sequential f64 chain per candidate: FP64-issue-bound, nothing to gain from staging.  This variant uses NO
// shared memory and 16 lanes per query (two queries per warp), so that (a) twice as many chains share every FP64
// warp-instruction as with one query per warp, and (b) its blocks run beside the resident screen CTA of the next batch
// (the staged variant's 6 KB per warp let only two warps per SM in, and the re-rank took 0.58 ms instead of 0.1 ms
// whenever it overlapped a screen -- SDB_TRACE timeline, round 2).  Every lane streams its own row (16-byte loads; the
// second half of each 32-byte sector comes from L1) and reads the query from global memory (one address per half-warp).
template <bool COSINE>
__global__ void __launch_bounds__(128) cand_rerank_packed_kernel(
    const float* __restrict__ rows, uint32_t dim, const double* __restrict__ mag, const double* __restrict__ q64,
    const double* __restrict__ qmag, const uint32_t* __restrict__ qflags, const Cand* __restrict__ cand,
    const uint32_t* __restrict__ cnt, uint32_t cap, const uint32_t* __restrict__ special, uint32_t n_special, uint32_t nq,
    uint64_t* __restrict__ rr_key, double* __restrict__ rr_dist, uint32_t* __restrict__ rr_row, uint32_t rr_stride) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t q = (blockIdx.x * 4 + warp) * 2 + (lane >> 4), l16 = lane & 15u;
  if (q >= nq) return;
  const uint32_t n_c = cnt[q] < cap ? cnt[q] : cap;
  const uint32_t n_e = n_c + n_special;
  const bool q_nan = (qflags[q] & 2u) != 0;
  const double qm = qmag[q];
  const double* qv = q64 + (size_t)q * dim;
  const bool vec4 = (dim & 3u) == 0;
  for (uint32_t e = l16; e < n_e; e += 16) {
    const uint32_t my_row = e < n_c ? cand[(size_t)q * cap + e].row : special[e - n_c];
    const float* x = rows + (size_t)my_row * dim;
    ExactAcc acc;
    if (vec4) {
#pragma unroll 2
      for (uint32_t j = 0; j < dim; j += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + j));
        const double2 qa = __ldg(reinterpret_cast<const double2*>(qv + j));
        const double2 qb = __ldg(reinterpret_cast<const double2*>(qv + j + 2));
        if (COSINE) {
          acc.cosine_step((double)v.x, qa.x);
          acc.cosine_step((double)v.y, qa.y);
          acc.cosine_step((double)v.z, qb.x);
          acc.cosine_step((double)v.w, qb.y);
        } else {
          acc.euclid_step((double)v.x, qa.x);
          acc.euclid_step((double)v.y, qa.y);
          acc.euclid_step((double)v.z, qb.x);
          acc.euclid_step((double)v.w, qb.y);
        }
      }
    } else {
      for (uint32_t j = 0; j < dim; j++) {
        if (COSINE) acc.cosine_step((double)__ldg(x + j), __ldg(qv + j));
        else acc.euclid_step((double)__ldg(x + j), __ldg(qv + j));
      }
    }
    const double d = COSINE ? cosine_finish(acc, mag[my_row], qm, q_nan) : euclid_finish(acc, q_nan);
    const size_t o = (size_t)q * rr_stride + e;
    rr_key[o] = dist_key(d);
    rr_dist[o] = d;
    rr_row[o] = my_row;
  }
}

sdb_status cand_rerank(Corpus* c, uint32_t nq, cudaStream_t st, bool small_sets) {
  const dim3 grid(nq, RR_GROUPS_Y);
  static const bool no_v4 = getenv("SDB_RERANK_SCALAR") != nullptr;
  static const bool no_packed = getenv("SDB_RERANK_STAGED") != nullptr;
  if (small_sets && c->dtype == SDB_F32 && !no_packed && (c->metric == SDB_COSINE || c->metric == SDB_EUCLIDEAN)) {
    const unsigned g = (nq + 7) / 8;
    if (c->metric == SDB_COSINE)
      cand_rerank_packed_kernel<true><<<g, 128, 0, st>>>((const float*)c->d_rows, c->dim, c->d_mag, c->d_q64, c->d_qmag, c->d_qflags,
                                                        c->d_cand, c->d_cand_cnt, c->sc_cap, c->d_special, c->n_special, nq,
                                                        c->d_rr_key, c->d_rr_dist, c->d_rr_row, c->rr_stride);
    else
      cand_rerank_packed_kernel<false><<<g, 128, 0, st>>>((const float*)c->d_rows, c->dim, c->d_mag, c->d_q64, c->d_qmag, c->d_qflags,
                                                         c->d_cand, c->d_cand_cnt, c->sc_cap, c->d_special, c->n_special, nq,
                                                         c->d_rr_key, c->d_rr_dist, c->d_rr_row, c->rr_stride);
  } else if (small_sets && c->dtype == SDB_F32)
    // (kept for A/B: one warp per query, rows transposed through 6 KB of shared memory)
    cand_rerank_kernel<float, 1, 32, 256><<<grid, 32, 0, st>>>((const float*)c->d_rows, c->dim, (int)c->metric, c->d_mag,
                                                               c->d_q64, c->d_qmag, c->d_qflags, c->d_cand, c->d_cand_cnt,
                                                               c->sc_cap, c->d_special, c->n_special, c->d_rr_key,
                                                               c->d_rr_dist, c->d_rr_row, c->rr_stride);
  else if (c->dtype == SDB_F32 && c->dim % 4 == 0 && !no_v4)
    cand_rerank_v4_kernel<4><<<grid, 128, RRV_SMEM, st>>>((const float*)c->d_rows, c->dim, (int)c->metric, c->d_mag, c->d_q64,
                                                          c->d_qmag, c->d_qflags, c->d_cand, c->d_cand_cnt, c->sc_cap,
                                                          c->d_special, c->n_special, c->d_rr_key, c->d_rr_dist,
                                                          c->d_rr_row, c->rr_stride);
  else if (c->dtype == SDB_F32)
    cand_rerank_kernel<float, 4, 32><<<grid, 128, 0, st>>>((const float*)c->d_rows, c->dim, (int)c->metric, c->d_mag,
                                                           c->d_q64, c->d_qmag, c->d_qflags, c->d_cand, c->d_cand_cnt,
                                                           c->sc_cap, c->d_special, c->n_special, c->d_rr_key,
                                                           c->d_rr_dist, c->d_rr_row, c->rr_stride);
  else
    cand_rerank_kernel<double, 4, 32><<<grid, 128, 0, st>>>((const double*)c->d_rows, c->dim, (int)c->metric, c->d_mag,
                                                            c->d_q64, c->d_qmag, c->d_qflags, c->d_cand, c->d_cand_cnt,
                                                            c->sc_cap, c->d_special, c->n_special, c->d_rr_key,
                                                            c->d_rr_dist, c->d_rr_row, c->rr_stride);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

// ------------------------------------------------------------------------------------------------
// final ordering + proof.  Entries sorted ascending by (Number::cmp key, row) -- exactly the
// DistanceEntry order of KnnTopK (knn_topk.rs:61-73): nearest first, earlier scan position wins ties.
// The candidate count varies per query (a handful on spread-out data, thousands inside a tight cluster), so the sort
// works on a fixed 512-entry window: up to 512 entries are sorted in one go; longer lists are folded in chunks of
// 256 into the running best 256 (k <= 256 on the screened path).
constexpr uint32_t FIN_WIN = 512, FIN_KEEP = 256;  // 8 KB of shared memory, 256 threads: co-resident with a screen CTA

__device__ __forceinline__ void bitonic_pairs(uint64_t* s_key, uint64_t* s_idx, uint32_t p2) {
  for (uint32_t kk = 2; kk <= p2; kk <<= 1) {
    for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
          const uint64_t ka = s_key[i], kb = s_key[ixj], ia = s_idx[i], ib = s_idx[ixj];
          const bool a_gt_b = ka > kb || (ka == kb && ia > ib);
          const bool up = ((i & kk) == 0);
          if (up ? a_gt_b : !a_gt_b) {
            s_key[i] = kb; s_key[ixj] = ka;
            s_idx[i] = ib; s_idx[ixj] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) cand_final_kernel(
    const uint64_t* __restrict__ rr_key, const double* __restrict__ rr_dist, const uint32_t* __restrict__ rr_row,
    uint32_t rr_stride, const uint32_t* __restrict__ cnt, uint32_t cap, uint32_t n_special,
    const float* __restrict__ tau, const double* __restrict__ qmag, const float* __restrict__ bscale,
    const float* __restrict__ beps, const float* __restrict__ tau2, const float* __restrict__ beps2,
    uint32_t* __restrict__ flags, const uint32_t* __restrict__ qflags, uint32_t* __restrict__ stat, int metric,
    uint32_t k, uint64_t row_base, uint64_t* __restrict__ out_rows,
    double* __restrict__ out_dist, uint32_t* __restrict__ out_count, int debug) {
  __shared__ uint64_t s_key[FIN_WIN];  // distance key
  __shared__ uint64_t s_idx[FIN_WIN];  // (row << 32 | entry): secondary order by row (unique), entry = index into rr_*
  const uint32_t q = blockIdx.x;
  const uint32_t n_c = cnt[q] < cap ? cnt[q] : cap;
  const uint32_t n_e = n_c + n_special;
  const uint64_t* qkey = rr_key + (size_t)q * rr_stride;
  const uint32_t* qrow = rr_row + (size_t)q * rr_stride;
  if (n_e <= FIN_WIN) {
    uint32_t p2 = 1;
    while (p2 < n_e) p2 <<= 1;
    for (uint32_t i = threadIdx.x; i < p2; i += blockDim.x) {
      if (i < n_e) {
        s_key[i] = qkey[i];
        s_idx[i] = ((uint64_t)qrow[i] << 32) | i;
      } else {
        s_key[i] = ~0ull;
        s_idx[i] = ~0ull;
      }
    }
    __syncthreads();
    bitonic_pairs(s_key, s_idx, p2);
  } else {
    for (uint32_t i = threadIdx.x; i < FIN_KEEP; i += blockDim.x) {
      s_key[i] = ~0ull;
      s_idx[i] = ~0ull;
    }
    for (uint32_t c0 = 0; c0 < n_e; c0 += FIN_WIN - FIN_KEEP) {
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < FIN_WIN - FIN_KEEP; i += blockDim.x) {
        const uint32_t e = c0 + i;
        s_key[FIN_KEEP + i] = e < n_e ? qkey[e] : ~0ull;
        s_idx[FIN_KEEP + i] = e < n_e ? (((uint64_t)qrow[e] << 32) | e) : ~0ull;
      }
      __syncthreads();
      bitonic_pairs(s_key, s_idx, FIN_WIN);  // the best FIN_KEEP so far end up in front
    }
  }
  const uint32_t n_out = n_e < k ? n_e : k;
  for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
    const uint32_t e = (uint32_t)s_idx[i];
    out_rows[(size_t)q * k + i] = row_base + (uint64_t)(s_idx[i] >> 32);
    out_dist[(size_t)q * k + i] = rr_dist[(size_t)q * rr_stride + e];
  }
  if (threadIdx.x == 0) {
    out_count[q] = n_out;
    // ---- proof that no row outside the candidate set can enter the top-k ----
    uint32_t fl = flags[q];
    const float t = tau[q];
    if (t > __int_as_float(0xff800000) && n_e >= k && k > 0) {  // tau == -inf: every screened-in row is a candidate
      const double qm = qmag[q];
      const uint64_t kth = s_key[k - 1];
      bool ok;
      if (metric == SDB_COSINE) {
        // non-candidate: score <= tau  =>  sim <= tau * bscale / |q| + eps  =>  dist >= 1 - tau * bscale / |q| - eps
        const double bound = 1.0 - (double)t * (double)bscale[q] / qm - (double)beps[q] - 1e-9;
        ok = dist_key(bound) > kth;
      } else {
        // score = 2 dot~ - |x|^2~ <= tau  =>  d^2 = |x|^2 - 2 dot + |q|^2 >= -tau + |q|^2 - eps_e
        const double L = -(double)t + qm * qm - (double)beps[q];
        ok = L > 0.0 && dist_key(sqrt(L) * (1.0 - 1e-12)) > kth;
      }
      // stage B dropped candidates whose f32 score is below tau2: the same proof with the f32 bound
      const float t2 = tau2[q];
      if (ok && t2 > __int_as_float(0xff800000)) {
        if (metric == SDB_COSINE) {
          const double bound2 = 1.0 - (double)t2 / qm - (double)beps2[q] - 1e-9;
          ok = dist_key(bound2) > kth;
        } else {
          const double L2 = -(double)t2 + qm * qm - (double)beps2[q];
          ok = L2 > 0.0 && dist_key(sqrt(L2) * (1.0 - 1e-12)) > kth;
        }
      }
      if (!ok) fl |= 2u;
      if (debug && q == 0)
        printf("[sdb final] q0 metric=%d tau=%g qmag=%g beps=%g n_e=%u kth_key=%llx ok=%d\n", metric, (double)t, qm,
               (double)beps[q], n_e, (unsigned long long)kth, (int)ok);
    }
    if (fl & 1u) fl |= 2u;  // overflowed candidate buffer => exact re-run
    flags[q] = fl;
    if ((fl & 2u) || (qflags[q] & 1u)) atomicAdd(stat + 0, 1u);  // queries the host still has to repair
    atomicAdd(stat + 1, n_e);
    atomicMax(stat + 2, n_e);
  }
}

sdb_status cand_final(Corpus* c, uint32_t nq, uint32_t k, uint64_t row_base, uint64_t* d_out_rows, double* d_out_dist,
                      uint32_t* d_out_count, cudaStream_t st) {
  if (k > FIN_KEEP) {
    set_error("cand_final: k = %u exceeds the screened path's limit of %u", k, FIN_KEEP);
    return SDB_EINVAL;
  }
  static const int debug = getenv("SDB_DEBUG") != nullptr;
  cand_final_kernel<<<nq, 256, 0, st>>>(c->d_rr_key, c->d_rr_dist, c->d_rr_row, c->rr_stride, c->d_cand_cnt, c->sc_cap,
                                        c->n_special, c->d_tau, c->d_qmag, c->d_bscale, c->d_beps, c->d_tau2, c->d_beps2, c->d_flags, c->d_qflags, c->d_stat,
                                        (int)c->metric, k, row_base, d_out_rows, d_out_dist, d_out_count, debug);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

sdb_status candidates_init_device() {
  SDB_CUDA(cudaFuncSetAttribute(cand_rerank_v4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RRV_SMEM));
  return SDB_OK;
}

}  // namespace sdb
