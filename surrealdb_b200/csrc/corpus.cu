// corpus.cu -- device-resident vector column: per-row exact magnitudes, screening norms, bf16 screen
// copy, special-row list.  Data layout in HBM:
//   rows   [cap][dim]        f32|f64  master copy (exact re-rank reads it; the SIMT screen streams it)
//   bf16   [cap][dim_pad]    bf16     screen copy, K-major rows = tcgen05 "B" operand via TMA
//   mag    [cap]             f64      sqrt(sum x^2), the reference's `magnitude()` arithmetic
//   snorm  [cap]             f32      cosine: 1/|x|, euclid: |x|^2, NaN => row never screened in
#include <algorithm>

#include "internal.cuh"
#include "rowwalk.cuh"

namespace sdb {

// magnitude_squared: v.iter().map(|a| a.to_float().powi(2)).sum::<f64>()   fnc/util/math/vector.rs:301-303
template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) finalize_rows_kernel(const T* __restrict__ rows, uint32_t dim, uint64_t n,
                                                            int metric, const uint8_t* __restrict__ skip,
                                                            double* __restrict__ mag, float* __restrict__ snorm,
                                                            uint32_t* __restrict__ special, uint32_t* special_cnt,
                                                            uint32_t* max_norm_bits) {
  __shared__ T tile[WARPS][32][33];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t warps_total = (uint64_t)gridDim.x * WARPS;
  for (uint64_t base = ((uint64_t)blockIdx.x * WARPS + warp) * 32; base < n; base += warps_total * 32) {
    const uint64_t r = base + lane;
    const uint32_t my_row = r < n ? (uint32_t)r : NO_ROW;
    double s = 0.0;
    warp_walk_rows<T>(rows, dim, my_row, tile[warp], [&](uint32_t, T x) {
      const double xd = (double)x;
      s = __dadd_rn(s, __dmul_rn(xd, xd));
    });
    if (my_row == NO_ROW) continue;
    const double m = __dsqrt_rn(s);
    mag[r] = m;
    const bool skipped = skip && skip[r];
    float sn;
    bool is_special;
    if (metric == SDB_COSINE) {
      is_special = !(m > 0.0) || !isfinite(m);  // zero / NaN / inf norm: distance is NaN or needs exact care
      sn = (float)(1.0 / m);
    } else {
      is_special = !isfinite(s);
      sn = (float)s;
      if (!isfinite(sn)) is_special = true;  // |x|^2 overflows f32: rank exactly
    }
    if (skipped) {
      sn = __int_as_float(0x7fc00000);
    } else if (is_special) {
      sn = __int_as_float(0x7fc00000);
      const uint32_t pos = atomicAdd(special_cnt, 1u);
      if (pos < (uint32_t)SPECIAL_CAP) special[pos] = (uint32_t)r;
    } else {
      atomicMax(max_norm_bits, __float_as_uint((float)m) + 1u);  // +1 ulp: upper bound after rounding
    }
    snorm[r] = sn;
  }
}

// rows in [n, cap_pad) are TMA padding of the last screening tile: NaN norm => never a candidate
__global__ void pad_snorm_kernel(float* __restrict__ snorm, uint64_t n, uint64_t n_pad) {
  const uint64_t i = n + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) snorm[i] = __int_as_float(0x7fc00000);
}

// bf16 screen copy, one warp per row; also measures e_x = max over valid rows of |x - bf16(x)| / |x| (the residual
// norm that enters the screen's error bound; at most 2^-8 by construction of round-to-nearest, usually ~0.6 of that)
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ rows, uint32_t dim, uint32_t dim_pad, uint64_t n,
                                                      uint64_t n_pad, const double* __restrict__ mag,
                                                      const float* __restrict__ snorm, __nv_bfloat16* __restrict__ out,
                                                      uint32_t* max_rel_bits) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * 8;
  for (uint64_t r = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_pad; r += warps) {
    __nv_bfloat16* o = out + r * dim_pad;
    float err2 = 0.f;
    for (uint32_t c = lane; c < dim_pad; c += 32) {
      const float v = (r < n && c < dim) ? rows[r * dim + c] : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      o[c] = h;
      const float d = v - __bfloat162float(h);
      if (d == d) err2 = fmaf(d, d, err2);
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, o2);
    if (lane == 0 && r < n) {
      const float sn = snorm[r];
      const double m = mag[r];
      if (sn == sn && m > 0.0 && isfinite(m) && isfinite(err2))  // skipped / special rows never reach the screen
        atomicMax(max_rel_bits, __float_as_uint((sqrtf(err2) / (float)m) * 1.0001f + 1e-9f));
    }
  }
}

// int8 screen copy (cosine): the NORMALISED rows x/|x| are quantised with ONE global scale s = gmax/127, so the
// integer dot product q8.x8 is itself the screening score (no per-row weight in the epilogue):
//   sim(q,x) = s_q * s * (q8.x8) / |q|  +  err,   |err| <= (1 + e_q) * e_x + e_q,   e_x = |x/|x| - s*x8|  (per row)
// pass 1: gmax = max over valid rows of max_i |x_i| / |x|
// It also keeps every row's figure (rmax[r]) and a 4096-bin histogram of them over the float bit pattern (8 exponent
// + 4 mantissa bits), from which the host picks the scale: a handful of outlier rows (one dominant component) must not
// dictate the quantisation step of the other ten million.
constexpr uint32_t RMAX_BINS = 4096;
__device__ __host__ inline uint32_t rmax_bin(float v) {  // v > 0
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(v);
#else
  memcpy(&u, &v, 4);
#endif
  return (u >> 19) & (RMAX_BINS - 1);
}
__global__ void __launch_bounds__(256) quantize_scan_kernel(const float* __restrict__ rows, uint32_t dim, uint64_t n,
                                                            const double* __restrict__ mag, const float* __restrict__ snorm,
                                                            uint32_t* gmax_bits, float* __restrict__ rmax,
                                                            uint32_t* __restrict__ hist) {
  __shared__ uint32_t s_hist[RMAX_BINS];
  for (uint32_t i = threadIdx.x; i < RMAX_BINS; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * 8;
  float best = 0.f;
  for (uint64_t r = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += warps) {
    const float sn = snorm[r];
    if (!(sn == sn)) {  // skipped / special
      if (lane == 0) rmax[r] = 0.f;
      continue;
    }
    const float* x = rows + r * dim;
    float mx = 0.f;
    for (uint32_t c = lane; c < dim; c += 32) mx = fmaxf(mx, fabsf(x[c]));
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
    const float v = mx / (float)mag[r];
    best = fmaxf(best, v);
    if (lane == 0) {
      rmax[r] = v;
      if (v > 0.f && isfinite(v)) atomicAdd(&s_hist[rmax_bin(v)], 1u);
    }
  }
  if (lane == 0 && best > 0.f) atomicMax(gmax_bits, __float_as_uint(best));
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < RMAX_BINS; i += blockDim.x)
    if (s_hist[i]) atomicAdd(hist + i, s_hist[i]);
}
// rows whose largest normalised component reaches `thr` become special rows (ranked exactly on every query, never a
// screen candidate): the int8 scale is then set by the remaining rows
__global__ void mark_outliers_kernel(const float* __restrict__ rmax, uint64_t n, float thr, float* __restrict__ snorm,
                                     uint32_t* __restrict__ special, uint32_t* special_cnt) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (rmax[r] >= thr) {
    const float sn = snorm[r];
    if (sn == sn) {
      const uint32_t pos = atomicAdd(special_cnt, 1u);
      if (pos < (uint32_t)SPECIAL_CAP) special[pos] = (uint32_t)r;
      snorm[r] = __int_as_float(0x7fc00000);
    }
  }
}
// pass 2: x8 = clamp(rn(x / (|x| s)), +-127), e_x accumulated exactly as the residual norm (clipping included)
__global__ void __launch_bounds__(256) quantize_rows_kernel(const float* __restrict__ rows, uint32_t dim, uint32_t dim_pad8,
                                                            uint64_t n, uint64_t n_pad, const double* __restrict__ mag,
                                                            const float* __restrict__ snorm, const uint32_t* gmax_bits,
                                                            int8_t* __restrict__ out, uint32_t* max_rel_bits) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warps = (uint64_t)gridDim.x * 8;
  const float gmax = __uint_as_float(*gmax_bits);
  const float s = gmax > 0.f ? gmax / 127.f : 1.f;
  for (uint64_t r = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_pad; r += warps) {
    int8_t* o = out + r * dim_pad8;
    const float sn = r < n ? snorm[r] : __int_as_float(0x7fc00000);
    const bool ok = (sn == sn) && gmax > 0.f;
    if (!ok) {  // zero row: scores 0, and compaction drops it through its NaN screening norm
      for (uint32_t c = lane; c < dim_pad8; c += 32) o[c] = 0;
      continue;
    }
    const float* x = rows + r * dim;
    const float inv_norm = 1.f / (float)mag[r];
    const float inv = 1.f / s;
    float err2 = 0.f;
    for (uint32_t c = lane; c < dim_pad8; c += 32) {
      int q = 0;
      if (c < dim) {
        const float xn = x[c] * inv_norm;
        q = __float2int_rn(xn * inv);
        q = q > 127 ? 127 : (q < -127 ? -127 : q);
        const float d = xn - (float)q * s;
        err2 = fmaf(d, d, err2);
      }
      o[c] = (int8_t)q;
    }
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, o2);
    // + 2^-22: the f32 normalisation x * (1/|x|) is itself rounded; the whole figure is rounded up
    if (lane == 0) atomicMax(max_rel_bits, __float_as_uint(sqrtf(err2) * 1.0001f + 5e-7f));
  }
}

// tombstones (sdb_corpus_remove): the row is skipped by every path from now on -- skip mask for the exact kernel, NaN
// screening norm for the screens and the re-rank's special list, and an all-zero int8 row so that the integer screen
// scores it exactly 0 (the only score for which that screen looks up a row's validity).  No re-finalize needed.
__global__ void or_mask_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && src[i]) dst[i] = 1;
}
sdb_status corpus_reapply_tombstones(Corpus* c, cudaStream_t st) {  // after the caller replaced the skip mask
  if (!c->d_removed || !c->n) return SDB_OK;
  if (!c->d_skip) {
    SDB_CUDA(cudaMalloc(&c->d_skip, c->cap));
    SDB_CUDA(cudaMemsetAsync(c->d_skip, 0, c->cap, st));
  }
  or_mask_kernel<<<(unsigned)((c->n + 255) / 256), 256, 0, st>>>(c->d_skip, c->d_removed, c->n);
  count_launch(c->ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}
__global__ void remove_rows_kernel(const uint64_t* __restrict__ ids, uint64_t n, uint8_t* __restrict__ skip,
                                   uint8_t* __restrict__ removed, float* __restrict__ snorm, int8_t* __restrict__ i8,
                                   uint32_t dim_pad8) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n) return;
  const uint64_t r = ids[w];
  if (lane == 0) {
    skip[r] = 1;
    removed[r] = 1;
    if (snorm) snorm[r] = __int_as_float(0x7fc00000);
  }
  if (i8)
    for (uint32_t c = lane; c < dim_pad8; c += 32) i8[r * dim_pad8 + c] = 0;
}
sdb_status corpus_remove_device(Corpus* c, const uint64_t* h_ids, uint64_t n) {
  Ctx* ctx = c->ctx;
  cudaStream_t st = ctx->stream;
  if (!c->d_skip) {
    SDB_CUDA(cudaMalloc(&c->d_skip, c->cap));
    SDB_CUDA(cudaMemsetAsync(c->d_skip, 0, c->cap, st));
  }
  if (!c->d_removed) {
    SDB_CUDA(cudaMalloc(&c->d_removed, c->cap));
    SDB_CUDA(cudaMemsetAsync(c->d_removed, 0, c->cap, st));
  }
  uint64_t* d_ids = nullptr;
  SDB_CUDA(cudaMallocAsync(&d_ids, sizeof(uint64_t) * n, st));
  SDB_CUDA(cudaMemcpyAsync(d_ids, h_ids, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st));
  remove_rows_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(d_ids, n, c->d_skip, c->d_removed, c->finalized ? c->d_snorm : nullptr,
                                                                       c->finalized ? c->d_i8 : nullptr, c->dim_pad8);
  count_launch(ctx);
  SDB_CUDA(cudaFreeAsync(d_ids, st));
  if (c->finalized && c->n_special) {  // a removed special row leaves the always-exact list
    std::vector<uint32_t> sp(c->n_special);
    SDB_CUDA(cudaMemcpyAsync(sp.data(), c->d_special, sizeof(uint32_t) * c->n_special, cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaStreamSynchronize(st));
    std::vector<uint64_t> sorted(h_ids, h_ids + n);
    std::sort(sorted.begin(), sorted.end());
    std::vector<uint32_t> keep;
    for (uint32_t r : sp)
      if (!std::binary_search(sorted.begin(), sorted.end(), (uint64_t)r)) keep.push_back(r);
    if (keep.size() != sp.size()) {
      if (!keep.empty())
        SDB_CUDA(cudaMemcpyAsync(c->d_special, keep.data(), sizeof(uint32_t) * keep.size(), cudaMemcpyHostToDevice, st));
      c->n_special = (uint32_t)keep.size();
    }
  }
  SDB_CUDA(cudaStreamSynchronize(st));
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

sdb_status corpus_finalize_device(Corpus* c) {
  Ctx* ctx = c->ctx;
  cudaStream_t st = ctx->stream;
  uint32_t* d_tmp = nullptr;  // [0] special count, [1] max-norm bits, [2] max relative int8 error bits, [3] int8 gmax
  SDB_CUDA(cudaMalloc(&d_tmp, 32));                                       // [4] max relative bf16 residual bits
  SDB_CUDA(cudaMemsetAsync(d_tmp, 0, 32, st));
  if (!c->d_special) SDB_CUDA(cudaMalloc(&c->d_special, sizeof(uint32_t) * SPECIAL_CAP));
  {
    const uint64_t n_pad = (c->n + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
    if (n_pad > c->n) {
      pad_snorm_kernel<<<(unsigned)((n_pad - c->n + 255) / 256), 256, 0, st>>>(c->d_snorm, c->n, n_pad);
      count_launch(ctx);
    }
  }
  if (c->n) {
    const int grid = ctx->sm_count * 8;
    if (c->dtype == SDB_F32)
      finalize_rows_kernel<float, 8><<<grid, 256, 0, st>>>((const float*)c->d_rows, c->dim, c->n, (int)c->metric,
                                                        c->d_skip, c->d_mag, c->d_snorm, c->d_special, d_tmp,
                                                        d_tmp + 1);
    else
      finalize_rows_kernel<double, 4><<<grid * 2, 128, 0, st>>>((const double*)c->d_rows, c->dim, c->n, (int)c->metric,
                                                         c->d_skip, c->d_mag, c->d_snorm, c->d_special, d_tmp,
                                                         d_tmp + 1);
    count_launch(ctx);
    SDB_CUDA(cudaGetLastError());
    if (c->dtype == SDB_F32 && c->d_bf16) {
      const uint64_t n_pad = (c->n + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
      to_bf16_kernel<<<ctx->sm_count * 16, 256, 0, st>>>((const float*)c->d_rows, c->dim, c->dim_pad, c->n, n_pad,
                                                         c->d_mag, c->d_snorm, c->d_bf16, d_tmp + 4);
      count_launch(ctx);
      SDB_CUDA(cudaGetLastError());
    }
  }
  c->n_outliers = 0;
  if (c->n && c->dtype == SDB_F32 && c->d_i8 && c->metric == SDB_COSINE) {
    const uint64_t n_pad = (c->n + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
    float* d_rmax = nullptr;
    uint32_t* d_hist = nullptr;
    SDB_CUDA(cudaMallocAsync(&d_rmax, sizeof(float) * c->n, st));
    SDB_CUDA(cudaMallocAsync(&d_hist, sizeof(uint32_t) * RMAX_BINS, st));
    SDB_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * RMAX_BINS, st));
    quantize_scan_kernel<<<ctx->sm_count * 8, 256, 0, st>>>((const float*)c->d_rows, c->dim, c->n, c->d_mag, c->d_snorm,
                                                            d_tmp + 3, d_rmax, d_hist);
    count_launch(ctx);
    // ---- pick the scale: if at most 64 rows sit far above the rest (their largest normalised component is more
    //      than 1.5 x that of the 65th), make THEM special rows and quantise for the others ----
    std::vector<uint32_t> h_hist(RMAX_BINS);
    uint32_t h4[4] = {0, 0, 0, 0};
    SDB_CUDA(cudaMemcpyAsync(h_hist.data(), d_hist, sizeof(uint32_t) * RMAX_BINS, cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaMemcpyAsync(h4, d_tmp, 16, cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaStreamSynchronize(st));
    float gmax;
    memcpy(&gmax, &h4[3], 4);
    if (gmax > 0.f && !getenv("SDB_NO_OUTLIER_ROWS")) {
      // walk the occupied bins from the top while at most 64 rows lie above; an outlier group ends at a GAP: the next
      // occupied bin below starts at less than 2/3 of the group's lowest bin.  The lowest such gap wins (largest
      // group of rows that are all clearly detached from the bulk); without a gap nothing is special.
      uint32_t cum = 0;
      int cut = -1, below = -1;
      int last = -1;  // lowest occupied bin seen so far
      for (int b = (int)RMAX_BINS - 1; b >= 0; b--) {
        if (!h_hist[b]) continue;
        if (last >= 0 && cum <= 64u) {
          // edges: bin x covers [edge(x), edge(x+1)); gap test between the top of bin b and the bottom of bin `last`
          uint32_t ul = (uint32_t)last << 19, ub = (uint32_t)(b + 1) << 19;
          float lo_last, hi_b;
          memcpy(&lo_last, &ul, 4);
          memcpy(&hi_b, &ub, 4);
          if (lo_last > 1.5f * hi_b) {
            cut = last;
            below = b;
          }
        }
        cum += h_hist[b];
        if (cum > 64u) break;
        last = b;
      }
      if (cut >= 0) {
        uint32_t n_out = 0;
        for (int b = cut; b < (int)RMAX_BINS; b++) n_out += h_hist[b];
        if (n_out > 0 && h4[0] + n_out <= (uint32_t)SPECIAL_CAP) {
          const uint32_t ut = (uint32_t)cut << 19, us = (uint32_t)(below + 1) << 19;
          float thr, new_gmax;
          memcpy(&thr, &ut, 4);        // rows with rmax >= thr are the outliers
          memcpy(&new_gmax, &us, 4);   // every remaining row has rmax < this: the scale of the int8 copy
          mark_outliers_kernel<<<(unsigned)((c->n + 255) / 256), 256, 0, st>>>(d_rmax, c->n, thr, c->d_snorm, c->d_special, d_tmp);
          count_launch(ctx);
          SDB_CUDA(cudaMemcpyAsync(d_tmp + 3, &new_gmax, 4, cudaMemcpyHostToDevice, st));
          SDB_CUDA(cudaStreamSynchronize(st));  // `new_gmax` lives on this stack frame
          c->n_outliers = n_out;
        }
      }
    }
    quantize_rows_kernel<<<ctx->sm_count * 8, 256, 0, st>>>((const float*)c->d_rows, c->dim, c->dim_pad8, c->n, n_pad,
                                                            c->d_mag, c->d_snorm, d_tmp + 3, c->d_i8, d_tmp + 2);
    count_launch(ctx);
    SDB_CUDA(cudaFreeAsync(d_rmax, st));
    SDB_CUDA(cudaFreeAsync(d_hist, st));
    SDB_CUDA(cudaGetLastError());
  }
  uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  SDB_CUDA(cudaMemcpyAsync(h, d_tmp, 32, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaStreamSynchronize(st));
  SDB_CUDA(cudaFree(d_tmp));
  c->special_overflow = h[0] > (uint32_t)SPECIAL_CAP;
  c->n_special = h[0] > (uint32_t)SPECIAL_CAP ? (uint32_t)SPECIAL_CAP : h[0];
  float mn;
  memcpy(&mn, &h[1], 4);
  c->max_norm = mn;
  memcpy(&c->max_rel_qerr, &h[2], 4);
  {
    float e;
    memcpy(&e, &h[4], 4);
    // never trust a figure above the analytic worst case of round-to-nearest (2^-8 per element => 2^-8 in norm)
    c->bf16_rel_err = (e > 0.f && e < 0.00390625f) ? e : 0.00390625f;
  }
  {
    float gmax;
    memcpy(&gmax, &h[3], 4);
    c->i8_scale = gmax > 0.f ? gmax / 127.f : 1.f;
  }
  c->finalized = true;
  return SDB_OK;
}

}  // namespace sdb
