// screen_simt.cu -- K1: streaming f32 screen for small query batches (<= 8 queries per pass over the corpus).
//
// HBM-bound by construction: every corpus row (dim * 4 bytes) is read exactly once per launch with
// coalesced 16-byte loads, QB dot products are accumulated per row in registers (queries live in shared
// memory), reduced with warp shuffles, scaled by the per-row screening norm and compared with the
// query's running threshold tau; survivors are appended to the query's candidate buffer.
// Algorithmic bytes per row: dim*4 + 4 (snorm).  Replaces the scan loop of KnnTopK::execute
// (exec/operators/knn_topk.rs:185-228) as the *screen*; exactness comes from candidates.cu.
#include "internal.cuh"

namespace sdb {

constexpr int SIMT_QB = 8;       // queries per launch
constexpr int SIMT_WARPS = 8;    // warps per block
constexpr int SIMT_ROWS_PER_ITER = 4;  // rows in flight per warp (memory-level parallelism)

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

template <int QB, bool VEC4>
__global__ void __launch_bounds__(SIMT_WARPS * 32) screen_simt_kernel(
    const float* __restrict__ rows, const float* __restrict__ snorm, uint32_t dim, uint64_t n_rows,
    const float* __restrict__ q32, uint32_t q0, uint32_t nqb, int metric, PassDesc pass,
    const float* __restrict__ tau, Cand* __restrict__ cand, uint32_t* __restrict__ cand_cnt, uint32_t cap) {
  extern __shared__ float s_q[];  // [QB][dim]
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < QB * dim; i += blockDim.x) {
    const uint32_t qi = i / dim;
    s_q[i] = qi < nqb ? q32[(size_t)(q0 + qi) * dim + (i - qi * dim)] : 0.f;
  }
  __syncthreads();
  float my_tau = __int_as_float(0x7f800000);  // +inf: lanes >= nqb never append
  if (lane < nqb) my_tau = tau[q0 + lane];

  for (uint32_t w = blockIdx.x; w < pass.count; w += gridDim.x) {
    const uint64_t tile_row0 = (uint64_t)pass_tile(pass, w) * TILE_ROWS;
    // warp handles rows tile_row0 + warp + SIMT_WARPS * j
    for (uint32_t j0 = 0; j0 < TILE_ROWS / SIMT_WARPS; j0 += SIMT_ROWS_PER_ITER) {
      float acc[SIMT_ROWS_PER_ITER][QB];
#pragma unroll
      for (int rr = 0; rr < SIMT_ROWS_PER_ITER; rr++)
#pragma unroll
        for (int qi = 0; qi < QB; qi++) acc[rr][qi] = 0.f;
      uint64_t rws[SIMT_ROWS_PER_ITER];
#pragma unroll
      for (int rr = 0; rr < SIMT_ROWS_PER_ITER; rr++) rws[rr] = tile_row0 + warp + (uint64_t)SIMT_WARPS * (j0 + rr);
      if (VEC4) {
        const uint32_t nv = dim >> 2;
        for (uint32_t v0 = 0; v0 < nv; v0 += 32 * 2) {
          float4 x[SIMT_ROWS_PER_ITER][2];
#pragma unroll
          for (int rr = 0; rr < SIMT_ROWS_PER_ITER; rr++)
#pragma unroll
            for (int u = 0; u < 2; u++) {
              const uint32_t v = v0 + u * 32 + lane;
              x[rr][u] = (rws[rr] < n_rows && v < nv)
                             ? ld_stream_f4(reinterpret_cast<const float4*>(rows + rws[rr] * dim) + v)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
          for (int u = 0; u < 2; u++) {
            const uint32_t v = v0 + u * 32 + lane;
            if (v < nv) {
#pragma unroll
              for (int qi = 0; qi < QB; qi++) {
                const float4 qv = reinterpret_cast<const float4*>(s_q + (size_t)qi * dim)[v];
#pragma unroll
                for (int rr = 0; rr < SIMT_ROWS_PER_ITER; rr++) {
                  acc[rr][qi] = fmaf(x[rr][u].x, qv.x, acc[rr][qi]);
                  acc[rr][qi] = fmaf(x[rr][u].y, qv.y, acc[rr][qi]);
                  acc[rr][qi] = fmaf(x[rr][u].z, qv.z, acc[rr][qi]);
                  acc[rr][qi] = fmaf(x[rr][u].w, qv.w, acc[rr][qi]);
                }
              }
            }
          }
        }
      } else {
        for (uint32_t c = lane; c < dim; c += 32) {
#pragma unroll
          for (int rr = 0; rr < SIMT_ROWS_PER_ITER; rr++) {
            const float x = rws[rr] < n_rows ? __ldg(rows + rws[rr] * dim + c) : 0.f;
#pragma unroll
            for (int qi = 0; qi < QB; qi++) acc[rr][qi] = fmaf(x, s_q[(size_t)qi * dim + c], acc[rr][qi]);
          }
        }
      }
      // warp reduction; afterwards lane qi keeps query qi's dot product
#pragma unroll
      for (int rr = 0; rr < SIMT_ROWS_PER_ITER; rr++) {
        float mine = 0.f;
#pragma unroll
        for (int qi = 0; qi < QB; qi++) {
          float v = acc[rr][qi];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if ((int)lane == qi) mine = v;
        }
        if (lane < nqb && rws[rr] < n_rows) {
          const float sn = __ldg(snorm + rws[rr]);
          const float s = metric == SDB_COSINE ? mine * sn : fmaf(2.f, mine, -sn);
          if (s >= my_tau) {  // NaN (skipped / special rows) never passes
            const uint32_t pos = atomicAdd(cand_cnt + q0 + lane, 1u);
            if (pos < cap) {
              Cand cd;
              cd.score = s;
              cd.row = (uint32_t)rws[rr];
              cand[(size_t)(q0 + lane) * cap + pos] = cd;
            }
          }
        }
      }
    }
  }
}

sdb_status screen_simt_pass(Corpus* c, uint32_t nq, const PassDesc& p, cudaStream_t st) {
  if (p.count == 0) return SDB_OK;
  Ctx* ctx = c->ctx;
  const size_t smem = sizeof(float) * SIMT_QB * c->dim;
  if (smem > 200 * 1024) {
    set_error("screen_simt: dim %u too large for the shared-memory query tile", c->dim);
    return SDB_EUNSUPPORTED;
  }
  const bool vec4 = (c->dim % 4 == 0);
  auto kern = vec4 ? screen_simt_kernel<SIMT_QB, true> : screen_simt_kernel<SIMT_QB, false>;
  SDB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  SDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SIMT_WARPS * 32, smem));
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = (uint32_t)(ctx->sm_count * per_sm);
  if (grid > p.count) grid = p.count;
  for (uint32_t q0 = 0; q0 < nq; q0 += SIMT_QB) {
    const uint32_t nqb = nq - q0 < (uint32_t)SIMT_QB ? nq - q0 : (uint32_t)SIMT_QB;
    kern<<<grid, SIMT_WARPS * 32, smem, st>>>((const float*)c->d_rows, c->d_snorm, c->dim, c->n, c->d_q32, q0, nqb,
                                              (int)c->metric, p, c->d_tau, c->d_cand, c->d_cand_cnt, c->sc_cap);
    count_launch(ctx);
  }
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // namespace sdb
