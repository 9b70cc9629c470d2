// micro-benchmark: TMEM -> register read throughput (tcgen05.ld 32x32b.x32) per SM, with 4 / 8 / 16 warps
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__global__ void k(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t s_tmem;
  const uint32_t warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_tmem + (((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t v[32];
      tmem_ld32(base + (c * 32) % 128, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; i++) acc ^= v[i];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_tmem) : "memory");
}
int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 8 * 256); cudaMalloc(&d_sink, 4 * 256 * 1024);
  for (int warps : {4, 8, 16}) {
    for (int grid : {1, 148}) {
      k<<<grid, warps * 32>>>(1000, d_out, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[256]; cudaMemcpy(h, d_out, 8 * grid, cudaMemcpyDeviceToHost);
      double bytes = 1000.0 * 4 * warps * 4096;
      printf("warps=%d grid=%d: %s cycles=%lld -> %.1f B/clk/SM (%.0f cycles per 128KB accumulator)\n", warps, grid,
             cudaGetErrorString(e), h[0], bytes / h[0], 131072.0 / (bytes / h[0]));
    }
  }
  return 0;
}
