// gen.cu -- counter-based synthetic data generator (bench / test inputs only).
// Bit-for-bit the same function as orc_gen_f32 in oracle/sdb_oracle.c, so the CPU oracle and any number
// of GPUs see identical corpora without moving 30 GB over PCIe.
#include "internal.cuh"

namespace sdb {

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline float gen_value(uint64_t seed, uint64_t index) {
  uint64_t h = mix64(seed * 0x9E3779B97F4A7C15ull + index + 0x632BE59BD9B4E019ull);
  uint32_t m = (uint32_t)(h >> 40);
  return (float)m * (1.0f / 8388608.0f) - 1.0f;
}

__global__ void gen_fill_kernel(float* __restrict__ out, uint64_t seed, uint64_t first, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) out[i] = gen_value(seed, first + i);
}

sdb_status gen_fill_f32(Ctx* ctx, float* d_out, uint64_t seed, uint64_t first, uint64_t n, cudaStream_t st) {
  if (n == 0) return SDB_OK;
  int grid = ctx->sm_count * 16;
  gen_fill_kernel<<<grid, 256, 0, st>>>(d_out, seed, first, n);
  count_launch(ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // namespace sdb
