// rowwalk.cuh -- warp-cooperative sequential walk over up to 32 rows (one row per lane).
// Loads are coalesced (each row chunk of 32 elements is one 128/256-byte request) and transposed through
// shared memory so that every lane then visits ITS row's elements strictly in column order -- which is
// what the reference's sequential f64 sums require (fnc/util/math/vector.rs:279-303).
#pragma once
#include <cstdint>

namespace sdb {

constexpr uint32_t NO_ROW = 0xFFFFFFFFu;

// tile: per-warp shared scratch T[32][33].  f(col, value) is called for col = 0..dim-1 in order, by the
// lane owning a valid row.  All 32 lanes must call (convergent).
template <typename T, typename F>
__device__ __forceinline__ void warp_walk_rows(const T* __restrict__ base, uint32_t dim, uint32_t my_row,
                                               T (*tile)[33], F&& f) {
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t c0 = 0; c0 < dim; c0 += 32) {
    const uint32_t c = c0 + lane;
#pragma unroll 8
    for (int r = 0; r < 32; r++) {
      const uint32_t row = __shfl_sync(0xffffffffu, my_row, r);
      T v = T(0);
      if (row != NO_ROW && c < dim) v = __ldg(base + (size_t)row * dim + c);
      tile[r][lane] = v;
    }
    __syncwarp();
    if (my_row != NO_ROW) {
      const uint32_t lim = (dim - c0 < 32u) ? dim - c0 : 32u;
      for (uint32_t j = 0; j < lim; j++) f(c0 + j, tile[lane][j]);
    }
    __syncwarp();
  }
}

}  // namespace sdb
