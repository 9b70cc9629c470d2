// screen_tc.cu -- K2: tcgen05 bf16 GEMM screen with a fused threshold-filter epilogue (sm_100a only).
//
//   scores[q][x] = <q~, x~>  (bf16 operands, fp32 accumulation in TMEM), q = queries (M), x = corpus rows (N)
//
// One CTA = one SM, persistent over work items (corpus tile of 256 rows) x (block of 128 queries):
//   warp 0      TMA producer : cp.async.bulk.tensor 2-D tiles (SWIZZLE_128B) of A (128 x 64) and B (256 x 64)
//                              into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16)
//                              accumulating into one of two 256-column TMEM stages; tcgen05.commit frees the
//                              smem stage / publishes the accumulator
//   warps 2..9  epilogue     : tcgen05.ld the accumulator (thread = query, 256 columns = corpus rows), scale by
//                              the row's screening norm, compare with the query's threshold tau and append the
//                              rare survivors to the query's candidate list -- the 128 x 256 score tile is never
//                              written to memory (at 1024 x 10M it would be 41 GB).
//   warp 10     refiner      : (streaming mode) raises the thresholds WHILE the kernel runs.  Every survivor is also
//                              counted (fire-and-forget RED) in its query's 256-bin score histogram in L2; the refiner
//                              of CTA b owns the queries q = b (mod grid), reads their histograms, finds the bin in
//                              which the count from the top reaches k -- a proof that the k-th best score seen so far
//                              is at least that bin's lower edge -- and publishes tau = edge - margin.  The epilogue
//                              threads re-read tau (ld.cg, L2) at every work item.  One launch therefore covers the
//                              whole corpus: no per-pass launches, compactions or host round trips.
// The epilogue of tile i overlaps the MMAs of tile i+1 through the two TMEM stages.
//
// Replaces, as the *screen*, the distance loop of KnnTopK::execute (exec/operators/knn_topk.rs:185-228);
// exactness is restored by candidates.cu (f64 re-rank + error-bound proof) and exact.cu.
#include <cuda.h>

#include "internal.cuh"

namespace sdb {

namespace tc {
constexpr uint32_t BLOCK_M = 128;   // queries per work item
constexpr uint32_t BLOCK_N = 256;   // corpus rows per work item (= TILE_ROWS)
constexpr uint32_t BLOCK_K = 64;    // bf16 elements per smem stage row = 128 bytes = one swizzle atom
constexpr uint32_t UMMA_K = 16;
constexpr uint32_t STAGES = 4;
constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t ACC_STAGES = 2;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t EPI_WARPS = 8;             // two warps per TMEM lane quarter, each takes half the columns
constexpr uint32_t REFINE_WARP = 2 + EPI_WARPS;  // warp 10
constexpr uint32_t THREADS = 64 + EPI_WARPS * 32 + 32;
constexpr uint32_t MAX_MBLOCKS = 16;           // queries per launch <= 2048 (the driver splits larger batches)
constexpr uint32_t SUBCAP = 16;                // private candidate slots per (query, CTA, column half) and pass
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + ACC_STAGES * BLOCK_N * 4 + 256 + MAX_MBLOCKS * 256 * 4 +
                                EPI_WARPS * 32 * 4 /* survivor scratch */ + 1024;
static_assert(BLOCK_N == TILE_ROWS, "screen tile must match the pass schedule tile");

// instruction descriptor (cute::UMMA::InstrDescriptor bit layout): D=f32, A=B=bf16, both K-major
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((BLOCK_N >> 3) << 17) | ((BLOCK_M >> 4) << 24);
// kind::i8: D = s32 (c_format 2), A = B = signed 8-bit (format 1), K-major; UMMA_K = 32 (still 32 bytes per step)
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((BLOCK_N >> 3) << 17) | ((BLOCK_M >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// work item w -> (corpus tile w / n_mb, query block): the query block is skewed by the tile index so that every CTA
// serves every query block equally often (keeps the per-(query, CTA) private candidate sub-lists evenly filled)
__device__ __forceinline__ uint32_t item_mb(uint32_t w, uint32_t n_mb) { return (w % n_mb + w / n_mb) % n_mb; }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address
  d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major; canonical 1)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: next 8-row core-matrix group
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ---- survivors: warp-cooperative, control-flow uniform --------------------------------------------------------------
// A chunk is 32 columns (corpus rows) of ONE query per lane.  When some lane's chunk maximum reaches its threshold the
// whole warp handles that lane's chunk together: the lane publishes its 32 values to a 128-byte scratch line, every
// lane picks up one column, compares it with the owner's threshold, and the survivors (ballot) are written side by
// side into the owner query's private sub-list (position = popcount prefix), each also counted in the query's
// histogram (streaming mode; a fire-and-forget RED).  No per-lane divergent scan, no calls: an event costs ~40 warp
// instructions whatever the number of survivors -- the per-lane version cost ~130 per survivor, and since the
// accumulator stage is released only when the slowest of the 8 epilogue warps is done, that doubled the time of a
// short (1.25M-row) launch.
// vals: this lane's 32 values (int accumulators or scaled float scores as bits); all 32 lanes must call.
template <bool INT8, int MODE>
__device__ __forceinline__ void warp_survivors(uint32_t hits, const uint32_t (&vals)[32], uint32_t* scratch /* [32] per warp */,
                                               uint32_t lane, float my_tau, int tau_i, float2 my_hp, uint32_t q_base,
                                               uint32_t row_first, const float* __restrict__ snorm, Cand* __restrict__ cand,
                                               uint32_t* __restrict__ cand_cnt, uint32_t cap, Cand* __restrict__ sub,
                                               uint32_t n_slots, uint32_t slot, uint32_t* s_cnt_warp /* [32] */,
                                               uint32_t* hist) {
  while (hits) {
    const uint32_t L = __ffs(hits) - 1;
    hits &= hits - 1;
    if (lane == L) {
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<uint4*>(scratch + i) = make_uint4(vals[i], vals[i + 1], vals[i + 2], vals[i + 3]);
    }
    __syncwarp();
    const uint32_t xb = scratch[lane];  // column `lane` of the owner's chunk
    const float tau_f = __shfl_sync(0xffffffffu, my_tau, L);
    const int tau_n = __shfl_sync(0xffffffffu, tau_i, L);
    const float hp_lo = __shfl_sync(0xffffffffu, my_hp.x, L);  // the owner query's histogram geometry (streaming mode)
    const float hp_inv = __shfl_sync(0xffffffffu, my_hp.y, L);
    const uint32_t qL = q_base + L;
    const uint32_t r = row_first + lane;
    bool surv = INT8 ? ((int)xb >= tau_n) : (__uint_as_float(xb) >= tau_f);  // NaN scores never pass
    if (INT8 && surv && (int)xb == 0) {
      // invalid rows (skipped / special / padding) are all-zero in the int8 copy and score exactly 0: only a zero
      // score needs the look-up of the row's screening norm
      const float sn = __ldg(snorm + r);
      surv = sn == sn;
    }
    const uint32_t smask = __ballot_sync(0xffffffffu, surv);
    if (smask) {  // uniform
      const uint32_t base = s_cnt_warp[L];  // broadcast read: appends of this (query, CTA, column half) so far
      if (surv) {
        const uint32_t pos = base + __popc(smask & ((1u << lane) - 1u));
        const float score = INT8 ? __int2float_rn((int)xb) : __uint_as_float(xb);
        const uint2 cd = make_uint2(__float_as_uint(score), r);
        if (pos < SUBCAP) {
          *reinterpret_cast<uint2*>(sub + ((size_t)qL * n_slots + slot) * SUBCAP + pos) = cd;
        } else {  // private slots full: spill to the query's shared list
          const uint32_t p2 = atomicAdd(cand_cnt + qL, 1u);
          if (p2 < cap) *reinterpret_cast<uint2*>(cand + (size_t)qL * cap + p2) = cd;
        }
        if (MODE == 2) {
          HistParam hp;
          hp.lo = hp_lo;
          hp.inv_w0 = hp_inv;
          hp.w0 = 0.f;
          hp.margin = 0.f;
          atomicAdd(hist + (size_t)qL * HIST_BINS + hist_bin(hp, score), 1u);
        }
      }
      __syncwarp();
      if (lane == 0) s_cnt_warp[L] = base + __popc(smask);
    }
    __syncwarp();  // scratch and counter are reused by the next event
  }
}

// MODE 0: pass 0 -- every score of the pass's tiles goes to a fixed slot of the query's main list (tau = -inf)
// MODE 1: threshold pass -- survivors of a fixed tau (legacy multi-pass schedule)
// MODE 2: streaming pass -- tau is re-read at every work item and raised by the refiner warps while the kernel runs
// MODE 3: probe -- nothing is appended; every epilogue thread writes the maximum of each 32-column chunk it sees
//         (probe[q][tile * 8 + chunk]).  The chunk maxima belong to DISJOINT row sets, so the k-th largest of them is a
//         lower bound of the k-th best score of the corpus: the seed of the streaming pass's thresholds, at the cost
//         of one round of MMAs and no candidate traffic.
template <bool COSINE, bool INT8, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
screen_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const float* __restrict__ snorm, uint32_t k_blocks, uint32_t n_mblocks, uint32_t nq,
                 PassDesc pass, float* tau, Cand* __restrict__ cand,
                 uint32_t* __restrict__ cand_cnt, uint32_t cap, Cand* __restrict__ sub, uint32_t* __restrict__ sub_cnt,
                 uint32_t k, const HistParam* __restrict__ hparam, uint32_t* hist, float* __restrict__ probe,
                 uint32_t probe_stride, uint32_t sleep_min_ns, uint32_t sleep_max_ns) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                  // [STAGES][128][64] bf16
  uint8_t* smem_b = smem + STAGES * A_BYTES;               // [STAGES][256][64] bf16
  float* s_snorm = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);  // [ACC_STAGES][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_snorm + ACC_STAGES * BLOCK_N);
  uint64_t* full_bar = bars;                      // [STAGES]
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;        // [ACC_STAGES]
  uint64_t* tempty_bar = bars + 2 * STAGES + ACC_STAGES;  // [ACC_STAGES]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_STAGES);
  uint32_t* s_done = s_tmem + 1;  // epilogue warps that have finished (the refiner's exit condition)
  uint32_t* s_cnt = s_tmem + 4;   // [n_mblocks][256] private append counters of the epilogue threads
  uint32_t* s_scratch = s_cnt + MAX_MBLOCKS * 256;  // [EPI_WARPS][32] one chunk of one lane, for the survivor hand-over

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_items = pass.count * n_mblocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (uint32_t s = 0; s < STAGES; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (uint32_t a = 0; a < ACC_STAGES; a++) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), EPI_WARPS);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *s_done = 0;
  }
  for (uint32_t i = threadIdx.x; i < n_mblocks * 256; i += blockDim.x) s_cnt[i] = 0;
  if (warp == 1) {  // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const uint32_t tile = pass_tile(pass, w / n_mblocks);
        const uint32_t mb = item_mb(w, n_mblocks);
        for (uint32_t kb = 0; kb < k_blocks; kb++, it++) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, STAGE_BYTES);
          tma_load_2d(smem_u32(smem_a + s * A_BYTES), &map_a, kb * (INT8 ? 2 * BLOCK_K : BLOCK_K), mb * BLOCK_M, fb);
          tma_load_2d(smem_u32(smem_b + s * B_BYTES), &map_b, kb * (INT8 ? 2 * BLOCK_K : BLOCK_K), tile * BLOCK_N, fb);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0, j = 0;
      for (uint32_t w = blockIdx.x; w < n_items; w += gridDim.x, j++) {
        const uint32_t a = j % ACC_STAGES, pa = (j / ACC_STAGES) & 1;
        mbar_wait(smem_u32(&tempty_bar[a]), pa ^ 1);  // epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + a * BLOCK_N;
        for (uint32_t kb = 0; kb < k_blocks; kb++, it++) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_desc(smem_u32(smem_a + s * A_BYTES));
          const uint64_t db = make_desc(smem_u32(smem_b + s * B_BYTES));
#pragma unroll
          for (uint32_t k = 0; k < BLOCK_K / UMMA_K; k++) {
            // advance 32 bytes (16 bf16) inside the 128-byte swizzle atom: +2 in the (>>4) start-address field
            if (INT8) umma_i8(d_tmem, da + 2 * k, db + 2 * k, IDESC_I8, (kb | k) != 0);
            else umma_bf16(d_tmem, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0);
          }
          umma_commit(smem_u32(&empty_bar[s]));  // smem stage reusable once these MMAs retire
        }
        umma_commit(smem_u32(&tfull_bar[a]));  // accumulator complete
      }
    }
    __syncwarp();
  } else if (warp < REFINE_WARP) {
    // ===================== epilogue (8 warps; thread = query row, warp pair splits the 256 columns) ==========
    const uint32_t wq = warp & 3;                 // TMEM lane quarter this warp may access
    const uint32_t half = (warp - 2) >> 2;        // 0: columns 0..127, 1: columns 128..255
    const uint32_t row_in_tile = wq * 32 + lane;
    const uint32_t et = threadIdx.x - 64;         // 0..255
    constexpr bool pass0 = MODE == 0;             // pass 0: tau = -inf everywhere, positions are deterministic
    const uint32_t cbase = half * (BLOCK_N / 2);
    uint32_t j = 0;
    // the threshold of the first item; later items prefetch theirs while the current one is processed
    uint32_t w = blockIdx.x;
    float next_tau = __int_as_float(0x7f800000);
    float2 next_hp = make_float2(0.f, 0.f);  // streaming mode: (lo, 1/w0) of the query's histogram
    if (w < n_items) {
      const uint32_t q0 = item_mb(w, n_mblocks) * BLOCK_M + row_in_tile;
      if (q0 < nq) {
        next_tau = __ldcg(tau + q0);
        if (MODE == 2) next_hp = __ldg(reinterpret_cast<const float2*>(hparam + q0));
      }
    }
    for (; w < n_items; w += gridDim.x, j++) {
      const uint32_t tidx = w / n_mblocks;
      const uint32_t tile = pass_tile(pass, tidx);
      const uint32_t mb = item_mb(w, n_mblocks);
      const uint32_t a = j % ACC_STAGES, pa = (j / ACC_STAGES) & 1;
      const uint32_t q = mb * BLOCK_M + row_in_tile;
      const float my_tau = next_tau;
      const float2 my_hp = next_hp;
      {  // prefetch the next item's threshold (a global load whose latency would otherwise sit in front of the wait)
        const uint32_t wn = w + gridDim.x;
        next_tau = __int_as_float(0x7f800000);
        if (wn < n_items) {
          const uint32_t qn = item_mb(wn, n_mblocks) * BLOCK_M + row_in_tile;
          if (qn < nq) {
            next_tau = __ldcg(tau + qn);  // L2: sees the refiners' updates
            if (MODE == 2) next_hp = __ldg(reinterpret_cast<const float2*>(hparam + qn));
          }
        }
      }
      const size_t row0 = (size_t)tile * BLOCK_N;
      float* sn = s_snorm + a * BLOCK_N;
      if (!INT8 || MODE == 3) {
        // stage this tile's screening norms (safe: every epilogue thread passed the named barrier of item j-1
        // only after finishing item j-2, the previous user of s_snorm[a])
        sn[et] = __ldg(snorm + row0 + et);
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      // integer threshold for the int8 screen: acc >= tau  <=>  acc >= ceil(tau)
      int tau_i = 0x7fffffff;
      if (INT8) {
        const float ct = ceilf(my_tau);
        tau_i = ct >= 2147483520.f ? 0x7fffffff : (ct <= -2147483520.f ? (int)0x80000000 : (int)ct);
      }
      mbar_wait(smem_u32(&tfull_bar[a]), pa);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((wq * 32) << 16) + a * BLOCK_N + cbase;
      Cand* my_cand = cand + (size_t)q * cap;
      const uint32_t n_slots = gridDim.x * 2;
      const uint32_t my_cnt = smem_u32(s_cnt + mb * 256 + et);
      Cand* my_sub = sub + ((size_t)q * n_slots + blockIdx.x * 2 + half) * SUBCAP;
      {
      uint32_t va[32], vb[32];
      tmem_ld32(taddr, va);
#pragma unroll
      for (uint32_t cc = 0; cc < BLOCK_N / 2 / 32; cc++) {
        const uint32_t c0 = cc * 32;
        uint32_t(&v)[32] = (cc & 1) ? vb : va;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cc + 1 < BLOCK_N / 2 / 32) tmem_ld32(taddr + c0 + 32, (cc & 1) ? va : vb);  // next chunk in flight
        if (INT8) {
          int gm[4];  // maxima of the four groups of 8 columns: the slow path only scans groups that hold a survivor
#pragma unroll
          for (int g = 0; g < 4; g++) {
            gm[g] = (int)v[8 * g];
#pragma unroll
            for (int i = 1; i < 8; i++) gm[g] = max(gm[g], (int)v[8 * g + i]);
          }
          const int m = max(max(gm[0], gm[1]), max(gm[2], gm[3]));
          if (MODE == 3) {
            // invalid rows (NaN screening norm) score 0 in the integer screen: they must not pose as a real score
            int pm = (int)0x80000000;
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const float snv = sn[cbase + c0 + i];
              if (snv == snv) pm = max(pm, (int)v[i]);
            }
            if (q < nq)
              probe[(size_t)q * probe_stride + tidx * 8 + half * 4 + cc] =
                  pm == (int)0x80000000 ? __int_as_float(0xff800000) : __int2float_rd(pm);
          } else if (pass0) {
            if (q < nq) {
#pragma unroll
              for (int i = 0; i < 32; i++) {
                Cand cd;
                cd.score = __int2float_rn((int)v[i]);
                cd.row = (uint32_t)(row0 + cbase + c0 + i);
                my_cand[(size_t)tidx * BLOCK_N + cbase + c0 + i] = cd;
              }
            }
          } else {
            const uint32_t hits = __ballot_sync(0xffffffffu, m >= tau_i);
            if (hits)  // rare, uniform across the warp
              warp_survivors<true, MODE>(hits, v, s_scratch + (warp - 2) * 32, lane, my_tau, tau_i, my_hp,
                                         mb * BLOCK_M + wq * 32, (uint32_t)(row0 + cbase + c0), snorm, cand, cand_cnt, cap,
                                         sub, gridDim.x * 2, blockIdx.x * 2 + half, s_cnt + mb * 256 + (et - lane), hist);
          }
        } else {
          float sc[32];
          float gmf[4];
#pragma unroll
          for (int g = 0; g < 4; g++) gmf[g] = __int_as_float(0xff800000);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 n4 = *reinterpret_cast<const float4*>(sn + cbase + c0 + i);
            sc[i + 0] = COSINE ? __uint_as_float(v[i + 0]) * n4.x : fmaf(2.f, __uint_as_float(v[i + 0]), -n4.x);
            sc[i + 1] = COSINE ? __uint_as_float(v[i + 1]) * n4.y : fmaf(2.f, __uint_as_float(v[i + 1]), -n4.y);
            sc[i + 2] = COSINE ? __uint_as_float(v[i + 2]) * n4.z : fmaf(2.f, __uint_as_float(v[i + 2]), -n4.z);
            sc[i + 3] = COSINE ? __uint_as_float(v[i + 3]) * n4.w : fmaf(2.f, __uint_as_float(v[i + 3]), -n4.w);
            gmf[i >> 3] = fmaxf(gmf[i >> 3], fmaxf(fmaxf(sc[i + 0], sc[i + 1]), fmaxf(sc[i + 2], sc[i + 3])));  // fmaxf drops NaNs
          }
          const float m = fmaxf(fmaxf(gmf[0], gmf[1]), fmaxf(gmf[2], gmf[3]));
          if (MODE == 3) {
            if (q < nq) probe[(size_t)q * probe_stride + tidx * 8 + half * 4 + cc] = m;  // -inf: no valid row in the chunk
          } else if (pass0) {
            if (q < nq) {  // every (finite or NaN) score goes to its fixed slot; compaction drops the NaNs
#pragma unroll
              for (int i = 0; i < 32; i++) {
                Cand cd;
                cd.score = sc[i];
                cd.row = (uint32_t)(row0 + cbase + c0 + i);
                my_cand[(size_t)tidx * BLOCK_N + cbase + c0 + i] = cd;
              }
            }
          } else {
            const uint32_t hits = __ballot_sync(0xffffffffu, m >= my_tau);
            if (hits) {  // rare, uniform across the warp
              uint32_t sb[32];
#pragma unroll
              for (int i = 0; i < 32; i++) sb[i] = __float_as_uint(sc[i]);
              warp_survivors<false, MODE>(hits, sb, s_scratch + (warp - 2) * 32, lane, my_tau, 0, my_hp,
                                          mb * BLOCK_M + wq * 32, (uint32_t)(row0 + cbase + c0), snorm, cand, cand_cnt, cap,
                                          sub, gridDim.x * 2, blockIdx.x * 2 + half, s_cnt + mb * 256 + (et - lane), hist);
            }
          }
        }
      }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[a]));
    }
    // publish the private append counters: slot (CTA, column half) of every query this thread served
    if (MODE == 1 || MODE == 2) {
      const uint32_t n_slots = gridDim.x * 2;
      for (uint32_t mb = 0; mb < n_mblocks; mb++) {
        const uint32_t qq = mb * BLOCK_M + row_in_tile;
        if (qq < nq) {
          const uint32_t cnt = s_cnt[mb * 256 + et];
          sub_cnt[(size_t)qq * n_slots + blockIdx.x * 2 + half] = cnt;
        }
      }
    }
    __syncwarp();
    if (lane == 0) atomicAdd(s_done, 1u);
  } else {
    // ===================== refiner (streaming mode): raise the thresholds of the queries this CTA owns ==========
    if (MODE == 2) {
      volatile uint32_t* done = s_done;
      const uint32_t stride = gridDim.x;
      uint32_t n_own = nq > blockIdx.x ? (nq - blockIdx.x + stride - 1) / stride : 0;
      if (n_own > 32) n_own = 32;  // (tiny grids only; the rest keeps its seed threshold)
      float my_tau = __int_as_float(0xff800000);  // lane j: the threshold last published for owned query j
      if (lane < n_own) my_tau = __ldcg(tau + blockIdx.x + lane * stride);
      uint32_t sleep_ns = sleep_min_ns;
      while (*done < EPI_WARPS) {
        bool any = false;
        for (uint32_t j = 0; j < n_own; j++) {
          const uint32_t q = blockIdx.x + j * stride;
          const uint4* hq = reinterpret_cast<const uint4*>(hist + (size_t)q * HIST_BINS) + lane * 2;
          const uint4 ha = __ldcg(hq), hb = __ldcg(hq + 1);  // bins lane*8 .. lane*8+7
          const uint32_t c[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
          uint32_t sum = 0;
#pragma unroll
          for (int i = 0; i < 8; i++) sum += c[i];
          uint32_t incl = sum;  // suffix sum over lanes: higher lanes hold higher bins
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += t;
          }
          const uint32_t above = incl - sum;
          const bool mine = above < k && k <= incl;  // the count from the top reaches k inside this lane's bins
          const uint32_t who = __ballot_sync(0xffffffffu, mine);
          if (who == 0) continue;  // fewer than k survivors counted so far
          uint32_t bstar = 0;
          if (mine) {
            uint32_t acc = above;
#pragma unroll
            for (int i = 7; i >= 0; i--) {
              acc += c[i];
              if (acc >= k) {
                bstar = lane * 8 + i;
                break;
              }
            }
          }
          bstar = __shfl_sync(0xffffffffu, bstar, __ffs(who) - 1);
          const float4 hv = __ldg(reinterpret_cast<const float4*>(hparam + q));
          HistParam hp;
          hp.lo = hv.x;
          hp.inv_w0 = hv.y;
          hp.w0 = hv.z;
          hp.margin = hv.w;
          // >= k rows with score >= edge(bstar) exist (2 % of a bin + 2e-6 relative absorb the rounding of hist_bin)
          const double e0 = hist_edge(hp, bstar), e1 = hist_edge(hp, bstar + 1);
          const float tn = __double2float_rd(e0 - (double)hp.margin - 0.02 * (e1 - e0) - 2e-6 * fabs(e0));
          const float told = __shfl_sync(0xffffffffu, my_tau, j);
          if (tn > told) {
            if (lane == j) my_tau = tn;
            if (lane == 0) __stcg(tau + q, tn);
            any = true;
          }
        }
        sleep_ns = any ? sleep_min_ns : (sleep_ns * 2 < sleep_max_ns ? sleep_ns * 2 : sleep_max_ns);
        __nanosleep(sleep_ns);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}
}  // namespace tc

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode(Ctx* ctx) {
  if (!ctx->encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      fn = nullptr;
    ctx->encode_tiled = fn;
  }
  return reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
}

static sdb_status make_map(Ctx* ctx, CUtensorMap* map, const void* base, uint64_t rows, uint32_t dim_pad,
                           uint32_t box_rows, bool stream_once, bool int8 = false) {
  EncodeTiledFn enc = get_encode(ctx);
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return SDB_ECUDA;
  }
  cuuint64_t gdim[2] = {dim_pad, rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dim_pad * (int8 ? 1 : 2)};
  cuuint32_t box[2] = {int8 ? 2 * tc::BLOCK_K : tc::BLOCK_K, box_rows};  // 128 bytes per row either way
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, int8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   stream_once ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SDB_ECUDA;
  }
  return SDB_OK;
}

bool screen_tc_available() { return true; }

sdb_status screen_tc_init_device() {
#define SET_SMEM(COS, I8, MODE) \
  SDB_CUDA(cudaFuncSetAttribute(tc::screen_tc_kernel<COS, I8, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES))
  SET_SMEM(true, false, 0); SET_SMEM(true, false, 1); SET_SMEM(true, false, 2); SET_SMEM(true, false, 3);
  SET_SMEM(false, false, 0); SET_SMEM(false, false, 1); SET_SMEM(false, false, 2); SET_SMEM(false, false, 3);
  SET_SMEM(true, true, 0); SET_SMEM(true, true, 1); SET_SMEM(true, true, 2); SET_SMEM(true, true, 3);
#undef SET_SMEM
  return SDB_OK;
}

sdb_status screen_tc_pass(Corpus* c, uint32_t nq, uint32_t k, const PassDesc& p, bool int8, int mode, cudaStream_t st) {
  if (p.count == 0) return SDB_OK;
  Ctx* ctx = c->ctx;
  if (int8 ? (!c->d_i8 || c->metric != SDB_COSINE) : !c->d_bf16) {
    set_error("tcgen05 screen: the %s screen copy is not available for this corpus", int8 ? "int8 (cosine only)" : "bf16");
    return SDB_EUNSUPPORTED;
  }
  const uint64_t n_pad = (c->n + TILE_ROWS - 1) / TILE_ROWS * TILE_ROWS;
  // refiner pacing (tuning knobs): it re-reads its queries' histograms at most every sleep_min ns while thresholds
  // move, backing off to sleep_max when they do not
  uint32_t sleep_min = 512, sleep_max = 8192;
  if (const char* e = getenv("SDB_REFINE_SLEEP_MIN")) sleep_min = (uint32_t)atoi(e);
  if (const char* e = getenv("SDB_REFINE_SLEEP_MAX")) sleep_max = (uint32_t)atoi(e);
  if (sleep_min < 32) sleep_min = 32;
  if (sleep_max < sleep_min) sleep_max = sleep_min;
  CUtensorMap map_b;
  if (int8) SDB_TRY(make_map(ctx, &map_b, c->d_i8, n_pad, c->dim_pad8, tc::BLOCK_N, true, true));
  else SDB_TRY(make_map(ctx, &map_b, c->d_bf16, n_pad, c->dim_pad, tc::BLOCK_N, true));
  const uint32_t k_blocks = int8 ? c->dim_pad8 / (2 * tc::BLOCK_K) : c->dim_pad / tc::BLOCK_K;
  // every launch uses the same grid so that the (CTA, half) slot numbering of the private sub-lists is stable
  const uint32_t chunk_q = tc::MAX_MBLOCKS * tc::BLOCK_M;
  uint32_t grid = (uint32_t)ctx->sm_count;
  {
    const uint32_t nq0 = nq < chunk_q ? nq : chunk_q;
    const uint64_t items0 = (uint64_t)p.count * ((nq0 + tc::BLOCK_M - 1) / tc::BLOCK_M);
    if (grid > items0) grid = (uint32_t)items0;
  }
  if (mode == 3 && p.count * 8 > PROBE_STRIDE) {
    set_error("screen_tc_pass: probe of %u tiles exceeds the probe buffer", p.count);
    return SDB_EINVAL;
  }
  if (mode != 3) c->last_slots = grid * 2;
  const uint32_t slots = grid * 2;
  for (uint32_t q0 = 0; q0 < nq; q0 += chunk_q) {
    const uint32_t nqc = nq - q0 < chunk_q ? nq - q0 : chunk_q;
    const uint32_t nq_pad = (nqc + tc::BLOCK_M - 1) / tc::BLOCK_M * tc::BLOCK_M;
    const uint32_t n_mblocks = nq_pad / tc::BLOCK_M;
    CUtensorMap map_a;
    if (int8) SDB_TRY(make_map(ctx, &map_a, c->d_q8 + (size_t)q0 * c->dim_pad8, nq_pad, c->dim_pad8, tc::BLOCK_M, false, true));
    else SDB_TRY(make_map(ctx, &map_a, c->d_qbf16 + (size_t)q0 * c->dim_pad, nq_pad, c->dim_pad, tc::BLOCK_M, false));
    float* tau = c->d_tau + q0;
    Cand* cand = c->d_cand + (size_t)q0 * c->sc_cap;
    uint32_t* ccnt = c->d_cand_cnt + q0;
    Cand* sub = c->d_sub + (size_t)q0 * slots * tc::SUBCAP;
    uint32_t* scnt = c->d_sub_cnt + (size_t)q0 * slots;
    const HistParam* hp = c->d_hparam + q0;
    uint32_t* hist = c->d_hist + (size_t)q0 * HIST_BINS;
    // the screen is launched at the highest priority: when the previous batch's screen retires, the blocks of THIS
    // launch are placed before the queued blocks of that batch's tail kernels (which fit beside a screen CTA anyway),
    // instead of waiting behind two 13 KB selection blocks per SM
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = tc::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributePriority;
    attr[0].val.priority = ctx->prio_high;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    float* probe_ptr = c->d_probe + (size_t)q0 * PROBE_STRIDE;
    const uint32_t probe_stride = PROBE_STRIDE, cap_arg = c->sc_cap;
    const float* snorm_arg = c->d_snorm;
#define LAUNCH_TC1(COS, I8, MODE)                                                                                    \
  SDB_CUDA(cudaLaunchKernelEx(&cfg, tc::screen_tc_kernel<COS, I8, MODE>, map_a, map_b, snorm_arg, k_blocks, n_mblocks, \
                              nqc, p, tau, cand, ccnt, cap_arg, sub, scnt, k, hp, hist, probe_ptr, probe_stride,       \
                              sleep_min, sleep_max))
#define LAUNCH_TC(COS, I8)                   \
  do {                                       \
    if (mode == 0) LAUNCH_TC1(COS, I8, 0);   \
    else if (mode == 1) LAUNCH_TC1(COS, I8, 1); \
    else if (mode == 2) LAUNCH_TC1(COS, I8, 2); \
    else LAUNCH_TC1(COS, I8, 3);             \
  } while (0)
    if (int8) LAUNCH_TC(true, true);
    else if (c->metric == SDB_COSINE) LAUNCH_TC(true, false);
    else LAUNCH_TC(false, false);
#undef LAUNCH_TC
#undef LAUNCH_TC1
    count_launch(ctx);
  }
  if (mode == 0) {
    SDB_TRY(cand_set_count(c, nq, p.count * TILE_ROWS, st));  // pass 0 wrote fixed slots of the main lists
    c->last_slots = 0;                                         // ... and no private sub-lists
  }
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

}  // namespace sdb
