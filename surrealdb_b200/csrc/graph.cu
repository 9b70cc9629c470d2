// graph.cu -- K4: `->edge->node` expansion over device-resident CSR adjacency.
//
// Replaces the per-source KV prefix scans of GraphEdgeScan::execute (exec/operators/scan/graph.rs:214-279)
// driven by LookupPart::evaluate_lookup (exec/parts/lookup.rs:139-170): for every frontier element, in
// frontier order, emit its targets in stored (KV key) order -- duplicates kept, per-source limit honoured
// (graph.rs:83,238,261).  Output position = exclusive prefix sum of the (limited) degrees, so the result is
// order-identical to the reference no matter how the work is split.  The expand kernel is OUTPUT-centric
// (each block owns a contiguous slice of the output and finds its sources by binary search), which balances
// power-law degree distributions at warp/block level without any per-vertex special casing.
// `+collect` (exec/operators/recursion/collect.rs:74-143) adds a first-seen de-duplication per BFS level.
//
// Algorithmic bytes per hop: 16|F| (two row_ptr reads per source) + 4|E_h| (col_idx) + 4|E_h| (output).
#include "internal.cuh"

namespace sdb {

struct Graph {
  Ctx* ctx = nullptr;
  uint64_t n_rows = 0, n_edges = 0;
  uint64_t* d_row_ptr = nullptr;
  uint32_t* d_col_idx = nullptr;
  bool targets_in_rows = true;  // every col_idx < n_rows (required by +collect, which indexes per-row state by target)
  // row-sharded adjacency: this rank holds rows [row_lo, row_hi) of the n_rows-row CSR; d_row_ptr is the
  // slice rebased to 0.  An unsharded graph is the shard [0, n_rows).
  uint64_t row_lo = 0, row_hi = 0;
  bool sharded = false;
  // +collect state, kept between calls (a 50M-node graph needs 450 MB of it: allocating it per call cost 145 ms)
  uint8_t* d_seen = nullptr;
  uint32_t* d_first = nullptr;
  uint32_t* d_res = nullptr;
  uint64_t res_cap = 0;
  std::mutex mu;
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ---- exclusive prefix sum (u64), hierarchical ------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out,
                                                                 uint64_t n, uint64_t* __restrict__ tile_sums) {
  __shared__ uint64_t s_warp[SCAN_THREADS / 32];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint64_t v[SCAN_ITEMS];
  uint64_t sum = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    v[i] = base + i < n ? in[base + i] : 0;
    sum += v[i];
  }
  // inclusive scan of per-thread sums inside the block
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= (uint32_t)o) w += t;
    }
    if (lane < SCAN_THREADS / 32) s_warp[lane] = w;
  }
  __syncthreads();
  uint64_t excl = inc - sum + (warp ? s_warp[warp - 1] : 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    if (base + i < n) out[base + i] = excl;
    excl += v[i];
  }
  if (threadIdx.x == SCAN_THREADS - 1 && tile_sums) tile_sums[blockIdx.x] = excl;
}
__global__ void scan_add_kernel(uint64_t* __restrict__ out, uint64_t n, const uint64_t* __restrict__ tile_off) {
  const uint64_t i = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
  const uint64_t add = tile_off[blockIdx.x];
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const uint64_t j = i + (uint64_t)k * SCAN_THREADS;
    if (j < n) out[j] += add;
  }
}
// out[0..n) = exclusive scan of in[0..n); *d_total = sum.  in/out may alias.
sdb_status exclusive_scan(Ctx* ctx, const uint64_t* d_in, uint64_t* d_out, uint64_t n, uint64_t* d_total,
                                 cudaStream_t st) {
  if (n == 0) {
    SDB_CUDA(cudaMemsetAsync(d_total, 0, 8, st));
    return SDB_OK;
  }
  const uint64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  uint64_t* d_sums = nullptr;
  SDB_CUDA(cudaMallocAsync(&d_sums, sizeof(uint64_t) * (tiles + 1), st));
  scan_tile_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(d_in, d_out, n, d_sums);
  count_launch(ctx);
  if (tiles > 1) {
    SDB_TRY(exclusive_scan(ctx, d_sums, d_sums, tiles, d_total, st));
    scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(d_out, n, d_sums);
    count_launch(ctx);
  } else {
    SDB_CUDA(cudaMemcpyAsync(d_total, d_sums, 8, cudaMemcpyDeviceToDevice, st));
  }
  SDB_CUDA(cudaFreeAsync(d_sums, st));
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
}

// ---- one hop -----------------------------------------------------------------------------------------
// row_ptr is this rank's slice [row_lo, row_hi) rebased to 0; sources owned by another rank contribute 0 here and
// their degree arrives through the all-reduce
__global__ void degree_kernel(const uint64_t* __restrict__ row_ptr, uint64_t n_rows, uint64_t row_lo, uint64_t row_hi,
                              const uint32_t* __restrict__ frontier, uint64_t n_f, uint32_t limit,
                              uint64_t* __restrict__ deg, uint32_t* __restrict__ err) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_f) return;
  const uint32_t v = frontier[i];
  if (v >= n_rows) {
    *err = 1;
    deg[i] = 0;
    return;
  }
  uint64_t d = 0;
  if (v >= row_lo && v < row_hi) {
    d = row_ptr[v - row_lo + 1] - row_ptr[v - row_lo];
    if (limit && d > limit) d = limit;
  }
  deg[i] = d;
}

constexpr int EXP_THREADS = 256;
constexpr int EXP_PER_THREAD = 8;
constexpr int EXP_TILE = EXP_THREADS * EXP_PER_THREAD;  // outputs per block
constexpr int EXP_SRC_MAX = EXP_TILE + 2;

__device__ __forceinline__ uint64_t upper_bound_minus1(const uint64_t* a, uint64_t lo, uint64_t hi, uint64_t x) {
  // largest i in [lo, hi) with a[i] <= x   (a is non-decreasing, a[lo] <= x)
  while (hi - lo > 1) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (a[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(EXP_THREADS) expand_kernel(const uint64_t* __restrict__ row_ptr,
                                                             const uint32_t* __restrict__ col_idx,
                                                             const uint32_t* __restrict__ frontier, uint64_t n_f,
                                                             const uint64_t* __restrict__ off /* n_f + 1 */,
                                                             uint64_t total, uint32_t* __restrict__ out,
                                                             uint64_t row_lo, uint64_t row_hi) {
  __shared__ uint64_t s_off[EXP_SRC_MAX];
  __shared__ uint64_t s_row[EXP_SRC_MAX];
  __shared__ uint64_t s_i0, s_i1;
  const uint64_t o0 = (uint64_t)blockIdx.x * EXP_TILE;
  const uint64_t o1 = o0 + EXP_TILE < total ? o0 + EXP_TILE : total;
  if (threadIdx.x == 0) {
    s_i0 = upper_bound_minus1(off, 0, n_f + 1, o0);
    s_i1 = upper_bound_minus1(off, 0, n_f + 1, o1 - 1);
  }
  __syncthreads();
  const uint64_t i0 = s_i0, i1 = s_i1;
  const bool in_smem = (i1 - i0 + 2) <= (uint64_t)EXP_SRC_MAX;
  if (in_smem) {
    for (uint64_t t = threadIdx.x; t < i1 - i0 + 2; t += EXP_THREADS) {
      const uint64_t i = i0 + t;
      s_off[t] = off[i];  // i <= i1 + 1 <= n_f
      uint64_t rb = ~0ull;  // ~0: the source belongs to another rank's rows -- its output slots stay zero here
      if (i < n_f) {
        const uint64_t v = frontier[i];
        if (v >= row_lo && v < row_hi) rb = row_ptr[v - row_lo];
      }
      s_row[t] = rb;
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < EXP_PER_THREAD; u++) {
    const uint64_t o = o0 + (uint64_t)u * EXP_THREADS + threadIdx.x;  // coalesced output
    if (o >= o1) break;
    uint64_t src, start, rbeg;
    if (in_smem) {
      const uint64_t t = upper_bound_minus1(s_off, 0, i1 - i0 + 2, o);
      src = i0 + t;
      start = s_off[t];
      rbeg = s_row[t];
    } else {  // pathological: thousands of zero-degree sources inside this slice
      src = upper_bound_minus1(off, i0, i1 + 2, o);
      start = off[src];
      const uint64_t v = frontier[src];
      rbeg = (v >= row_lo && v < row_hi) ? row_ptr[v - row_lo] : ~0ull;
    }
    if (rbeg != ~0ull) out[o] = __ldg(col_idx + rbeg + (o - start));
    (void)src;
  }
}

static sdb_status hop_device(Graph* g, const uint32_t* d_frontier, uint64_t n_f, uint32_t limit, uint32_t** d_out,
                             uint64_t* n_out, cudaStream_t st) {
  Ctx* ctx = g->ctx;
  *d_out = nullptr;
  *n_out = 0;
  if (n_f == 0) return SDB_OK;
  uint64_t* d_off = nullptr;
  uint64_t* d_total = nullptr;
  uint32_t* d_err = nullptr;
  SDB_CUDA(cudaMallocAsync(&d_off, sizeof(uint64_t) * (n_f + 2), st));
  d_total = d_off + n_f;  // off[n_f] = total: exactly the sentinel the expand kernel wants
  SDB_CUDA(cudaMallocAsync(&d_err, 4, st));
  SDB_CUDA(cudaMemsetAsync(d_err, 0, 4, st));
  degree_kernel<<<(unsigned)((n_f + 255) / 256), 256, 0, st>>>(g->d_row_ptr, g->n_rows, g->row_lo, g->row_hi, d_frontier, n_f,
                                                               limit, d_off, d_err);
  count_launch(ctx);
  // sharded adjacency: every source is owned by exactly one rank, so the SUM over ranks is the full degree array
  const bool multi = g->sharded && comm_size(ctx) > 1;
  if (multi) SDB_TRY(comm_allreduce_sum(ctx, d_off, n_f, 8, st));
  SDB_TRY(exclusive_scan(ctx, d_off, d_off, n_f, d_total, st));
  uint64_t total = 0;
  uint32_t err = 0;
  SDB_CUDA(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaStreamSynchronize(st));
  if (err) {
    cudaFreeAsync(d_off, st);
    cudaFreeAsync(d_err, st);
    set_error("graph expand: frontier id out of range (graph has %llu rows)", (unsigned long long)g->n_rows);
    return SDB_EINVAL;
  }
  if (total > 0xFFFFFFF0ull) {
    cudaFreeAsync(d_off, st);
    cudaFreeAsync(d_err, st);
    set_error("graph expand: %llu results exceed the 2^32 frontier limit", (unsigned long long)total);
    return SDB_EOVERFLOW;
  }
  if (total) {
    SDB_CUDA(cudaMallocAsync(d_out, sizeof(uint32_t) * total, st));
    if (multi) SDB_CUDA(cudaMemsetAsync(*d_out, 0, sizeof(uint32_t) * total, st));
    expand_kernel<<<(unsigned)((total + EXP_TILE - 1) / EXP_TILE), EXP_THREADS, 0, st>>>(
        g->d_row_ptr, g->d_col_idx, d_frontier, n_f, d_off, total, *d_out, g->row_lo, g->row_hi);
    count_launch(ctx);
    // every output slot was written by exactly one rank (the owner of its source), the others hold 0: the sum over
    // ranks IS the next frontier, in the reference's order, on every rank -- ONE exchange per hop
    if (multi) SDB_TRY(comm_allreduce_sum(ctx, *d_out, total, 4, st));
  }
  SDB_CUDA(cudaFreeAsync(d_off, st));
  SDB_CUDA(cudaFreeAsync(d_err, st));
  SDB_CUDA(cudaGetLastError());
  *n_out = total;
  return SDB_OK;
}

// ---- +collect: first-seen de-duplication of one BFS level ----------------------------------------------
__global__ void collect_mark_kernel(const uint32_t* __restrict__ lvl, uint64_t n, const uint8_t* __restrict__ seen,
                                    uint32_t* __restrict__ first_pos) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t v = lvl[p];
  // hubs occur millions of times per level: only positions that could still lower the minimum issue the atomic
  // (a plain load first -- positions are handed out in increasing order, so later duplicates almost always skip it)
  if (!seen[v] && __ldcg(first_pos + v) > (uint32_t)p) atomicMin(first_pos + v, (uint32_t)p);
}
__global__ void collect_flag_kernel(const uint32_t* __restrict__ lvl, uint64_t n, const uint8_t* __restrict__ seen,
                                    const uint32_t* __restrict__ first_pos, uint64_t* __restrict__ keep) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t v = lvl[p];
  keep[p] = (!seen[v] && first_pos[v] == (uint32_t)p) ? 1 : 0;
}
__global__ void collect_compact_kernel(const uint32_t* __restrict__ lvl, uint64_t n, const uint64_t* __restrict__ pos /* n+1 */,
                                       uint8_t* __restrict__ seen, uint32_t* __restrict__ first_pos,
                                       uint32_t* __restrict__ next) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (pos[p + 1] != pos[p]) {  // kept
    const uint32_t v = lvl[p];
    next[pos[p]] = v;
    seen[v] = 1;
    first_pos[v] = 0xFFFFFFFFu;
  }
}

}  // namespace sdb

struct sdb_graph : sdb::Graph {};
using namespace sdb;

extern "C" {

void sdb_graph_destroy(sdb_graph* g);

static sdb_status graph_load(sdb_ctx* ctx, uint64_t n_rows, uint64_t row_lo, uint64_t row_hi, const uint64_t* row_ptr,
                             const uint32_t* col_idx, bool sharded, sdb_graph** out) {
  if (!ctx || !out || !row_ptr || n_rows >= 0xFFFFFFF0ull || row_lo > row_hi || row_hi > n_rows) return SDB_EINVAL;
  *out = nullptr;
  const uint64_t n_local = row_hi - row_lo;
  if (row_ptr[0] != 0) {
    set_error("graph load: row_ptr[0] must be 0 (a shard's slice is rebased to its first row)");
    return SDB_EINVAL;
  }
  const uint64_t n_edges = row_ptr[n_local];
  if (n_edges && !col_idx) return SDB_EINVAL;
  for (uint64_t i = 0; i < n_local; i++)
    if (row_ptr[i + 1] < row_ptr[i]) {
      set_error("row_ptr is not non-decreasing at %llu", (unsigned long long)i);
      return SDB_EINVAL;
    }
  std::lock_guard<std::mutex> guard(ctx->mu);
  SDB_CUDA(cudaSetDevice(ctx->device));
  sdb_graph* g = new sdb_graph();
  g->ctx = ctx;
  g->n_rows = n_rows;
  g->n_edges = n_edges;
  g->row_lo = row_lo;
  g->row_hi = row_hi;
  g->sharded = sharded;
  cudaError_t e = cudaMalloc(&g->d_row_ptr, sizeof(uint64_t) * (n_local + 1));
  if (e == cudaSuccess) e = cudaMalloc(&g->d_col_idx, sizeof(uint32_t) * (n_edges ? n_edges : 1));
  if (e != cudaSuccess) {
    set_error("graph allocation failed: %s", cudaGetErrorString(e));
    sdb_graph_destroy(g);
    return SDB_ENOMEM;
  }
  SDB_CUDA(cudaMemcpyAsync(g->d_row_ptr, row_ptr, sizeof(uint64_t) * (n_local + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (n_edges)
    SDB_CUDA(cudaMemcpyAsync(g->d_col_idx, col_idx, sizeof(uint32_t) * n_edges, cudaMemcpyHostToDevice, ctx->stream));
  SDB_CUDA(cudaStreamSynchronize(ctx->stream));
  {  // range check on the device copy: row_ptr consistent with the edge count; do targets stay inside the rows?
    unsigned long long bad[2] = {0, 0};
    const sdb_status rc = csr_check(ctx, g->d_row_ptr, g->d_col_idx, n_local, n_edges, n_rows, bad, "sdb_graph_load_csr", ctx->stream);
    if (rc != SDB_OK || bad[0]) {
      if (rc == SDB_OK) set_error("sdb_graph_load_csr: malformed row_ptr (%llu violations)", bad[0]);
      sdb_graph_destroy(g);
      return rc != SDB_OK ? rc : SDB_EINVAL;
    }
    g->targets_in_rows = bad[1] == 0;  // targets of another table may exceed this table's rows: fine for plain hops
  }
  *out = g;
  return SDB_OK;
}

sdb_status sdb_graph_load_csr(sdb_ctx* ctx, uint64_t n_rows, const uint64_t* row_ptr, const uint32_t* col_idx,
                              sdb_graph** out) {
  return graph_load(ctx, n_rows, 0, n_rows, row_ptr, col_idx, false, out);
}

sdb_status sdb_graph_load_csr_shard(sdb_ctx* ctx, uint64_t n_rows_total, uint64_t row_lo, uint64_t row_hi,
                                    const uint64_t* row_ptr, const uint32_t* col_idx, sdb_graph** out) {
  return graph_load(ctx, n_rows_total, row_lo, row_hi, row_ptr, col_idx, true, out);
}

void sdb_graph_destroy(sdb_graph* g) {
  if (!g) return;
  cudaSetDevice(g->ctx->device);
  cudaFree(g->d_row_ptr);
  cudaFree(g->d_col_idx);
  cudaFree(g->d_seen);
  cudaFree(g->d_first);
  cudaFree(g->d_res);
  delete g;
}

// device-resident core: frontier and result stay in HBM (the result is library-owned: sdb_device_free)
static sdb_status graph_expand_dev(sdb_graph* const* hops, uint32_t n_hops, const uint32_t* d_frontier, uint64_t n_frontier,
                                   uint32_t per_source_limit, uint32_t** d_out, uint64_t* out_n, cudaStream_t st) {
  uint32_t* d_f = nullptr;
  uint64_t n_f = n_frontier;
  bool owned = false;  // the caller's frontier is never freed
  const uint32_t* cur = d_frontier;
  for (uint32_t h = 0; h < n_hops && n_f; h++) {
    uint32_t* d_next = nullptr;
    uint64_t n_next = 0;
    if (ctx_cancelled(hops[h]->ctx)) {  // polled once per hop
      if (owned) cudaFreeAsync(d_f, st);
      set_error("query cancelled");
      return SDB_ECANCELLED;
    }
    sdb_status s = hop_device(hops[h], cur, n_f, per_source_limit, &d_next, &n_next, st);
    if (owned) cudaFreeAsync(d_f, st);
    if (s != SDB_OK) {
      if (d_next) cudaFreeAsync(d_next, st);
      return s;
    }
    d_f = d_next;
    cur = d_next;
    owned = true;
    n_f = n_next;
  }
  if (!owned && n_f) {  // zero hops: hand back a copy
    SDB_CUDA(cudaMallocAsync(&d_f, sizeof(uint32_t) * n_f, st));
    SDB_CUDA(cudaMemcpyAsync(d_f, d_frontier, sizeof(uint32_t) * n_f, cudaMemcpyDeviceToDevice, st));
  }
  if (n_f == 0 && owned && d_f) {
    cudaFreeAsync(d_f, st);
    d_f = nullptr;
  }
  *d_out = n_f ? d_f : nullptr;
  *out_n = n_f;
  return SDB_OK;
}

sdb_status sdb_graph_expand_device(sdb_graph* const* hops, uint32_t n_hops, const uint32_t* d_frontier, uint64_t n_frontier,
                                   uint32_t per_source_limit, uint32_t** d_out_ids, uint64_t* out_n) {
  if (!hops || !n_hops || !d_out_ids || !out_n || (n_frontier && !d_frontier)) return SDB_EINVAL;
  *d_out_ids = nullptr;
  *out_n = 0;
  for (uint32_t h = 0; h < n_hops; h++)
    if (!hops[h]) return SDB_EINVAL;
  Ctx* ctx = hops[0]->ctx;
  std::lock_guard<std::mutex> guard(ctx->mu);
  SDB_CUDA(cudaSetDevice(ctx->device));
  SDB_TRY(graph_expand_dev(hops, n_hops, d_frontier, n_frontier, per_source_limit, d_out_ids, out_n, ctx->stream));
  SDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SDB_OK;
}

void sdb_device_free(sdb_ctx* ctx, void* d_ptr) {
  if (!ctx || !d_ptr) return;
  cudaSetDevice(ctx->device);
  cudaFreeAsync(d_ptr, ctx->stream);
}

sdb_status sdb_graph_expand(sdb_graph* const* hops, uint32_t n_hops, const uint32_t* frontier, uint64_t n_frontier,
                            uint32_t per_source_limit, uint32_t** out_ids, uint64_t* out_n) {
  if (!hops || !n_hops || !out_ids || !out_n || (n_frontier && !frontier)) return SDB_EINVAL;
  *out_ids = nullptr;
  *out_n = 0;
  for (uint32_t h = 0; h < n_hops; h++)
    if (!hops[h]) return SDB_EINVAL;
  Ctx* ctx = hops[0]->ctx;
  std::lock_guard<std::mutex> guard(ctx->mu);
  SDB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  uint32_t* d_in = nullptr;
  if (n_frontier) {
    SDB_CUDA(cudaMallocAsync(&d_in, sizeof(uint32_t) * n_frontier, st));
    SDB_CUDA(cudaMemcpyAsync(d_in, frontier, sizeof(uint32_t) * n_frontier, cudaMemcpyHostToDevice, st));
  }
  uint32_t* d_f = nullptr;
  uint64_t n_f = 0;
  sdb_status s = graph_expand_dev(hops, n_hops, d_in, n_frontier, per_source_limit, &d_f, &n_f, st);
  if (d_in) cudaFreeAsync(d_in, st);
  if (s != SDB_OK) return s;
  if (n_f) {
    uint32_t* h_out = (uint32_t*)malloc(sizeof(uint32_t) * n_f);
    if (!h_out) {
      cudaFreeAsync(d_f, st);
      return SDB_ENOMEM;
    }
    // device -> pinned staging (full PCIe rate) -> caller-owned pageable buffer
    const size_t bytes = sizeof(uint32_t) * n_f;
    if (ctx->h_stage_bytes < bytes) {
      if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
      ctx->h_stage = nullptr;
      ctx->h_stage_bytes = 0;
      if (cudaHostAlloc(&ctx->h_stage, bytes, cudaHostAllocDefault) == cudaSuccess) ctx->h_stage_bytes = bytes;
    }
    void* dst = ctx->h_stage_bytes >= bytes ? ctx->h_stage : (void*)h_out;
    SDB_CUDA(cudaMemcpyAsync(dst, d_f, bytes, cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaFreeAsync(d_f, st));
    SDB_CUDA(cudaStreamSynchronize(st));
    if (dst != (void*)h_out) memcpy(h_out, dst, bytes);
    *out_ids = h_out;
    *out_n = n_f;
  } else {
    SDB_CUDA(cudaStreamSynchronize(st));
  }
  return SDB_OK;
}

sdb_status sdb_graph_collect(sdb_graph* g, const uint32_t* start, uint64_t n_start, uint32_t min_depth,
                             uint32_t max_depth, int inclusive, uint32_t** out_ids, uint64_t* out_n) {
  if (!g || !out_ids || !out_n || (n_start && !start)) return SDB_EINVAL;
  *out_ids = nullptr;
  *out_n = 0;
  Ctx* ctx = g->ctx;
  std::lock_guard<std::mutex> guard(ctx->mu);
  SDB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (!g->targets_in_rows) {
    set_error("graph collect: this CSR has targets outside its own rows (an edge table into another node table); "
              "+collect needs source and target ids in one id space");
    return SDB_EINVAL;
  }
  for (uint64_t i = 0; i < n_start; i++)
    if (start[i] >= g->n_rows) {
      set_error("graph collect: start id out of range");
      return SDB_EINVAL;
    }
  // stream-ordered temporaries are released on EVERY return path
  struct Temps {
    cudaStream_t st;
    std::vector<void*> ptrs;
    ~Temps() {
      for (void* p : ptrs)
        if (p) cudaFreeAsync(p, st);
      cudaStreamSynchronize(st);
    }
    void drop(void* p) {
      for (auto& q : ptrs)
        if (q == p) q = nullptr;
      if (p) cudaFreeAsync(p, st);
    }
  } tmp{st, {}};
  auto dalloc = [&](void** p, size_t bytes) -> sdb_status {
    SDB_CUDA(cudaMallocAsync(p, bytes ? bytes : 1, st));
    tmp.ptrs.push_back(*p);
    return SDB_OK;
  };
  const uint64_t nr = g->n_rows ? g->n_rows : 1;
  uint8_t* d_seen = nullptr;
  uint32_t *d_first = nullptr, *d_f = nullptr, *d_res = nullptr;
  // every node is emitted at most once (+ the start values): the result is accumulated on the device
  const uint64_t res_cap = g->n_rows + n_start;
  if (!g->d_seen) SDB_CUDA(cudaMalloc(&g->d_seen, nr));
  if (!g->d_first) SDB_CUDA(cudaMalloc(&g->d_first, sizeof(uint32_t) * nr));
  if (g->res_cap < res_cap) {
    cudaFree(g->d_res);
    g->d_res = nullptr;
    g->res_cap = 0;
    SDB_CUDA(cudaMalloc(&g->d_res, sizeof(uint32_t) * (res_cap ? res_cap : 1)));
    g->res_cap = res_cap;
  }
  d_seen = g->d_seen;
  d_first = g->d_first;
  d_res = g->d_res;
  uint64_t n_res = 0;
  SDB_CUDA(cudaMemsetAsync(d_seen, 0, nr, st));
  SDB_CUDA(cudaMemsetAsync(d_first, 0xFF, sizeof(uint32_t) * nr, st));
  uint64_t n_f = n_start;
  if (n_f) {
    SDB_TRY(dalloc((void**)&d_f, sizeof(uint32_t) * n_f));
    SDB_CUDA(cudaMemcpyAsync(d_f, start, sizeof(uint32_t) * n_f, cudaMemcpyHostToDevice, st));
  }
  if (inclusive && n_start) {  // collect.rs:83-86: the start value is emitted and marked seen only when inclusive
    const uint8_t one = 1;
    for (uint64_t i = 0; i < n_start; i++)
      SDB_CUDA(cudaMemcpyAsync(d_seen + start[i], &one, 1, cudaMemcpyHostToDevice, st));
    SDB_CUDA(cudaMemcpyAsync(d_res, d_f, sizeof(uint32_t) * n_start, cudaMemcpyDeviceToDevice, st));
    n_res = n_start;
    SDB_CUDA(cudaStreamSynchronize(st));  // `one` lives on this stack frame
  }
  uint32_t depth = 0;
  while (n_f && (max_depth == 0 || depth < max_depth)) {
    if (ctx_cancelled(ctx)) {  // polled once per BFS level
      set_error("query cancelled");
      return SDB_ECANCELLED;
    }
    uint32_t* d_lvl = nullptr;
    uint64_t n_lvl = 0;
    SDB_TRY(hop_device(g, d_f, n_f, 0, &d_lvl, &n_lvl, st));
    tmp.ptrs.push_back(d_lvl);
    tmp.drop(d_f);
    d_f = nullptr;
    n_f = 0;
    if (n_lvl) {
      uint64_t* d_pos = nullptr;
      SDB_TRY(dalloc((void**)&d_pos, sizeof(uint64_t) * (n_lvl + 2)));
      const unsigned grid = (unsigned)((n_lvl + 255) / 256);
      collect_mark_kernel<<<grid, 256, 0, st>>>(d_lvl, n_lvl, d_seen, d_first);
      collect_flag_kernel<<<grid, 256, 0, st>>>(d_lvl, n_lvl, d_seen, d_first, d_pos);
      count_launch(ctx, 2);
      SDB_TRY(exclusive_scan(ctx, d_pos, d_pos, n_lvl, d_pos + n_lvl, st));
      uint64_t n_next = 0;
      SDB_CUDA(cudaMemcpyAsync(&n_next, d_pos + n_lvl, 8, cudaMemcpyDeviceToHost, st));
      SDB_CUDA(cudaStreamSynchronize(st));
      if (n_next) {
        SDB_TRY(dalloc((void**)&d_f, sizeof(uint32_t) * n_next));
        collect_compact_kernel<<<grid, 256, 0, st>>>(d_lvl, n_lvl, d_pos, d_seen, d_first, d_f);
        count_launch(ctx);
        n_f = n_next;
        if (depth + 1 >= min_depth) {  // nodes below min_depth are traversed but not emitted
          if (n_res + n_next > res_cap) {
            set_error("graph collect: internal result overflow");
            return SDB_EOVERFLOW;
          }
          SDB_CUDA(cudaMemcpyAsync(d_res + n_res, d_f, sizeof(uint32_t) * n_next, cudaMemcpyDeviceToDevice, st));
          n_res += n_next;
        }
      }
      tmp.drop(d_pos);
    }
    tmp.drop(d_lvl);
    depth++;
  }
  SDB_CUDA(cudaGetLastError());
  if (n_res) {
    uint32_t* h_out = (uint32_t*)malloc(sizeof(uint32_t) * n_res);
    if (!h_out) return SDB_ENOMEM;
    // large results go through the context's pinned staging buffer (pageable D2H is several times slower)
    const size_t bytes = sizeof(uint32_t) * n_res;
    cudaError_t e = cudaSuccess;
    if (bytes >= (1u << 20)) {
      if (ctx->h_stage_bytes < bytes) {
        if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
        ctx->h_stage = nullptr;
        ctx->h_stage_bytes = 0;
        if (cudaHostAlloc(&ctx->h_stage, bytes, cudaHostAllocDefault) == cudaSuccess) ctx->h_stage_bytes = bytes;
      }
    }
    if (ctx->h_stage_bytes >= bytes && bytes >= (1u << 20)) {
      e = cudaMemcpyAsync(ctx->h_stage, d_res, bytes, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e == cudaSuccess) memcpy(h_out, ctx->h_stage, bytes);
    } else {
      e = cudaMemcpyAsync(h_out, d_res, bytes, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (e != cudaSuccess) {
      free(h_out);
      set_error("graph collect: result copy failed: %s", cudaGetErrorString(e));
      return SDB_ECUDA;
    }
    *out_ids = h_out;
    *out_n = n_res;
  }
  return SDB_OK;
}

}  // extern "C"
