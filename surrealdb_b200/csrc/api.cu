// api.cu -- the extern "C" boundary (include/sdbgpu.h): contexts, corpus lifecycle, brute-force KNN driver.
#include <cmath>

#include <chrono>

#include "internal.cuh"

namespace sdb {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}


// schedule ratio of the LEGACY multi-pass schedule (f32 SIMT screen; tensor-core screens when streaming refinement is
// switched off): every pass looks at (R-1) x the rows seen so far; SDB_PASS_RATIO overrides it (tuning only)
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
  PassDesc p0{(uint32_t)stride, 0u, (uint32_t)((T + stride - 1) / stride), 0u};
  v.push_back(p0);
  for (uint64_t s = stride / PASS_RATIO; s >= 1; s /= PASS_RATIO) {
    const uint64_t M = (T + s - 1) / s;
    PassDesc p{(uint32_t)s, (uint32_t)PASS_RATIO, (uint32_t)(M - (M + PASS_RATIO - 1) / PASS_RATIO), 0u};
    v.push_back(p);
    if (s == 1) break;
  }
  return v;
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) {
  while (b) {
    const uint32_t t = a % b;
    a = b;
    b = t;
  }
  return a;
}
// streaming schedule: a PROBE launch scores a few tiles spread over the corpus and keeps only chunk maxima (seed of the
// thresholds), then ONE streaming launch covers every tile, visiting them in a golden-ratio stride order so that any
// stretch of the launch samples the whole corpus (sorted / clustered corpora do not fool the early thresholds).
// Corpora that fit the candidate lists entirely are scored in one pass-0 launch instead.
static void build_stream_passes(uint64_t n_rows, uint32_t cand_cap, uint32_t k, PassDesc* probe, PassDesc* main) {
  const uint64_t T = (n_rows + TILE_ROWS - 1) / TILE_ROWS;
  const uint64_t max0 = cand_cap / TILE_ROWS;
  *probe = PassDesc{1u, 0u, 0u, 0u};
  *main = PassDesc{1u, 0u, 0u, 0u};
  if (T == 0) return;
  if (T <= max0) {  // *probe doubles as the single pass-0 launch: main stays empty
    probe->count = (uint32_t)T;
    return;
  }
  uint64_t P = k <= 32 ? 16 : PROBE_TILES_MAX;  // 8 chunk maxima per tile: 128 / 512 values >= 4 k / 2 k
  if (const char* e = getenv("SDB_PROBE_TILES")) {  // tuning knob: a larger probe starts the thresholds higher
    const int v = atoi(e);
    if (v >= (int)P && v <= (int)PROBE_TILES_MAX) P = (uint64_t)v;
  }
  const uint64_t stride = T / P;                       // T > max0 >= 16; for P = 64 and T < 64 every tile is probed
  if (stride == 0) *probe = PassDesc{1u, 0u, (uint32_t)T, 0u};
  else *probe = PassDesc{(uint32_t)stride, 0u, (uint32_t)P, 0u};
  const uint32_t cnt = (uint32_t)T;
  uint32_t perm = 0;
  if (cnt >= 8 && !getenv("SDB_STREAM_INORDER")) {
    perm = (uint32_t)(cnt * 0.6180339887498949) | 1u;
    while (gcd_u32(perm, cnt) != 1) perm += 2;
    if (perm >= cnt) perm = 0;
  }
  *main = PassDesc{1u, 0u, cnt, perm};
}

struct Rung {
  sdb_screen scr;
  uint32_t cap;  // candidate-list capacity per query
};
// The precision ladder, cheapest first.  The candidate set of a query is "every row whose screened score is within the
// screen's error margin of the k-th best", so its size adapts to the data (a handful on spread-out data, a whole
// cluster on tightly packed data); a rung fails for a query only when that set overflows the list.  When that happens
// to more than a handful of queries the batch is re-screened with a tighter screen / longer lists instead of paying
// one exact pass over the corpus per failed query; the rung that worked is remembered per corpus and k.
static std::vector<Rung> build_rungs(Corpus* c, uint32_t k, sdb_screen* first) {
  const bool int8_ok = c->d_i8 && c->metric == SDB_COSINE && screen_tc_available();
  sdb_screen scr = c->screen;
  if (scr == SDB_SCREEN_AUTO)
    scr = !screen_tc_available() ? SDB_SCREEN_SIMT_F32
                                 : (int8_ok && c->max_rel_qerr <= 0.02f ? SDB_SCREEN_TC_INT8 : SDB_SCREEN_TC_BF16);
  if (scr == SDB_SCREEN_TC_INT8 && !int8_ok) scr = SDB_SCREEN_TC_BF16;
  if (scr == SDB_SCREEN_TC_BF16 && (!screen_tc_available() || !c->d_bf16)) scr = SDB_SCREEN_SIMT_F32;
  const bool screenable = c->metric == SDB_COSINE || c->metric == SDB_EUCLIDEAN;
  if (c->dtype == SDB_F64 || c->special_overflow || k > 256 || !screenable) scr = SDB_SCREEN_NONE_EXACT;
  *first = scr;
  std::vector<Rung> r;
  if (scr == SDB_SCREEN_TC_INT8)
    r = {{SDB_SCREEN_TC_INT8, 4096}, {SDB_SCREEN_TC_INT8, 16384}, {SDB_SCREEN_TC_BF16, 4096}, {SDB_SCREEN_TC_BF16, 16384}};
  else if (scr == SDB_SCREEN_TC_BF16) r = {{SDB_SCREEN_TC_BF16, 4096}, {SDB_SCREEN_TC_BF16, 16384}};
  else if (scr == SDB_SCREEN_SIMT_F32) r = {{SDB_SCREEN_SIMT_F32, 4096}};
  // the f32 stream (error bound ~500x tighter than bf16) as the last rung before the exact kernel -- only ever used
  // for the few queries of a batch that every tensor-core rung failed to prove (finish_local), never for a whole batch
  if (!r.empty() && r.back().scr != SDB_SCREEN_SIMT_F32 && c->dtype == SDB_F32) r.push_back({SDB_SCREEN_SIMT_F32, 4096});
  return r;
}
static uint32_t n_batch_rungs(const std::vector<Rung>& r) {  // rungs a WHOLE batch may be re-screened on
  uint32_t n = (uint32_t)r.size();
  if (n > 1 && r.back().scr == SDB_SCREEN_SIMT_F32) n--;
  return n;
}

// swap the ticket's scratch set into the corpus' active fields (see Scratch in internal.cuh)
static void activate_set(Corpus* c, int set) {
  if (c->active_set == set) return;
  c->sets[c->active_set] = static_cast<Scratch&>(*c);
  static_cast<Scratch&>(*c) = c->sets[set];
  c->active_set = set;
}
static cudaError_t drain(Ctx* ctx) {  // both batch streams idle
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  const cudaError_t e2 = cudaStreamSynchronize(ctx->stream2);
  return e != cudaSuccess ? e : e2;
}

// ---- one batch = enqueue (no host synchronisation) + finish (event wait, ladder, exact fallbacks) ----------------
static sdb_status ticket_prepare(Corpus* c, Ticket& t, uint32_t nq) {
  if (!t.ev_begin) {
    SDB_CUDA(cudaEventCreate(&t.ev_begin));
    SDB_CUDA(cudaEventCreate(&t.ev_screen0));
    SDB_CUDA(cudaEventCreate(&t.ev_screen1));
    SDB_CUDA(cudaEventCreate(&t.ev_end));
    SDB_CUDA(cudaEventCreateWithFlags(&t.ev_h2d, cudaEventDisableTiming));
    SDB_CUDA(cudaEventCreateWithFlags(&t.ev_out, cudaEventDisableTiming));
    SDB_CUDA(cudaEventCreateWithFlags(&t.ev_main, cudaEventDisableTiming));
  }
  if (t.h_cap < nq) {
    if (t.h_flags) cudaFreeHost(t.h_flags);
    t.h_flags = nullptr;
    t.h_cap = 0;
    const uint32_t cap = (nq + 1023) / 1024 * 1024;
    SDB_CUDA(cudaHostAlloc(&t.h_flags, sizeof(uint32_t) * (2 * (size_t)cap + 8), cudaHostAllocDefault));
    t.h_qflags = t.h_flags + cap;
    t.h_stat = t.h_qflags + cap;
    t.h_cap = cap;
  }
  return SDB_OK;
}

bool trace_enabled() {
  static const bool on = getenv("SDB_TRACE") != nullptr;
  return on;
}
void trace_mark(Ctx* ctx, Ticket& t, const char* name, cudaStream_t st) {
  if (!trace_enabled()) return;
  if (!ctx->trace_epoch) {
    cudaEventCreate(&ctx->trace_epoch);
    cudaEventRecord(ctx->trace_epoch, st);
    cudaEventSynchronize(ctx->trace_epoch);
    ctx->trace_host0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  t.trace.emplace_back(name, e);
}
void trace_host(Ctx* ctx, uint32_t ticket, const char* name) {
  if (!trace_enabled() || !ctx->trace_epoch) return;
  const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  fprintf(stderr, "TRACE dev%d ticket%u host %-12s %10.3f\n", ctx->device, ticket, name, (now - ctx->trace_host0) * 1e3);
}
void trace_dump(Ctx* ctx, Ticket& t) {
  if (!trace_enabled()) return;
  for (auto& m : t.trace) {
    float ms = 0.f;
    cudaEventSynchronize(m.second);
    cudaEventElapsedTime(&ms, ctx->trace_epoch, m.second);
    fprintf(stderr, "TRACE dev%d ticket%u set%d %-12s %10.3f\n", ctx->device, t.id, t.set, m.first, ms);
    cudaEventDestroy(m.second);
  }
  t.trace.clear();
}

static sdb_status enqueue_batch(Corpus* c, Ticket& t) {
  Ctx* ctx = c->ctx;
  cudaStream_t st = t.stream;
  activate_set(c, t.set);
  const uint32_t nq = t.nq, k = t.k;
  sdb_screen first;
  const std::vector<Rung> rungs = build_rungs(c, k, &first);
  t.n_rungs = (uint32_t)rungs.size();
  t.n_passes = 0;
  SDB_CUDA(cudaEventRecord(t.ev_begin, st));
  trace_mark(ctx, t, "begin", st);
  if (rungs.empty()) {  // exact-only: the exact kernel needs the prepared queries (f64 copy, |q|, flags)
    t.screen = SDB_SCREEN_NONE_EXACT;
    SDB_TRY(scratch_for(c, nq, 4096));
    SDB_TRY(prep_queries(c, t.d_queries, nq, st));
    SDB_CUDA(cudaEventRecord(t.ev_screen0, st));
    SDB_CUDA(cudaEventRecord(t.ev_screen1, st));
    SDB_CUDA(cudaMemsetAsync(t.d_out_count, 0, sizeof(uint32_t) * nq, st));
    SDB_CUDA(cudaMemcpyAsync(t.h_qflags, c->d_qflags, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, st));
    for (uint32_t q = 0; q < nq; q++) t.h_flags[q] = 2u;  // every query takes the exact kernel
    t.h_stat[0] = nq;
    t.h_stat[1] = t.h_stat[2] = t.h_stat[3] = 0;
    SDB_CUDA(cudaEventRecord(t.ev_end, st));
    return SDB_OK;
  }
  if (t.rung >= rungs.size()) t.rung = (uint32_t)rungs.size() - 1;
  const Rung rg = rungs[t.rung];
  const sdb_screen rs = rg.scr;
  t.screen = (int)rs;
  const bool tc = rs == SDB_SCREEN_TC_INT8 || rs == SDB_SCREEN_TC_BF16;
  const bool int8 = rs == SDB_SCREEN_TC_INT8;
  SDB_TRY(scratch_for(c, nq, rg.cap));
  const uint32_t cap = c->sc_cap;
  SDB_TRY(prep_queries(c, t.d_queries, nq, st));
  SDB_TRY(cand_begin(c, nq, (int)rs, st));
  // Screens are persistent one-CTA-per-SM kernels: two of them in flight on different streams would split the SMs,
  // run in two waves and starve the refiners of the CTAs that are not resident yet.  So the screen of this batch waits
  // for the end of the previous batch's screen -- only the TAIL of the previous batch overlaps with it.
  if (c->last_main && c->last_main != t.ev_main) SDB_CUDA(cudaStreamWaitEvent(st, c->last_main, 0));
  SDB_CUDA(cudaEventRecord(t.ev_screen0, st));  // after the wait: screen_ms is this batch's screen, not the queueing
  trace_mark(ctx, t, "screen0", st);
  if (tc && c->stream_refine) {
    PassDesc p0, pm;
    build_stream_passes(c->n, cap, k, &p0, &pm);
    if (p0.count && !pm.count) {  // the whole corpus fits the lists: score everything once
      SDB_TRY(screen_tc_pass(c, nq, k, p0, int8, 0, st));
      SDB_CUDA(cudaEventRecord(t.ev_main, st));
      c->last_main = t.ev_main;
      SDB_TRY(cand_select(c, nq, k, int8, 0u, false, st));
      t.n_passes++;
    } else if (pm.count) {
      SDB_TRY(screen_tc_pass(c, nq, k, p0, int8, 3, st));           // probe: chunk maxima of a few tiles
      SDB_TRY(cand_seed_from_probe(c, nq, k, p0.count, st));         // thresholds + histogram geometry
      trace_mark(ctx, t, "seeded", st);
      SDB_TRY(screen_tc_pass(c, nq, k, pm, int8, 2, st));           // the streaming launch over every tile
      trace_mark(ctx, t, "main_end", st);
      SDB_CUDA(cudaEventRecord(t.ev_main, st));
      c->last_main = t.ev_main;
      SDB_TRY(cand_select(c, nq, k, int8, c->last_slots, false, st));
      t.n_passes += 2;
    }
  } else {
    const std::vector<PassDesc> passes = build_passes(c->n, cap, nq);
    bool first_pass = true;
    for (const PassDesc& p : passes) {
      if ((t.cancel && *t.cancel) || ctx_cancelled(ctx)) {
        cudaStreamSynchronize(st);
        set_error("query cancelled");
        return SDB_ECANCELLED;
      }
      if (!tc) SDB_TRY(screen_simt_pass(c, nq, p, st));
      else SDB_TRY(screen_tc_pass(c, nq, k, p, int8, first_pass ? 0 : 1, st));
      SDB_TRY(cand_select(c, nq, k, int8, tc ? c->last_slots : 0u, false, st));
      first_pass = false;
      t.n_passes++;
    }
    SDB_CUDA(cudaEventRecord(t.ev_main, st));
    c->last_main = t.ev_main;
  }
  SDB_CUDA(cudaEventRecord(t.ev_screen1, st));
  trace_mark(ctx, t, "selected", st);
  // stage B: the coarse screens' candidates are re-scored in f32 and narrowed before the (FP64-bound) exact re-rank
  static const bool no_refine = getenv("SDB_NO_REFINE") != nullptr;
  bool refined = false;
  if (tc && c->exact && c->dtype == SDB_F32 && !no_refine) {
    SDB_TRY(cand_refine(c, nq, st));
    SDB_TRY(cand_select(c, nq, k, false, 0u, false, st, 1));
    refined = true;
    trace_mark(ctx, t, "refined", st);
  }
  SDB_TRY(cand_rerank(c, nq, st, refined));
  trace_mark(ctx, t, "reranked", st);
  SDB_TRY(cand_final(c, nq, k, t.row_base, t.d_out_rows, t.d_out_dist, t.d_out_count, st));
  trace_mark(ctx, t, "final", st);
  SDB_CUDA(cudaMemcpyAsync(t.h_flags, c->d_flags, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaMemcpyAsync(t.h_qflags, c->d_qflags, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaMemcpyAsync(t.h_stat, c->d_stat, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaEventRecord(t.ev_end, st));
  trace_mark(ctx, t, "end", st);
  return SDB_OK;
}

static sdb_status copy_out(Corpus* c, Ticket& t) {  // host-buffer entry points: device result -> caller's buffers
  cudaStream_t st = t.stream;
  if (!t.h_out_count) return SDB_OK;
  if (t.k) {
    SDB_CUDA(cudaMemcpyAsync(t.h_out_rows, t.d_out_rows, sizeof(uint64_t) * (size_t)t.nq * t.k, cudaMemcpyDeviceToHost, st));
    SDB_CUDA(cudaMemcpyAsync(t.h_out_dist, t.d_out_dist, sizeof(double) * (size_t)t.nq * t.k, cudaMemcpyDeviceToHost, st));
  }
  SDB_CUDA(cudaMemcpyAsync(t.h_out_count, t.d_out_count, sizeof(uint32_t) * t.nq, cudaMemcpyDeviceToHost, st));
  SDB_CUDA(cudaEventRecord(t.ev_out, st));  // wait() blocks on THIS batch's copies, not on whatever was queued behind it
  return SDB_OK;
}

// local (this shard's) part of the completion: ladder re-runs and exact fallbacks.  *repaired = the device result
// changed after the batch's own kernels had produced it.
static sdb_status finish_local(Corpus* c, Ticket& t, uint32_t* n_fallback, bool* repaired) {
  Ctx* ctx = c->ctx;
  cudaStream_t st = t.stream;
  const uint32_t nq = t.nq, k = t.k;
  *repaired = false;
  *n_fallback = 0;
  SDB_CUDA(cudaEventSynchronize(t.ev_end));
  if (t.screen != SDB_SCREEN_NONE_EXACT && c->exact) {
    while (t.rung + 1 < t.n_batch_rungs) {  // many failures: the whole batch moves up one rung (and stays there)
      uint32_t n_fail = 0;
      for (uint32_t q = 0; q < nq; q++) n_fail += (t.h_flags[q] & 2u) ? 1u : 0u;
      if (n_fail <= 2 + nq / 64) break;
      SDB_CUDA(drain(ctx));  // later batches in flight share scratch sets and the ladder state: drain them first
      t.rung++;
      SDB_TRY(enqueue_batch(c, t));
      SDB_CUDA(cudaEventSynchronize(t.ev_end));
      *repaired = true;
    }
  }
  if (t.n_rungs) {
    sdb_screen first;
    build_rungs(c, k, &first);
    c->rung_scr = first;
    c->rung_k = k;
    c->rung = t.rung;
  }
  // ---- what the batch's rung could not prove ----
  std::vector<uint32_t> fails, exacts;
  for (uint32_t q = 0; q < nq; q++) {
    if (t.h_qflags[q] & 1u) exacts.push_back(q);  // zero / non-finite query norm: ranked by the exact kernel
    else if ((t.h_flags[q] & 2u) && (c->exact || t.screen == SDB_SCREEN_NONE_EXACT)) fails.push_back(q);
  }
  if (fails.empty() && exacts.empty()) return SDB_OK;
  SDB_CUDA(drain(ctx));  // the repair below shares scratch (and the exact kernel's keys) with every batch in flight
  // A few failures: only THOSE queries climb the remaining rungs, as a small batch of their own (a bf16 pass over the
  // corpus costs about as much for 60 queries as for 1, and far less than one sequential-f64 pass per query); the f32
  // stream is the last rung.  Whatever is still unproven after that goes to the exact kernel.
  if (!fails.empty() && t.screen != SDB_SCREEN_NONE_EXACT && c->exact && k) {
    const uint32_t save_rung = t.rung, save_nq = t.nq, save_passes = t.n_passes;
    const int save_screen = t.screen;
    const double* save_q = t.d_queries;
    uint64_t* save_rows = t.d_out_rows;
    double* save_dist = t.d_out_dist;
    uint32_t* save_cnt = t.d_out_count;
    uint32_t save_stat[4] = {t.h_stat[0], t.h_stat[1], t.h_stat[2], t.h_stat[3]};
    sdb_status rc = SDB_OK;
    for (uint32_t rung = save_rung + 1; rung < t.n_rungs && !fails.empty() && rc == SDB_OK; rung++) {
      if ((t.cancel && *t.cancel) || ctx_cancelled(ctx)) break;
      const uint32_t nf = (uint32_t)fails.size();
      const size_t need_q = (size_t)nf * c->dim, need_o = (size_t)nf * k;
      if (c->rp_cap_q < need_q || c->rp_cap_o < need_o || c->rp_cap_n < nf) {
        cudaFree(c->d_rp_q); cudaFree(c->d_rp_rows); cudaFree(c->d_rp_dist); cudaFree(c->d_rp_cnt);
        c->d_rp_q = nullptr; c->d_rp_rows = nullptr; c->d_rp_dist = nullptr; c->d_rp_cnt = nullptr;
        c->rp_cap_q = c->rp_cap_o = c->rp_cap_n = 0;
        if (cudaMalloc(&c->d_rp_q, sizeof(double) * need_q) != cudaSuccess || cudaMalloc(&c->d_rp_rows, sizeof(uint64_t) * need_o) != cudaSuccess ||
            cudaMalloc(&c->d_rp_dist, sizeof(double) * need_o) != cudaSuccess || cudaMalloc(&c->d_rp_cnt, sizeof(uint32_t) * nf) != cudaSuccess) {
          set_error("repair buffers: %s", cudaGetErrorString(cudaGetLastError()));
          rc = SDB_ENOMEM;
          break;
        }
        c->rp_cap_q = need_q; c->rp_cap_o = need_o; c->rp_cap_n = nf;
      }
      for (uint32_t i = 0; i < nf; i++)
        cudaMemcpyAsync(c->d_rp_q + (size_t)i * c->dim, save_q + (size_t)fails[i] * c->dim, sizeof(double) * c->dim, cudaMemcpyDeviceToDevice, st);
      t.d_queries = c->d_rp_q;
      t.nq = nf;
      t.d_out_rows = c->d_rp_rows;
      t.d_out_dist = c->d_rp_dist;
      t.d_out_count = c->d_rp_cnt;
      t.rung = rung;
      rc = enqueue_batch(c, t);
      if (rc == SDB_OK && cudaEventSynchronize(t.ev_end) != cudaSuccess) rc = SDB_ECUDA;
      if (rc != SDB_OK) break;
      std::vector<uint32_t> still;
      for (uint32_t i = 0; i < nf; i++) {
        const uint32_t q = fails[i];
        if ((t.h_flags[i] & 2u) || (t.h_qflags[i] & 1u)) {
          still.push_back(q);
          continue;
        }
        cudaMemcpyAsync(save_rows + (size_t)q * k, c->d_rp_rows + (size_t)i * k, sizeof(uint64_t) * k, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(save_dist + (size_t)q * k, c->d_rp_dist + (size_t)i * k, sizeof(double) * k, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(save_cnt + q, c->d_rp_cnt + i, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
        t.n_repaired++;
      }
      SDB_CUDA(cudaStreamSynchronize(st));
      fails.swap(still);
      *repaired = true;
    }
    t.d_queries = save_q;
    t.nq = save_nq;
    t.d_out_rows = save_rows;
    t.d_out_dist = save_dist;
    t.d_out_count = save_cnt;
    t.rung = save_rung;
    t.screen = save_screen;
    t.n_passes = save_passes;
    for (int i = 0; i < 4; i++) t.h_stat[i] = save_stat[i];
    SDB_TRY(rc);
  }
  // ---- exact kernel: special queries and whatever no screen could prove ----
  exacts.insert(exacts.end(), fails.begin(), fails.end());
  for (uint32_t q : exacts) {
    if ((t.cancel && *t.cancel) || ctx_cancelled(ctx)) {
      cudaStreamSynchronize(st);
      set_error("query cancelled");
      return SDB_ECANCELLED;
    }
    SDB_TRY(prep_fallback_query(c, t.d_queries + (size_t)q * c->dim, st));
    SDB_TRY(exact_query(c, c->d_fb_q, c->d_fb_qmag, c->d_fb_qflags, k, t.row_base, t.d_out_rows + (size_t)q * k,
                        t.d_out_dist + (size_t)q * k, t.d_out_count + q, st));
    (*n_fallback)++;
    *repaired = true;
  }
  if (*repaired) SDB_CUDA(cudaStreamSynchronize(st));
  return SDB_OK;
}

static sdb_status finish_stats(Corpus* c, Ticket& t, uint32_t n_fallback) {
  sdb_knn_stats stt{};
  SDB_CUDA(cudaEventElapsedTime(&stt.screen_ms, t.ev_screen0, t.ev_screen1));
  SDB_CUDA(cudaEventElapsedTime(&stt.total_ms, t.ev_begin, t.ev_end));
  stt.screen_used = (uint32_t)t.screen;
  stt.n_passes = t.n_passes;
  stt.n_fallback = n_fallback;
  stt.n_repaired = t.n_repaired;
  stt.n_special_rows = c->n_special;
  stt.n_candidates = t.h_stat[2];  // largest candidate set of the batch
  stt.n_reranked = t.h_stat[1];
  stt.n_survivors = t.h_stat[3];
  stt.kernel_launches = c->ctx->launches - t.launches0;
  c->stats = stt;
  trace_host(c->ctx, t.id, "waited");
  trace_dump(c->ctx, t);
  return SDB_OK;
}

static Ticket* find_ticket(Corpus* c, uint32_t id) {
  for (Ticket& t : c->tickets)
    if (t.busy && t.id == id) return &t;
  return nullptr;
}
static Ticket* free_ticket(Corpus* c) {
  for (Ticket& t : c->tickets)
    if (!t.busy) return &t;
  return nullptr;
}

static sdb_status submit_locked(Corpus* c, Ticket* t, const double* d_queries, uint32_t nq, uint32_t k, uint64_t row_base,
                                uint64_t* d_out_rows, double* d_out_dist, uint32_t* d_out_count,
                                const volatile int* cancel) {
  if (!c->finalized) {
    set_error("corpus not finalized (call sdb_corpus_finalize after the last append)");
    return SDB_EINVAL;
  }
  if ((cancel && *cancel) || ctx_cancelled(c->ctx)) {  // the poll of knn_topk.rs:186 before any work is queued
    set_error("query cancelled");
    return SDB_ECANCELLED;
  }
  SDB_TRY(ticket_prepare(c, *t, nq ? nq : 1));
  {  // slot parity picks the stream and the scratch set: consecutive batches overlap (screen of i+1 || tail of i)
    const int slot = (int)(t - c->tickets);
    const bool one_stream = getenv("SDB_ONE_STREAM") != nullptr;  // (A/B knob)
    t->set = one_stream ? 0 : (slot & 1);
    t->stream = t->set ? c->ctx->stream2 : c->ctx->stream;
    if (t->wait_h2d) SDB_CUDA(cudaStreamWaitEvent(t->stream, t->ev_h2d, 0));
    t->wait_h2d = false;
  }
  t->id = c->next_ticket++;
  if (c->next_ticket == 0) c->next_ticket = 1;
  t->d_queries = d_queries;
  t->nq = nq;
  t->k = k;
  t->row_base = row_base;
  t->d_out_rows = d_out_rows;
  t->d_out_dist = d_out_dist;
  t->d_out_count = d_out_count;
  t->cancel = cancel;
  t->launches0 = c->ctx->launches;
  t->n_repaired = 0;
  sdb_screen first;
  const std::vector<Rung> rungs = build_rungs(c, k, &first);
  t->rung = (c->rung_scr == first && c->rung_k == k && c->rung < n_batch_rungs(rungs)) ? c->rung : 0;
  t->n_rungs = (uint32_t)rungs.size();
  t->n_batch_rungs = n_batch_rungs(rungs);
  if (nq == 0 || k == 0) {  // nothing to search: counts are zero
    cudaStream_t st = t->stream;
    SDB_CUDA(cudaEventRecord(t->ev_begin, st));
    SDB_CUDA(cudaEventRecord(t->ev_screen0, st));
    SDB_CUDA(cudaEventRecord(t->ev_screen1, st));
    if (nq) SDB_CUDA(cudaMemsetAsync(d_out_count, 0, sizeof(uint32_t) * nq, st));
    for (uint32_t q = 0; q < nq; q++) t->h_flags[q] = t->h_qflags[q] = 0;
    t->h_stat[0] = t->h_stat[1] = t->h_stat[2] = t->h_stat[3] = 0;
    t->screen = SDB_SCREEN_NONE_EXACT;
    t->n_rungs = 0;
    t->n_passes = 0;
    SDB_CUDA(cudaEventRecord(t->ev_end, st));
    t->busy = true;
    return SDB_OK;
  }
  const sdb_status rc = enqueue_batch(c, *t);
  if (rc == SDB_OK) t->busy = true;
  trace_host(c->ctx, t->id, "submitted");
  return rc;
}

static sdb_status wait_locked(Corpus* c, Ticket* t) {
  uint32_t n_fb = 0;
  bool repaired = false;
  sdb_status rc = finish_local(c, *t, &n_fb, &repaired);
  if (rc == SDB_OK && t->h_out_count) {
    if (repaired) rc = copy_out(c, *t);  // the copies enqueued at submit time predate the repair
    if (rc == SDB_OK && cudaEventSynchronize(t->ev_out) != cudaSuccess) {
      set_error("sdb_knn_wait: %s", cudaGetErrorString(cudaGetLastError()));
      rc = SDB_ECUDA;
    }
  }
  if (rc == SDB_OK) rc = finish_stats(c, *t, n_fb);
  t->busy = false;
  t->h_out_rows = nullptr;
  t->h_out_dist = nullptr;
  t->h_out_count = nullptr;
  return rc;
}

// ---- hooks for comm.cu (sharded search).  The caller holds c->mu. ---------------------------------------------------
sdb_status knn_submit_for_shard(Corpus* c, const double* d_queries, const double* h_queries, uint32_t nq, uint32_t k,
                                uint64_t* d_out_rows, double* d_out_dist, uint32_t* d_out_count, int* slot_index,
                                uint32_t* ticket, const double** d_queries_used) {
  Ticket* t = free_ticket(c);
  if (!t) {
    set_error("too many batches in flight (%d): call the matching wait first", N_TICKETS);
    return SDB_EOVERFLOW;
  }
  *slot_index = (int)(t - c->tickets);
  bool h2d_pending = false;
  if (h_queries) {  // host queries: staged through the slot's device buffer on the copy stream
    SDB_TRY(ticket_prepare(c, *t, nq));
    const size_t need_q = (size_t)nq * c->dim;
    if (t->in_cap < need_q) {
      cudaFree(t->d_in_q);
      t->d_in_q = nullptr;
      t->in_cap = 0;
      SDB_CUDA(cudaMalloc(&t->d_in_q, sizeof(double) * need_q));
      t->in_cap = need_q;
    }
    cudaStream_t cs = c->ctx->copy_stream;
    SDB_CUDA(cudaMemcpyAsync(t->d_in_q, h_queries, sizeof(double) * need_q, cudaMemcpyHostToDevice, cs));
    SDB_CUDA(cudaEventRecord(t->ev_h2d, cs));
    d_queries = t->d_in_q;
    h2d_pending = true;
  }
  t->wait_h2d = h2d_pending;
  SDB_TRY(submit_locked(c, t, d_queries, nq, k, c->row_base, d_out_rows, d_out_dist, d_out_count, nullptr));
  *ticket = t->id;
  *d_queries_used = d_queries;
  return SDB_OK;
}
void knn_trace_mark(Corpus* c, uint32_t ticket, const char* name) {
  Ticket* t = find_ticket(c, ticket);
  if (t) trace_mark(c->ctx, *t, name, t->stream);
}
cudaStream_t knn_ticket_stream(Corpus* c, uint32_t ticket) {
  Ticket* t = find_ticket(c, ticket);
  return t ? t->stream : c->ctx->stream;
}
sdb_status knn_finish_for_shard(Corpus* c, uint32_t ticket, bool* repaired) {
  Ticket* t = find_ticket(c, ticket);
  if (!t) return SDB_EINVAL;
  uint32_t n_fb = 0;
  SDB_TRY(finish_local(c, *t, &n_fb, repaired));
  return finish_stats(c, *t, n_fb);
}
sdb_status knn_release_ticket(Corpus* c, uint32_t ticket) {
  Ticket* t = find_ticket(c, ticket);
  if (!t) return SDB_EINVAL;
  t->busy = false;
  t->h_out_rows = nullptr;
  t->h_out_dist = nullptr;
  t->h_out_count = nullptr;
  return SDB_OK;
}
const uint32_t* knn_ticket_stat_host(Corpus* c, uint32_t ticket, int* exact_only) {
  Ticket* t = find_ticket(c, ticket);
  if (!t) return nullptr;
  *exact_only = (t->n_rungs == 0) ? 1 : 0;
  return t->h_stat;
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


// <= 32 lists: one warp per query, lane l walks list l (each list is already in (distance, row) order).  Every step takes
// the warp-wide minimum head.  No shared memory and 128-thread blocks, so the merge runs beside the resident screen of
// the next batch (the sorter above needs 24 bytes of shared memory per entry and up to 1024 threads).
__global__ void __launch_bounds__(128) topk_kway_merge_kernel(uint32_t n_lists, uint32_t nq, uint32_t k,
                                                              const uint64_t* __restrict__ rows,
                                                              const double* __restrict__ dist,
                                                              const uint32_t* __restrict__ counts, uint64_t st_rows,
                                                              uint64_t st_dist, uint64_t st_cnt,
                                                              uint64_t* __restrict__ out_rows, double* __restrict__ out_dist,
                                                              uint32_t* __restrict__ out_count) {
  const uint32_t q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if (q >= nq) return;
  uint32_t cnt = 0, pos = 0;
  const uint64_t* lr = nullptr;
  const double* ld = nullptr;
  if (lane < n_lists) {
    cnt = counts[(size_t)lane * st_cnt + q];
    if (cnt > k) cnt = k;
    lr = rows + (size_t)lane * st_rows + (size_t)q * k;
    ld = dist + (size_t)lane * st_dist + (size_t)q * k;
  }
  uint64_t key = ~0ull, row = ~0ull;
  double d = 0.0;
  if (pos < cnt) {
    d = ld[0];
    key = dist_key(d);
    row = lr[0];
  }
  uint32_t n_out = 0;
  for (; n_out < k; n_out++) {
    uint64_t bk = key, br = row;
    uint32_t bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t ok = __shfl_xor_sync(0xffffffffu, bk, o), orow = __shfl_xor_sync(0xffffffffu, br, o);
      const uint32_t ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ok < bk || (ok == bk && (orow < br || (orow == br && ol < bl)))) {
        bk = ok;
        br = orow;
        bl = ol;
      }
    }
    if (bk == ~0ull && br == ~0ull) break;  // every list exhausted
    if (lane == bl) {
      out_rows[(size_t)q * k + n_out] = row;
      out_dist[(size_t)q * k + n_out] = d;
      pos++;
      if (pos < cnt) {
        d = ld[pos];
        key = dist_key(d);
        row = lr[pos];
      } else {
        key = ~0ull;
        row = ~0ull;
      }
    }
  }
  if (lane == 0) out_count[q] = n_out;
}

sdb_status topk_merge_launch(Ctx* ctx, uint32_t n_lists, uint32_t nq, uint32_t k, const uint64_t* d_rows,
                             const double* d_dist, const uint32_t* d_counts, uint64_t stride_rows, uint64_t stride_dist,
                             uint64_t stride_counts, uint64_t* d_out_rows, double* d_out_dist, uint32_t* d_out_count,
                             cudaStream_t st) {
  if (!stride_rows) stride_rows = (uint64_t)nq * k;
  if (!stride_dist) stride_dist = (uint64_t)nq * k;
  if (!stride_counts) stride_counts = nq;
  if (n_lists <= 32) {
    topk_kway_merge_kernel<<<(nq + 3) / 4, 128, 0, st>>>(n_lists, nq, k, d_rows, d_dist, d_counts, stride_rows, stride_dist,
                                                       stride_counts, d_out_rows, d_out_dist, d_out_count);
    count_launch(ctx);
    SDB_CUDA(cudaGetLastError());
    return SDB_OK;
  }
  uint32_t p2 = 1;
  while (p2 < n_lists * k) p2 <<= 1;
  const size_t smem = (size_t)p2 * 24;
  if (smem > 200 * 1024) {
    set_error("merge of %u lists x k=%u exceeds the shared-memory sorter", n_lists, k);
    return SDB_EUNSUPPORTED;
  }
  const uint32_t threads = p2 >= 1024 ? 1024 : (p2 < 64 ? 64 : p2);
  topk_merge_kernel<<<nq, threads, smem, st>>>(n_lists, nq, k, d_rows, d_dist, d_counts, stride_rows, stride_dist,
                                               stride_counts, d_out_rows, d_out_dist, d_out_count);
  count_launch(ctx);
  SDB_CUDA(cudaGetLastError());
  return SDB_OK;
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
  {
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) c->prio_high = greatest;
  }
  SDB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  SDB_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  SDB_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  {
    int* hc = nullptr;
    SDB_CUDA(cudaHostAlloc(&hc, sizeof(int), cudaHostAllocMapped));
    *hc = 0;
    c->h_cancel = hc;
    SDB_CUDA(cudaMalloc(&c->d_cancel, sizeof(int)));
    SDB_CUDA(cudaMemset(c->d_cancel, 0, sizeof(int)));
    SDB_CUDA(cudaStreamCreateWithFlags(&c->cancel_stream, cudaStreamNonBlocking));
  }
  // dynamic shared-memory limits are per device: set them for THIS device now (not behind a process-wide flag)
  SDB_TRY(screen_tc_init_device());
  SDB_TRY(candidates_init_device());
  SDB_TRY(exact_init_device());
  SDB_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
  comm_destroy(c);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->h_cancel) cudaFreeHost((void*)c->h_cancel);
  if (c->d_cancel) cudaFree(c->d_cancel);
  if (c->cancel_stream) cudaStreamDestroy(c->cancel_stream);
  delete c;
}
static void push_cancel_word(sdb_ctx* c) {  // any host thread; its own stream, so it overtakes running kernels
  if (!c->d_cancel || !c->cancel_stream) return;
  if (cudaSetDevice(c->device) != cudaSuccess) return;
  cudaMemcpyAsync(c->d_cancel, (const void*)c->h_cancel, sizeof(int), cudaMemcpyHostToDevice, c->cancel_stream);
  cudaStreamSynchronize(c->cancel_stream);
}
void sdb_ctx_cancel(sdb_ctx* c) {
  if (!c || !c->h_cancel) return;
  *c->h_cancel = 1;
  push_cancel_word(c);
}
void sdb_ctx_cancel_reset(sdb_ctx* c) {
  if (!c || !c->h_cancel) return;
  *c->h_cancel = 0;
  push_cancel_word(c);
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
  // COSINE / EUCLIDEAN: screened (K1/K2) + exact re-rank.  MANHATTAN / CHEBYSHEV / HAMMING / PEARSON / JACCARD /
  // MINKOWSKI: served by the exact kernel alone (sequential f64, Distance::compute op for op).  MINKOWSKI goes through
  // pow(), which CUDA's libm and Rust's (the platform libm) implement separately: within 1 ulp of each other per term,
  // so its distances are compared with a 1e-12 relative tolerance instead of bit equality (tests/test_gpu_knn.py).
  const bool screenable = m == SDB_COSINE || m == SDB_EUCLIDEAN;
  if ((int)m < 0 || (int)m > (int)SDB_PEARSON) {
    set_error("unknown metric %d", (int)m);
    return SDB_EINVAL;
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
  drain(c->ctx);
  comm_corpus_released(c);
  for (int si = 0; si < 2; si++) {  // the parked scratch set
    if (si == c->active_set) continue;
    Scratch& z = c->sets[si];
    void* sp[] = {z.d_q64, z.d_q32, z.d_qbf16, z.d_qmag, z.d_qflags, z.d_qbferr, z.d_q8, z.d_q8scale, z.d_q8err, z.d_sub,
                  z.d_sub_cnt, z.d_bscale, z.d_beps, z.d_margin, z.d_margin2, z.d_beps2, z.d_tau2, z.d_qlow, z.d_qcap,
                  z.d_hparam, z.d_hist, z.d_probe, z.d_tau, z.d_cand, z.d_cand_cnt, z.d_flags, z.d_stat, z.d_rr_key,
                  z.d_rr_dist, z.d_rr_row};
    for (void* p : sp) cudaFree(p);
  }
  void* ptrs[] = {c->d_i8, c->d_q8, c->d_q8scale, c->d_q8err, c->d_bscale, c->d_beps, c->d_margin, c->d_qlow, c->d_qcap,
                  c->d_hparam, c->d_hist, c->d_qbferr, c->d_stat, c->d_probe, c->d_margin2, c->d_beps2, c->d_tau2, c->d_sub, c->d_sub_cnt, c->d_rows, c->d_mag, c->d_snorm,
                  c->d_bf16, c->d_skip, c->d_removed, c->d_special, c->d_q64, c->d_q32, c->d_qbf16, c->d_qmag, c->d_qflags,
                  c->d_tau, c->d_cand, c->d_cand_cnt, c->d_flags, c->d_rr_key, c->d_rr_dist, c->d_rr_row, c->d_ex_key,
                  c->d_sel, c->d_fb_q, c->d_fb_qmag, c->d_fb_qflags, c->d_block, c->d_gather, c->d_rp_q, c->d_rp_rows,
                  c->d_rp_dist, c->d_rp_cnt};
  for (void* p : ptrs) cudaFree(p);
  for (Ticket& t : c->tickets) {
    cudaEvent_t evs[] = {t.ev_begin, t.ev_screen0, t.ev_screen1, t.ev_end, t.ev_h2d, t.ev_out, t.ev_main};
    for (cudaEvent_t e : evs)
      if (e) cudaEventDestroy(e);
    if (t.h_flags) cudaFreeHost(t.h_flags);
    cudaFree(t.d_in_q);
    cudaFree(t.d_res_rows);
    cudaFree(t.d_res_dist);
    cudaFree(t.d_res_count);
  }
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
  }
  SDB_TRY(corpus_reapply_tombstones(c, c->ctx->stream));  // rows removed earlier stay removed
  SDB_CUDA(cudaStreamSynchronize(c->ctx->stream));
  c->finalized = false;
  return SDB_OK;
}
sdb_status sdb_corpus_remove(sdb_corpus* c, const uint64_t* row_ids, uint64_t n) {
  if (!c || (n && !row_ids)) return SDB_EINVAL;
  if (n == 0) return SDB_OK;
  std::lock_guard<std::mutex> g(c->mu);
  for (uint64_t i = 0; i < n; i++)
    if (row_ids[i] >= c->n) {
      set_error("sdb_corpus_remove: row %llu outside the corpus (%llu rows)", (unsigned long long)row_ids[i],
                (unsigned long long)c->n);
      return SDB_EINVAL;
    }
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  SDB_CUDA(drain(c->ctx));  // no batch may be in flight while rows disappear
  SDB_TRY(corpus_remove_device(c, row_ids, n));
  c->n_removed += n;
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
sdb_status sdb_corpus_read_rows(sdb_corpus* c, uint64_t first_row, uint64_t n, void* out) {
  if (!c || (!out && n)) return SDB_EINVAL;
  std::lock_guard<std::mutex> g(c->mu);
  if (first_row > c->n || n > c->n - first_row) {
    set_error("sdb_corpus_read_rows: rows %llu..%llu outside the corpus (%llu rows)", (unsigned long long)first_row,
              (unsigned long long)(first_row + n), (unsigned long long)c->n);
    return SDB_EINVAL;
  }
  if (n == 0) return SDB_OK;
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  const size_t esz = c->dtype == SDB_F32 ? 4 : 8;
  SDB_CUDA(cudaMemcpyAsync(out, (const char*)c->d_rows + esz * first_row * c->dim, esz * n * c->dim,
                           cudaMemcpyDeviceToHost, c->ctx->copy_stream));
  SDB_CUDA(cudaStreamSynchronize(c->ctx->copy_stream));
  return SDB_OK;
}
sdb_status sdb_corpus_set_minkowski_order(sdb_corpus* c, double order) {
  if (!c || !(order == order)) return SDB_EINVAL;
  c->minkowski_p = order;
  return SDB_OK;
}
sdb_status sdb_corpus_set_schedule(sdb_corpus* c, int streaming) {
  if (!c) return SDB_EINVAL;
  c->stream_refine = streaming != 0;
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

// ---- asynchronous batches ------------------------------------------------------------------------------------------
sdb_status sdb_knn_submit_device(sdb_corpus* c, const double* d_queries, uint32_t nq, uint32_t k, uint64_t row_base,
                                 uint64_t* d_out_rows, double* d_out_dist, uint32_t* d_out_count, uint32_t* ticket) {
  if (!c || !ticket || (nq && (!d_queries || !d_out_count || (k && (!d_out_rows || !d_out_dist))))) return SDB_EINVAL;
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  Ticket* t = free_ticket(c);
  if (!t) {
    set_error("too many batches in flight (%d): call sdb_knn_wait first", N_TICKETS);
    return SDB_EOVERFLOW;
  }
  SDB_TRY(submit_locked(c, t, d_queries, nq, k, row_base, d_out_rows, d_out_dist, d_out_count, nullptr));
  *ticket = t->id;
  return SDB_OK;
}

static sdb_status ensure_slot_buffers(Corpus* c, Ticket* t, uint32_t nq, uint32_t k) {
  const size_t need_q = (size_t)nq * c->dim, need = (size_t)nq * (k ? k : 1);
  if (t->in_cap < need_q) {
    cudaFree(t->d_in_q);
    t->d_in_q = nullptr;
    t->in_cap = 0;
    SDB_CUDA(cudaMalloc(&t->d_in_q, sizeof(double) * need_q));
    t->in_cap = need_q;
  }
  if (t->res_cap < need || t->res_cap_q < nq) {
    cudaFree(t->d_res_rows);
    cudaFree(t->d_res_dist);
    cudaFree(t->d_res_count);
    t->d_res_rows = nullptr; t->d_res_dist = nullptr; t->d_res_count = nullptr;
    t->res_cap = t->res_cap_q = 0;
    SDB_CUDA(cudaMalloc(&t->d_res_rows, sizeof(uint64_t) * need));
    SDB_CUDA(cudaMalloc(&t->d_res_dist, sizeof(double) * need));
    SDB_CUDA(cudaMalloc(&t->d_res_count, sizeof(uint32_t) * nq));
    t->res_cap = need;
    t->res_cap_q = nq;
  }
  return SDB_OK;
}

static sdb_status submit_host_locked(sdb_corpus* c, const double* queries, uint32_t nq, uint32_t k, uint64_t* out_rows,
                                     double* out_dist, uint32_t* out_count, const volatile int* cancel, Ticket** out_t) {
  Ticket* t = free_ticket(c);
  if (!t) {
    set_error("too many batches in flight (%d): call sdb_knn_wait first", N_TICKETS);
    return SDB_EOVERFLOW;
  }
  SDB_TRY(ticket_prepare(c, *t, nq));
  SDB_TRY(ensure_slot_buffers(c, t, nq, k));
  // the queries travel on the copy stream, so the transfer of batch i+1 overlaps the kernels of batch i
  cudaStream_t cs = c->ctx->copy_stream;
  SDB_CUDA(cudaMemcpyAsync(t->d_in_q, queries, sizeof(double) * (size_t)nq * c->dim, cudaMemcpyHostToDevice, cs));
  SDB_CUDA(cudaEventRecord(t->ev_h2d, cs));
  t->wait_h2d = true;  // submit_locked makes the batch's stream wait for the transfer
  SDB_TRY(submit_locked(c, t, t->d_in_q, nq, k, c->row_base, t->d_res_rows, t->d_res_dist, t->d_res_count, cancel));
  t->h_out_rows = out_rows;
  t->h_out_dist = out_dist;
  t->h_out_count = out_count;
  const sdb_status rc = copy_out(c, *t);
  if (rc != SDB_OK) {
    cudaStreamSynchronize(t->stream);
    t->busy = false;
    return rc;
  }
  *out_t = t;
  return SDB_OK;
}

// Host-only diagnostic (no GPU needed): the corpus tiles (TILE_ROWS rows each) a screened search visits, in order.
// Every row that is not visited can never become a candidate and the exactness proof would not know, so "every tile
// exactly once" is a safety property of the schedules; tests/test_schedule_cover.py checks it exhaustively on the CPU.
sdb_status sdb_debug_schedule(uint64_t n_rows, uint32_t cand_cap, uint32_t k, uint32_t nq, int streaming,
                              uint32_t* out_tiles, uint64_t cap_tiles, uint64_t* out_n, uint32_t* out_probe_tiles,
                              uint32_t cap_probe, uint32_t* out_n_probe) {
  if (!out_n || !out_n_probe || cand_cap < TILE_ROWS) return SDB_EINVAL;
  uint64_t n = 0;
  uint32_t np = 0;
  auto emit = [&](uint32_t tile) {
    if (out_tiles && n < cap_tiles) out_tiles[n] = tile;
    n++;
  };
  if (streaming) {
    PassDesc p0, pm;
    build_stream_passes(n_rows, cand_cap, k, &p0, &pm);
    if (p0.count && !pm.count) {
      for (uint32_t i = 0; i < p0.count; i++) emit(pass_tile(p0, i));  // pass 0 scores everything
    } else {
      for (uint32_t i = 0; i < p0.count; i++) {
        if (out_probe_tiles && np < cap_probe) out_probe_tiles[np] = pass_tile(p0, i);
        np++;
      }
      for (uint32_t i = 0; i < pm.count; i++) emit(pass_tile(pm, i));
    }
  } else {
    for (const PassDesc& p : build_passes(n_rows, cand_cap, nq))
      for (uint32_t i = 0; i < p.count; i++) emit(pass_tile(p, i));
  }
  *out_n = n;
  *out_n_probe = np;
  return SDB_OK;
}

sdb_status sdb_knn_submit(sdb_corpus* c, const double* queries, uint32_t nq, uint32_t k, uint64_t* out_rows,
                          double* out_dist, uint32_t* out_count, uint32_t* ticket) {
  if (!c || !ticket || !nq || !queries || !out_count || (k && (!out_rows || !out_dist))) return SDB_EINVAL;
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  Ticket* t = nullptr;
  SDB_TRY(submit_host_locked(c, queries, nq, k, out_rows, out_dist, out_count, nullptr, &t));
  *ticket = t->id;
  return SDB_OK;
}

sdb_status sdb_knn_wait(sdb_corpus* c, uint32_t ticket) {
  if (!c) return SDB_EINVAL;
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  Ticket* t = find_ticket(c, ticket);
  if (!t) {
    set_error("sdb_knn_wait: unknown or already completed ticket %u", ticket);
    return SDB_EINVAL;
  }
  return wait_locked(c, t);
}

sdb_status sdb_knn_bruteforce_device(sdb_corpus* c, const double* d_queries, uint32_t nq, uint32_t k,
                                     uint64_t row_base, uint64_t* d_out_rows, double* d_out_dist,
                                     uint32_t* d_out_count) {
  if (!c || (nq && (!d_queries || !d_out_count || (k && (!d_out_rows || !d_out_dist))))) return SDB_EINVAL;
  if (nq == 0) return SDB_OK;
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  Ticket* t = free_ticket(c);
  if (!t) {
    set_error("too many batches in flight (%d): call sdb_knn_wait first", N_TICKETS);
    return SDB_EOVERFLOW;
  }
  SDB_TRY(submit_locked(c, t, d_queries, nq, k, row_base, d_out_rows, d_out_dist, d_out_count, nullptr));
  return wait_locked(c, t);
}

sdb_status sdb_knn_bruteforce(sdb_corpus* c, const double* queries, uint32_t nq, uint32_t k, uint64_t* out_rows,
                              double* out_dist, uint32_t* out_count, const volatile int* cancel_flag) {
  if (!c || (nq && (!queries || !out_count || (k && (!out_rows || !out_dist))))) return SDB_EINVAL;
  if (nq == 0) return SDB_OK;
  if (cancel_flag && *cancel_flag) {
    set_error("query cancelled");
    return SDB_ECANCELLED;
  }
  std::lock_guard<std::mutex> g(c->mu);
  SDB_CUDA(cudaSetDevice(c->ctx->device));
  Ticket* t = nullptr;
  SDB_TRY(submit_host_locked(c, queries, nq, k, out_rows, out_dist, out_count, cancel_flag, &t));
  return wait_locked(c, t);
}

sdb_status sdb_corpus_project(sdb_corpus* c, const double* query, int fn, double* out) {
  const bool is_metric = fn == SDB_COSINE || fn == SDB_EUCLIDEAN || fn == SDB_MANHATTAN || fn == SDB_CHEBYSHEV ||
                         fn == SDB_HAMMING || fn == SDB_PEARSON || fn == SDB_MINKOWSKI || fn == SDB_JACCARD;
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
  SDB_CUDA(drain(c->ctx));  // borrows the active scratch set for the prepared query
  cudaStream_t st = c->ctx->stream;
  double *d_q = nullptr, *d_vals = nullptr;
  auto run = [&]() -> sdb_status {
    SDB_CUDA(cudaMallocAsync(&d_q, sizeof(double) * c->dim, st));
    SDB_CUDA(cudaMallocAsync(&d_vals, sizeof(double) * c->n, st));
    if (query) SDB_CUDA(cudaMemcpyAsync(d_q, query, sizeof(double) * c->dim, cudaMemcpyHostToDevice, st));
    else SDB_CUDA(cudaMemsetAsync(d_q, 0, sizeof(double) * c->dim, st));
    SDB_TRY(scratch_for(c, 1, 4096));
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
  if (!ctx || !n_lists || !k || !d_rows || !d_dist || !d_counts || !d_out_rows || !d_out_dist || !d_out_count)
    return SDB_EINVAL;
  if (nq == 0) return SDB_OK;
  std::lock_guard<std::mutex> g(ctx->mu);
  SDB_CUDA(cudaSetDevice(ctx->device));
  SDB_TRY(topk_merge_launch(ctx, n_lists, nq, k, d_rows, d_dist, d_counts, stride_rows, stride_dist, stride_counts,
                            d_out_rows, d_out_dist, d_out_count, ctx->stream));
  SDB_CUDA(cudaStreamSynchronize(ctx->stream));
  return SDB_OK;
}

}  // extern "C"
