// swb200.cu — host side of libswb200.so: the C ABI declared in include/swb200.h.
//
// Host orchestration of the alignment hot path (no CPU fallback: every DP cell is computed by the
// kernels in sw_core.cuh).  What the host does, and the reference code it stands in for:
//   * tabulate/validate the scoring callback           smithwaterman.cpp:6-38, similaritymatrix.cpp:389-392
//   * cut the reference into overlapping pieces        plocalaligner.cpp:44-67 (_make_string_range)
//   * turn every (sequence, piece) into a task, pack tasks two by two into s16x2 "pairs", pick the
//     lane geometry (L lanes x R rows) and the checkpoint period B, size the HBM work buffers
//   * launch pass 1 (score), the piece selection of plocalaligner.cpp:122-129, pass 2 (arg-max +
//     traceback) and, for custom scoring, the default-scoring re-alignment of plocalaligner.cpp:132-136
#include "../../include/swb200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#define SWB_HELPER_KERNELS 1
#include "sw_core.cuh"
#include "sw_qs.cuh"

using namespace swb;

// one translation unit per rows-per-lane value R (csrc/sw_inst.cu compiled with -DSWB_R=<R>)
#define SWB_DECL(RR)                                                                                                  \
  cudaError_t swb_launch_score_r##RR(int C, int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const PassParams& p); \
  cudaError_t swb_launch_trace_r##RR(int C, int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const TraceParams& p); \
  cudaError_t swb_launch_dump_r##RR(int am, bool profile, size_t smem, cudaStream_t st, const DumpParams& p);       \
  cudaError_t swb_launch_qs_score_r##RR(bool sat, bool p16, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const QsParams& p); \
  cudaError_t swb_launch_qs_trace_r##RR(bool sat, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const QsTraceParams& p);
SWB_DECL(2) SWB_DECL(4) SWB_DECL(5) SWB_DECL(8) SWB_DECL(12) SWB_DECL(16) SWB_DECL(19) SWB_DECL(24) SWB_DECL(32)
#undef SWB_DECL

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Geometry { int L = 0, logL = 0, R = 0, nstrips = 1; };
const int kRSet[] = {2, 4, 5, 8, 12, 16, 19, 24, 32};
constexpr int kNumR = sizeof(kRSet) / sizeof(kRSet[0]);

// A launch class: tasks that share lane geometry and block size.
struct LaunchClass {
  Geometry geo;
  std::vector<PairDesc> pairs;
  std::vector<TaskDesc> tasks;      // task id order: local_read * pieces + piece
  int pieces = 1;                   // tasks per read (1 = plain SWAligner)
  int nreads = 0;
  size_t blk_words = 0, ck_words = 0, q_words = 0, bnd_words = 0;
  int max_m = 0, max_strips = 1;
  // device copies
  DevBuf d_pairs, d_tasks;
};

struct HostScoring {
  int mode = SWB_MODE_SAT_U8;
  bool match_shaped = true;         // table is a == b ? M : X
  int M = 3, X = 3, G = 2;          // SAT_U8: saturated; EXACT match-shaped: M = match, X = -mismatch
  std::vector<int32_t> table;       // EXACT, 256x256, only when !match_shaped
  int max_pos = 3;                  // largest positive score
  bool is_default() const { return match_shaped && M == 3 && X == 3 && G == 2; }
};

}  // namespace

struct swb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
  HostScoring sc;
  // reference
  std::vector<uint8_t> y;
  DevBuf d_ref_raw, d_ref_code, d_table;
  int KP = 0;
  uint8_t code_of[256];
  bool table_dirty = true;
  // staged batch
  bool staged = false;
  size_t n_seqs = 0;
  int npiece = 0;
  float ratio = 0.f;
  unsigned flags = 0;
  size_t cons_stride = 0;
  int B = 64, logB = 6;
  bool force_l32 = false;           // swb_matrix: always the 32-lane geometry (one pair per warp)
  int C = 1;                        // columns per wavefront step (2 = more ILP per warp; measured slower on B200, kept selectable)
  std::vector<LaunchClass> classes;
  DevBuf d_reads, d_qpairs, d_blkmax, d_ckpt, d_bnd, d_scratch, d_taskmax, d_winner, d_units, d_progress, d_next_task;
  DevBuf d_score, d_pos, d_end, d_cx, d_cy, d_len, d_flags;
  std::vector<DevBuf> desc_pool;      // descriptor buffers of finished launch classes, reused by the next stage
  DevBuf d_check;                     // SWB_CHECKED builds: one word, the highest failing bounds-check site
  // query-stationary mode (sw_qs.cuh): database search against a short reference, computed transposed
  bool qs = false;
  int qs_KP = 0;
  uint64_t batch_residues = 0, batch_max_m = 0;   // of the staged batch (swb_batch_rebind_reference)
  bool wide = false;                  // EXACT scores beyond the 16-bit lanes: one alignment per 32-bit lane
  DevBuf d_xcode, d_qs_table;
  swb_stats stats{};
  std::vector<cudaEvent_t> ev_pool;   // event pairs around every pass-2 launch of the current run
  size_t ev_used = 0;
};

namespace {

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                        \
      return SWB_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

int fail(swb_ctx* ctx, int code, const std::string& msg) { ctx->err = msg; return code; }

// SWB_DEBUG=1: synchronise after every kernel and print its wall time (diagnostics only; distorts timings)
struct DebugTimer {
  bool on; cudaStream_t st; double t0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
  explicit DebugTimer(cudaStream_t s) : on(getenv("SWB_DEBUG") != nullptr), st(s), t0(0) { if (on) { cudaStreamSynchronize(st); t0 = now(); } }
  void mark(const char* what, int L, int R, size_t n) {
    if (!on) return;
    cudaStreamSynchronize(st);
    const double t1 = now();
    fprintf(stderr, "[swb200] %-14s L=%-2d R=%-2d items=%-8zu %9.3f ms\n", what, L, R, n, t1 - t0);
    t0 = t1;
  }
};

// pairs per sub-batch (see build_classes): four waves of 148 SMs x 4 CTAs x 16 pairs
size_t chunk_pairs(const swb_ctx*) {
  size_t v = 37888;
  if (const char* e = getenv("SWB_CHUNK_PAIRS")) v = (size_t)std::max(64L, atol(e));
  return v;
}

// Symbol-score selection of the kernels: "profile" = per-warp query profile in shared memory (one LDS per
// cell pair, any scoring table), "compare" = HSET2 + LOP3 on packed symbols (match/mismatch scoring only).
bool use_profile(const swb_ctx* ctx, bool force_default);

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// _saturate, similaritymatrix.cpp:376-384
int saturate_u8(float a) { if (a < 0) return 0; if (a > 255) return 255; return (int)(uint8_t)a; }

// Lane geometry for sequences of m rows when `npairs` pairs are in flight.
bool choose_geometry(int m, size_t npairs, int r_cap, Geometry* out) {
  struct Cand { int L, R; double eff; double warps; };
  std::vector<Cand> c;
  int l_max = 32, l_min = 1;
  if (const char* e = getenv("SWB_FORCE_L")) { l_max = l_min = std::max(1, std::min(32, atoi(e))); }
  const int r_forced = getenv("SWB_FORCE_R") ? atoi(getenv("SWB_FORCE_R")) : 0;
  for (int L = l_max; L >= l_min; L >>= 1) {
    const int need = (m + L - 1) / L;
    int R = 0;
    for (int i = 0; i < kNumR; ++i) if (kRSet[i] >= need && kRSet[i] <= r_cap) { R = kRSet[i]; break; }
    if (!R || (r_forced && R != r_forced)) continue;
    c.push_back({L, R, (double)m / (L * R), (double)npairs * L / 32.0});
  }
  if (c.empty()) return false;
  double best_eff = 0; for (auto& k : c) best_eff = std::max(best_eff, k.eff);
  const double target = 148.0 * 8.0;
  const Cand* pick = nullptr;
  for (auto& k : c) {          // c is ordered by decreasing L: the last admissible candidate has the largest R
    if (k.eff < best_eff - 0.04) continue;
    if (!pick) { pick = &k; continue; }
    if (k.warps >= target) pick = &k;
  }
  out->L = pick->L; out->logL = ilog2(pick->L); out->R = pick->R;
  return true;
}

bool use_profile(const swb_ctx* ctx, bool force_default) {
  const HostScoring& hs = ctx->sc;
  if (!hs.match_shaped && !force_default) return true;
  if (force_default && !hs.is_default()) return false;      // the profile table holds the constructor's scoring
  if (const char* e = getenv("SWB_SELECT")) { if (!strcmp(e, "profile")) return ctx->KP <= 64; if (!strcmp(e, "compare")) return false; }
  // match/mismatch scoring: the profile costs no ALU instruction per cell (one LDS instead of HSET2 + LOP3)
  // and wins whenever the alphabet is small enough to keep 16 warps per SM resident (measured: +22 % on DNA)
  return ctx->KP <= 6;
}

Scoring device_scoring(const HostScoring& hs, bool force_default, bool wide = false) {
  int M = force_default ? 3 : hs.M, X = force_default ? 3 : hs.X, G = force_default ? 2 : hs.G;
  // two s16 halves per register, or one s32 value (wide lanes, AM_WIDE)
  auto pk = [wide](int v) { return wide ? (uint32_t)v : (uint32_t)(uint16_t)(int16_t)v * 0x00010001u; };
  Scoring s;
  s.G = G;
  s.negG2 = pk(-G);
  const int sm = M + G, sx = G - X;      // (s + G) for match / mismatch
  s.sel_xor = pk(sx);
  s.sel_and = pk(sm) ^ pk(sx);
  s.ceil2 = pk(255 - G);
  return s;
}

// Pass-2 ring geometry (sw_core.cuh, trace_body step 3): the ring of the last Wc steps lives in shared memory, packed
// (SAT_U8: one byte per cell, EXACT: 16 bits), for the NB lanes at and above the walker's lane; nlc local checkpoints
// of the lane state per warp live in HBM scratch.
struct TraceGeom { int Wc, logWc, NB, nlc; size_t ring_words, scratch_words; };
TraceGeom trace_geometry(int L, int R, int C, bool sat, bool wide = false) {
  const int PW = sat ? (R + 3) / 4 : (wide ? R : (R + 1) / 2);
  const int groups = 32 / L;
  // a diagonal walk of Wc columns climbs Wc rows: the walker's lane plus ceil(Wc / R) lanes above it
  // A lane k lanes above the walker's holds C * (Wc - k) columns left of the walker, and a diagonal walk reaches it after
  // k * R columns: a session ends after about C * Wc * R / (R + C) columns, in lane ceil(C * Wc / (R + C)) above.
  auto nb_for = [&](int wc) { return std::min(L, std::max(2, (wc * C + R + C - 1) / (R + C) + 1)); };
  auto bytes_for = [&](int wc) { return (size_t)wc * C * PW * groups * nb_for(wc) * 4; };
  int wc = bytes_for(64) <= 16 * 1024 ? 64 : 32;
  if (const char* e = getenv("SWB_TRACE_WC")) wc = std::max(32, std::min(256, 1 << ilog2(atoi(e))));   // >= 32: strip replays restart on 32-column chunks
  TraceGeom t;
  t.Wc = wc; t.logWc = ilog2(wc); t.NB = nb_for(wc);
  // wide strips with many columns per step: keep one warp's ring under 96 KB (fewer band lanes only shorten a session)
  while (t.NB > 2 && (size_t)wc * C * PW * groups * t.NB * 4 > 96 * 1024) --t.NB;
  if (const char* e = getenv("SWB_TRACE_NB")) t.NB = std::min(L, std::max(L > 1 ? 2 : 1, atoi(e)));
  t.nlc = 8;
  t.ring_words = (size_t)wc * C * PW * groups * t.NB;
  t.scratch_words = (size_t)t.nlc * (sat ? (R + C + 1) / 2 : R + C) * 32;
  return t;
}

// ---- kernel dispatch: one translation unit per R (sw_inst.cu compiled with -DSWB_R=<R>) ------------------

cudaError_t launch_score(int R, int C, int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const PassParams& p) {
  switch (R) {
#define SWB_CASE(RR) case RR: return swb_launch_score_r##RR(C, am, profile, grid, block, smem, st, p);
    SWB_CASE(2) SWB_CASE(4) SWB_CASE(5) SWB_CASE(8) SWB_CASE(12) SWB_CASE(16) SWB_CASE(19) SWB_CASE(24) SWB_CASE(32)
#undef SWB_CASE
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_trace(int R, int C, int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const TraceParams& p) {
  switch (R) {
#define SWB_CASE(RR) case RR: return swb_launch_trace_r##RR(C, am, profile, grid, block, smem, st, p);
    SWB_CASE(2) SWB_CASE(4) SWB_CASE(5) SWB_CASE(8) SWB_CASE(12) SWB_CASE(16) SWB_CASE(19) SWB_CASE(24) SWB_CASE(32)
#undef SWB_CASE
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_qs_score(int R, bool sat, bool p16, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const QsParams& p) {
  switch (R) {
#define SWB_CASE(RR) case RR: return swb_launch_qs_score_r##RR(sat, p16, grid, block, smem, st, p);
    SWB_CASE(2) SWB_CASE(4) SWB_CASE(5) SWB_CASE(8) SWB_CASE(12) SWB_CASE(16) SWB_CASE(19) SWB_CASE(24) SWB_CASE(32)
#undef SWB_CASE
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_qs_trace(int R, bool sat, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const QsTraceParams& p) {
  switch (R) {
#define SWB_CASE(RR) case RR: return swb_launch_qs_trace_r##RR(sat, grid, block, smem, st, p);
    SWB_CASE(2) SWB_CASE(4) SWB_CASE(5) SWB_CASE(8) SWB_CASE(12) SWB_CASE(16) SWB_CASE(19) SWB_CASE(24) SWB_CASE(32)
#undef SWB_CASE
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_dump(int R, int am, bool profile, size_t smem, cudaStream_t st, const DumpParams& p) {
  switch (R) {
#define SWB_CASE(RR) case RR: return swb_launch_dump_r##RR(am, profile, smem, st, p);
    SWB_CASE(2) SWB_CASE(4) SWB_CASE(5) SWB_CASE(8) SWB_CASE(12) SWB_CASE(16) SWB_CASE(19) SWB_CASE(24) SWB_CASE(32)
#undef SWB_CASE
    default: return cudaErrorInvalidValue;
  }
}

// Upload the (s + G) table of the profile select: [257][KP] int16, row 256 / column KP-1 = sentinels.
int upload_profile_table(swb_ctx* ctx) {
  if (!ctx->table_dirty) return SWB_OK;
  const HostScoring& hs = ctx->sc;
  if (ctx->y.empty()) { ctx->table_dirty = false; return SWB_OK; }
  const int KP = ctx->KP;
  std::vector<int16_t> t((size_t)257 * KP);
  uint8_t byte_of[256]; memset(byte_of, 0, sizeof byte_of);
  for (int b = 0; b < 256; ++b) if (ctx->code_of[b] != 0xFF) byte_of[ctx->code_of[b]] = (uint8_t)b;
  const int16_t never = (int16_t)(-16000);
  for (int a = 0; a <= 256; ++a)
    for (int c = 0; c < KP; ++c) {
      int16_t v = never;
      if (a < 256 && c < KP - 1) {
        const int sc = hs.match_shaped ? (a == byte_of[c] ? hs.M : -hs.X) : hs.table[(size_t)a * 256 + byte_of[c]];
        v = (int16_t)(sc + hs.G);
      }
      t[(size_t)a * KP + c] = v;
    }
  CUDA_TRY(ctx->d_table.ensure(t.size() * sizeof(int16_t)));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_table.p, t.data(), t.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->table_dirty = false;
  return SWB_OK;
}

// Profile words per (symbol, lane) of the score kernels' 128-bit layout: prof_stride<R>() of sw_core.cuh.
int prof_stride_rows(int R) { const int s = (R + 3) / 4 * 4; return s % 8 == 4 ? s : s + 4; }
// Largest rows-per-lane whose per-warp score profile (KP symbols) fits `bytes` of shared memory.
int max_rows_for_profile(int KP, size_t bytes) {
  int r = 2;
  for (int R = 2; R <= 32; ++R) if ((size_t)KP * prof_stride_rows(R) * 128 <= bytes) r = R;
  return r;
}

// Rows per lane of the strip geometry (L = 32) for a batch whose longest sequence has m_max rows.  With few long
// pairs (the long-pair config) the strips of a pair run concurrently, one warp each, so thin strips are preferred — as
// thin as the HBM budget for the strip boundary rows allows — to put a warp on every SM sub-partition; *few_long tells
// the caller that this mode applies.
int strip_rows(const swb_ctx* ctx, uint32_t m_max, uint32_t n_max, size_t npairs_est, bool profile, bool* few_long) {
  const int r_hard = profile ? max_rows_for_profile(ctx->KP, 200 * 1024) : 32;
  const int r_pref = profile ? std::min(r_hard, max_rows_for_profile(ctx->KP, 14336)) : 32;
  int r_strip = kRSet[0];
  const int cap = profile ? r_pref : 32;
  for (int i = 0; i < kNumR; ++i) if (kRSet[i] <= cap) r_strip = kRSet[i];
  if (few_long) *few_long = false;
  if (npairs_est < 148 * 8 && (int)m_max > 32 * r_strip) {
    if (few_long) *few_long = true;
    // boundary rows: up to three quarters of the free HBM (at least 32 GiB), SWB_BND_BUDGET_MB overrides.  10 kbp x 51 Mbp:
    // 4 rows per lane need 127 GB (79 strips per pair, 632 units), 5 rows 101 GB (504 units), 8 rows 64 GB (320 units)
    size_t budget_mb = 32768;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget_mb = std::max<size_t>(budget_mb, (free_b + ctx->d_bnd.cap) / 4 * 3 / 1048576);
    if (const char* e = getenv("SWB_BND_BUDGET_MB")) budget_mb = (size_t)std::max(64L, atol(e));
    // ... until a round of tasks holds at least one unit per SM sub-partition (the persistent warps of score_units_kernel)
    for (int i = kNumR - 1; i >= 0; --i) {
      const int R = kRSet[i];
      if (R > r_strip || R < 4) continue;
      const double strips = std::ceil((double)m_max / (32.0 * R));
      const double bytes = (strips - 1) * ((double)n_max + 1) * 4.0 * (double)npairs_est;
      if (bytes > (double)budget_mb * 1048576.0) break;
      r_strip = R;
      if (strips * (double)npairs_est >= 4.0 * ctx->sm_count) break;
    }
  }
  if (const char* e = getenv("SWB_STRIP_R")) {
    const int want = std::max(2, std::min(r_hard, atoi(e)));
    for (int i = 0; i < kNumR; ++i) if (kRSet[i] <= want) r_strip = kRSet[i];
  }
  return r_strip;
}

struct TaskSeed { uint32_t read; uint32_t piece; uint32_t y_off; uint32_t n; uint32_t m; uint32_t x_off; };

// Build launch classes from task seeds (task id = position in `seeds`, grouped by read: pieces contiguous).
int build_classes(swb_ctx* ctx, const std::vector<TaskSeed>& seeds, int pieces, std::vector<LaunchClass>* out) {
  out->clear();
  const bool profile = use_profile(ctx, false);
  // rows per lane: up to 32 registers; with the profile select prefer a per-warp profile <= 14 KB (16 warps / SM)
  // and never exceed what one warp's profile can hold in shared memory
  const int r_hard = profile ? max_rows_for_profile(ctx->KP, 200 * 1024) : 32;
  const int r_pref = profile ? std::min(r_hard, max_rows_for_profile(ctx->KP, 14336)) : 32;
  // geometry per distinct m
  std::map<uint32_t, size_t> count_by_m;
  for (auto& s : seeds) count_by_m[s.m]++;
  std::map<uint32_t, Geometry> geo_by_m;
  uint32_t m_max = 0, n_max = 0;
  for (auto& sd : seeds) { m_max = std::max(m_max, sd.m); n_max = std::max(n_max, sd.n); }
  int r_strip = strip_rows(ctx, m_max, n_max, (seeds.size() + 1) / 2, profile, nullptr);   // rows per lane of the strip geometry (L = 32)
  // Geometry per distinct length.  First pass: the best (L, R) for every length on its own.
  const int r_cap = profile ? std::max(r_pref, kRSet[0]) : 32;
  std::map<std::pair<int, int>, size_t> tasks_per_geo;
  for (auto& kv : count_by_m) {
    Geometry g;
    if (ctx->force_l32 || (int)kv.first > 32 * r_strip || !choose_geometry((int)kv.first, (seeds.size() + 1) / 2, r_cap, &g)) {
      g.L = 32; g.logL = 5; g.R = 0; g.nstrips = 0;     // R = 0: strip geometry, filled in below
    }
    geo_by_m[kv.first] = g;
    tasks_per_geo[std::make_pair(g.L, g.R)] += kv.second;
  }
  // Ragged batches (a protein database): (L, R) classes that hold only a sliver of the batch would each cost
  // their own short launches; they are folded into ONE strip-geometry class (32 lanes x thin strips, strip
  // count per pair, longest sequences first).  Classes that hold a real share of the batch keep their geometry
  // (reads trimmed to 100..150 bp still run as 8 x 16 and 8 x 19).
  const size_t small_class = std::max<size_t>(1024, seeds.size() / 50);
  size_t folded = 0;
  if (tasks_per_geo.size() > 2)
    for (auto& kv : geo_by_m) {
      Geometry& g = kv.second;
      if (g.R != 0 && tasks_per_geo[std::make_pair(g.L, g.R)] < small_class) { g.L = 32; g.logL = 5; g.R = 0; g.nstrips = 0; ++folded; }
    }
  if (folded > 0 && !getenv("SWB_STRIP_R")) {
    const int cap = 4;      // thin strips: pass 2 replays O(rows-in-strip) columns per strip (measured best on B200)
    r_strip = kRSet[0];
    for (int i = 0; i < kNumR; ++i) if (kRSet[i] <= cap && kRSet[i] <= r_hard) r_strip = kRSet[i];
  }
  for (auto& kv : geo_by_m) if (kv.second.R == 0) kv.second.R = r_strip;
  std::map<std::pair<int, int>, int> class_of;   // (L, R) -> class index
  std::vector<int> cls(seeds.size());
  for (size_t i = 0; i < seeds.size(); i += (size_t)pieces) {
    const Geometry g = geo_by_m[seeds[i].m];
    auto key = std::make_pair(g.L, g.R);
    auto it = class_of.find(key);
    if (it == class_of.end()) { it = class_of.emplace(key, (int)out->size()).first; out->emplace_back(); out->back().geo = g; out->back().pieces = pieces; }
    for (int pc = 0; pc < pieces; ++pc) cls[i + pc] = it->second;
  }
  // tasks in id order per class; remember ids for pairing
  std::vector<std::vector<uint32_t>> ids(out->size());
  for (size_t i = 0; i < seeds.size(); ++i) ids[cls[i]].push_back((uint32_t)i);
  // Sub-batches: a class with more pairs than a few waves of the GPU is cut into consecutive sub-classes that
  // run one after the other and reuse the same HBM work buffers, so the checkpoint period B (and with it the
  // cost of pass 2) does not grow with the batch size.  Not for the chunked-reference path (piece selection
  // needs every piece of a read in one class).
  if (pieces == 1) {
    const size_t chunk_tasks = 2 * chunk_pairs(ctx);
    std::vector<LaunchClass> split;
    std::vector<std::vector<uint32_t>> split_ids;
    for (size_t c = 0; c < out->size(); ++c) {
      const auto& id = ids[c];
      for (size_t o = 0; o < id.size(); o += chunk_tasks) {
        split.emplace_back();
        split.back().geo = (*out)[c].geo; split.back().pieces = 1;
        split_ids.emplace_back(id.begin() + o, id.begin() + std::min(id.size(), o + chunk_tasks));
      }
    }
    out->swap(split);
    ids.swap(split_ids);
  }
  for (size_t c = 0; c < out->size(); ++c) {
    LaunchClass& lc = (*out)[c];
    const int L = lc.geo.L, R = lc.geo.R;
    auto& id = ids[c];
    lc.tasks.resize(id.size());
    lc.nreads = (int)(id.size() / (size_t)pieces);
    // pairing order: by (y_off, n), then id — tasks of one pair must share the y-range
    std::vector<uint32_t> order(id.size());
    std::iota(order.begin(), order.end(), 0u);
    bool uniform = true, same_m = true;
    for (size_t k = 1; k < id.size() && (uniform || same_m); ++k) {
      uniform = uniform && seeds[id[k]].y_off == seeds[id[0]].y_off && seeds[id[k]].n == seeds[id[0]].n;
      same_m = same_m && seeds[id[k]].m == seeds[id[0]].m;
    }
    if (!uniform || (!same_m && L == 32))
      std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const TaskSeed &sa = seeds[id[a]], &sb = seeds[id[b]];
        if (sa.y_off != sb.y_off) return sa.y_off < sb.y_off;
        if (sa.n != sb.n) return sa.n < sb.n;
        return L == 32 && sa.m > sb.m;       // longest first: pair-mates of similar length, big work units scheduled first
      });
    size_t k = 0;
    while (k < order.size()) {
      const TaskSeed& a = seeds[id[order[k]]];
      PairDesc pd{};
      pd.y_off = a.y_off; pd.n = a.n;
      pd.nblk = (uint32_t)(((a.n + ctx->C - 1) / ctx->C + L - 1 + ctx->B - 1) / ctx->B);
      pd.mA = a.m; pd.xA = a.x_off;
      const uint32_t pair_idx = (uint32_t)lc.pairs.size();
      lc.tasks[order[k]] = TaskDesc{pair_idx, 0u, a.read, a.y_off};
      lc.max_m = std::max(lc.max_m, (int)a.m);
      if (k + 1 < order.size() && !ctx->wide) {          // wide lanes hold one alignment per register: no pair-mate
        const TaskSeed& b = seeds[id[order[k + 1]]];
        if (b.y_off == a.y_off && b.n == a.n) {
          pd.mB = b.m; pd.xB = b.x_off;
          lc.tasks[order[k + 1]] = TaskDesc{pair_idx, 1u, b.read, b.y_off};
          lc.max_m = std::max(lc.max_m, (int)b.m);
          ++k;
        }
      }
      ++k;
      const size_t ns = (L == 32) ? ((size_t)std::max(pd.mA, pd.mB) + (size_t)L * R - 1) / ((size_t)L * R) : 1;
      pd.nstrips = (uint32_t)ns;
      lc.max_strips = std::max(lc.max_strips, (int)ns);
      pd.q_off = (uint32_t)lc.q_words; lc.q_words += ns * L * R;
      pd.blk_off = lc.blk_words; lc.blk_words += ns * pd.nblk;
      const size_t ckw = ctx->sc.mode == SWB_MODE_SAT_U8 ? (size_t)(R + ctx->C + 1) / 2 : (size_t)(R + ctx->C);   // state_words<R, C, SAT>
      pd.ck_off = lc.ck_words; lc.ck_words += ns * pd.nblk * ckw * L;
      pd.bnd_off = lc.bnd_words; lc.bnd_words += (ns - 1) * ((size_t)pd.n + 1);
      lc.pairs.push_back(pd);
    }
    if (lc.q_words > 0xFFFFFFFFull) return fail(ctx, SWB_ERR_UNSUPPORTED, "batch too large for one launch class (split the batch)");
  }
  return SWB_OK;
}

// The descriptor buffers of the classes come from a pool in the context and go back to it: cudaFree / cudaMalloc per staged
// batch cost 1 .. 500 ms on a GPU that holds gigabytes of work buffers (measured with SWB_DEBUG_STAGE: "free classes").
DevBuf take_desc_buf(swb_ctx* ctx, size_t bytes) {
  int best = -1;
  for (size_t i = 0; i < ctx->desc_pool.size(); ++i) {
    const DevBuf& b = ctx->desc_pool[i];
    if (b.cap >= bytes && (best < 0 || b.cap < ctx->desc_pool[best].cap)) best = (int)i;
  }
  if (best < 0 && !ctx->desc_pool.empty()) best = (int)ctx->desc_pool.size() - 1;   // too small: ensure() regrows it
  DevBuf b;
  if (best >= 0) { b = ctx->desc_pool[best]; ctx->desc_pool.erase(ctx->desc_pool.begin() + best); }
  return b;
}

int upload_classes(swb_ctx* ctx, std::vector<LaunchClass>& classes) {
  for (auto& lc : classes) {
    if (!lc.d_pairs.p) lc.d_pairs = take_desc_buf(ctx, lc.pairs.size() * sizeof(PairDesc));
    if (!lc.d_tasks.p) lc.d_tasks = take_desc_buf(ctx, lc.tasks.size() * sizeof(TaskDesc));
    CUDA_TRY(lc.d_pairs.ensure(lc.pairs.size() * sizeof(PairDesc)));
    CUDA_TRY(lc.d_tasks.ensure(lc.tasks.size() * sizeof(TaskDesc)));
    CUDA_TRY(cudaMemcpyAsync(lc.d_pairs.p, lc.pairs.data(), lc.pairs.size() * sizeof(PairDesc), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(lc.d_tasks.p, lc.tasks.data(), lc.tasks.size() * sizeof(TaskDesc), cudaMemcpyHostToDevice, ctx->stream));
  }
  return SWB_OK;
}

void free_classes(swb_ctx* ctx, std::vector<LaunchClass>& classes) {
  for (auto& lc : classes) {
    if (lc.d_pairs.p) ctx->desc_pool.push_back(lc.d_pairs);
    if (lc.d_tasks.p) ctx->desc_pool.push_back(lc.d_tasks);
    lc.d_pairs = DevBuf{}; lc.d_tasks = DevBuf{};
  }
  classes.clear();
  while (ctx->desc_pool.size() > 64) { ctx->desc_pool.back().release(); ctx->desc_pool.pop_back(); }
}

// Run pass 1 (+ piece selection) + pass 2 for every class.  `force_default`: plocalaligner.cpp:135 re-alignment.
int run_classes(swb_ctx* ctx, LaunchClass* classes, size_t nclasses, bool force_default, bool select_pieces, bool trace) {
  const HostScoring& hs = ctx->sc;
  const bool sat = hs.mode == SWB_MODE_SAT_U8;
  const bool wide = ctx->wide && !sat;
  const int am = sat ? AM_SAT : (wide ? AM_WIDE : AM_EXACT);
  const bool profile = use_profile(ctx, force_default);
  DebugTimer dbg(ctx->stream);
  for (size_t ci = 0; ci < nclasses; ++ci) {
    LaunchClass& lc = classes[ci];
    const int L = lc.geo.L, R = lc.geo.R;
    CUDA_TRY(ctx->d_qpairs.ensure(lc.q_words * 4));
    CUDA_TRY(ctx->d_blkmax.ensure(lc.blk_words * 4));
    CUDA_TRY(ctx->d_ckpt.ensure(lc.ck_words * 4));
    CUDA_TRY(ctx->d_bnd.ensure(lc.bnd_words * 4 + 256));
    PassParams pp{};
    pp.ref_raw = ctx->d_ref_raw.as<uint8_t>();
    pp.ref_code = ctx->d_ref_code.as<uint8_t>();
    pp.reads_raw = ctx->d_reads.as<uint8_t>();
    pp.qpairs = ctx->d_qpairs.as<uint32_t>();
    pp.table = ctx->d_table.as<int16_t>();
    pp.KP = ctx->KP;
    pp.pairs = lc.d_pairs.as<PairDesc>();
    pp.npairs = (int)lc.pairs.size();
    pp.blkmax = ctx->d_blkmax.as<uint32_t>();
    pp.ckpt = ctx->d_ckpt.as<uint32_t>();
    pp.bnd = ctx->d_bnd.as<uint32_t>();
    pp.L = L; pp.logL = lc.geo.logL; pp.B = ctx->B; pp.logB = ctx->logB;
    pp.strips = lc.max_strips > 1 ? 1 : 0;
    pp.check = ctx->d_check.as<uint32_t>();
    pp.ck_words = lc.ck_words; pp.blk_words = lc.blk_words; pp.bnd_words = lc.bnd_words + 64; pp.ref_len = ctx->y.size();
    pp.sc = device_scoring(hs, force_default, wide);
    if (profile) pp.sc.G = hs.G;

    // pack rows
    {
      const int rows_per_pair = L * R * lc.max_strips;
      const long long total = (long long)lc.pairs.size() * rows_per_pair;
      const int thr = 256;
      pack_rows_kernel<<<(unsigned)((total + thr - 1) / thr), thr, 0, ctx->stream>>>(pp.reads_raw, pp.pairs, pp.npairs, rows_per_pair, L * R, ctx->d_qpairs.as<uint32_t>());
      CUDA_TRY(cudaGetLastError());
      ctx->stats.kernel_launches++;
    }
    // pass 1
    int warps_per_cta = 4;
    size_t smem = 0;
    if (profile) {
      const size_t per_warp = (size_t)ctx->KP * prof_stride_rows(R) * 32 * 4;
      while (warps_per_cta > 1 && per_warp * warps_per_cta > 200 * 1024) warps_per_cta >>= 1;
      smem = per_warp * warps_per_cta;
      if (const char* e = getenv("SWB_EXTRA_SMEM")) smem += (size_t)atol(e);   // occupancy experiments
    }
    const int groups_per_warp = 32 / L;
    {
      size_t warps = (lc.pairs.size() + groups_per_warp - 1) / groups_per_warp;
      // few long pairs: run the strips of a pair concurrently, one warp per (pair, strip) unit
      size_t total_units = 0;
      for (auto& pd : lc.pairs) total_units += pd.nstrips;
      const bool pipelined = L == 32 && lc.max_strips > 1 && lc.pairs.size() < 148 * 8 && !getenv("SWB_NO_PIPELINE");
      pp.units = nullptr; pp.nunits = 0; pp.progress = nullptr; pp.abort_flag = nullptr; pp.ticket = nullptr;
      if (pipelined) {
        std::vector<uint2> units;
        units.reserve(total_units);
        for (size_t pi = 0; pi < lc.pairs.size(); ++pi)
          for (uint32_t st = 0; st < lc.pairs[pi].nstrips; ++st) units.push_back(make_uint2((unsigned)pi, st));
        CUDA_TRY(ctx->d_units.ensure(units.size() * sizeof(uint2)));
        CUDA_TRY(ctx->d_progress.ensure((2 * units.size() + 2) * 4));
        CUDA_TRY(cudaMemcpyAsync(ctx->d_units.p, units.data(), units.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(ctx->d_progress.p, 0, (2 * units.size() + 2) * 4, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));      // `units` is a host temporary
        pp.units = ctx->d_units.as<uint2>(); pp.nunits = (int)units.size();
        pp.progress = ctx->d_progress.as<uint32_t>();
        pp.abort_flag = ctx->d_progress.as<uint32_t>() + units.size();
        pp.ticket = ctx->d_progress.as<uint32_t>() + units.size() + 1;
        pp.blocks_done = ctx->d_progress.as<uint32_t>() + units.size() + 2;
        // a task = one unit x enough checkpoint blocks for 256 steps (64 .. 128 when the reference is so short that the
        // 2 * strips rounds it takes to fill the pipeline would show); tickets run round-major, task block = round - 2 * strip
        uint32_t nblk_max = 1;
        for (auto& pd : lc.pairs) nblk_max = std::max(nblk_max, pd.nblk);
        int task_steps = 256;
        while (task_steps > 64 && (uint64_t)nblk_max * ctx->B / task_steps < 16ull * lc.max_strips) task_steps >>= 1;
        pp.task_blocks = std::max(1, task_steps / ctx->B);
        if (const char* e = getenv("SWB_TASK_BLOCKS")) pp.task_blocks = std::max(1, atoi(e));
        uint64_t rounds = 0;
        for (auto& pd : lc.pairs) rounds = std::max<uint64_t>(rounds, 2ull * (pd.nstrips - 1) + (pd.nblk + pp.task_blocks - 1) / pp.task_blocks);
        if (rounds * units.size() >= 0xffffffffull) return fail(ctx, SWB_ERR_UNSUPPORTED, "pipelined strips: too many tasks");
        pp.ntickets = (uint32_t)(rounds * units.size());
        // persistent warps: one per SM sub-partition, never more than there are units in a round
        warps = std::min<size_t>(units.size(), (size_t)4 * (size_t)ctx->sm_count);
        if (const char* e = getenv("SWB_UNIT_WARPS")) warps = std::max<size_t>(1, std::min<size_t>(units.size(), (size_t)atol(e)));
        warps_per_cta = 1;                                   // spread the warps over all SMs
        if (profile) smem = (size_t)ctx->KP * prof_stride_rows(R) * 32 * 4;
      }
      const unsigned grid = (unsigned)((warps + warps_per_cta - 1) / warps_per_cta);
      CUDA_TRY(launch_score(R, ctx->C, am, profile, dim3(grid), dim3(32 * warps_per_cta), smem, ctx->stream, pp));
      ctx->stats.kernel_launches++;
      dbg.mark("score", L, R, lc.pairs.size());
      for (auto& pd : lc.pairs) ctx->stats.cells_executed += (uint64_t)pd.nstrips * pd.nblk * ctx->B * ctx->C * L * R * (wide ? 1ull : 2ull);
      if (pipelined) {
        uint32_t aborted = 0;
        CUDA_TRY(cudaMemcpyAsync(&aborted, pp.abort_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (aborted) return fail(ctx, SWB_ERR_CUDA, "pipelined strips: a strip timed out waiting for its producer");
        pp.units = nullptr; pp.nunits = 0;
        warps_per_cta = 4;
      }
    }
    if (!trace && !select_pieces) continue;

    const uint32_t* task_list = nullptr;
    int ntrace = (int)lc.tasks.size();
    if (select_pieces) {
      CUDA_TRY(ctx->d_taskmax.ensure(lc.tasks.size() * 4));
      CUDA_TRY(ctx->d_winner.ensure((size_t)lc.nreads * 4));
      const int thr = 128;
      task_max_kernel<<<(unsigned)((lc.tasks.size() + thr - 1) / thr), thr, 0, ctx->stream>>>(pp.pairs, lc.d_tasks.as<TaskDesc>(), (int)lc.tasks.size(), pp.blkmax, pp.sc.G, wide ? 1 : 0, ctx->d_taskmax.as<int32_t>());
      CUDA_TRY(cudaGetLastError());
      select_piece_kernel<<<(unsigned)((lc.nreads + thr - 1) / thr), thr, 0, ctx->stream>>>(ctx->d_taskmax.as<int32_t>(), lc.nreads, lc.pieces, ctx->d_winner.as<uint32_t>());
      CUDA_TRY(cudaGetLastError());
      ctx->stats.kernel_launches += 2;
      task_list = ctx->d_winner.as<uint32_t>();
      ntrace = lc.nreads;
    }
    if (!trace) continue;

    // pass 2
    TraceParams tp{};
    tp.pp = pp;
    tp.tasks = lc.d_tasks.as<TaskDesc>();
    tp.task_list = task_list;
    tp.ntasks = ntrace;
    tp.mode = hs.mode;
    tp.max_pos = force_default ? 3 : std::max(1, hs.max_pos);
    const TraceGeom tg = trace_geometry(L, R, ctx->C, sat, wide);
    tp.Wc = tg.Wc; tp.logWc = tg.logWc; tp.NB = tg.NB; tp.nlc = tg.nlc;
    // Pass 2 is latency-bound, not ALU-bound: with match/mismatch scoring it gives up the per-warp profile (10 KB of shared
    // memory at 16 rows x 5 symbols, rebuilt per task) for two more ALU instructions per cell pair and a fourth resident
    // block per SM (measured: pass 2 of C1x64 1.90 -> 1.69 ms, C3 24.5 -> 20.4 ms; SWB_TRACE_SELECT=profile is the old way).
    bool tprofile = profile;
    if (profile && (hs.match_shaped || force_default))
      tprofile = getenv("SWB_TRACE_SELECT") && !strcmp(getenv("SWB_TRACE_SELECT"), "profile");
    const size_t prof_words_warp = tprofile ? (size_t)ctx->KP * R * 32 : 0;
    warps_per_cta = 4;
    while (warps_per_cta > 1 && (prof_words_warp + tg.ring_words) * 4 * warps_per_cta > 200 * 1024) warps_per_cta >>= 1;
    tp.ring_off = (int)(prof_words_warp * warps_per_cta);
    smem = (prof_words_warp + tg.ring_words) * 4 * warps_per_cta;
    if (smem > 220 * 1024) return fail(ctx, SWB_ERR_UNSUPPORTED, "pass-2 ring does not fit shared memory (lower SWB_TRACE_WC / SWB_TRACE_NB)");
    const size_t max_groups = (size_t)ctx->sm_count * 24 * groups_per_warp;   // up to 6 blocks per SM; the warps take tasks from a counter, surplus blocks exit at once
    size_t groups = std::min<size_t>((size_t)ntrace, max_groups);
    size_t warps = (groups + groups_per_warp - 1) / groups_per_warp;
    const unsigned grid = (unsigned)((warps + warps_per_cta - 1) / warps_per_cta);
    CUDA_TRY(ctx->d_scratch.ensure((size_t)grid * warps_per_cta * tg.scratch_words * 4));
    tp.scratch = ctx->d_scratch.as<uint32_t>();
    tp.scratch_words = (unsigned long long)grid * warps_per_cta * tg.scratch_words;
    tp.out_score = ctx->d_score.as<int32_t>();
    tp.out_pos = ctx->d_pos.as<uint32_t>();
    tp.out_end = ctx->d_end.as<uint32_t>();
    tp.out_cx = ctx->d_cx.as<uint8_t>();
    tp.out_cy = ctx->d_cy.as<uint8_t>();
    tp.out_len = ctx->d_len.as<uint32_t>();
    tp.out_flags = ctx->d_flags.as<uint32_t>();
    tp.cons_cap = (uint32_t)ctx->cons_stride;
    tp.want_consensus = (ctx->flags & SWB_FLAG_CONSENSUS) ? 1 : 0;
    if (!tp.want_consensus) tp.cons_cap = 0x7FFFFFFFu;
    CUDA_TRY(ctx->d_next_task.ensure(256));
    CUDA_TRY(cudaMemsetAsync(ctx->d_next_task.p, 0, 4, ctx->stream));
    tp.next_task = ctx->d_next_task.as<uint32_t>();
    DevBuf d_cnt;
    tp.counters = nullptr;
    tp.dbg_flags = getenv("SWB_DEBUG_FLAGS") ? atoi(getenv("SWB_DEBUG_FLAGS")) : 0;
    if (dbg.on) { CUDA_TRY(d_cnt.ensure(128)); CUDA_TRY(cudaMemsetAsync(d_cnt.p, 0, 128, ctx->stream)); tp.counters = d_cnt.as<unsigned long long>(); }
    while (ctx->ev_pool.size() < ctx->ev_used + 2) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); ctx->ev_pool.push_back(e); }
    CUDA_TRY(cudaEventRecord(ctx->ev_pool[ctx->ev_used], ctx->stream));
    CUDA_TRY(launch_trace(R, ctx->C, am, tprofile, dim3(grid), dim3(32 * warps_per_cta), smem, ctx->stream, tp));
    CUDA_TRY(cudaEventRecord(ctx->ev_pool[ctx->ev_used + 1], ctx->stream));
    ctx->ev_used += 2;
    ctx->stats.kernel_launches++;
    dbg.mark("trace", L, R, (size_t)ntrace);
    if (dbg.on) {
      unsigned long long c[16] = {0};
      cudaMemcpy(c, d_cnt.p, sizeof c, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[swb200]   scan replays %llu, sessions %llu, scan rounds (warps) %llu, session rounds (warps) %llu, session steps %llu\n", c[0], c[1], c[2], c[3], c[4]);
      fprintf(stderr, "[swb200]   warp-cycles (M): prepare %.1f vmax %.1f search %.1f scan %.1f session %.1f walk %.1f\n", c[8] / 1e6, c[9] / 1e6, c[10] / 1e6, c[11] / 1e6, c[12] / 1e6, c[13] / 1e6);
      d_cnt.release();
    }
    ctx->stats.cells_pass2 += (uint64_t)ntrace * (uint64_t)(3 * ctx->B * ctx->C + lc.max_m + 16) * L * R * 2ull;
    ctx->stats.lanes_per_pair = L; ctx->stats.rows_per_lane = R; ctx->stats.block_steps = ctx->B;
    ctx->stats.cols_per_step = ctx->C; ctx->stats.kernel_kind = (L == 32 && lc.max_strips > 1 && lc.pairs.size() < 148 * 8 && !getenv("SWB_NO_PIPELINE")) ? 1u : 0u;
  }
  return SWB_OK;
}


// ---- query-stationary mode (sw_qs.cuh) -----------------------------------------------------------------------
// Database search: many sequences against ONE short reference (mpi_sw_solve_uniprot.cpp: x = database protein,
// y = query).  The matrix is computed transposed (rows = y, shared by every alignment; columns = x), see sw_qs.cuh.
// Used when the batch is large (>= 2048 sequences), y fits one strip of at most 608 rows and the batch's alphabet
// is large (proteins: more than 6 symbols);
// SWB_QSTAT=0 / 1 forces it off / on (where applicable).  Returns 1 when the mode does not apply (caller falls back).
int stage_qs(swb_ctx* ctx, const char* seqs, const uint64_t* offsets, size_t n_seqs, unsigned flags) {
  const HostScoring& hs = ctx->sc;
  const size_t N = ctx->y.size();
  const char* env = getenv("SWB_QSTAT");
  if (env && atoi(env) == 0) return 1;
  const bool forced = env && atoi(env) != 0;
  if (!forced && (n_seqs < 2048 || N > 608)) return 1;
  if (ctx->force_l32 || N > 1024) return 1;
  // geometry: L*R >= N rows with the least padding (ties: fewer lanes, i.e. more pairs per warp)
  Geometry geo;
  size_t best_rows = ~(size_t)0;
  for (int L = 8; L <= 32; L <<= 1)
    for (int i = 0; i < kNumR; ++i)
      if ((size_t)L * kRSet[i] >= N) { if ((size_t)L * kRSet[i] < best_rows) { best_rows = (size_t)L * kRSet[i]; geo.L = L; geo.logL = ilog2(L); geo.R = kRSet[i]; } break; }
  if (!geo.L) return 1;
  const int L = geo.L, R = geo.R;
  // alphabet of the batch (the kernel's column symbols)
  const size_t blob = offsets[n_seqs] - offsets[0];
  uint8_t xcode_of[256]; memset(xcode_of, 0xFF, sizeof xcode_of);
  uint8_t byte_of[256]; memset(byte_of, 0, sizeof byte_of);
  int KX = 0;
  std::vector<uint8_t> codes(blob);
  const uint8_t* src = (const uint8_t*)seqs + offsets[0];
  for (size_t i = 0; i < blob; ++i) {
    const uint8_t b = src[i];
    if (xcode_of[b] == 0xFF) { if (KX >= 254) return 1; byte_of[KX] = b; xcode_of[b] = (uint8_t)KX++; }
    codes[i] = xcode_of[b];
  }
  const int KP = KX + 1;
  if ((size_t)KP * R * 128 > 200 * 1024) return 1;      // one profile per thread block
  // small alphabets (DNA) are served better by the batched kernels: their per-warp profile is small enough for
  // 16-19 rows per lane and costs one LDS per cell pair instead of two
  if (!forced && KX <= 6) return 1;
  uint32_t max_m = 0;
  uint64_t cells_ref = 0;
  for (size_t r = 0; r < n_seqs; ++r) {
    const uint64_t m64 = offsets[r + 1] - offsets[r];
    if (m64 == 0) return fail(ctx, SWB_ERR_ARG, "empty sequence at index " + std::to_string(r));
    max_m = std::max(max_m, (uint32_t)m64);
    cells_ref += m64 * N;
  }
  {
    const uint64_t reach = (uint64_t)std::min<size_t>(max_m, N) * (uint64_t)std::max(1, hs.max_pos) + (uint64_t)hs.G + 16;
    if (hs.mode == SWB_MODE_EXACT && (reach > 32000 || getenv("SWB_FORCE_WIDE"))) return 1;   // wide lanes: the batched kernels
  }
  ctx->wide = false;
  ctx->C = 1;
  // (s + G) table: rows = reference (query) byte, 256 = padding row; columns = batch code, KP-1 = padding column
  std::vector<int16_t> t((size_t)257 * KP, (int16_t)(-16000));
  for (int a = 0; a < 256; ++a)
    for (int c = 0; c < KP - 1; ++c) {
      const int xb = byte_of[c];
      const int sc = hs.match_shaped ? (xb == a ? hs.M : -hs.X) : hs.table[(size_t)xb * 256 + a];
      t[(size_t)a * KP + c] = (int16_t)(sc + hs.G);
    }
  // sequences by decreasing length: pair-mates of similar length, long ones scheduled first
  std::vector<uint32_t> order(n_seqs);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b]; });
  const size_t ckw = hs.mode == SWB_MODE_SAT_U8 ? (size_t)(R + 2) / 2 : (size_t)(R + 1);     // state_words<R, 1, SAT>
  size_t budget_mb = 8192;
  if (const char* e = getenv("SWB_CKPT_BUDGET_MB")) budget_mb = (size_t)std::max(64L, atol(e));
  const size_t chunk_tasks = 2 * chunk_pairs(ctx);
  int B = 16;       // short checkpoint period: pass 2 replays O(B) columns per scan and per session
  if (const char* e = getenv("SWB_QS_B")) B = std::max(8, std::min(1024, 1 << ilog2(atoi(e))));
  for (;; B <<= 1) {
    size_t worst = 0;
    for (size_t o = 0; o < n_seqs; o += chunk_tasks) {
      size_t words = 0;
      for (size_t k = o; k < std::min(n_seqs, o + chunk_tasks); k += 2) {
        const size_t n = offsets[order[k] + 1] - offsets[order[k]];
        words += ((n + L - 1 + B - 1) / B) * ckw * L;
      }
      worst = std::max(worst, words);
    }
    if (worst * 4 <= budget_mb * 1048576 || B >= 65536) break;
  }
  ctx->B = B; ctx->logB = ilog2(B);
  free_classes(ctx, ctx->classes);
  for (size_t o = 0; o < n_seqs; o += chunk_tasks) {
    ctx->classes.emplace_back();
    LaunchClass& lc = ctx->classes.back();
    lc.geo = geo; lc.pieces = 1; lc.max_m = (int)N; lc.max_strips = 1;
    const size_t end = std::min(n_seqs, o + chunk_tasks);
    lc.nreads = (int)(end - o);
    for (size_t k = o; k < end; k += 2) {
      const uint32_t ra = order[k];
      PairDesc pd{};
      pd.mA = (uint32_t)(offsets[ra + 1] - offsets[ra]); pd.xA = (uint32_t)(offsets[ra] - offsets[0]);
      const uint32_t pair_idx = (uint32_t)lc.pairs.size();
      lc.tasks.push_back(TaskDesc{pair_idx, 0u, ra, 0u});
      if (k + 1 < end) {
        const uint32_t rb = order[k + 1];
        pd.mB = (uint32_t)(offsets[rb + 1] - offsets[rb]); pd.xB = (uint32_t)(offsets[rb] - offsets[0]);
        lc.tasks.push_back(TaskDesc{pair_idx, 1u, rb, 0u});
      } else { pd.mB = 0; pd.xB = pd.xA; }
      pd.y_off = 0; pd.n = std::max(pd.mA, pd.mB);
      pd.nblk = (uint32_t)((pd.n + L - 1 + B - 1) / B);
      pd.nstrips = 1;
      pd.blk_off = lc.blk_words; lc.blk_words += pd.nblk;
      pd.ck_off = lc.ck_words; lc.ck_words += (size_t)pd.nblk * ckw * L;
      lc.pairs.push_back(pd);
    }
  }
  CUDA_TRY(ctx->d_reads.ensure(blob + 64));
  CUDA_TRY(ctx->d_xcode.ensure(blob + 64));
  CUDA_TRY(ctx->d_qs_table.ensure(t.size() * sizeof(int16_t)));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_reads.p, src, blob, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_xcode.p, codes.data(), blob, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_qs_table.p, t.data(), t.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  int rc = upload_classes(ctx, ctx->classes);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));          // codes / t are host temporaries
  ctx->qs = true; ctx->qs_KP = KP;
  ctx->batch_residues = blob; ctx->batch_max_m = max_m;
  ctx->stats = swb_stats{};
  ctx->stats.cells_reference = cells_ref;
  return SWB_OK;
}

int run_qs(swb_ctx* ctx) {
  const HostScoring& hs = ctx->sc;
  const bool sat = hs.mode == SWB_MODE_SAT_U8;
  DebugTimer dbg(ctx->stream);
  for (auto& lc : ctx->classes) {
    const int L = lc.geo.L, R = lc.geo.R;
    const int groups_per_warp = 32 / L;
    CUDA_TRY(ctx->d_blkmax.ensure(lc.blk_words * 4));
    CUDA_TRY(ctx->d_ckpt.ensure(lc.ck_words * 4));
    QsParams qp{};
    PassParams& pp = qp.pp;
    pp.ref_raw = ctx->d_reads.as<uint8_t>();             // the kernel's columns: the batch
    pp.ref_code = ctx->d_xcode.as<uint8_t>();
    pp.reads_raw = ctx->d_ref_raw.as<uint8_t>();         // the kernel's rows: the reference (query)
    pp.table = ctx->d_qs_table.as<int16_t>();
    pp.KP = ctx->qs_KP;
    pp.pairs = lc.d_pairs.as<PairDesc>();
    pp.npairs = (int)lc.pairs.size();
    pp.blkmax = ctx->d_blkmax.as<uint32_t>();
    pp.ckpt = ctx->d_ckpt.as<uint32_t>();
    pp.L = L; pp.logL = lc.geo.logL; pp.B = ctx->B; pp.logB = ctx->logB;
    pp.sc = device_scoring(hs, false);
    pp.sc.G = hs.G;
    pp.check = ctx->d_check.as<uint32_t>();
    pp.ck_words = lc.ck_words; pp.blk_words = lc.blk_words; pp.bnd_words = 0; pp.ref_len = ctx->batch_residues;
    qp.m = (int)ctx->y.size();
    const size_t smem = (size_t)ctx->qs_KP * R * 32 * 4;      // score pass: 32-bit profile, one per thread block
    const size_t smem_trace = (size_t)ctx->qs_KP * ((R + 1) / 2) * 32 * 4;   // pass 2: 16-bit profile, two rows per word (qs_build_profile_paired)
    int warps_per_cta = 4;
    {
      const size_t warps = (lc.pairs.size() + groups_per_warp - 1) / groups_per_warp;
      const unsigned grid = (unsigned)((warps + warps_per_cta - 1) / warps_per_cta);
      // pass 1 on the paired 16-bit profile too: 40 instead of 76 LDS per two steps, half the shared memory
      // (measured on C4: 1.95 -> 1.75 ms per 75 776 proteins); SWB_QS_PROF16=0 is the 32-bit profile
      const bool p16 = !(getenv("SWB_QS_PROF16") && atoi(getenv("SWB_QS_PROF16")) == 0);
      CUDA_TRY(launch_qs_score(R, sat, p16, dim3(grid), dim3(32 * warps_per_cta), p16 ? smem_trace : smem, ctx->stream, qp));
      ctx->stats.kernel_launches++;
      dbg.mark("qs score", L, R, lc.pairs.size());
      for (auto& pd : lc.pairs) ctx->stats.cells_executed += (uint64_t)pd.nblk * ctx->B * L * R * 2ull;
    }
    QsTraceParams qt{};
    TraceParams& tp = qt.tp;
    qt.m = qp.m;
    tp.pp = pp;
    tp.tasks = lc.d_tasks.as<TaskDesc>();
    tp.task_list = nullptr;
    tp.ntasks = (int)lc.tasks.size();
    tp.mode = hs.mode;
    tp.max_pos = std::max(1, hs.max_pos);
    const TraceGeom tg = trace_geometry(L, R, 1, sat);
    tp.Wc = tg.Wc; tp.logWc = tg.logWc; tp.NB = tg.NB; tp.nlc = tg.nlc;
    tp.ring_off = (int)(smem_trace / 4);                      // the rings follow the block's 16-bit profile
    while (warps_per_cta > 1 && smem_trace + tg.ring_words * 4 * warps_per_cta > 200 * 1024) warps_per_cta >>= 1;
    const size_t smem_trace_total = smem_trace + tg.ring_words * 4 * warps_per_cta;
    if (smem_trace_total > 220 * 1024) return fail(ctx, SWB_ERR_UNSUPPORTED, "pass-2 ring does not fit shared memory (lower SWB_TRACE_WC / SWB_TRACE_NB)");
    const size_t max_groups = (size_t)ctx->sm_count * 24 * groups_per_warp;   // up to 6 blocks per SM; the warps take tasks from a counter, surplus blocks exit at once
    const size_t groups = std::min<size_t>(lc.tasks.size(), max_groups);
    const size_t warps = (groups + groups_per_warp - 1) / groups_per_warp;
    const unsigned grid = (unsigned)((warps + warps_per_cta - 1) / warps_per_cta);
    CUDA_TRY(ctx->d_scratch.ensure((size_t)grid * warps_per_cta * tg.scratch_words * 4));
    tp.scratch = ctx->d_scratch.as<uint32_t>();
    tp.scratch_words = (unsigned long long)grid * warps_per_cta * tg.scratch_words;
    tp.out_score = ctx->d_score.as<int32_t>();
    tp.out_pos = ctx->d_pos.as<uint32_t>();
    tp.out_end = ctx->d_end.as<uint32_t>();
    tp.out_cx = ctx->d_cx.as<uint8_t>();
    tp.out_cy = ctx->d_cy.as<uint8_t>();
    tp.out_len = ctx->d_len.as<uint32_t>();
    tp.out_flags = ctx->d_flags.as<uint32_t>();
    tp.cons_cap = (uint32_t)ctx->cons_stride;
    tp.want_consensus = (ctx->flags & SWB_FLAG_CONSENSUS) ? 1 : 0;
    if (!tp.want_consensus) tp.cons_cap = 0x7FFFFFFFu;
    CUDA_TRY(ctx->d_next_task.ensure(256));
    CUDA_TRY(cudaMemsetAsync(ctx->d_next_task.p, 0, 4, ctx->stream));
    tp.next_task = ctx->d_next_task.as<uint32_t>();
    DevBuf d_cnt;
    tp.counters = nullptr;
    tp.dbg_flags = 0;
    if (dbg.on) { CUDA_TRY(d_cnt.ensure(128)); CUDA_TRY(cudaMemsetAsync(d_cnt.p, 0, 128, ctx->stream)); tp.counters = d_cnt.as<unsigned long long>(); }
    while (ctx->ev_pool.size() < ctx->ev_used + 2) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); ctx->ev_pool.push_back(e); }
    CUDA_TRY(cudaEventRecord(ctx->ev_pool[ctx->ev_used], ctx->stream));
    CUDA_TRY(launch_qs_trace(R, sat, dim3(grid), dim3(32 * warps_per_cta), smem_trace_total, ctx->stream, qt));
    CUDA_TRY(cudaEventRecord(ctx->ev_pool[ctx->ev_used + 1], ctx->stream));
    ctx->ev_used += 2;
    ctx->stats.kernel_launches++;
    dbg.mark("qs trace", L, R, lc.tasks.size());
    if (dbg.on) {
      unsigned long long c[16] = {0};
      cudaMemcpy(c, d_cnt.p, sizeof c, cudaMemcpyDeviceToHost);
      fprintf(stderr, "[swb200]   scan replays %llu, sessions %llu, scan rounds (warps) %llu, session rounds (warps) %llu, session steps %llu\n", c[0], c[1], c[2], c[3], c[4]);
      fprintf(stderr, "[swb200]   warp-cycles (M): prepare %.1f vmax %.1f search %.1f scan %.1f session %.1f walk %.1f\n", c[8] / 1e6, c[9] / 1e6, c[10] / 1e6, c[11] / 1e6, c[12] / 1e6, c[13] / 1e6);
      d_cnt.release();
    }
    ctx->stats.cells_pass2 += (uint64_t)lc.tasks.size() * (uint64_t)(3 * ctx->B + lc.max_m + 16) * L * R * 2ull;
    ctx->stats.lanes_per_pair = L; ctx->stats.rows_per_lane = R; ctx->stats.block_steps = ctx->B;
    ctx->stats.cols_per_step = 1; ctx->stats.kernel_kind = 2u;
  }
  return SWB_OK;
}

}  // namespace

// =========================================================================================================
extern "C" {

#ifdef SWB_CHECKED
const char* swb_version(void) { return "swb200 0.2 (sm_100a, SWB_CHECKED bounds checks)"; }
#else
const char* swb_version(void) { return "swb200 0.2 (sm_100a)"; }
#endif

int swb_device_count(void) {
  int ndev = 0;
  return cudaGetDeviceCount(&ndev) == cudaSuccess ? ndev : 0;
}

int swb_create(int device, swb_ctx** out) {
  if (!out) return SWB_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return SWB_ERR_CUDA;  // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return SWB_ERR_CUDA;
  swb_ctx* ctx = new swb_ctx();
  ctx->device = device;
  if (cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->sm_count <= 0) ctx->sm_count = 148;
  memset(ctx->code_of, 0xFF, sizeof ctx->code_of);
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SWB_ERR_CUDA; }
  for (auto& e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return SWB_ERR_CUDA; }
  *out = ctx;
  return SWB_OK;
}

void swb_destroy(swb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  free_classes(ctx, ctx->classes);
  for (auto& b : ctx->desc_pool) b.release();
  for (DevBuf* b : {&ctx->d_ref_raw, &ctx->d_ref_code, &ctx->d_table, &ctx->d_reads, &ctx->d_qpairs, &ctx->d_blkmax, &ctx->d_ckpt, &ctx->d_bnd, &ctx->d_units, &ctx->d_progress, &ctx->d_next_task,
                    &ctx->d_scratch, &ctx->d_taskmax, &ctx->d_winner, &ctx->d_score, &ctx->d_pos, &ctx->d_end, &ctx->d_cx, &ctx->d_cy,
                    &ctx->d_len, &ctx->d_flags, &ctx->d_xcode, &ctx->d_qs_table, &ctx->d_check}) b->release();
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* swb_last_error(const swb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int swb_set_scoring(swb_ctx* ctx, int mode, const float* table, float gap) {
  if (!ctx || !table) return SWB_ERR_ARG;
  if (mode != SWB_MODE_SAT_U8 && mode != SWB_MODE_EXACT) return fail(ctx, SWB_ERR_ARG, "unknown mode");
  HostScoring hs;
  hs.mode = mode;
  if (mode == SWB_MODE_SAT_U8) {
    // similaritymatrix.cpp:389-392: only two probes of the callback are used
    hs.M = saturate_u8(table['A' * 256 + 'A']);
    hs.X = saturate_u8(-table['A' * 256 + 'T']);
    hs.G = saturate_u8(gap);
    hs.match_shaped = true;
    hs.max_pos = hs.M;
  } else {
    if (gap != std::floor(gap) || gap < 0 || gap > 4000) return fail(ctx, SWB_ERR_SCORING, "EXACT mode needs an integer gap penalty in [0, 4000]");
    hs.G = (int)gap;
    hs.table.resize(65536);
    int mx = 0;
    for (int i = 0; i < 65536; ++i) {
      const float v = table[i];
      if (v != std::floor(v) || std::fabs(v) > 4000.f) return fail(ctx, SWB_ERR_SCORING, "EXACT mode needs integer-valued scores with |s| <= 4000");
      hs.table[i] = (int32_t)v;
      mx = std::max(mx, (int)v);
    }
    hs.max_pos = mx;
    const int32_t d = hs.table[0], o = hs.table[1];
    bool shaped = true;
    for (int a = 0; a < 256 && shaped; ++a)
      for (int b = 0; b < 256; ++b)
        if (hs.table[a * 256 + b] != (a == b ? d : o)) { shaped = false; break; }
    hs.match_shaped = shaped;
    if (shaped) { hs.M = d; hs.X = -o; hs.table.clear(); hs.table.shrink_to_fit(); }
  }
  ctx->sc = hs;
  ctx->table_dirty = true;
  ctx->staged = false;
  return SWB_OK;
}

int swb_set_scoring_match(swb_ctx* ctx, int mode, float match, float mismatch, float gap) {
  if (!ctx) return SWB_ERR_ARG;
  std::vector<float> t(65536, mismatch);
  for (int a = 0; a < 256; ++a) t[a * 256 + a] = match;
  return swb_set_scoring(ctx, mode, t.data(), gap);
}

static int set_reference_impl(swb_ctx* ctx, const char* y, size_t n);

int swb_set_reference(swb_ctx* ctx, const char* y, size_t n) {
  if (!ctx || (!y && n)) return SWB_ERR_ARG;
  const int rc = set_reference_impl(ctx, y, n);
  if (rc == SWB_OK) ctx->staged = false;
  return rc;
}

// Database search runs MANY references (queries) against ONE staged batch (mpi_sw_solve_uniprot.cpp: every query meets
// the whole database): when the batch is staged in query-stationary mode, its HBM-resident sequences, pairing and work
// buffers do not depend on the reference, only the kernel's row count does.  Swaps the reference without re-staging.
int swb_batch_rebind_reference(swb_ctx* ctx, const char* y, size_t n) {
  if (!ctx || !y || n == 0) return SWB_ERR_ARG;
  if (!ctx->staged || !ctx->qs || ctx->classes.empty()) return fail(ctx, SWB_ERR_STATE, "rebind needs a batch staged in query-stationary mode");
  const Geometry& geo = ctx->classes[0].geo;
  if (n > (size_t)geo.L * geo.R) return fail(ctx, SWB_ERR_STATE, "the new reference does not fit the staged lane geometry (stage again)");
  const uint64_t reach = std::min<uint64_t>(ctx->batch_max_m, n) * (uint64_t)std::max(1, ctx->sc.max_pos) + (uint64_t)ctx->sc.G + 16;
  if (ctx->sc.mode == SWB_MODE_EXACT && reach > 32000) return fail(ctx, SWB_ERR_STATE, "the new reference may leave the 16-bit lane range (stage again)");
  const int rc = set_reference_impl(ctx, y, n);
  if (rc != SWB_OK) { ctx->staged = false; return rc; }
  ctx->stats.cells_reference = ctx->batch_residues * (uint64_t)n;
  return SWB_OK;
}

static int set_reference_impl(swb_ctx* ctx, const char* y, size_t n) {
  if (n == 0) return fail(ctx, SWB_ERR_ARG, "empty reference");
  if (n > 0x7FF00000ull) return fail(ctx, SWB_ERR_UNSUPPORTED, "reference longer than 2^31 - 2^20 (step counters are 32-bit)");
  CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->y.assign((const uint8_t*)y, (const uint8_t*)y + n);
  memset(ctx->code_of, 0xFF, sizeof ctx->code_of);
  int K = 0;
  std::vector<uint8_t> codes(n);
  for (size_t i = 0; i < n; ++i) {
    uint8_t b = ctx->y[i];
    if (ctx->code_of[b] == 0xFF) ctx->code_of[b] = (uint8_t)K++;
    codes[i] = ctx->code_of[b];
  }
  if (K > 254) { K = 254; }   // profile select is not used with such alphabets (checked at stage time)
  ctx->KP = K + 1;
  CUDA_TRY(ctx->d_ref_raw.ensure(n + 64));
  CUDA_TRY(ctx->d_ref_code.ensure(n + 64));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_ref_raw.p, ctx->y.data(), n, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_ref_code.p, codes.data(), n, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->table_dirty = true;
  return SWB_OK;
}

int swb_make_string_range(int npiece, int64_t shortlen, int64_t longlen, float ratio, int64_t* left, int64_t* right) {
  // plocalaligner.cpp:44-67; the three asserts become SWB_ERR_RANGE
  if (npiece < 1 || !left || !right) return SWB_ERR_RANGE;
  const int64_t ov = (int64_t)((float)shortlen * ratio);
  if (npiece == 1) { left[0] = 0; right[0] = longlen; return 1; }
  const int64_t piece = (longlen + (int64_t)(npiece - 1) * ov) / npiece;
  if (ov > piece) return SWB_ERR_RANGE;
  int64_t l = 0, r = piece;
  int k = 0;
  left[k] = l; right[k] = r; ++k;
  while (k < npiece - 1) {
    l = std::max<int64_t>(0, r - ov);
    r = std::min(l + piece, longlen);
    left[k] = l; right[k] = r; ++k;
  }
  if (!(r < longlen)) return SWB_ERR_RANGE;
  left[k] = std::max<int64_t>(0, r - ov); right[k] = longlen; ++k;
  return k;
}

int swb_batch_stage(swb_ctx* ctx, const char* seqs, const uint64_t* offsets, size_t n_seqs, int npiece, float ratio, unsigned flags, size_t cons_stride) {
  if (!ctx) return SWB_ERR_ARG;
  if (ctx->y.empty()) return fail(ctx, SWB_ERR_STATE, "swb_set_reference has not been called");
  if (!seqs || !offsets || n_seqs == 0) return fail(ctx, SWB_ERR_ARG, "empty batch");
  if ((flags & SWB_FLAG_CONSENSUS) && cons_stride == 0) return fail(ctx, SWB_ERR_ARG, "cons_stride must be > 0 when consensus is requested");
  CUDA_TRY(cudaSetDevice(ctx->device));
  ctx->staged = false;
  // SWB_DEBUG_STAGE=1: host wall time of the stage's parts on stderr (where end-to-end time goes besides the kernels)
  const bool dbg_stage = getenv("SWB_DEBUG_STAGE") != nullptr;
  double ts_prev = dbg_stage ? DebugTimer::now() : 0.0;
  auto lap = [&](const char* what) { if (dbg_stage) { const double t = DebugTimer::now(); fprintf(stderr, "[swb200 stage] %-18s %8.3f ms\n", what, t - ts_prev); ts_prev = t; } };
  const size_t N = ctx->y.size();
  const size_t blob = offsets[n_seqs] - offsets[0];
  if (blob > 0xFFFF0000ull) return fail(ctx, SWB_ERR_UNSUPPORTED, "sequence blob larger than 4 GiB (split the batch)");
  const HostScoring& hs = ctx->sc;
  if (!hs.match_shaped && ctx->KP > 200) return fail(ctx, SWB_ERR_UNSUPPORTED, "tabulated scoring needs a reference alphabet of at most 199 symbols");

  ctx->n_seqs = n_seqs; ctx->npiece = npiece; ctx->ratio = ratio; ctx->flags = flags; ctx->cons_stride = cons_stride;
  const bool chunked = npiece >= 1;
  ctx->qs = false;
  if (!chunked) {
    const int qrc = stage_qs(ctx, seqs, offsets, n_seqs, flags);
    if (qrc < 0) return qrc;
    if (qrc == 0) {
      CUDA_TRY(ctx->d_score.ensure(n_seqs * 4));
      CUDA_TRY(ctx->d_pos.ensure(n_seqs * 4));
      CUDA_TRY(ctx->d_end.ensure(n_seqs * 8));
      CUDA_TRY(ctx->d_len.ensure(n_seqs * 4));
      CUDA_TRY(ctx->d_flags.ensure(n_seqs * 4));
      if (flags & SWB_FLAG_CONSENSUS) {
        CUDA_TRY(ctx->d_cx.ensure(n_seqs * cons_stride));
        CUDA_TRY(ctx->d_cy.ensure(n_seqs * cons_stride));
      }
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      ctx->staged = true;
      return SWB_OK;
    }
  }

  // tasks
  const int pieces = chunked ? npiece : 1;
  std::vector<TaskSeed> seeds;
  seeds.reserve(n_seqs * (size_t)pieces);
  std::map<uint32_t, std::vector<std::pair<int64_t, int64_t>>> ranges_by_m;
  uint64_t cells_ref = 0;
  size_t max_n = 0;
  uint32_t max_m = 0, min_m = 0xFFFFFFFFu;
  for (size_t r = 0; r < n_seqs; ++r) {
    const uint64_t m64 = offsets[r + 1] - offsets[r];
    if (m64 == 0) return fail(ctx, SWB_ERR_ARG, "empty sequence at index " + std::to_string(r));
    const uint32_t m = (uint32_t)m64, xo = (uint32_t)(offsets[r] - offsets[0]);
    max_m = std::max(max_m, m); min_m = std::min(min_m, m);
    cells_ref += (uint64_t)m * N;
    if (!chunked) { seeds.push_back({(uint32_t)r, 0u, 0u, (uint32_t)N, m, xo}); max_n = N; continue; }
    auto it = ranges_by_m.find(m);
    if (it == ranges_by_m.end()) {
      std::vector<int64_t> l(npiece), rr(npiece);
      const int k = swb_make_string_range(npiece, m, (int64_t)N, ratio, l.data(), rr.data());
      if (k < 0) return fail(ctx, SWB_ERR_RANGE, "_make_string_range precondition failed (plocalaligner.cpp:52,63,65) for a sequence of length " + std::to_string(m));
      std::vector<std::pair<int64_t, int64_t>> v;
      for (int i = 0; i < k; ++i) v.emplace_back(l[i], rr[i]);
      it = ranges_by_m.emplace(m, std::move(v)).first;
    }
    for (int pc = 0; pc < npiece; ++pc) {
      const auto& rg = it->second[pc];
      seeds.push_back({(uint32_t)r, (uint32_t)pc, (uint32_t)rg.first, (uint32_t)(rg.second - rg.first), m, xo});
      max_n = std::max(max_n, (size_t)(rg.second - rg.first));
    }
  }
  // Lane width: two s16 cells per register while the largest reachable score stays below 2^15, one s32 cell per register
  // beyond that (AM_WIDE).  The reference's f32 matrix is exact to 2^24 (similaritymatrix.cpp:49-54); past that its own
  // results are rounded, so there is nothing exact to reproduce and the batch is refused.
  {
    const uint64_t reach = (uint64_t)std::min<size_t>(max_m, max_n) * (uint64_t)std::max(1, hs.max_pos) + (uint64_t)hs.G + 16;
    ctx->wide = hs.mode == SWB_MODE_EXACT && (reach > 32000 || getenv("SWB_FORCE_WIDE"));
    if (ctx->wide && reach > (1ull << 24))
      return fail(ctx, SWB_ERR_UNSUPPORTED, "EXACT-mode scores may exceed 2^24 (min(len) * max score = " + std::to_string(reach) + "): the reference's f32 matrix is not exact there either");
  }
  // checkpoint period B (steps between register-state checkpoints): as small as the HBM budget for the
  // checkpoints allows (pass 2 recomputes O(B) columns per alignment), at least 32.
  {
    size_t budget_mb = 8192;
    if (const char* e = getenv("SWB_CKPT_BUDGET_MB")) budget_mb = (size_t)std::max(64L, atol(e));
    const size_t pairs_in_flight = chunked ? (seeds.size() + 1) / 2 : std::min((seeds.size() + 1) / 2, chunk_pairs(ctx));
    const double words_per_block = (double)pairs_in_flight * ((double)max_m * 1.08 + 40.0) * (hs.mode == SWB_MODE_SAT_U8 ? 0.5 : 1.0);
    int B = 32;
    // columns per step: 1 everywhere (the software pipeline gives the ILP) except for few long pairs, where one
    // warp per SM sub-partition is latency bound and two columns per step amortise the per-step latency
    // (few long pairs), and for batches too small to put more than one warp on an SM sub-partition
    ctx->C = (!ctx->force_l32 && (((seeds.size() + 1) / 2 < 148 * 8 && max_m > 1024) || (seeds.size() + 1) / 2 <= 148 * 4)) ? 2 : 1;
    // four columns per step for the pipelined strips of few long pairs (score_units_kernel; instantiated for R <= 8):
    // every sequence must span several strips, so that no other kernel sees the batch
    bool few_long = false;
    const int r_long = strip_rows(ctx, max_m, (uint32_t)max_n, (seeds.size() + 1) / 2, use_profile(ctx, false), &few_long);
    const bool c4_ok = !ctx->force_l32 && few_long && r_long <= 8 && min_m > (uint32_t)(32 * r_long) && !getenv("SWB_NO_PIPELINE");
    if (c4_ok) ctx->C = 8;      // measured at 10 kbp x 51 Mbp: 2.20 / 2.40 TCUPS at 4 / 8 columns per step (8 rows per lane), 2.43 / 2.55 (4 rows)
    if (ctx->force_l32) ctx->C = 1;
    else if (const char* e = getenv("SWB_COLS")) { const int c = atoi(e); ctx->C = ((c == 4 || c == 8) && c4_ok) ? c : (c >= 2 ? 2 : 1); }
    while (B < 65536 && words_per_block * 4.0 * ((double)max_n / ((double)B * ctx->C) + 1.0) > (double)budget_mb * 1048576.0) B <<= 1;
    // SWB_FORCE_B: run small test batches with the checkpoint period a large batch would get (parity of the timed geometry)
    if (const char* e = getenv("SWB_FORCE_B")) B = std::max(32, std::min(65536, 1 << ilog2(atoi(e))));
    ctx->B = B; ctx->logB = ilog2(B);
  }
  lap("seeds + geometry");
  int rc = upload_profile_table(ctx);
  if (rc) return rc;
  free_classes(ctx, ctx->classes);
  lap("free classes");
  rc = build_classes(ctx, seeds, pieces, &ctx->classes);
  if (rc) return rc;
  lap("build classes");
  CUDA_TRY(ctx->d_reads.ensure(blob + 64));
  CUDA_TRY(cudaMemcpyAsync(ctx->d_reads.p, seqs + offsets[0], blob, cudaMemcpyHostToDevice, ctx->stream));
  lap("H2D sequences");
  rc = upload_classes(ctx, ctx->classes);
  if (rc) return rc;
  lap("upload classes");
  // outputs
  CUDA_TRY(ctx->d_score.ensure(n_seqs * 4));
  CUDA_TRY(ctx->d_pos.ensure(n_seqs * 4));
  CUDA_TRY(ctx->d_end.ensure(n_seqs * 8));
  CUDA_TRY(ctx->d_len.ensure(n_seqs * 4));
  CUDA_TRY(ctx->d_flags.ensure(n_seqs * 4));
  if (flags & SWB_FLAG_CONSENSUS) {
    CUDA_TRY(ctx->d_cx.ensure(n_seqs * cons_stride));
    CUDA_TRY(ctx->d_cy.ensure(n_seqs * cons_stride));
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  lap("outputs + sync");
  ctx->stats = swb_stats{};
  ctx->stats.cells_reference = cells_ref;
  ctx->staged = true;
  return SWB_OK;
}

int swb_batch_run(swb_ctx* ctx, float* device_us) {
  if (!ctx) return SWB_ERR_ARG;
  if (!ctx->staged) return fail(ctx, SWB_ERR_STATE, "no staged batch");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const uint64_t cells_ref = ctx->stats.cells_reference;
  ctx->stats = swb_stats{};
  ctx->stats.cells_reference = cells_ref;
  const bool chunked = ctx->npiece >= 1;
  const bool realign = chunked && !ctx->sc.is_default();
#ifdef SWB_CHECKED
  CUDA_TRY(ctx->d_check.ensure(4));
  CUDA_TRY(cudaMemsetAsync(ctx->d_check.p, 0, 4, ctx->stream));
#endif
  CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
  ctx->ev_used = 0;
  if (ctx->qs) {
    int rc = run_qs(ctx);
    if (rc) return rc;
  } else if (!realign) {
    int rc = run_classes(ctx, ctx->classes.data(), ctx->classes.size(), false, chunked, true);
    if (rc) return rc;
  } else {
    // custom scoring + chunking: pick the piece with the constructor's scoring (plocalaligner.cpp:113-129),
    // then re-align the winner with the DEFAULT scoring (plocalaligner.cpp:132-136).
    std::vector<TaskSeed> winners(ctx->n_seqs);
    for (size_t c = 0; c < ctx->classes.size(); ++c) {
      LaunchClass& lc = ctx->classes[c];
      int rc = run_classes(ctx, &lc, 1, false, true, false);
      if (rc) return rc;
      std::vector<uint32_t> win(lc.nreads);
      CUDA_TRY(cudaMemcpyAsync(win.data(), ctx->d_winner.p, win.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      for (int lr = 0; lr < lc.nreads; ++lr) {
        const TaskDesc& td = lc.tasks[win[lr]];
        const PairDesc& pd = lc.pairs[td.pair];
        winners[td.out] = TaskSeed{td.out, 0u, pd.y_off, pd.n, td.half ? pd.mB : pd.mA, td.half ? pd.xB : pd.xA};
      }
    }
    std::vector<LaunchClass> second;
    int rc = build_classes(ctx, winners, 1, &second);
    if (!rc) rc = upload_classes(ctx, second);
    if (!rc) rc = run_classes(ctx, second.data(), second.size(), true, false, true);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    free_classes(ctx, second);
    if (rc) return rc;
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return SWB_ERR_CUDA; }
  }
  CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(cudaGetLastError());
#ifdef SWB_CHECKED
  {
    uint32_t site = 0;
    CUDA_TRY(cudaMemcpy(&site, ctx->d_check.p, 4, cudaMemcpyDeviceToHost));
    if (site) return fail(ctx, SWB_ERR_CUDA, "SWB_CHECKED: bounds check failed at site " + std::to_string(site));
  }
#endif
  float ms = 0, ms2 = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) { float t = 0; CUDA_TRY(cudaEventElapsedTime(&t, ctx->ev_pool[i], ctx->ev_pool[i + 1])); ms2 += t; }
  const float total_us = ms * 1000.f;
  ctx->stats.pass2_us = ms2 * 1000.f;
  ctx->stats.pass1_us = total_us - ctx->stats.pass2_us;
  if (device_us) *device_us = total_us;
  return SWB_OK;
}

int swb_batch_fetch(swb_ctx* ctx, int32_t* score, uint32_t* pos, uint32_t* end_xy, char* cons_x, char* cons_y, uint32_t* cons_len, uint32_t* out_flags) {
  if (!ctx) return SWB_ERR_ARG;
  if (!ctx->staged) return fail(ctx, SWB_ERR_STATE, "no staged batch");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t n = ctx->n_seqs;
  if (score) CUDA_TRY(cudaMemcpyAsync(score, ctx->d_score.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (pos) CUDA_TRY(cudaMemcpyAsync(pos, ctx->d_pos.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (end_xy) CUDA_TRY(cudaMemcpyAsync(end_xy, ctx->d_end.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (cons_len) CUDA_TRY(cudaMemcpyAsync(cons_len, ctx->d_len.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_flags) CUDA_TRY(cudaMemcpyAsync(out_flags, ctx->d_flags.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if ((ctx->flags & SWB_FLAG_CONSENSUS) && cons_x) CUDA_TRY(cudaMemcpyAsync(cons_x, ctx->d_cx.p, n * ctx->cons_stride, cudaMemcpyDeviceToHost, ctx->stream));
  if ((ctx->flags & SWB_FLAG_CONSENSUS) && cons_y) CUDA_TRY(cudaMemcpyAsync(cons_y, ctx->d_cy.p, n * ctx->cons_stride, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return SWB_OK;
}

int swb_batch_device_results(swb_ctx* ctx, const int32_t** d_score, const uint32_t** d_pos) {
  if (!ctx || !ctx->staged) return SWB_ERR_STATE;
  if (d_score) *d_score = ctx->d_score.as<int32_t>();
  if (d_pos) *d_pos = ctx->d_pos.as<uint32_t>();
  return SWB_OK;
}

int swb_align_batch(swb_ctx* ctx, const char* seqs, const uint64_t* offsets, size_t n_seqs, int npiece, float ratio, unsigned flags,
                    int32_t* score, uint32_t* pos, uint32_t* end_xy, char* cons_x, char* cons_y, uint32_t* cons_len, size_t cons_stride,
                    uint32_t* out_flags, float* device_us) {
  int rc = swb_batch_stage(ctx, seqs, offsets, n_seqs, npiece, ratio, flags, cons_stride);
  if (rc) return rc;
  rc = swb_batch_run(ctx, device_us);
  if (rc) return rc;
  return swb_batch_fetch(ctx, score, pos, end_xy, cons_x, cons_y, cons_len, out_flags);
}

int swb_last_stats(const swb_ctx* ctx, swb_stats* out) {
  if (!ctx || !out) return SWB_ERR_ARG;
  *out = ctx->stats;
  return SWB_OK;
}

int swb_matrix(swb_ctx* ctx, const char* x, size_t m, int32_t* out) {
  // Dense H through the device path: stage the single sequence (L = 32 geometry, strips as needed), run the
  // score pass once (it writes the strip boundary rows the replay reads), then replay with a store hook.
  if (!ctx || !x || !out || m == 0) return SWB_ERR_ARG;
  if (ctx->y.empty()) return fail(ctx, SWB_ERR_STATE, "swb_set_reference has not been called");
  const size_t n = ctx->y.size();
  if ((m + 1) * (n + 1) > ((size_t)1 << 28)) return fail(ctx, SWB_ERR_UNSUPPORTED, "swb_matrix is for small inputs: (m+1)*(n+1) must not exceed 2^28 cells");
  const uint64_t offs[2] = {0, m};
  ctx->force_l32 = true;
  int rc = swb_batch_stage(ctx, x, offs, 1, 0, 0.f, 0, 0);
  ctx->force_l32 = false;
  if (rc) return rc;
  rc = swb_batch_run(ctx, nullptr);
  if (rc) return rc;
  LaunchClass& lc = ctx->classes[0];
  if (lc.geo.L != 32) {
    // re-stage with the strip geometry forced to 32 lanes: cheap, the input is small
    return fail(ctx, SWB_ERR_UNSUPPORTED, "internal: swb_matrix expects the 32-lane geometry");
  }
  DevBuf d_out;
  CUDA_TRY(d_out.ensure((m + 1) * (n + 1) * sizeof(int32_t)));
  CUDA_TRY(cudaMemsetAsync(d_out.p, 0, (m + 1) * (n + 1) * sizeof(int32_t), ctx->stream));
  const bool sat = ctx->sc.mode == SWB_MODE_SAT_U8;
  const bool profile = use_profile(ctx, false);
  DumpParams dp{};
  PassParams& pp = dp.pp;
  pp.ref_raw = ctx->d_ref_raw.as<uint8_t>(); pp.ref_code = ctx->d_ref_code.as<uint8_t>(); pp.reads_raw = ctx->d_reads.as<uint8_t>();
  pp.qpairs = ctx->d_qpairs.as<uint32_t>(); pp.table = ctx->d_table.as<int16_t>(); pp.KP = ctx->KP;
  pp.pairs = lc.d_pairs.as<PairDesc>(); pp.npairs = 1;
  pp.blkmax = ctx->d_blkmax.as<uint32_t>(); pp.ckpt = ctx->d_ckpt.as<uint32_t>(); pp.bnd = ctx->d_bnd.as<uint32_t>();
  pp.L = lc.geo.L; pp.logL = lc.geo.logL; pp.B = ctx->B; pp.logB = ctx->logB;
  const bool wide = ctx->wide && !sat;
  pp.sc = device_scoring(ctx->sc, false, wide);
  dp.out = d_out.as<int32_t>(); dp.m = (int)m; dp.n = (int)n;
  const size_t smem = profile ? (size_t)ctx->KP * lc.geo.R * 32 * 4 : 0;
  cudaError_t e = launch_dump(lc.geo.R, sat ? AM_SAT : (wide ? AM_WIDE : AM_EXACT), profile, smem, ctx->stream, dp);
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out.p, (m + 1) * (n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  d_out.release();
  if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return SWB_ERR_CUDA; }
  return SWB_OK;
}

}  // extern "C"
