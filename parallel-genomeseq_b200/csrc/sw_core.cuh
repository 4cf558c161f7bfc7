// sw_core.cuh — sm_100a Smith-Waterman wavefront core (device code).
//
// Replaces, for the alignment hot path of kosta777/parallel-genomeseq (citations relative to
// /root/reference/):
//   Similarity_Matrix_Skewed::iterate        src/aligner/similaritymatrix.cpp:386-561   (SAT_U8 mode)
//   Similarity_Matrix::iterate               src/aligner/similaritymatrix.cpp:99-116,247-257 (EXACT mode)
//   ..::find_index_of_maximum                src/aligner/similaritymatrix.cpp:291-299 / :21-28
//   SWAligner::traceback                     src/aligner/smithwaterman.cpp:40-78
//
// Design (B200-first, not a port of the AVX2 skewed-matrix code):
//   * The full H matrix is never materialised.  A "pair" packs TWO alignments against the same
//     y-range into the two signed 16-bit halves of every 32-bit register (s16x2); one DP cell pair
//     costs three DPX instructions plus half a VIMNMX3 — see step() below.
//   * L lanes (a power of two, a sub-warp "group") stripe the rows of one pair, R rows per lane in
//     registers.  Lane g works on column j = t - g at step t (a skewed wavefront); the row carry
//     between neighbouring lanes is one __shfl_up_sync per step.
//   * Values are carried as E = H - G (G = gap penalty): then
//         d    = max(E_nw + (s+G), E_w, 0)            one VIADDMNMX.S16x2.RELU   (= max(NW+s, W-G, 0))
//         dG   = min(d - G, 255 - G)                  one VIADDMNMX (min form)   (u8 saturation, SAT_U8)
//         E    = max(E_n - G, dG)                     one VIADDMNMX              (= H - G)
//         bmax = max(bmax, E, E')                     one VIMNMX3.S16x2 per two cells
//   * Symbol scores come from a per-warp query profile in shared memory (one LDS per cell pair, fetched
//     one step ahead into registers) or, for big alphabets with match/mismatch scoring, from HSET2 + LOP3.
//   * Every B steps each group stores its block maximum and every lane checkpoints its register state.
//     Pass 2 (locate + traceback) restarts the same wavefront from a checkpoint, so the arg-max
//     tie-break of the reference and its value-greedy traceback are reproduced exactly while only
//     O(B * m) cells are ever recomputed.
//   * Sequences longer than one warp holds are cut into row strips that hand their last row over
//     through HBM; with few long pairs the strips of a pair run concurrently, one warp each.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace swb {

constexpr int MODE_SAT_U8 = 0;
constexpr int MODE_EXACT = 1;

// One packed pair of alignments (halves A = low 16 bits, B = high 16 bits).
struct PairDesc {
  uint32_t q_off;     // word offset into qpairs: L*R packed rows
  uint32_t y_off;     // first y position of this pair's range (0-based)
  uint32_t n;         // columns in the range
  uint32_t nblk;      // ceil((L - 1 + ceil(n / C)) / B) blocks of B steps
  uint32_t mA, mB;    // real row counts (0 = empty half)
  uint32_t xA, xB;    // byte offsets of the raw sequences in reads_raw
  uint64_t blk_off;   // word offset into blkmax: nstrips * nblk words (maximum over the L lanes of the group)
  uint64_t ck_off;    // word offset into ckpt  : nstrips * nblk * (R+C) * L words
  uint64_t bnd_off;   // word offset into bnd   : (nstrips-1) * (n+1) words (bottom row of every strip but the last)
  uint32_t nstrips;   // row strips of L*R rows (1 unless the sequences are longer than one warp can hold; then L == 32)
  uint32_t pad_;
};

struct Scoring {
  uint32_t negG2;     // pack(-G, -G)
  uint32_t sel_and;   // (s_match+G) ^ (s_mismatch+G), both halves
  uint32_t sel_xor;   // pack(s_mismatch+G, ...)
  uint32_t ceil2;     // pack(255-G, 255-G)   (SAT_U8)
  int32_t G;
};

// Own bounds checks (-DSWB_CHECKED, libswb200_checked.so): compute-sanitizer is closed on the GPU pool this was built on, so
// every store into a work buffer, every ring access of pass 2 and the unmasked reference loads carry an index check that
// records a site code in PassParams::check (the host turns a non-zero code into an error).  Compiled out otherwise.
#ifdef SWB_CHECKED
#define SWB_CHECK(chk, cond, code) do { if (!(cond) && (chk)) atomicMax((chk), (uint32_t)(code)); } while (0)
#else
#define SWB_CHECK(chk, cond, code) do { } while (0)
#endif

struct PassParams {
  const uint8_t* ref_raw;      // y bytes
  const uint8_t* ref_code;     // y alphabet codes (profile select)
  const uint8_t* reads_raw;    // concatenated x bytes
  const uint32_t* qpairs;      // packed per-row symbols (compare select) or x bytes (profile select)
  const int16_t* table;        // [256][KP] (s + G) per (x byte, y code), profile select
  int KP;                      // y alphabet size + 1 (sentinel code = KP-1)
  const PairDesc* pairs;
  int npairs;
  uint32_t* blkmax;
  uint32_t* ckpt;
  uint32_t* bnd;               // strip boundary rows (packed E words), see PairDesc::bnd_off
  // pipelined strips (few long pairs): one warp per (pair, strip) unit, strips of a pair run concurrently and
  // hand their boundary rows over through HBM, synchronised by a per-unit progress counter
  const uint2* units;          // (pair, strip) per warp, null = one warp per pair
  int nunits;
  uint32_t* progress;          // per unit: boundary columns published so far
  uint32_t* abort_flag;        // set when a consumer gave up waiting (never expected; avoids a hang)
  uint32_t* ticket;            // tasks are handed out in the order in which warps ask for them (score_units_kernel)
  uint32_t* blocks_done;       // per unit: tasks finished so far
  int task_blocks;             // checkpoint blocks per task
  uint32_t ntickets;           // rounds * nunits
  int strips;                  // host: some pair of the class has more than one row strip (score_strips_kernel instead of score_kernel)
  uint32_t* check;             // SWB_CHECKED builds: highest failing check site (0 = none), null otherwise
  unsigned long long ck_words, blk_words, bnd_words, ref_len;   // sizes of the work buffers (for the checks)
  int L, logL, B, logB;
  Scoring sc;
};

// ---- DPX / packed-half primitives -----------------------------------------------------------------
__device__ __forceinline__ uint32_t hset2_eq(uint32_t a, uint32_t b) {
  // HSET2.EQ: per 16-bit half 0xFFFF where the halves are equal (operands: SYM_BASE | symbol).
  uint32_t r;
  asm("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t viaddmin_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }

constexpr uint32_t NEG_INF2 = 0x80008000u;   // pack(-32768, -32768); as one s32 (wide lanes) it is about -2^31: "minus infinity" for a max either way

// Arithmetic of a lane register (template parameter AM of everything below):
//   AM_EXACT  two s16 cells per register, no upper clamp          (Similarity_Matrix, scores < 2^15)
//   AM_SAT    two s16 cells per register, clamped to [0, 255]     (Similarity_Matrix_Skewed)
//   AM_WIDE   ONE s32 cell per register, no upper clamp           (Similarity_Matrix with scores >= 2^15: the reference's f32
//             matrix is exact to 2^24, similaritymatrix.cpp:49-54; same instruction count as AM_EXACT, half the cells per
//             instruction; only half A of a pair exists)
constexpr int AM_EXACT = 0, AM_SAT = 1, AM_WIDE = 2;
template <bool WIDE> __device__ __forceinline__ uint32_t lane_max(uint32_t a, uint32_t b) { return WIDE ? (uint32_t)max((int)a, (int)b) : __vmaxs2(a, b); }
// value of one alignment's cell in a packed word: the task's half (s16) or the whole word (wide lanes)
template <bool WIDE> __device__ __forceinline__ int lane_val(uint32_t v, uint32_t half) { return WIDE ? (int)v : (int)(int16_t)(half ? (v >> 16) : (v & 0xFFFFu)); }
// Symbols are compared as fp16 bit patterns by HSET2; 0x4000 | byte is a NORMAL fp16 number (2.0 .. 2.5),
// so the comparison never touches subnormals, signed zeros or NaNs.
constexpr uint32_t SYM_BASE = 0x4000u;
constexpr uint32_t SENT_Y = SYM_BASE | 0x0100u;   // never equals a byte symbol
constexpr uint32_t SENT_X = SYM_BASE | 0x0101u;   // never equals a byte symbol nor SENT_Y

__device__ __forceinline__ int warp_max_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Per-lane register state of the wavefront for one pair.  A step advances every lane by C columns.
template <int R, int C>
struct LaneState {
  uint32_t E[R];        // H - G of this lane's R rows at the last column of the previous step
  uint32_t up_prev;     // H - G of the row above this lane's first row at that same column
  uint32_t bot[C];      // H - G of this lane's LAST row at each of the C columns of the previous step
                        // (bot[C-1] == E[R-1]); the lane below shuffles them in as its "north" inputs
};
// Checkpoint words per lane: E[R], up_prev, bot[0..C-2] = R + C registers.  In SAT_U8 mode every value is
// E + G in [0, 255], so two registers (four bytes) share one word and the checkpoints take half the HBM.
template <int R, int C, int AM> __host__ __device__ constexpr int state_words() { return AM == AM_SAT ? (R + C + 1) / 2 : R + C; }

// Geometry of the skewed wavefront: at step t (1-based) lane g works on columns col_of(t,g,0..C-1).
template <int C> __device__ __forceinline__ int col_of(int t, int g, int c) { return C * (t - g - 1) + 1 + c; }
template <int C> __device__ __forceinline__ int step_of(int j, int g) { return g + (j + C - 1) / C; }

// Symbol selection: returns pack(s+G) for row k against column c of the current step.
template <int R, int C, bool WIDE = false>
struct CompareSelect {
  uint32_t q[R];        // pack(symA[row], symB[row])
  uint32_t r2[C];       // pack(y[j_c], y[j_c])
  uint32_t sel_and, sel_xor;
  __device__ __forceinline__ void set_column(int c, uint32_t ysym) { r2[c] = ysym * 0x00010001u; }
  __device__ __forceinline__ uint32_t operator()(int k, int c) const {
    if (WIDE) return (((q[k] ^ r2[c]) & 0xFFFFu) == 0u ? sel_and : 0u) ^ sel_xor;      // half A only, one s32 score
    return (hset2_eq(q[k], r2[c]) & sel_and) ^ sel_xor;
  }
};

// Rows per lane padded for 128-bit profile loads (VEC layout below): a multiple of 4 words that is 4 mod 8, so that the
// 8 lanes of a quarter warp — one LDS.128 wavefront — hit 32 distinct banks.
template <int R> __host__ __device__ constexpr int prof_stride() { return ((R + 3) / 4 * 4) % 8 == 4 ? (R + 3) / 4 * 4 : (R + 3) / 4 * 4 + 4; }

// Profile selection: per-warp query profile in shared memory.
//   scalar layout (pass 2, dense dump): word (code*R + k)*32 + lane — bank = lane, one LDS per cell pair;
//   VEC layout (the score kernels): word (code*32 + lane)*S + k, S = prof_stride<R>() — the R scores of a lane's rows for
//   one column symbol are contiguous and fetched four rows per LDS.128 into registers (RegSelect::fetch).
template <int R, int C, bool VEC = false>
struct ProfileSelect {
  const uint32_t* prof;   // shared memory, already offset by lane (VEC: by lane * S)
  const uint32_t* col[C];
  __device__ __forceinline__ void set_column(int c, uint32_t ycode) { col[c] = prof + ycode * ((VEC ? prof_stride<R>() : R) * 32); }
  __device__ __forceinline__ uint32_t operator()(int k, int c) const { return VEC ? col[c][k] : col[c][k * 32]; }
};

// One wavefront step for one lane: C columns of rows [g*R, (g+1)*R).  The C column chains are independent
// up to a one-row skew, which is where the instruction-level parallelism comes from.
//   hook(k, c, E_new): every new cell, E-space (H - G), packed s16x2.
template <int R, int C, int AM, class Select, class Hook>
__device__ __forceinline__ void step(LaneState<R, C>& st, const Select& sel, const Scoring& sc, const uint32_t (&upv)[C],
                                     uint32_t& bmax, Hook&& hook) {
  constexpr bool SAT = AM == AM_SAT, WIDE = AM == AM_WIDE;
  uint32_t above[C + 1];          // row k-1: [0] = last column of the previous step, [c+1] = column c of this step
  above[0] = st.up_prev;
#pragma unroll
  for (int c = 0; c < C; ++c) above[c + 1] = upv[c];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    uint32_t cur[C + 1];
    cur[0] = st.E[k];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const uint32_t sG = sel(k, c);
      if (WIDE) {                                                                       // the same three operations on one s32 cell
        const int d = __viaddmax_s32_relu((int)above[c], (int)sG, (int)cur[c]);
        cur[c + 1] = (uint32_t)__viaddmax_s32((int)above[c + 1], (int)sc.negG2, d + (int)sc.negG2);
      } else {
        const uint32_t d = __viaddmax_s16x2_relu(above[c], sG, cur[c]);                  // max(NW + s, W - G, 0)
        const uint32_t dG = SAT ? viaddmin_s16x2(d, sc.negG2, sc.ceil2) : __vadd2(d, sc.negG2);   // min(., 255) - G
        cur[c + 1] = __viaddmax_s16x2(above[c + 1], sc.negG2, dG);                       // max(N - G, .) = H - G
      }
      hook(k, c, cur[c + 1]);
    }
    if (C % 2 == 0) {
#pragma unroll
      for (int c = 0; c < C; c += 2) bmax = WIDE ? (uint32_t)__vimax3_s32((int)bmax, (int)cur[c + 1], (int)cur[c + 2]) : __vimax3_s16x2(bmax, cur[c + 1], cur[c + 2]);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) bmax = lane_max<WIDE>(bmax, cur[c + 1]);
    }
    st.E[k] = cur[C];
#pragma unroll
    for (int c = 0; c <= C; ++c) above[c] = cur[c];
  }
#pragma unroll
  for (int c = 0; c < C; ++c) st.bot[c] = above[c + 1];
  st.up_prev = upv[C - 1];
}


// y symbol for column j (1-based) of a pair's range, or the sentinel outside [1, n].  MASKED = false skips the
// range test (interior blocks of the score pass, where every lane's column is known to be inside the range).
template <bool PROFILE, bool MASKED = true>
__device__ __forceinline__ uint32_t load_y(const PassParams& p, const PairDesc& pd, int j) {
  if (!MASKED) {
    const uint32_t idx = pd.y_off + (uint32_t)(j - 1);
    SWB_CHECK(p.check, j >= 1 && (unsigned long long)idx < p.ref_len, 1);
    if (PROFILE) return __ldg(p.ref_code + idx);
    return SYM_BASE | __ldg(p.ref_raw + idx);
  }
  const bool in = (j >= 1) && (j <= (int)pd.n);
  const uint32_t idx = in ? (pd.y_off + (uint32_t)(j - 1)) : pd.y_off;
  if (PROFILE) { uint32_t c = __ldg(p.ref_code + idx); return in ? c : (uint32_t)(p.KP - 1); }
  uint32_t c = __ldg(p.ref_raw + idx);
  return in ? (SYM_BASE | c) : SENT_Y;
}

// Scores of one step held in registers (filled from the shared-memory profile one step ahead of their use).
template <int R, int C>
struct RegSelect {
  uint32_t s[R][C];
  __device__ __forceinline__ uint32_t operator()(int k, int c) const { return s[k][c]; }
  template <class Sel>
  __device__ __forceinline__ void fetch(const Sel& sel) {
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
      for (int c = 0; c < C; ++c) s[k][c] = sel(k, c);
  }
  // VEC profile: four rows per LDS.128
  __device__ __forceinline__ void fetch(const ProfileSelect<R, C, true>& sel) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const uint4* q = reinterpret_cast<const uint4*>(sel.col[c]);
#pragma unroll
      for (int k4 = 0; k4 < (R + 3) / 4; ++k4) {
        const uint4 v = q[k4];
        s[4 * k4][c] = v.x;
        if (4 * k4 + 1 < R) s[4 * k4 + 1][c] = v.y;
        if (4 * k4 + 2 < R) s[4 * k4 + 2][c] = v.z;
        if (4 * k4 + 3 < R) s[4 * k4 + 3][c] = v.w;
      }
    }
  }
};

template <int R, int C>
__device__ __forceinline__ void init_state(LaneState<R, C>& st, const Scoring& sc) {
#pragma unroll
  for (int k = 0; k < R; ++k) st.E[k] = sc.negG2;
  st.up_prev = sc.negG2;
#pragma unroll
  for (int c = 0; c < C; ++c) st.bot[c] = sc.negG2;
}

// Checkpoint layout: word ((unit * state_words + w) * L + g).  Register order: E[0..R-1], up_prev, bot[0..C-2].
template <int R, int C>
__device__ __forceinline__ uint32_t state_reg(const LaneState<R, C>& st, int i) {
  return i < R ? st.E[i] : (i == R ? st.up_prev : st.bot[i - R - 1]);
}
template <int R, int C, int AM>
__device__ __forceinline__ void save_state(const LaneState<R, C>& st, const Scoring& sc, uint32_t* ck, int L, int g) {
  constexpr bool SAT = AM == AM_SAT;
  constexpr int N = R + C;
  if (SAT) {
    const uint32_t g2 = (uint32_t)sc.G * 0x00010001u;   // pack(+G, +G): E + G is in [0, 255]
#pragma unroll
    for (int w = 0; w < (N + 1) / 2; ++w) {
      const uint32_t a = __vadd2(state_reg<R, C>(st, 2 * w), g2);
      const uint32_t b = (2 * w + 1 < N) ? __vadd2(state_reg<R, C>(st, 2 * w + 1), g2) : 0u;
      ck[w * L + g] = __byte_perm(a, b, 0x6420);   // bytes: a.lo, a.hi, b.lo, b.hi
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) ck[i * L + g] = state_reg<R, C>(st, i);
  }
}
// value of checkpoint register i (packed E word) from a checkpoint
template <int R, int C, int AM>
__device__ __forceinline__ uint32_t ck_reg(const Scoring& sc, const uint32_t* ck, int L, int g, int i) {
  if (AM == AM_SAT) {
    const uint32_t w = ck[(i >> 1) * L + g];
    const uint32_t v = (i & 1) ? __byte_perm(w, 0u, 0x4342) : __byte_perm(w, 0u, 0x4140);   // two bytes -> two u16 halves
    return __vadd2(v, sc.negG2);
  }
  return ck[i * L + g];
}
template <int R, int C, int AM>
__device__ __forceinline__ void load_state(LaneState<R, C>& st, const Scoring& sc, const uint32_t* ck, int L, int g) {
#pragma unroll
  for (int k = 0; k < R; ++k) st.E[k] = ck_reg<R, C, AM>(sc, ck, L, g, k);
  st.up_prev = ck_reg<R, C, AM>(sc, ck, L, g, R);
#pragma unroll
  for (int c = 0; c < C - 1; ++c) st.bot[c] = ck_reg<R, C, AM>(sc, ck, L, g, R + 1 + c);
  st.bot[C - 1] = st.E[R - 1];
}

template <int R, int C, bool WIDE>
__device__ __forceinline__ void load_compare_rows(CompareSelect<R, C, WIDE>& sel, const PassParams& p, const PairDesc& pd, int g) {
  const uint32_t* q = p.qpairs + pd.q_off + (uint32_t)g * R;
#pragma unroll
  for (int k = 0; k < R; ++k) sel.q[k] = __ldg(q + k);
  sel.sel_and = p.sc.sel_and;
  sel.sel_xor = p.sc.sel_xor;
}

// Build this lane's slice of the per-warp profile: prof[(code*R + k)*32 + lane] = pack(T[xA][code], T[xB][code]).
template <int R, bool WIDE = false, bool VEC = false>
__device__ __forceinline__ void build_profile(uint32_t* prof_warp, const PassParams& p, const PairDesc& pd, int g, int lane) {
  const uint32_t* q = p.qpairs + pd.q_off + (uint32_t)g * R;
  for (int k = 0; k < R; ++k) {
    const uint32_t w = __ldg(q + k);
    const uint32_t a = (w & 0xFFFFu) - SYM_BASE, b = (w >> 16) - SYM_BASE;
    for (int c = 0; c < p.KP; ++c) {
      // sentinel rows (a/b >= 256) and the sentinel column use the table's last row/column: never a match
      const int16_t sa = p.table[(a < 256 ? a : 256) * p.KP + c];
      const int16_t sb = p.table[(b < 256 ? b : 256) * p.KP + c];
      prof_warp[VEC ? (c * 32 + lane) * prof_stride<R>() + k : (c * R + k) * 32 + lane] = WIDE ? (uint32_t)(int32_t)sa : ((uint32_t)(uint16_t)sa | ((uint32_t)(uint16_t)sb << 16));
    }
  }
}

// maximum of a packed s16x2 word over the L lanes of a group (all 32 lanes call it)
template <bool WIDE = false>
__device__ __forceinline__ uint32_t group_max_s16x2(uint32_t v, int L) {
  for (int o = L >> 1; o > 0; o >>= 1) v = lane_max<WIDE>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Shared driver of both passes: restore (or initialise) the lane state, then run `nsteps` steps after
// step t0.  All 32 lanes execute it together (full-mask shuffles); hook(k, c, t, j, E_new) fires for steps <= t1.
//
// Row strips: a pair whose sequences do not fit L*R rows is cut into nstrips strips of L*R rows (L == 32).
// Strips run top to bottom; the last row of strip s is written to a boundary row in HBM (one packed word
// per column) and read back as the "north" input of strip s+1 — 32 columns per coalesced load, handed to
// lane 0 with one shuffle per step.  BND selects that code path (compiled out for single-strip launches).
template <int R, int C, int AM, bool PROFILE, bool VEC = false>
struct Wavefront {
  static constexpr bool SAT = AM == AM_SAT, WIDE = AM == AM_WIDE;
  static constexpr int PROF_ROWS = VEC ? prof_stride<R>() : R;     // profile words per (symbol, lane)
  const PassParams& p;
  CompareSelect<R, C, WIDE> csel;
  ProfileSelect<R, C, VEC> psel;
  LaneState<R, C> st;
  int L, g, lane;
  int strip = 0;
  const uint32_t* bnd_in = nullptr;   // boundary row above this strip (indexed by column), null for strip 0
  uint32_t* bnd_out = nullptr;        // boundary row below this strip, null for the last strip
  uint32_t chunk_cur = 0, chunk_next = 0;
  volatile const uint32_t* wait_on = nullptr;   // progress counter of the strip above (pipelined strips)
  mutable uint32_t seen_progress = 0;           // last value read from *wait_on
  uint32_t* publish_to = nullptr;               // this strip's progress counter
  __device__ __forceinline__ Wavefront(const PassParams& p_) : p(p_) {}

  __device__ __forceinline__ size_t blk_index(const PairDesc& pd, int b) const { return (size_t)strip * pd.nblk + b; }
  __device__ __forceinline__ size_t ck_index(const PairDesc& pd, int b) const { return ((size_t)strip * pd.nblk + b) * state_words<R, C, AM>() * L; }

  // select the strip and load its rows (registers or shared-memory profile)
  __device__ __forceinline__ void prepare(const PairDesc& pd, int s, uint32_t* prof_warp) {
    strip = s;
    bnd_in = s > 0 ? p.bnd + pd.bnd_off + (size_t)(s - 1) * (pd.n + 1) : nullptr;
    bnd_out = (s + 1 < (int)pd.nstrips) ? p.bnd + pd.bnd_off + (size_t)s * (pd.n + 1) : nullptr;
    PairDesc sp = pd;
    sp.q_off = pd.q_off + (uint32_t)s * (uint32_t)(L * R);
    if (PROFILE) {
      build_profile<R, WIDE, VEC>(prof_warp, p, sp, g, lane);     // every lane fills (and later reads) only its own column
      psel.prof = prof_warp + (VEC ? lane * prof_stride<R>() : lane);
    } else {
      load_compare_rows<R, C>(csel, p, sp, g);
    }
  }
  // Pass 2 may restart from a LOCAL checkpoint (the lane state a previous replay of the same strip saved into the
  // warp's scratch, layout word w * 32 + lane) instead of a pass-1 checkpoint; per-lane pointer, null = pass 1.
  const uint32_t* restore_from = nullptr;
  __device__ __forceinline__ void restore(const PairDesc& pd, int t0) {
    if (t0 == 0) init_state<R, C>(st, p.sc);
    else if (restore_from) load_state<R, C, AM>(st, p.sc, restore_from, 32, lane);
    else load_state<R, C, AM>(st, p.sc, p.ckpt + pd.ck_off + ck_index(pd, (t0 >> p.logB) - 1), L, g);
  }
  template <bool MASKED>
  __device__ __forceinline__ void load_symbols_m(const PairDesc& pd, int t, uint32_t (&y)[C]) const {
#pragma unroll
    for (int c = 0; c < C; ++c) y[c] = load_y<PROFILE, MASKED>(p, pd, col_of<C>(t, g, c));
  }
  // 32 boundary columns starting at column 32*k + 1, one per lane
  __device__ __forceinline__ uint32_t load_chunk(const PairDesc& pd, int k) const {
    const int j = 32 * k + 1 + lane;
    if (wait_on && bnd_in) {
      // pipelined strips: wait until the strip above has published the last column of this chunk.  The last
      // value read from the progress counter is remembered, so most chunks need no poll at all.
      const uint32_t need = (uint32_t)min((int)pd.n, 32 * k + 32);
      if (32 * k + 1 <= (int)pd.n && seen_progress < need) {
        unsigned spins = 0;
        while ((seen_progress = *wait_on) < need) {
          __nanosleep(400);
          if (++spins > (1u << 24)) { if (p.abort_flag) *p.abort_flag = 1u; break; }
        }
        __threadfence();
      }
    }
    return (bnd_in && j <= (int)pd.n) ? ((volatile const uint32_t*)bnd_in)[j] : p.sc.negG2;
  }
  // ---- software-pipelined stepping -------------------------------------------------------------------------
  // Symbols are loaded two steps ahead; with the profile select the scores of step t+1 are fetched from shared
  // memory (into ra / rb) while step t computes, so no step waits on an LDS or an LDG.
  uint32_t py0[C], py1[C];
  RegSelect<R, C> ra, rb;

  template <bool BND, class Sel, class Hook>
  __device__ __forceinline__ void step_sel(const PairDesc& pd, int t, const Sel& sel, uint32_t& bmax, Hook&& hook) {
    uint32_t upv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      upv[c] = __shfl_up_sync(0xffffffffu, st.bot[c], 1, L);
      if (BND) {                                          // L == 32: lane 0 works on columns C*(t-1)+1 .. C*t
        const int base = C * (t - 1);                     // 0-based column of this step's first column
        if (c == 0 && (base & 31) == 0) { chunk_cur = chunk_next; chunk_next = load_chunk(pd, (base >> 5) + 1); }
        const uint32_t north = __shfl_sync(0xffffffffu, chunk_cur, (base & 31) + c);
        if (g == 0) upv[c] = north;
      } else {
        if (g == 0) upv[c] = p.sc.negG2;                  // row 0 of H is zero: E = -G
      }
    }
    step<R, C, AM>(st, sel, p.sc, upv, bmax, [&](int k, int c, uint32_t e_new) { hook(k, c, t, col_of<C>(t, g, c), e_new); });
    if (BND) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int j = col_of<C>(t, g, c);
        if (bnd_out && g == L - 1 && j >= 1 && j <= (int)pd.n) {
          SWB_CHECK(p.check, pd.bnd_off + (unsigned long long)strip * (pd.n + 1) + j < p.bnd_words, 6);
          bnd_out[j] = st.bot[c];
          if (publish_to && ((j & 511) == 0 || j == (int)pd.n)) { __threadfence(); *(volatile uint32_t*)publish_to = (uint32_t)j; }
        }
      }
    }
  }
  // make the pipeline ready to execute step t
  __device__ __forceinline__ void prime(const PairDesc& pd, int t) {
    if (PROFILE) {
      load_symbols_m<true>(pd, t, py0);
#pragma unroll
      for (int c = 0; c < C; ++c) psel.set_column(c, py0[c]);
      ra.fetch(psel);
      load_symbols_m<true>(pd, t + 1, py0);               // py0 = symbols of step t+1
    } else {
      load_symbols_m<true>(pd, t, py0);                   // py0 = symbols of step t
      load_symbols_m<true>(pd, t + 1, py1);               // py1 = symbols of step t+1
    }
  }
  // steps t and t+1; hook(k, c, step, j, E_new)
  template <bool MASKED, bool BND, class Hook>
  __device__ __forceinline__ void two_steps(const PairDesc& pd, int t, uint32_t& bmax, Hook&& hook) {
    if (PROFILE) {
      load_symbols_m<MASKED>(pd, t + 2, py1);
#pragma unroll
      for (int c = 0; c < C; ++c) psel.set_column(c, py0[c]);
      rb.fetch(psel);                                     // scores of step t+1
      step_sel<BND>(pd, t, ra, bmax, hook);
      load_symbols_m<MASKED>(pd, t + 3, py0);
#pragma unroll
      for (int c = 0; c < C; ++c) psel.set_column(c, py1[c]);
      ra.fetch(psel);                                     // scores of step t+2
      step_sel<BND>(pd, t + 1, rb, bmax, hook);
    } else {
      uint32_t y2[C];
      load_symbols_m<MASKED>(pd, t + 2, y2);
#pragma unroll
      for (int c = 0; c < C; ++c) csel.set_column(c, py0[c]);
      step_sel<BND>(pd, t, csel, bmax, hook);
      load_symbols_m<MASKED>(pd, t + 3, py0);
#pragma unroll
      for (int c = 0; c < C; ++c) csel.set_column(c, py1[c]);
      step_sel<BND>(pd, t + 1, csel, bmax, hook);
#pragma unroll
      for (int c = 0; c < C; ++c) { const uint32_t tmp = py0[c]; py0[c] = y2[c]; py1[c] = tmp; }
    }
  }
  template <bool BND>
  __device__ __forceinline__ void begin(const PairDesc& pd, int t0) {
    restore(pd, t0);
    if (BND) chunk_next = load_chunk(pd, (C * t0) >> 5);  // B >= 32, so C * t0 is a multiple of 32
    prime(pd, t0 + 1);
  }
  // Pass-2 replay: a plain one-step loop (scores straight from the select, symbols one step ahead).  Pass 2
  // is short and has several replay sites per kernel; the compact body keeps them inside the instruction
  // cache, which matters more there than the software pipelining of the score pass.
  template <bool BND, class Hook, class Post>
  __device__ __forceinline__ void replay_impl(const PairDesc& pd, int t0, int t1, int nsteps, Hook&& hook, Post&& post) {
    restore(pd, t0);
    if (BND) chunk_next = load_chunk(pd, (C * t0) >> 5);  // B >= 32, so C * t0 is a multiple of 32
    // symbols are loaded DEPTH steps ahead (a register queue shifted once per step): with few resident warps a step
    // is shorter than the latency of a global load, and a two-step look-ahead left every step waiting for its symbol
    constexpr int DEPTH = C >= 4 ? 3 : 6;
    uint32_t yq[DEPTH][C];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) load_symbols_m<true>(pd, t0 + 1 + d, yq[d]);
    for (int s = 1; s <= nsteps; ++s) {
      const int t = t0 + s;
      uint32_t ycur[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        ycur[c] = yq[0][c];
#pragma unroll
        for (int d = 0; d + 1 < DEPTH; ++d) yq[d][c] = yq[d + 1][c];
      }
      load_symbols_m<true>(pd, t + DEPTH, yq[DEPTH - 1]);
      const bool on = t <= t1;
      auto h = [&](int k, int c, int tt, int j, uint32_t e_new) { if (on) hook(k, c, tt, j, e_new); };
      uint32_t smax = NEG_INF2;                            // maximum over this step's cells
      if (PROFILE) {
#pragma unroll
        for (int c = 0; c < C; ++c) psel.set_column(c, ycur[c]);
        step_sel<BND>(pd, t, psel, smax, h);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) csel.set_column(c, ycur[c]);
        step_sel<BND>(pd, t, csel, smax, h);
      }
      post(t, on, smax);
    }
  }
  // replay never writes boundary rows or progress counters again, it only reads boundary rows.
  // hook(k, c, t, j, E_new) fires for every cell of steps <= t1, post(t, on, step_max) after every step.
  template <class Hook, class Post>
  __device__ __forceinline__ void replay(const PairDesc& pd, bool multi, int t0, int t1, int nsteps, Hook&& hook, Post&& post) {
    bnd_out = nullptr; wait_on = nullptr; publish_to = nullptr;
    if (multi) replay_impl<true>(pd, t0, t1, nsteps, hook, post);
    else replay_impl<false>(pd, t0, t1, nsteps, hook, post);
  }
  template <class Hook>
  __device__ __forceinline__ void replay(const PairDesc& pd, bool multi, int t0, int t1, int nsteps, Hook&& hook) {
    replay(pd, multi, t0, t1, nsteps, hook, [](int, bool, uint32_t) {});
  }
};

// ======================================================================================================
// Pass 1: score pass.  One group of L lanes per pair, 32/L pairs per warp (or one warp per strip).
// Blocks whose columns (plus the two-step look-ahead) are inside [1, n] for every lane skip the range tests.
// ======================================================================================================
template <int R, int C, int AM, bool PROFILE, bool BND, class WF>
__device__ __forceinline__ void score_pass(WF& wf, const PassParams& p, const PairDesc& pd,
                                           int steps, int n_min, bool live) {
  const int L = wf.L, g = wf.g;
  uint32_t* blk = p.blkmax + pd.blk_off;
  uint32_t* ck = p.ckpt + pd.ck_off;
  wf.template begin<BND>(pd, 0);
  uint32_t bmax = NEG_INF2;
  auto nohook = [](int, int, int, int, uint32_t) {};
  const int nb = steps >> p.logB;
  for (int b = 0; b < nb; ++b) {
    const int t0 = b << p.logB;
    const bool interior = (t0 + 1 >= L) && ((t0 + p.B + 2) * C <= n_min);
    if (interior) { for (int t = t0 + 1; t <= t0 + p.B; t += 2) wf.template two_steps<false, BND>(pd, t, bmax, nohook); }
    else { for (int t = t0 + 1; t <= t0 + p.B; t += 2) wf.template two_steps<true, BND>(pd, t, bmax, nohook); }
    const uint32_t gm = group_max_s16x2<AM == AM_WIDE>(bmax, L);
    if (live && b < (int)pd.nblk) {
      SWB_CHECK(p.check, pd.blk_off + wf.blk_index(pd, b) < p.blk_words, 2);
      SWB_CHECK(p.check, (pd.ck_off + wf.ck_index(pd, b) + (unsigned long long)state_words<R, C, AM>() * L <= p.ck_words), 3);
      if (g == 0) blk[wf.blk_index(pd, b)] = gm;
      save_state<R, C, AM>(wf.st, p.sc, ck + wf.ck_index(pd, b), L, g);
    }
    bmax = NEG_INF2;
  }
}

// score_kernel must keep 4 thread blocks of 128 threads resident per SM (16 warps): at most 128 registers per thread up to
// 19 rows per lane.  ptxas is left free to pick its allocation (forcing the bound with __launch_bounds__(128, 4) or
// __maxnreg__(128) costs 5-18 instructions in the hot loop and 2 % of its scheduled cycles: 206 / 219 vs 201 instructions
// per 38 cell pairs, measured 8.44 vs 8.63 TCUPS on one box); instead the register count of
// the built kernels is CHECKED by tools/sass_counts.py --check (a CPU test): round 2 once let an unrelated edit take it
// to 138 registers, 3 blocks per SM and 84 % instead of 88 % ALU-pipe busy without anything failing.
// The batched kernel exists twice: score_kernel for classes whose pairs all fit ONE strip (every read-mapping batch: the
// hot kernel of the C3 benchmark) and score_strips_kernel for classes with row strips run top to bottom by one warp.
// ptxas allocates registers and schedules per kernel, so keeping the strip loop out of score_kernel keeps its hot loop
// independent of edits to the strip path (tools/sass_counts.py --check pins its instruction count).
template <int R, int C, int AM, bool PROFILE, bool STRIPS>
__device__ __forceinline__ void score_batched(const PassParams& p, uint32_t* smem_prof) {
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
  const int L = p.L;
  const int g = lane & (L - 1);
  const int groups_per_warp = 32 >> p.logL;
  uint32_t* prof_warp = smem_prof + (size_t)warp_in_cta * p.KP * prof_stride<R>() * 32;
  int pair = gwarp * groups_per_warp + (lane >> p.logL);
  const bool live = pair < p.npairs;
  if (!live) pair = p.npairs - 1;        // keep the lane in the shuffles; it stores nothing
  const PairDesc pd = p.pairs[pair];

  Wavefront<R, C, AM, PROFILE, true> wf(p);
  wf.L = L; wf.g = g; wf.lane = lane;

  // run whole blocks so that every lane flushes together; lanes past their range see sentinel columns
  const int steps = warp_max_i32((int)pd.nblk << p.logB);
  int n_min = (int)pd.n;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_min = min(n_min, __shfl_xor_sync(0xffffffffu, n_min, o));
  if (STRIPS) {
    const int nstrips = warp_max_i32((int)pd.nstrips);    // > 1 only with L == 32 (one pair per warp)
    for (int s = 0; s < nstrips; ++s) {
      wf.prepare(pd, s, prof_warp);
      if (!live) wf.bnd_out = nullptr;   // a padding warp shadows the last pair: it must not write that pair's boundary rows again
      score_pass<R, C, AM, PROFILE, true>(wf, p, pd, steps, n_min, live);
      __syncwarp();                       // boundary row of strip s is complete before strip s+1 reads it
      __threadfence_block();
    }
  } else {
    wf.prepare(pd, 0, prof_warp);
    score_pass<R, C, AM, PROFILE, false>(wf, p, pd, steps, n_min, live);
  }
}

template <int R, int C, int AM, bool PROFILE>
__global__ void __launch_bounds__(128, 1) score_kernel(const PassParams p) {
  extern __shared__ uint32_t smem_prof[];
  score_batched<R, C, AM, PROFILE, false>(p, smem_prof);
}

template <int R, int C, int AM, bool PROFILE>
__global__ void __launch_bounds__(128) score_strips_kernel(const PassParams p) {
  extern __shared__ uint32_t smem_prof[];
  score_batched<R, C, AM, PROFILE, true>(p, smem_prof);
}

// ======================================================================================================
// Pipelined strips (few long pairs): the strips of a pair run concurrently, each some columns behind the strip above.
// A strip reads the boundary row of the strip above as soon as that strip has published it (progress
// counters), 32 columns per coalesced load.  The work is cut into tasks — (pair, strip) unit x a few checkpoint
// blocks — that persistent warps, one per SM sub-partition, take from an atomic ticket (score_units_kernel below).
//
// This path has a kernel and a stepping routine of its own (UnitsWavefront adds to Wavefront, it changes
// nothing in it): ptxas allocates registers and schedules per kernel, and the batched kernel above is
// sensitive to it (A/B on one box: +-3-5 % on the hot loop from unrelated edits to shared code).  Differences
// to Wavefront::step_sel<true>: the boundary row is stored by a precomputed writer lane without range tests in
// interior blocks, and progress is published once per few blocks instead of being tested per column.
// ======================================================================================================
template <int R, int C, int AM, bool PROFILE>
struct UnitsWavefront : Wavefront<R, C, AM, PROFILE, true> {
  using Base = Wavefront<R, C, AM, PROFILE, true>;
  bool writer = false;     // lane 31 of a strip that has a strip below it
  uint32_t chunk_next2 = 0;   // boundary chunks are fetched TWO chunks (64 columns) ahead of their use
  __device__ __forceinline__ UnitsWavefront(const PassParams& p_) : Base(p_) {}

  template <bool MASKED, class Sel>
  __device__ __forceinline__ void step_units(const PairDesc& pd, int t, const Sel& sel, uint32_t& bmax) {
    uint32_t upv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      upv[c] = __shfl_up_sync(0xffffffffu, this->st.bot[c], 1);
      const int base = C * (t - 1);                       // 0-based column of lane 0's first column in this step
      if (c == 0 && (base & 31) == 0) { this->chunk_cur = this->chunk_next; this->chunk_next = chunk_next2; chunk_next2 = this->load_chunk(pd, (base >> 5) + 2); }
      const uint32_t north = __shfl_sync(0xffffffffu, this->chunk_cur, (base & 31) + c);
      if (this->lane == 0) upv[c] = north;
    }
    step<R, C, AM>(this->st, sel, this->p.sc, upv, bmax, [](int, int, uint32_t) {});
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = col_of<C>(t, this->lane, c);
      if (writer && (!MASKED || (j >= 1 && j <= (int)pd.n))) {
        SWB_CHECK(this->p.check, j >= 1 && j <= (int)pd.n && pd.bnd_off + (unsigned long long)this->strip * (pd.n + 1) + j < this->p.bnd_words, 7);
        this->bnd_out[j] = this->st.bot[c];
      }
    }
  }
  // steps t and t+1, software-pipelined like Wavefront::two_steps
  template <bool MASKED>
  __device__ __forceinline__ void two_steps_units(const PairDesc& pd, int t, uint32_t& bmax) {
    if (PROFILE) {
      this->template load_symbols_m<MASKED>(pd, t + 2, this->py1);
#pragma unroll
      for (int c = 0; c < C; ++c) this->psel.set_column(c, this->py0[c]);
      this->rb.fetch(this->psel);
      step_units<MASKED>(pd, t, this->ra, bmax);
      this->template load_symbols_m<MASKED>(pd, t + 3, this->py0);
#pragma unroll
      for (int c = 0; c < C; ++c) this->psel.set_column(c, this->py1[c]);
      this->ra.fetch(this->psel);
      step_units<MASKED>(pd, t + 1, this->rb, bmax);
    } else {
      uint32_t y2[C];
      this->template load_symbols_m<MASKED>(pd, t + 2, y2);
#pragma unroll
      for (int c = 0; c < C; ++c) this->csel.set_column(c, this->py0[c]);
      step_units<MASKED>(pd, t, this->csel, bmax);
      this->template load_symbols_m<MASKED>(pd, t + 3, this->py0);
#pragma unroll
      for (int c = 0; c < C; ++c) this->csel.set_column(c, this->py1[c]);
      step_units<MASKED>(pd, t + 1, this->csel, bmax);
#pragma unroll
      for (int c = 0; c < C; ++c) { const uint32_t tmp = this->py0[c]; this->py0[c] = y2[c]; this->py1[c] = tmp; }
    }
  }
};

template <int R, int C, int AM, bool PROFILE>
__global__ void __launch_bounds__(32, 1) score_units_kernel(const PassParams p) {
  extern __shared__ uint32_t smem_prof[];
  const int lane = threadIdx.x & 31;
  uint32_t* prof_warp = smem_prof;                 // one warp per thread block
  UnitsWavefront<R, C, AM, PROFILE> wf(p);
  wf.L = 32; wf.g = lane; wf.lane = lane;
  int cur_unit = -1;
  // Persistent warps, one per SM sub-partition, work through TASKS = (unit, task_blocks checkpoint blocks) handed out by
  // an atomic ticket in round-major order: ticket = round * nunits + unit, task block = round - 2 * strip.  A task reads
  // what (strip, block-1) wrote (checkpoint), the boundary row (strip-1, block) wrote and, in its last few dozen steps,
  // the first columns of (strip-1, block+1): all three belong to earlier ROUNDS.  Tickets are taken in order, so whatever
  // a task could wait for is finished or running (forward progress never depends on the dispatch order of thread
  // blocks), and with no more warps than units per round it is normally finished: nobody polls in the steady state, and
  // a slow SM delays nobody — it simply takes fewer tickets.  (One warp per strip for the whole pass, the round-1
  // design, made every strip of a pair run at the pace of the slowest warp above it: 25 % of all issue slots of the
  // 10 kbp x 51 Mbp run went into polling, profiles/score_units_kernel_r02_c5_ncu.txt.)
  for (;;) {
    uint32_t ticket = 0;
    if (lane == 0) ticket = atomicAdd(p.ticket, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= p.ntickets) break;
    const int unit = (int)(ticket % (uint32_t)p.nunits);
    const int round = (int)(ticket / (uint32_t)p.nunits);
    const uint2 u = p.units[unit];
    const int tb = round - 2 * (int)u.y;                         // task block of this unit in this round
    const PairDesc pd = p.pairs[u.x];
    const int nb = (int)pd.nblk;
    if (tb < 0 || tb * p.task_blocks >= nb) continue;
    if (unit != cur_unit) {
      cur_unit = unit;
      wf.prepare(pd, (int)u.y, prof_warp);
      wf.writer = wf.bnd_out != nullptr && lane == 31;
      wf.wait_on = u.y > 0 ? p.progress + (unit - 1) : nullptr;   // units of a pair are consecutive, strips ascending
      wf.seen_progress = 0;
    }
    volatile uint32_t* const done = p.blocks_done + unit;
    if (tb > 0) {                                                 // this strip's previous task (checkpoint + boundary row position)
      unsigned spins = 0;
      while (*done < (uint32_t)tb) {
        __nanosleep(200);
        if (++spins > (1u << 24)) { if (p.abort_flag) *p.abort_flag = 1u; break; }
      }
      __threadfence();
    }
    uint32_t* const publish_to = (u.y + 1 < pd.nstrips) ? p.progress + unit : nullptr;   // wf.publish_to stays null: no per-column publishing
    uint32_t* blk = p.blkmax + pd.blk_off;
    uint32_t* ck = p.ckpt + pd.ck_off;
    const int n = (int)pd.n;
    const int b0 = tb * p.task_blocks, b1 = min(nb, b0 + p.task_blocks);
    wf.template begin<true>(pd, b0 << p.logB);
    wf.chunk_next2 = wf.load_chunk(pd, ((C * (b0 << p.logB)) >> 5) + 1);
    uint32_t bmax = NEG_INF2;
    for (int b = b0; b < b1; ++b) {
      const int t0 = b << p.logB;
      const bool interior = (t0 + 1 >= 32) && ((t0 + p.B + 2) * C <= n);
      if (interior) { for (int t = t0 + 1; t <= t0 + p.B; t += 2) wf.template two_steps_units<false>(pd, t, bmax); }
      else { for (int t = t0 + 1; t <= t0 + p.B; t += 2) wf.template two_steps_units<true>(pd, t, bmax); }
      const uint32_t gm = group_max_s16x2<AM == AM_WIDE>(bmax, 32);
      SWB_CHECK(p.check, pd.blk_off + wf.blk_index(pd, b) < p.blk_words, 4);
      SWB_CHECK(p.check, (pd.ck_off + wf.ck_index(pd, b) + (unsigned long long)state_words<R, C, AM>() * 32 <= p.ck_words), 5);
      if (lane == 0) blk[wf.blk_index(pd, b)] = gm;
      save_state<R, C, AM>(wf.st, p.sc, ck + wf.ck_index(pd, b), 32, lane);
      bmax = NEG_INF2;
      if (b == b1 - 1) {
        // end of the task: every lane fences its own stores (boundary row: lane 31, checkpoints: all lanes), then one lane
        // publishes the boundary columns finished so far (for the strip below) and the task count (for this strip's next task)
        __threadfence();
        __syncwarp();
        if (lane == 31) {
          const int jd = min(n, col_of<C>(t0 + p.B, 31, C - 1));    // last column lane 31 has finished
          if (publish_to && jd >= 1) *(volatile uint32_t*)publish_to = (uint32_t)jd;
          *done = (uint32_t)(tb + 1);
        }
      }
    }
  }
}

// ======================================================================================================
// Pass 2: locate the reference's arg-max cell and trace back.  One group of L lanes per task; the 32/L
// groups of a warp run in LOCKSTEP (same instruction stream, per-group predicates), so a round costs the
// longest group's work, not the sum.
// ======================================================================================================
struct TaskDesc {
  uint32_t pair;      // index into pairs
  uint32_t half;      // 0 = low half (A), 1 = high half (B)
  uint32_t out;       // output slot (read index)
  uint32_t pos_add;   // added to pos (left edge of the chunk, plocalaligner.cpp:137)
};

struct TraceParams {
  PassParams pp;
  const TaskDesc* tasks;      // indexed through task_list when non-null
  const uint32_t* task_list;
  int ntasks;
  uint32_t* next_task;        // zeroed by the host before the launch: first task nobody has taken yet
  int mode;                   // MODE_SAT_U8: skewed raw-order tie-break; MODE_EXACT: column-major
  int max_pos;                // largest positive substitution score: a cell of score V needs row >= ceil(V / max_pos)
  uint32_t* scratch;          // per warp: nlc local checkpoints (lane states), word (slot * state_words + w) * 32 + lane
  int Wc, logWc;              // ring depth in steps (power of two) = period of the local checkpoints
  int NB;                     // band: lanes per group kept in the ring (the lane of the current row and NB-1 above it)
  int nlc;                    // local checkpoint slots per warp (power of two)
  int ring_off;               // word offset of the rings in dynamic shared memory (after the profiles)
  unsigned long long scratch_words;   // size of scratch (for the SWB_CHECKED bounds checks)
  int32_t* out_score;
  uint32_t* out_pos;
  uint32_t* out_end;          // 2 per task: index_x, index_y of the arg-max
  uint8_t* out_cx;
  uint8_t* out_cy;
  uint32_t* out_len;
  uint32_t cons_cap;          // bytes per task in out_cx / out_cy
  uint32_t* out_flags;        // bit0: consensus overflowed cons_cap
  int want_consensus;
  int dbg_flags;                  // diagnostics: bit1 skip the walk (results are then wrong)
  unsigned long long* counters;   // diagnostics (SWB_DEBUG): [0] scan replays, [1] sessions, [2] scan lockstep rounds, [3] session rounds, [4] session steps (warp max)
};

__device__ __forceinline__ int half_of(uint32_t v, uint32_t half) { return (int)(int16_t)(half ? (v >> 16) : (v & 0xFFFFu)); }

// Raw key of the skewed storage, _trueindex2rawindex similaritymatrix.cpp:353-364 with the constructor's
// role swap (:274-289): ti = column j, tj = row i, len_x = n+1, len_y = m+1.  Key = (rj << 32) | ri.
__device__ __forceinline__ uint64_t skew_key(int i, int j, int m, int n) {
  const int len_x = n + 1, len_y = m + 1;
  const int nrows = min(len_x, len_y), ncols = max(len_x, len_y);
  const int ti = j, tj = i;
  int ri, rj;
  if (ti + tj < nrows - 1) { ri = ti; rj = ti + tj; }
  else if (ti + tj > ncols - 1) { ri = ti - ncols + len_y; rj = ti + tj - (ncols - 1) - 1; }
  else { ri = (len_x <= len_y) ? ti : len_y - 1 - tj; rj = ti + tj; }
  return ((uint64_t)(uint32_t)rj << 32) | (uint32_t)ri;
}
__device__ __forceinline__ uint64_t colmajor_key(int i, int j) { return ((uint64_t)(uint32_t)j << 32) | (uint32_t)i; }

__device__ __forceinline__ uint64_t group_min_u64(uint64_t v, int L) {
  for (int o = L >> 1; o > 0; o >>= 1) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, o), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), o);
    uint64_t w = ((uint64_t)hi << 32) | lo;
    v = w < v ? w : v;
  }
  return v;
}
__device__ __forceinline__ int group_max_i32(int v, int L) {
  for (int o = L >> 1; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// QS = query-stationary frame (sw_qs.cuh): the kernel's rows are the reference's y (the shared query, qs_m rows)
// and its columns the reference's x (a database sequence per half); WF is then QsWavefront and the arg-max order
// and the traceback priority are expressed in that transposed frame.
#ifndef SWB_TRACE_MINBLOCKS
#define SWB_TRACE_MINBLOCKS 4
#endif
template <int R, int C, int AM, bool PROFILE, bool QS, class WF>
__device__ __forceinline__ void trace_body(const TraceParams& tp, uint32_t* smem_prof, int qs_m) {
  constexpr bool SAT = AM == AM_SAT, WIDE = AM == AM_WIDE;
  const PassParams& p = tp.pp;
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  const int L = p.L;
  const int g = lane & (L - 1);
  const int grp_in_warp = lane >> p.logL;
  const int groups_per_warp = 32 >> p.logL;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
  const int wmask = tp.Wc - 1;
  const int G = p.sc.G;
  const int S = L * R;                 // rows per strip
  const uint32_t gshift = (uint32_t)(grp_in_warp * L);
  const uint32_t gbits = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
  uint32_t* prof_warp = QS ? smem_prof : smem_prof + (size_t)warp_in_cta * p.KP * R * 32;   // QS: one profile per CTA

  WF wf(p);
  wf.L = L; wf.g = g; wf.lane = lane;

  // A warp takes the next 32/L tasks from a counter when it has finished its last ones: alignments differ in length
  // (sessions, walk) and not every thread block of the launch is resident from the start, so a fixed stride over the
  // tasks left the SMs idle for 13-22 % of the kernel at its end (ncu: sm__cycles_elapsed.max vs smsp__cycles_active.avg).
  for (;;) {
    int base = 0;
    if (lane == 0) base = (int)atomicAdd(tp.next_task, (uint32_t)groups_per_warp);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= tp.ntasks) break;                                 // warp-uniform
    const int ti = base + grp_in_warp;
    bool active = ti < tp.ntasks;
    const TaskDesc td = tp.tasks[tp.task_list ? tp.task_list[active ? ti : base] : (active ? ti : base)];
    const PairDesc pd = p.pairs[td.pair];
    const int m = QS ? qs_m : (int)(td.half ? pd.mB : pd.mA);     // rows of the kernel frame
    const int n = QS ? (int)(td.half ? pd.mB : pd.mA) : (int)pd.n; // columns of the kernel frame
    const int nblk = (int)pd.nblk;
    const int nunits = nblk * (int)pd.nstrips;       // (strip, block) units
    const bool multi = warp_max_i32((int)pd.nstrips) > 1;   // > 1 only with L == 32: one group per warp
    const uint32_t half = td.half;
    // characters along the kernel's rows / columns (QS: rows = the query in reads_raw, columns = a database sequence)
    const uint8_t* xraw = QS ? p.reads_raw : p.reads_raw + (td.half ? pd.xB : pd.xA);
    const uint8_t* yraw = QS ? p.ref_raw + (td.half ? pd.xB : pd.xA) : p.ref_raw + pd.y_off;
    long long tk0 = clock64();
    int cur_strip = 0;
    wf.prepare(pd, 0, prof_warp);
    if (tp.counters && lane == 0) { const long long tk1 = clock64(); atomicAdd(tp.counters + 8, (unsigned long long)(tk1 - tk0)); tk0 = tk1; }

    // ---- 1. maximum over the block maxima (E-space) ------------------------------------------------------
    const uint32_t* blk = p.blkmax + pd.blk_off;
    int vmax = WIDE ? (int)NEG_INF2 : -32768;
    for (int w = g; w < nunits; w += L) vmax = max(vmax, lane_val<WIDE>(blk[w], half));
    vmax = group_max_i32(vmax, L);
    const int score = vmax + G;
    if (tp.counters && lane == 0) { const long long tk1 = clock64(); atomicAdd(tp.counters + 9, (unsigned long long)(tk1 - tk0)); tk0 = tk1; }
    if (active && (m == 0 || score <= 0)) {
      // all-zero matrix: the reference reads H(-1,-1) (undefined); we return score 0, pos 0, "".
      if (g == 0) {
        tp.out_score[td.out] = 0; tp.out_pos[td.out] = 0; tp.out_len[td.out] = 0;
        tp.out_end[2 * td.out] = 0; tp.out_end[2 * td.out + 1] = 0; tp.out_flags[td.out] = 0;
      }
      active = false;
    }

    // ---- 2. arg-max with the reference's tie-break --------------------------------------------------------
    // Candidate units: some lane's block maximum equals vmax.  Their cells are recomputed and keyed; the
    // smallest key wins (skewed raw order for SAT_U8, column-major for EXACT).  Units whose smallest possible
    // key already exceeds the current winner are skipped; units that may hold wrapped (lower-triangle) cells
    // sort first in the skewed order and are visited first (phase 0), the others in phase 1.
    const int ncols_raw = max(n + 1, m + 1);
    const int row_min = (score + tp.max_pos - 1) / max(tp.max_pos, 1);   // a cell worth `score` cannot sit above this row
    uint64_t best = ~0ull;
    int cursor = active ? 0 : 2 * nunits;            // [0, nunits): phase 0, [nunits, 2*nunits): phase 1
    const uint32_t vmax2 = WIDE ? (uint32_t)vmax : (uint32_t)(uint16_t)(int16_t)vmax * 0x00010001u;
    const uint32_t hmask = WIDE ? 0xFFFFFFFFu : (half ? 0xFFFF0000u : 0x0000FFFFu);
    while (true) {
      int myu = -1;
      while (true) {                                   // warp-uniform loop; the body is predicated per group
        const bool searching = cursor < 2 * nunits && myu < 0;
        if (!__any_sync(0xffffffffu, searching)) break;
        const int pos = cursor + g;
        bool cand = false;
        int u = -1;
        if (searching && pos < 2 * nunits) {
          const int phase = pos >= nunits;
          u = pos - phase * nunits;
          const int us = u / nblk, ub = u - us * nblk;
          if (lane_val<WIDE>(blk[u], half) == vmax) {
            const int t0 = ub << p.logB;
            const int jmin = max(1, C * (t0 - (L - 1)) + 1), jmax = min(n, C * (t0 + p.B));
            const int imin = max(us * S + 1, row_min), imax = min(m, (us + 1) * S);
            if (jmin <= jmax && imin <= imax) {
              const bool wraps = tp.mode == MODE_SAT_U8 && (jmax + imax >= ncols_raw);
              uint64_t lb;
              // smallest raw column of the unit: d_min for ordinary cells, d_min - ncols (>= 0) for wrapped ones
              if (tp.mode == MODE_SAT_U8) lb = (uint64_t)(uint32_t)(wraps ? max(0, jmin + imin - ncols_raw) : jmin + imin) << 32;
              else lb = (uint64_t)(uint32_t)(QS ? imin : jmin) << 32;
              cand = (lb <= best) && (wraps == (phase == 0));
            }
          }
        }
        const uint32_t cm = (__ballot_sync(0xffffffffu, cand) >> gshift) & gbits;
        const int q = cm ? __ffs(cm) - 1 : 0;
        const int usel = __shfl_sync(0xffffffffu, u, (int)gshift + q);
        if (searching) {
          if (cm) { myu = usel; cursor += q + 1; } else { cursor += L; }
        }
      }
      if (tp.counters && lane == 0) { const long long tk1 = clock64(); atomicAdd(tp.counters + 10, (unsigned long long)(tk1 - tk0)); tk0 = tk1; }
      if (!__any_sync(0xffffffffu, myu >= 0)) break;
      const bool has = myu >= 0;
      if (tp.counters) { if (has && g == 0) atomicAdd(tp.counters + 0, 1ull); if (lane == 0) atomicAdd(tp.counters + 2, 1ull); }
      const int us = has ? myu / nblk : cur_strip;
      const int t0 = has ? ((myu - us * nblk) << p.logB) : 0;
      if (multi && us != cur_strip) { cur_strip = us; wf.prepare(pd, us, prof_warp); }
      const int row0 = cur_strip * S + g * R + 1;
      uint64_t mine = ~0ull;
      auto consider = [&](int i, int j) {
        if (i <= m && j >= 1 && j <= n) {
          const uint64_t key = QS ? ((tp.mode == MODE_SAT_U8) ? skew_key(j, i, n, m) : colmajor_key(j, i))
                                  : ((tp.mode == MODE_SAT_U8) ? skew_key(i, j, m, n) : colmajor_key(i, j));
          mine = key < mine ? key : mine;
        }
      };
      if (C == 1) {
        // one column per step: the step's cells are still in the lane state afterwards, so a step is examined
        // only when its maximum reaches vmax (no cell of this half can exceed it)
        wf.replay(pd, multi, t0, has ? t0 + p.B : -1, p.B, [](int, int, int, int, uint32_t) {}, [&](int t, bool on, uint32_t smax) {
          if (on && ((smax ^ vmax2) & hmask) == 0) {
            const int j = col_of<C>(t, g, 0);
#pragma unroll
            for (int k = 0; k < R; ++k)
              if (((wf.st.E[k] ^ vmax2) & hmask) == 0) consider(row0 + k, j);   // unrolled: saturated plateaus hit often
          }
        });
      } else {
        wf.replay(pd, multi, t0, has ? t0 + p.B : -1, p.B, [&](int k, int, int, int j, uint32_t e_new) {
          if (((e_new ^ vmax2) & hmask) == 0) consider(row0 + k, j);
        });
      }
      mine = group_min_u64(mine, L);
      best = mine < best ? mine : best;
      if (tp.counters && lane == 0) { const long long tk1 = clock64(); atomicAdd(tp.counters + 11, (unsigned long long)(tk1 - tk0)); tk0 = tk1; }
    }
    int ie = 1, je = 1;
    if (active) {
      if (tp.mode == MODE_SAT_U8) {
        // invert skew_key: _rawindex2trueindex, similaritymatrix.cpp:330-346
        const int rj = (int)(best >> 32), ri = (int)(uint32_t)best;
        const int len_x = (QS ? m : n) + 1, len_y = (QS ? n : m) + 1, nrows = min(len_x, len_y);
        int t_i, t_j;
        if (rj < nrows - 1) {
          if (ri <= rj) { t_i = ri; t_j = rj - ri; } else { t_i = len_x - nrows + ri; t_j = len_y - ri + rj; }
        } else {
          if (len_x <= len_y) { t_i = ri; t_j = rj - ri; } else { t_i = rj - (nrows - 1) + ri; t_j = nrows - 1 - ri; }
        }
        if (QS) { ie = t_i; je = t_j; } else { je = t_i; ie = t_j; }
      } else if (QS) { ie = (int)(best >> 32); je = (int)(uint32_t)best; }
      else { je = (int)(best >> 32); ie = (int)(uint32_t)best; }
      if (g == 0) {
        tp.out_score[td.out] = score;
        tp.out_end[2 * td.out] = (uint32_t)(QS ? je : ie); tp.out_end[2 * td.out + 1] = (uint32_t)(QS ? ie : je);
      }
    }

    // ---- 3. traceback -----------------------------------------------------------------------------------------
    // SWAligner::traceback, smithwaterman.cpp:40-78, literally: compare the three neighbours' VALUES.
    // A SESSION replays the strip that holds row ix up to the step in which (ix, iy) is computed and keeps the last
    // Wc steps of the task's own half in a ring in SHARED memory — one byte per cell in SAT_U8 (H fits a byte), 16
    // bits in EXACT, packed with PRMT, and only for the NB lanes at and above the lane of row ix (a walk moves up and
    // to the left).  The group's lane 0 then walks in shared memory until it needs a cell the ring does not hold
    // (left of the ring, above the band, or in the strip above); the next session starts there.  Every session also
    // saves the lane state every Wc steps into the warp's scratch (local checkpoints), so a follow-up session
    // restarts Wc..2*Wc steps back instead of at a pass-1 checkpoint up to B steps away.
    constexpr int PW = SAT ? (R + 3) / 4 : (WIDE ? R : (R + 1) / 2);   // packed ring words per lane and column
    constexpr int EPW = SAT ? 4 : (WIDE ? 1 : 2);              // ring elements (cells) per word
    constexpr int RP = PW * EPW;                               // element slots per lane: R rounded up to whole words
    constexpr int SW = state_words<R, C, AM>();
    const int NB = tp.NB;
    const int cmask = tp.Wc * C - 1;                           // ring columns - 1 (Wc and C are powers of two)
    const int cstride = NB * PW;                               // words per ring column: the band's rows, lane by lane
    // this group's ring: word ((j - 1) & cmask) * cstride + (lane - band_lo) * PW + w  (indexed by COLUMN, so that the
    // walker's address is a multiply-add of its own coordinates; lane g of a step stores column t - g)
    uint32_t* const ring = smem_prof + tp.ring_off + ((size_t)warp_in_cta * groups_per_warp + grp_in_warp) * ((size_t)(cmask + 1) * cstride);
    uint32_t* const lck = tp.scratch + (size_t)gwarp * tp.nlc * SW * 32;
    const uint32_t sel2 = SAT ? (half ? 0x6262u : 0x4040u) : (half ? 0x7632u : 0x5410u);
    int ix = ie, iy = je;
    uint32_t len = 0, flags = 0, pos = 0;
    uint8_t* cx = tp.out_cx + (size_t)td.out * tp.cons_cap;
    uint8_t* cy = tp.out_cy + (size_t)td.out * tp.cons_cap;
    bool done = !active;
    int lc_lo = 1, lc_hi = 0, lc_strip = -1;                   // valid local checkpoints: multiples of Wc in [lc_lo, lc_hi]
    while (!__all_sync(0xffffffffu, done)) {
      const int ss = (ix - 1) / S;                             // strip of row ix
      const int il = ix - ss * S;                              // row within the strip, 1..S
      const int l_e = (il - 1) / R;
      const int t_hi = done ? 0 : step_of<C>(iy, l_e);
      const int band_lo = max(0, l_e - NB + 1);
      if (multi && !done && ss != cur_strip) { cur_strip = ss; wf.prepare(pd, ss, prof_warp); }
      if (ss != lc_strip) { lc_lo = 1; lc_hi = 0; lc_strip = ss; }
      // restart: the latest checkpoint (pass 1: every B steps; local: every Wc steps) at least Wc steps before t_hi
      const int t_want = max(0, t_hi - tp.Wc);
      int t_start = (t_want >> p.logB) << p.logB;
      const uint32_t* from = nullptr;
      {
        const int t_l = min(lc_hi, (t_want >> tp.logWc) << tp.logWc);
        if (lc_lo <= lc_hi && t_l >= lc_lo && t_l > t_start) { t_start = t_l; from = lck + (size_t)((t_l >> tp.logWc) & (tp.nlc - 1)) * SW * 32; }
      }
      if (done) t_start = 0;
      wf.restore_from = done ? nullptr : from;
      const int nsteps = warp_max_i32(done ? 0 : t_hi - t_start);
      const int t_store = max(t_start + 1, t_hi - tp.Wc + 1);  // oldest step this session leaves in the ring
      const bool in_band = !done && g >= band_lo && g < band_lo + NB;
      uint32_t* const ring_lane = ring + (g - band_lo) * PW;
      if (tp.counters) { if (!done && g == 0) atomicAdd(tp.counters + 1, 1ull); if (lane == 0) { atomicAdd(tp.counters + 3, 1ull); atomicAdd(tp.counters + 4, (unsigned long long)nsteps); } }
      uint32_t colv[C][R];                                     // the step's cells per column (C > 1; with C == 1 they are the lane state)
      auto keep = [&](int k, int c, int, int, uint32_t e_new) { if (C > 1) colv[c][k] = e_new; };
      wf.replay(pd, multi, t_start, t_hi, nsteps, keep, [&](int t, bool on, uint32_t) {
        if (on && in_band && t >= t_store) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const uint32_t* v = C == 1 ? wf.st.E : colv[c];
            uint32_t* dst = ring_lane + ((col_of<C>(t, g, c) - 1) & cmask) * cstride;
            SWB_CHECK(p.check, g - band_lo >= 0 && g - band_lo < NB && (dst - ring) + PW <= (cmask + 1) * cstride, 8);
            if (SAT) {
#pragma unroll
              for (int w = 0; w < PW; ++w) {                   // four rows per word: the low byte of the half (E mod 256)
                const uint32_t a = __byte_perm(v[4 * w], 4 * w + 1 < R ? v[4 * w + 1] : 0u, sel2);
                const uint32_t b = 4 * w + 2 < R ? __byte_perm(v[4 * w + 2], 4 * w + 3 < R ? v[4 * w + 3] : 0u, sel2) : 0u;
                dst[w] = __byte_perm(a, b, 0x5410);
              }
            } else if (WIDE) {
#pragma unroll
              for (int w = 0; w < PW; ++w) dst[w] = v[w];       // wide lanes: the s32 cell itself
            } else {
#pragma unroll
              for (int w = 0; w < PW; ++w)                      // two rows per word: the half's 16 bits
                dst[w] = __byte_perm(v[2 * w], 2 * w + 1 < R ? v[2 * w + 1] : 0u, sel2);
            }
          }
        }
        if (on && (t & wmask) == 0) {
          SWB_CHECK(p.check, (size_t)gwarp * tp.nlc * SW * 32 + (size_t)((t >> tp.logWc) & (tp.nlc - 1)) * SW * 32 + SW * 32 <= tp.scratch_words, 9);
          save_state<R, C, AM>(wf.st, p.sc, lck + (size_t)((t >> tp.logWc) & (tp.nlc - 1)) * SW * 32, 32, lane);
        }
      });
      wf.restore_from = nullptr;
      if (!done) {
        // local checkpoints written by this session: multiples of Wc in (t_start, t_hi]; only the last nlc survive
        const int w_lo = ((t_start >> tp.logWc) + 1) << tp.logWc, w_hi = (t_hi >> tp.logWc) << tp.logWc;
        if (w_lo <= w_hi) {
          const int span = tp.nlc << tp.logWc;
          const int w_lo_eff = max(w_lo, w_hi - span + tp.Wc);
          const bool have = lc_lo <= lc_hi;
          if (!(have && w_lo_eff >= lc_lo && w_hi <= lc_hi)) {       // not a rewrite of checkpoints we already hold
            int nhi = w_hi;
            if (have && lc_lo <= w_hi + tp.Wc && lc_hi > w_hi) nhi = min(lc_hi, w_lo_eff + span - tp.Wc);
            lc_lo = w_lo_eff; lc_hi = nhi;
          }
        }
      }
      __syncwarp();
      if (tp.counters && lane == 0) { const long long tk1 = clock64(); atomicAdd(tp.counters + 12, (unsigned long long)(tk1 - tk0)); tk0 = tk1; }
      if (g == 0 && !done) {
        const int row_lo = ss * S;                           // last row of the strip above (0 for strip 0)
        const uint32_t* above = ss > 0 ? p.bnd + pd.bnd_off + (size_t)(ss - 1) * (n + 1) : nullptr;
        const int i_min = row_lo + band_lo * R + 1;          // first row the ring holds
        // first column the ring holds for row i: its lane gg kept the steps from t_store on, i.e. columns from
        // C * (t_store - gg - 1) + 1 (lower lanes started later, so the row ABOVE the walker is the binding one)
        auto j_min_of = [&](int i) -> int { return C * (t_store - band_lo - (i - i_min) / R - 1) + 1; };
        // H(i, j) from the ring (both inside it): one byte (SAT_U8: E mod 256) or one signed 16-bit (EXACT: E) load
        auto ring_val = [&](int i, int j) -> int {
          const int r = i - i_min;
          const int e = (RP == R) ? r : r + (r / R) * (RP - R);          // rows of a lane are padded to whole words
          const int col = ((j - 1) & cmask) * cstride;
          SWB_CHECK(p.check, r >= 0 && e >= 0 && e < NB * RP && j >= j_min_of(i), 10);   // j may be 0: a stored virtual column (H = 0)
          if (SAT) return (int)((reinterpret_cast<const uint8_t*>(ring)[col * 4 + e] + (uint32_t)G) & 0xFFu);
          if (WIDE) return (int)ring[col + e] + G;
          return (int)reinterpret_cast<const int16_t*>(ring)[col * 2 + e] + G;
        };
        // any cell: the zero border, the boundary row of the strip above (HBM, packed E words), the ring, or -1 when
        // this session does not hold it
        auto cell = [&](int i, int j) -> int {
          if (i <= 0 || j <= 0) return 0;
          if (i == row_lo) return lane_val<WIDE>(__ldcg(above + j), half) + G;
          if (i < i_min || j < j_min_of(i)) return -1;
          return ring_val(i, j);
        };
        const int ix0 = ix, iy0 = iy;
        while (!(tp.dbg_flags & 2)) {
          if (ix <= row_lo) break;                           // walked into the strip above: next session there
          int vd, vu, vl;                                    // H(ix-1, iy-1), H(ix-1, iy), H(ix, iy-1)
          if (ix - 1 >= i_min && iy - 1 >= j_min_of(ix - 1)) { vd = ring_val(ix - 1, iy - 1); vu = ring_val(ix - 1, iy); vl = ring_val(ix, iy - 1); }
          else {
            vd = cell(ix - 1, iy - 1); vu = cell(ix - 1, iy); vl = cell(ix, iy - 1);
            if ((vd | vu | vl) < 0) break;                   // left the ring: next session starts at (ix, iy)
          }
          // the reference's second neighbour is H(ix, iy-1) and its third H(ix-1, iy); in the QS frame (transposed)
          // those are the cell above and the cell to the left
          const int n1 = vd, n2 = QS ? vu : vl, n3 = QS ? vl : vu;
          const bool emit = tp.want_consensus && len < tp.cons_cap;
          if (len >= tp.cons_cap) flags |= 1u;               // consensus truncated; the walk goes on, so pos stays exact
          uint8_t xc = 0, yc = 0;
          SWB_CHECK(p.check, ix >= 1 && iy >= 1 && ix <= m && iy <= n && (!emit || len < tp.cons_cap), 11);
          if (emit) { xc = xraw[ix - 1]; yc = yraw[iy - 1]; }
          const uint8_t rx = QS ? yc : xc, ry = QS ? xc : yc;   // the reference's x / y characters of this cell
          if (n1 == 0 || n2 == 0 || n3 == 0) {
            if (emit) { cx[len] = rx; cy[len] = ry; }
            ++len; pos = (uint32_t)(QS ? ix : iy); done = true; break;
          }
          if (n1 >= n2 && n1 >= n3) { if (emit) { cx[len] = rx; cy[len] = ry; } --ix; --iy; }
          else if (n2 >= n1 && n2 >= n3) { if (emit) { cx[len] = '-'; cy[len] = ry; } if (QS) --ix; else --iy; }
          else { if (emit) { cx[len] = rx; cy[len] = '-'; } if (QS) --iy; else --ix; }
          ++len;
        }
        if (tp.dbg_flags & 2) done = true;
        if (!done && ix == ix0 && iy == iy0) { flags |= 2u; done = true; }   // no progress: never expected (internal error flag)
        if (done) {
          tp.out_pos[td.out] = pos + td.pos_add;
          tp.out_len[td.out] = len;
          tp.out_flags[td.out] = flags;
        }
      }
      __syncwarp();
      if (tp.counters && lane == 0) { const long long tk1 = clock64(); atomicAdd(tp.counters + 13, (unsigned long long)(tk1 - tk0)); tk0 = tk1; }
      // broadcast the walker's state to its group
      ix = __shfl_sync(0xffffffffu, ix, (int)gshift); iy = __shfl_sync(0xffffffffu, iy, (int)gshift);
      done = __shfl_sync(0xffffffffu, (int)done, (int)gshift) != 0;
    }
  }
}

template <int R, int C, int AM, bool PROFILE>
__global__ void __launch_bounds__(128, SWB_TRACE_MINBLOCKS) trace_kernel(const TraceParams tp) {
  extern __shared__ uint32_t smem_prof[];
  trace_body<R, C, AM, PROFILE, false, Wavefront<R, C, AM, PROFILE>>(tp, smem_prof, 0);
}

// ======================================================================================================
// Dense H for one pair (tests / small inputs): the Abstract_Similarity_Matrix::operator()(row, col) surface.
// One warp replays every strip from its initial state and stores H = E + G of half A, row-major (m+1) x (n+1).
// ======================================================================================================
struct DumpParams {
  PassParams pp;
  int32_t* out;     // (m+1) x (n+1), pre-zeroed
  int m, n;
};

template <int R, int C, int AM, bool PROFILE>
__global__ void __launch_bounds__(128) dump_kernel(const DumpParams dp) {
  extern __shared__ uint32_t smem_prof[];
  const PassParams& p = dp.pp;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32 || blockIdx.x > 0) return;
  const int L = p.L;                       // the host stages a single pair with L == 32
  const int g = lane & (L - 1);
  const PairDesc pd = p.pairs[0];
  Wavefront<R, C, AM, PROFILE> wf(p);
  wf.L = L; wf.g = g; wf.lane = lane;
  const int S = L * R;
  const int steps = (int)pd.nblk << p.logB;
  const bool multi = pd.nstrips > 1;
  for (int s = 0; s < (int)pd.nstrips; ++s) {
    wf.prepare(pd, s, smem_prof);
    const int row0 = s * S + g * R + 1;
    wf.replay(pd, multi, 0, steps, steps, [&](int k, int, int, int j, uint32_t e_new) {
      const int i = row0 + k;
      if (i <= dp.m && j >= 1 && j <= dp.n) dp.out[(size_t)i * (dp.n + 1) + j] = lane_val<AM == AM_WIDE>(e_new, 0) + p.sc.G;
    });
    __syncwarp();
  }
}

#ifdef SWB_HELPER_KERNELS
// ======================================================================================================
// Small helper kernels
// ======================================================================================================
// qpairs[q_off + row] = pack(symA(row), symB(row)); rows beyond a half's length hold SENT_X.
__global__ void pack_rows_kernel(const uint8_t* reads_raw, const PairDesc* pairs, int npairs, int max_rows_per_pair, int rows_per_strip, uint32_t* qpairs) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)npairs * max_rows_per_pair;
  if (idx >= total) return;
  const int pair = (int)(idx / max_rows_per_pair), row = (int)(idx % max_rows_per_pair);
  const PairDesc pd = pairs[pair];
  if (row >= (int)pd.nstrips * rows_per_strip) return;
  const uint32_t a = row < (int)pd.mA ? (SYM_BASE | reads_raw[pd.xA + row]) : SENT_X;
  const uint32_t b = row < (int)pd.mB ? (SYM_BASE | reads_raw[pd.xB + row]) : SENT_X;
  qpairs[pd.q_off + row] = a | (b << 16);
}

// Per-task maximum (E-space + G = score) from the block maxima; used by the chunked path to pick the
// best piece (plocalaligner.cpp:122-129) before any traceback.
__global__ void task_max_kernel(const PairDesc* pairs, const TaskDesc* tasks, int ntasks, const uint32_t* blkmax, int G, int wide, int32_t* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntasks) return;
  const TaskDesc td = tasks[t];
  const PairDesc pd = pairs[td.pair];
  const uint32_t* blk = blkmax + pd.blk_off;
  int v = wide ? (int)NEG_INF2 : -32768;
  for (uint32_t w = 0; w < pd.nblk * pd.nstrips; ++w) v = max(v, wide ? (int)blk[w] : half_of(blk[w], td.half));
  const int m = td.half ? pd.mB : pd.mA;
  out[t] = (m == 0) ? -1 : max(v + G, 0);
}

// Lowest-index piece with the strictly greatest maximum (serial semantic of plocalaligner.cpp:122-129).
__global__ void select_piece_kernel(const int32_t* task_max, int nreads, int npiece, uint32_t* winner_task) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nreads) return;
  int best = -1, bp = 0;      // max_score_l = -1.0, max_score_piece = 0  (:107-108)
  for (int pc = 0; pc < npiece; ++pc) {
    const int v = task_max[(size_t)r * npiece + pc];
    if (v > best) { best = v; bp = pc; }
  }
  winner_task[r] = (uint32_t)(r * npiece + bp);
}

#endif  // SWB_HELPER_KERNELS

}  // namespace swb
