// sw_qs.cuh — query-stationary kernels for database search (BASELINE config 4, mpi_sw_solve_uniprot.cpp:95-138).
//
// The reference aligns every database protein x (rows of its matrix) against ONE query y (columns).  With a
// 20+-letter alphabet the per-warp profile of the batched kernels (one copy per warp, rows = x) caps the rows
// per lane at 4 and cuts proteins into 128-row strips.  Here the matrix is computed TRANSPOSED: the kernel's
// rows are the query (the same for every alignment: one profile per thread block, L*R >= len(y) rows, no strips)
// and its columns are the database sequence.  The two 16-bit halves of a register still hold two independent
// alignments, now two database sequences: a cell pair's score is two LDS (one per half's column symbol) merged
// by one PRMT.  H(i, j) of the reference is the kernel's cell (row j, column i); pass 2 (trace_body<QS = true>)
// keys the arg-max and walks back in the reference's order expressed in this frame.
//
// PairDesc in this mode: xA / xB = offsets of the two database sequences in the batch (column streams),
// mA / mB = their lengths (columns per half), n = max(mA, mB) (columns stepped), one strip; PassParams:
// ref_code / ref_raw = the BATCH (codes / bytes), reads_raw = the QUERY bytes, table = [257][KP] (s + G) indexed
// by (query byte | 256 = padding row, batch code | KP-1 = padding column).
#pragma once
#include "sw_core.cuh"
#include <type_traits>

namespace swb {

struct QsParams { PassParams pp; int m; };              // m = len(y), the kernel's row count
struct QsTraceParams { TraceParams tp; int m; };

// prof[(code*R + k)*32 + lane] = score of row (lane & (L-1))*R + k against batch code `code`, as PT.
// PT = uint32_t: pack(s, s), one bank per lane — the score pass (measured 1.4x faster than the 16-bit form there).
// PT = uint16_t: half the shared memory (both halves of a register sit in the same row, so one 16-bit entry
// serves either) — pass 2, where more resident warps hide the latency of the replay and of the walk; stored in the
// PAIRED layout below (qs_build_profile_paired), this builder then only serves the 32-bit form.
template <int R, class PT>
__device__ __forceinline__ void qs_build_profile(uint32_t* prof32, const PassParams& p, int m) {
  PT* prof = reinterpret_cast<PT*>(prof32);
  const int total = p.KP * R * 32;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int lane = idx & 31, kc = idx >> 5;
    const int code = kc / R, k = kc - code * R;
    const int row = (lane & (p.L - 1)) * R + k;
    const int a = row < m ? (int)p.reads_raw[row] : 256;
    const uint32_t v = (uint32_t)(uint16_t)p.table[a * p.KP + code];
    prof[idx] = (PT)(sizeof(PT) == 4 ? v * 0x00010001u : v);
  }
  __syncthreads();
}

// Pass-2 profile (PT = uint16_t), PAIRED layout: word (code*RH + k/2)*32 + lane, RH = (R+1)/2, holds the 16-bit scores of the
// lane's rows k (low half) and k+1 (high half, k even) against batch code `code`.  Bank = lane whatever the codes of the
// lanes are (every lane of the skewed wavefront is on a different column), and one LDS serves two rows: 20 loads per step
// at 19 rows instead of 38.  (The plain 16-bit layout [(code*R + k)*32 + lane] put lanes 2i and 2i+1 into one bank with
// different words whenever their codes differed in parity: ncu showed 111 M conflict wavefronts in 245 M.)
template <int R>
__device__ __forceinline__ void qs_build_profile_paired(uint32_t* prof, const PassParams& p, int m) {
  constexpr int RH = (R + 1) / 2;
  const int total = p.KP * RH * 32;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int lane = idx & 31, kc = idx >> 5;
    const int code = kc / RH, k2 = kc - code * RH;
    uint32_t v = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 2 * k2 + h;
      if (k < R) {
        const int row = (lane & (p.L - 1)) * R + k;
        const int a = row < m ? (int)p.reads_raw[row] : 256;
        v |= (uint32_t)(uint16_t)p.table[a * p.KP + code] << (16 * h);
      }
    }
    prof[idx] = v;
  }
  __syncthreads();
}

template <int R, class PT>
struct DualProfileSelect {
  const PT* prof;     // shared memory, already offset by lane
  const PT* colA;
  const PT* colB;
  __device__ __forceinline__ void set_column(uint32_t ca, uint32_t cb) { colA = prof + ca * (R * 32); colB = prof + cb * (R * 32); }
  // low half: the A column's score, high half: the B column's score
  __device__ __forceinline__ uint32_t operator()(int k, int) const {
    return __byte_perm((uint32_t)colA[k * 32], (uint32_t)colB[k * 32], sizeof(PT) == 4 ? 0x7610 : 0x5410);
  }
};

template <int R>
struct DualProfileSelect<R, uint16_t> {
  static constexpr int RH = (R + 1) / 2;
  const uint32_t* prof;   // shared memory, already offset by lane (paired layout, qs_build_profile_paired)
  const uint32_t* colA;
  const uint32_t* colB;
  __device__ __forceinline__ void set_column(uint32_t ca, uint32_t cb) { colA = prof + ca * (RH * 32); colB = prof + cb * (RH * 32); }
  // rows k and k+1 (k even) share a word: low halves of A and B for the even row, high halves for the odd one
  __device__ __forceinline__ uint32_t operator()(int k, int) const {
    return __byte_perm(colA[(k >> 1) * 32], colB[(k >> 1) * 32], (k & 1) ? 0x7632 : 0x5410);
  }
};

template <int R, bool SAT, class PT>
struct QsWavefront {
  static constexpr int C = 1;
  const PassParams& p;
  DualProfileSelect<R, PT> dsel;
  LaneState<R, 1> st;
  int L, g, lane;
  RegSelect<R, 1> ra, rb;
  uint32_t a0, b0, a1, b1;       // column codes (half A / half B) of the next two steps
  __device__ __forceinline__ QsWavefront(const PassParams& p_) : p(p_) {}

  __device__ __forceinline__ size_t blk_index(const PairDesc&, int b) const { return (size_t)b; }
  __device__ __forceinline__ size_t ck_index(const PairDesc&, int b) const { return (size_t)b * state_words<R, 1, SAT>() * L; }
  __device__ __forceinline__ void prepare(const PairDesc&, int, uint32_t* prof_cta) { dsel.prof = reinterpret_cast<decltype(dsel.prof)>(prof_cta) + lane; }
  const uint32_t* restore_from = nullptr;   // local checkpoint of pass 2 (see Wavefront::restore_from)
  __device__ __forceinline__ void restore(const PairDesc& pd, int t0) {
    if (t0 == 0) init_state<R, 1>(st, p.sc);
    else if (restore_from) load_state<R, 1, SAT>(st, p.sc, restore_from, 32, lane);
    else load_state<R, 1, SAT>(st, p.sc, p.ckpt + pd.ck_off + ck_index(pd, (t0 >> p.logB) - 1), L, g);
  }
  template <bool MASKED>
  __device__ __forceinline__ void load_codes(const PairDesc& pd, int t, uint32_t& ca, uint32_t& cb) const {
    const int j = col_of<1>(t, g, 0);
    if (!MASKED) { ca = __ldg(p.ref_code + pd.xA + (uint32_t)(j - 1)); cb = __ldg(p.ref_code + pd.xB + (uint32_t)(j - 1)); return; }
    const bool ia = j >= 1 && j <= (int)pd.mA, ib = j >= 1 && j <= (int)pd.mB;
    const uint32_t va = __ldg(p.ref_code + (ia ? pd.xA + (uint32_t)(j - 1) : pd.xA));
    const uint32_t vb = __ldg(p.ref_code + (ib ? pd.xB + (uint32_t)(j - 1) : pd.xA));
    ca = ia ? va : (uint32_t)(p.KP - 1);
    cb = ib ? vb : (uint32_t)(p.KP - 1);
  }
  template <class Sel, class Hook>
  __device__ __forceinline__ void step_q(int t, const Sel& sel, uint32_t& bmax, Hook&& hook) {
    uint32_t upv[1];
    upv[0] = __shfl_up_sync(0xffffffffu, st.bot[0], 1, L);
    if (g == 0) upv[0] = p.sc.negG2;                      // row 0 of H is zero: E = -G
    step<R, 1, SAT>(st, sel, p.sc, upv, bmax, [&](int k, int c, uint32_t e_new) { hook(k, c, t, col_of<1>(t, g, 0), e_new); });
  }
  // software pipeline of the score pass, as Wavefront::two_steps (profile form)
  __device__ __forceinline__ void prime(const PairDesc& pd, int t) {
    load_codes<true>(pd, t, a0, b0);
    dsel.set_column(a0, b0);
    ra.fetch(dsel);
    load_codes<true>(pd, t + 1, a0, b0);
  }
  template <bool MASKED>
  __device__ __forceinline__ void two_steps(const PairDesc& pd, int t, uint32_t& bmax) {
    auto nohook = [](int, int, int, int, uint32_t) {};
    load_codes<MASKED>(pd, t + 2, a1, b1);
    dsel.set_column(a0, b0);
    rb.fetch(dsel);                                       // scores of step t+1
    step_q(t, ra, bmax, nohook);
    load_codes<MASKED>(pd, t + 3, a0, b0);
    dsel.set_column(a1, b1);
    ra.fetch(dsel);                                       // scores of step t+2
    step_q(t + 1, rb, bmax, nohook);
  }
  // pass-2 replay, same contract as Wavefront::replay (single strip: `multi` is ignored)
  template <class Hook, class Post>
  __device__ __forceinline__ void replay(const PairDesc& pd, bool, int t0, int t1, int nsteps, Hook&& hook, Post&& post) {
    restore(pd, t0);
    constexpr int DEPTH = 6;                               // look-ahead of the column codes, see Wavefront::replay_impl
    uint32_t qa[DEPTH], qb[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) load_codes<true>(pd, t0 + 1 + d, qa[d], qb[d]);
    for (int s = 1; s <= nsteps; ++s) {
      const int t = t0 + s;
      const uint32_t ca = qa[0], cb = qb[0];
#pragma unroll
      for (int d = 0; d + 1 < DEPTH; ++d) { qa[d] = qa[d + 1]; qb[d] = qb[d + 1]; }
      load_codes<true>(pd, t + DEPTH, qa[DEPTH - 1], qb[DEPTH - 1]);
      const bool on = t <= t1;
      auto h = [&](int k, int c, int tt, int j, uint32_t e_new) { if (on) hook(k, c, tt, j, e_new); };
      uint32_t smax = NEG_INF2;
      dsel.set_column(ca, cb);
      step_q(t, dsel, smax, h);
      post(t, on, smax);
    }
  }
  template <class Hook>
  __device__ __forceinline__ void replay(const PairDesc& pd, bool multi, int t0, int t1, int nsteps, Hook&& hook) {
    replay(pd, multi, t0, t1, nsteps, hook, [](int, bool, uint32_t) {});
  }
};

// Score pass: one group of L lanes per pair of database sequences, 32/L pairs per warp.
// P16: the paired 16-bit profile of pass 2 (half the loads and half the shared memory) instead of the 32-bit one; SWB_QS_PROF16=1.
template <int R, bool SAT, bool P16 = false>
__global__ void __launch_bounds__(128) qs_score_kernel(const QsParams qp) {
  extern __shared__ uint32_t smem_prof[];
  const PassParams& p = qp.pp;
  using PT = typename std::conditional<P16, uint16_t, uint32_t>::type;
  if (P16) qs_build_profile_paired<R>(smem_prof, p, qp.m);
  else qs_build_profile<R, uint32_t>(smem_prof, p, qp.m);
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int L = p.L;
  const int g = lane & (L - 1);
  int pair = gwarp * (32 >> p.logL) + (lane >> p.logL);
  const bool live = pair < p.npairs;
  if (!live) pair = p.npairs - 1;        // keep the lane in the shuffles; it stores nothing
  const PairDesc pd = p.pairs[pair];
  QsWavefront<R, SAT, PT> wf(p);
  wf.L = L; wf.g = g; wf.lane = lane;
  wf.prepare(pd, 0, smem_prof);
  const int steps = warp_max_i32((int)pd.nblk << p.logB);
  int n_min = (int)min(pd.mA, pd.mB);    // both halves' columns are in range up to here
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_min = min(n_min, __shfl_xor_sync(0xffffffffu, n_min, o));
  uint32_t* blk = p.blkmax + pd.blk_off;
  uint32_t* ck = p.ckpt + pd.ck_off;
  wf.restore(pd, 0);
  wf.prime(pd, 1);
  uint32_t bmax = NEG_INF2;
  const int nb = steps >> p.logB;
  for (int b = 0; b < nb; ++b) {
    const int t0 = b << p.logB;
    const bool interior = (t0 + 1 >= L) && (t0 + p.B + 2 <= n_min);
    if (interior) { for (int t = t0 + 1; t <= t0 + p.B; t += 2) wf.template two_steps<false>(pd, t, bmax); }
    else { for (int t = t0 + 1; t <= t0 + p.B; t += 2) wf.template two_steps<true>(pd, t, bmax); }
    const uint32_t gm = group_max_s16x2(bmax, L);
    if (live && b < (int)pd.nblk) {
      if (g == 0) blk[b] = gm;
      save_state<R, 1, SAT>(wf.st, p.sc, ck + wf.ck_index(pd, b), L, g);
    }
    bmax = NEG_INF2;
  }
}

template <int R, bool SAT>
__global__ void __launch_bounds__(128, SWB_TRACE_MINBLOCKS) qs_trace_kernel(const QsTraceParams qp) {
  extern __shared__ uint32_t smem_prof[];
  qs_build_profile_paired<R>(smem_prof, qp.tp.pp, qp.m);
  trace_body<R, 1, SAT, true, true, QsWavefront<R, SAT, uint16_t>>(qp.tp, smem_prof, qp.m);
}

}  // namespace swb
