// sw_inst.cu — explicit instantiation of the wavefront kernels for ONE rows-per-lane value (-DSWB_R=<R>).
// build.py compiles this file once per R in parallel and links the objects into libswb200.so.
#include "sw_core.cuh"
#include "sw_qs.cuh"

#ifndef SWB_R
#error "compile with -DSWB_R=<rows per lane>"
#endif
#define SWB_CAT2(a, b) a##b
#define SWB_CAT(a, b) SWB_CAT2(a, b)

using namespace swb;

namespace {
template <class K, class P>
cudaError_t go(K k, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const P& p) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k<<<grid, block, smem, st>>>(p);
  return cudaGetLastError();
}
// one instantiation per arithmetic mode (sw_core.cuh: AM_EXACT / AM_SAT / AM_WIDE) and symbol-score select
template <int C, int AM>
cudaError_t score_cm(bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const PassParams& p) {
  if (p.strips) return profile ? go(score_strips_kernel<SWB_R, C, AM, true>, grid, block, smem, st, p) : go(score_strips_kernel<SWB_R, C, AM, false>, grid, block, 0, st, p);
  return profile ? go(score_kernel<SWB_R, C, AM, true>, grid, block, smem, st, p) : go(score_kernel<SWB_R, C, AM, false>, grid, block, 0, st, p);
}
template <int C>
cudaError_t score_c(int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const PassParams& p) {
  return am == AM_SAT ? score_cm<C, AM_SAT>(profile, grid, block, smem, st, p) : am == AM_WIDE ? score_cm<C, AM_WIDE>(profile, grid, block, smem, st, p) : score_cm<C, AM_EXACT>(profile, grid, block, smem, st, p);
}
template <int C, int AM>
cudaError_t score_units_cm(bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const PassParams& p) {
  return profile ? go(score_units_kernel<SWB_R, C, AM, true>, grid, block, smem, st, p) : go(score_units_kernel<SWB_R, C, AM, false>, grid, block, 0, st, p);
}
template <int C>
cudaError_t score_units_c(int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const PassParams& p) {
  return am == AM_SAT ? score_units_cm<C, AM_SAT>(profile, grid, block, smem, st, p) : am == AM_WIDE ? score_units_cm<C, AM_WIDE>(profile, grid, block, smem, st, p) : score_units_cm<C, AM_EXACT>(profile, grid, block, smem, st, p);
}
template <int C, int AM>
cudaError_t trace_cm(bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const TraceParams& p) {   // smem: profiles + the pass-2 rings
  return profile ? go(trace_kernel<SWB_R, C, AM, true>, grid, block, smem, st, p) : go(trace_kernel<SWB_R, C, AM, false>, grid, block, smem, st, p);
}
template <int C>
cudaError_t trace_c(int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const TraceParams& p) {
  return am == AM_SAT ? trace_cm<C, AM_SAT>(profile, grid, block, smem, st, p) : am == AM_WIDE ? trace_cm<C, AM_WIDE>(profile, grid, block, smem, st, p) : trace_cm<C, AM_EXACT>(profile, grid, block, smem, st, p);
}
template <int AM>
cudaError_t dump_cm(bool profile, size_t smem, cudaStream_t st, const DumpParams& p) {
  return profile ? go(dump_kernel<SWB_R, 1, AM, true>, dim3(1), dim3(32), smem, st, p) : go(dump_kernel<SWB_R, 1, AM, false>, dim3(1), dim3(32), 0, st, p);
}
}  // namespace

cudaError_t SWB_CAT(swb_launch_dump_r, SWB_R)(int am, bool profile, size_t smem, cudaStream_t st, const DumpParams& p) {
  return am == AM_SAT ? dump_cm<AM_SAT>(profile, smem, st, p) : am == AM_WIDE ? dump_cm<AM_WIDE>(profile, smem, st, p) : dump_cm<AM_EXACT>(profile, smem, st, p);
}
// Four columns per step exist for the pipelined strips of few long pairs only (score_units_kernel and the pass-2 kernel
// that replays its checkpoints), and only for thin strips.
#if SWB_R <= 8
#define SWB_HAVE_C4 1      // and eight columns per step
#endif
cudaError_t SWB_CAT(swb_launch_score_r, SWB_R)(int C, int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const PassParams& p) {
  if (p.units) {
#ifdef SWB_HAVE_C4
    if (C == 4) return score_units_c<4>(am, profile, grid, block, smem, st, p);
    if (C == 8) return score_units_c<8>(am, profile, grid, block, smem, st, p);
#endif
    if (C > 2) return cudaErrorInvalidValue;
    return C == 1 ? score_units_c<1>(am, profile, grid, block, smem, st, p) : score_units_c<2>(am, profile, grid, block, smem, st, p);
  }
  if (C > 2) return cudaErrorInvalidValue;
  return C == 1 ? score_c<1>(am, profile, grid, block, smem, st, p) : score_c<2>(am, profile, grid, block, smem, st, p);
}
cudaError_t SWB_CAT(swb_launch_trace_r, SWB_R)(int C, int am, bool profile, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const TraceParams& p) {
#ifdef SWB_HAVE_C4
  if (C == 4) return trace_c<4>(am, profile, grid, block, smem, st, p);
  if (C == 8) return trace_c<8>(am, profile, grid, block, smem, st, p);
#endif
  if (C > 2) return cudaErrorInvalidValue;
  return C == 1 ? trace_c<1>(am, profile, grid, block, smem, st, p) : trace_c<2>(am, profile, grid, block, smem, st, p);
}

// query-stationary kernels (sw_qs.cuh): one column per step, profile select
cudaError_t SWB_CAT(swb_launch_qs_score_r, SWB_R)(bool sat, bool p16, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const QsParams& p) {
  if (p16) return sat ? go(qs_score_kernel<SWB_R, true, true>, grid, block, smem, st, p) : go(qs_score_kernel<SWB_R, false, true>, grid, block, smem, st, p);
  return sat ? go(qs_score_kernel<SWB_R, true>, grid, block, smem, st, p) : go(qs_score_kernel<SWB_R, false>, grid, block, smem, st, p);
}
cudaError_t SWB_CAT(swb_launch_qs_trace_r, SWB_R)(bool sat, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const QsTraceParams& p) {
  return sat ? go(qs_trace_kernel<SWB_R, true>, grid, block, smem, st, p) : go(qs_trace_kernel<SWB_R, false>, grid, block, smem, st, p);
}
