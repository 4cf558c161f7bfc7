// alu_peak.cu — instruction-throughput microbenchmark for the Smith-Waterman cell-update roofline.
//
// SURVEY.md §8(d) asks for P_int (peak simple-INT32 lane-ops/s) to be MEASURED because
// MEASURED_PEAKS.json only holds HBM and bf16 numbers.  Each kernel below runs NCHAIN
// independent dependency chains of one SASS instruction class per thread, so the issuing
// pipe (not latency) is the limit.  We report warp-instructions per clock per SM (from
// clock64 deltas) and lane-ops/s (from CUDA events).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o alu_peak alu_peak.cu
// Run  : ./alu_peak [json_out]
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(2);} } while (0)

constexpr int NCHAIN = 8;
constexpr int UNROLL = 16;
constexpr int ITERS  = 2048;

struct Out { unsigned long long cycles; unsigned sink; };
template <typename T> __device__ __forceinline__ unsigned p0_as_u(T p) { return *reinterpret_cast<unsigned*>(&p) * 2654435761u; }

#define KERNEL(NAME, TYPE, INIT, BODY)                                              \
__global__ void __launch_bounds__(256) NAME(Out* out, TYPE p0, TYPE p1, TYPE p2) {  \
  TYPE v[NCHAIN];                                                                   \
  _Pragma("unroll") for (int c = 0; c < NCHAIN; ++c) { v[c] = INIT; }               \
  __syncthreads();                                                                  \
  long long t0 = clock64();                                                         \
  for (int it = 0; it < ITERS; ++it) {                                              \
    _Pragma("unroll") for (int u = 0; u < UNROLL; ++u) {                            \
      _Pragma("unroll") for (int c = 0; c < NCHAIN; ++c) { BODY; }                  \
    }                                                                               \
  }                                                                                 \
  long long t1 = clock64();                                                         \
  unsigned s = 0;                                                                   \
  _Pragma("unroll") for (int c = 0; c < NCHAIN; ++c) s ^= *reinterpret_cast<unsigned*>(&v[c]); \
  if (threadIdx.x == 0) out[blockIdx.x].cycles = (unsigned long long)(t1 - t0);     \
  if (s == p0_as_u(p0)) out[blockIdx.x].sink = s;                                   \
}

// ---- scalar INT32 ---------------------------------------------------------------------------
KERNEL(k_iadd3,        int, (int)threadIdx.x + c, v[c] = v[c] + v[(c+1)&7] + v[(c+2)&7])              // IADD3 (3-input add cannot become IMAD)
KERNEL(k_imad,         int, (int)threadIdx.x + c, v[c] = v[c] * p0 + p1)                 // IMAD (fma pipe)
KERNEL(k_lop3,         unsigned, threadIdx.x + c, v[c] = (v[c] & v[(c+1)&7]) ^ v[(c+2)&7])            // LOP3
KERNEL(k_imnmx,        int, (int)threadIdx.x + c, v[c] = max(v[c], v[(c+1)&7]) )                // VIMNMX s32
KERNEL(k_viaddmax,     int, (int)threadIdx.x + c, v[c] = __viaddmax_s32(v[c], p0, p1))   // VIADDMNMX
KERNEL(k_viaddmax_relu,int, (int)threadIdx.x + c, v[c] = __viaddmax_s32_relu(v[c], p0, p1))
// ---- packed s16x2 ---------------------------------------------------------------------------
KERNEL(k_viadd16x2,    unsigned, threadIdx.x + c, v[c] = __vadd2(v[c], v[(c+1)&7]))             // VIADD.16x2
KERNEL(k_vimnmx16x2,   unsigned, threadIdx.x + c, v[c] = __vmaxs2(v[c], v[(c+1)&7]))            // VIMNMX.S16x2
KERNEL(k_vimnmx3_16x2, unsigned, threadIdx.x + c, v[c] = __vimax3_s16x2(v[c], p0, p1))   // VIMNMX3.S16x2
KERNEL(k_viaddmax16x2, unsigned, threadIdx.x + c, v[c] = __viaddmax_s16x2(v[c], p0, p1)) // VIADDMNMX.S16x2
KERNEL(k_viaddmax16x2_relu, unsigned, threadIdx.x + c, v[c] = __viaddmax_s16x2_relu(v[c], p0, p1))
// ---- half2 ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned hset2_eq(unsigned a, unsigned b) {
  unsigned r; asm("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
KERNEL(k_hset2,        unsigned, threadIdx.x + c, v[c] = hset2_eq(v[c], p0) ^ p1)        // HSET2 + LOP3
KERNEL(k_hfma2,        __half2, __floats2half2_rn(1.f + c, 2.f), v[c] = __hfma2(v[c], p0, p1))
KERNEL(k_hmnmx2,       __half2, __floats2half2_rn(1.f + c, 2.f), v[c] = __hmax2(v[c], v[(c+1)&7]))
KERNEL(k_fmnmx,        float, 1.f + c, v[c] = fmaxf(v[c], v[(c+1)&7]))
KERNEL(k_ffma,         float, 1.f + c, v[c] = fmaf(v[c], p0, p1))
// ---- mixes: can ALU-pipe and FMA-pipe instructions dual-issue to reach 1 warp-inst/clk/SMSP? -----
KERNEL(k_mix_viaddmax_imad, int, (int)threadIdx.x + c, v[c] = (c & 1) ? __viaddmax_s32(v[c], p0, p1) : v[c] * p0 + p1)
KERNEL(k_mix_vimnmx16_hfma2, unsigned, threadIdx.x + c, {
  if (c & 1) v[c] = __vmaxs2(v[c], v[(c+2)&7]);
  else { __half2 h = *reinterpret_cast<__half2*>(&v[c]); h = __hfma2(h, *reinterpret_cast<__half2*>(&p1), *reinterpret_cast<__half2*>(&p2)); v[c] = *reinterpret_cast<unsigned*>(&h); } })
KERNEL(k_mix_viaddmax16_hset2, unsigned, threadIdx.x + c, v[c] = (c & 1) ? __viaddmax_s16x2(v[c], p0, p1) : hset2_eq(v[c], p0))
// ---- warp shuffle and shared-memory load (per-column overheads of the wavefront kernel) --------------
KERNEL(k_shfl,         unsigned, threadIdx.x + c, v[c] = __shfl_up_sync(0xffffffffu, v[c], 1))
__global__ void __launch_bounds__(256) k_lds(Out* out, unsigned p0, unsigned p1, unsigned p2) {
  __shared__ unsigned tab[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = (i * 37 + p0) & 1023;
  unsigned v[NCHAIN];
  #pragma unroll
  for (int c = 0; c < NCHAIN; ++c) v[c] = (threadIdx.x + c * 32) & 1023;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    #pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      #pragma unroll
      for (int c = 0; c < NCHAIN; ++c) v[c] = tab[v[c]];
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
  #pragma unroll
  for (int c = 0; c < NCHAIN; ++c) s ^= v[c];
  if (threadIdx.x == 0) out[blockIdx.x].cycles = (unsigned long long)(t1 - t0);
  if (s == p0_as_u(p0)) out[blockIdx.x].sink = s;
}

template <typename T> struct Arg { static T make(float f); };
template <> int Arg<int>::make(float f) { return (int)f; }
template <> unsigned Arg<unsigned>::make(float f) { return (unsigned)f * 0x00010001u; }
template <> float Arg<float>::make(float f) { return f; }
template <> __half2 Arg<__half2>::make(float f) { return __floats2half2_rn(f, f); }

struct Row { std::string name; double winst_per_clk_sm; double tlaneops; double ms; double sm_mhz; };

template <typename T>
Row run(const char* name, void (*k)(Out*, T, T, T), int nsm, int ctas_per_sm, int ops_per_inst_extra = 1) {
  int grid = nsm * ctas_per_sm;
  Out* d; CK(cudaMalloc(&d, grid * sizeof(Out))); CK(cudaMemset(d, 0, grid * sizeof(Out)));
  T a = Arg<T>::make(1.f), b = Arg<T>::make(3.f), c = Arg<T>::make(2.f);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) k<<<grid, 256>>>(d, a, b, c);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0)); k<<<grid, 256>>>(d, a, b, c); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  std::vector<Out> h(grid); CK(cudaMemcpy(h.data(), d, grid * sizeof(Out), cudaMemcpyDeviceToHost));
  double cyc = 0; for (auto& o : h) cyc += (double)o.cycles; cyc /= grid;
  double winst_cta = (double)ITERS * UNROLL * NCHAIN * 8 /*warps per CTA*/ * ops_per_inst_extra;
  Row row; row.name = name;
  row.winst_per_clk_sm = winst_cta * ctas_per_sm / cyc;       // all resident CTAs share the SM for ~cyc clocks
  row.tlaneops = winst_cta * 32.0 * grid / (best * 1e-3) / 1e12;
  row.ms = best;
  row.sm_mhz = row.tlaneops * 1e12 / (row.winst_per_clk_sm * 32.0 * nsm) / 1e6;
  CK(cudaFree(d)); CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  return row;
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  const int CPS = 8;  // 8 CTAs x 256 thr = 64 warps/SM (full occupancy)
  std::vector<Row> rows;
  rows.push_back(run<int>("IADD3", k_iadd3, nsm, CPS));
  rows.push_back(run<int>("IMAD", k_imad, nsm, CPS));
  rows.push_back(run<unsigned>("LOP3", k_lop3, nsm, CPS));
  rows.push_back(run<int>("VIMNMX.S32", k_imnmx, nsm, CPS));
  rows.push_back(run<int>("VIADDMNMX.S32", k_viaddmax, nsm, CPS));
  rows.push_back(run<int>("VIADDMNMX.S32.RELU", k_viaddmax_relu, nsm, CPS));
  rows.push_back(run<unsigned>("VIADD.16x2", k_viadd16x2, nsm, CPS));
  rows.push_back(run<unsigned>("VIMNMX.S16x2", k_vimnmx16x2, nsm, CPS));
  rows.push_back(run<unsigned>("VIMNMX3.S16x2", k_vimnmx3_16x2, nsm, CPS));
  rows.push_back(run<unsigned>("VIADDMNMX.S16x2", k_viaddmax16x2, nsm, CPS));
  rows.push_back(run<unsigned>("VIADDMNMX.S16x2.RELU", k_viaddmax16x2_relu, nsm, CPS));
  rows.push_back(run<unsigned>("HSET2.EQ+LOP3 (2 inst)", k_hset2, nsm, CPS, 2));
  rows.push_back(run<__half2>("HFMA2", k_hfma2, nsm, CPS));
  rows.push_back(run<__half2>("HMNMX2", k_hmnmx2, nsm, CPS));
  rows.push_back(run<float>("FMNMX", k_fmnmx, nsm, CPS));
  rows.push_back(run<float>("FFMA", k_ffma, nsm, CPS));
  rows.push_back(run<int>("mix VIADDMNMX.S32 + IMAD", k_mix_viaddmax_imad, nsm, CPS));
  rows.push_back(run<unsigned>("mix VIMNMX.S16x2 + HFMA2", k_mix_vimnmx16_hfma2, nsm, CPS));
  rows.push_back(run<unsigned>("mix VIADDMNMX.S16x2 + HSET2", k_mix_viaddmax16_hset2, nsm, CPS));
  rows.push_back(run<unsigned>("SHFL.UP", k_shfl, nsm, CPS));
  rows.push_back(run<unsigned>("LDS.32 (dependent, conflict-free-ish)", k_lds, nsm, CPS));

  printf("device: %s, %d SMs, clockRate attr %d kHz\n", p.name, nsm, p.clockRate);
  printf("%-42s %14s %14s %10s %10s\n", "instruction", "winst/clk/SM", "Tlane-ops/s", "ms", "SM MHz*");
  for (auto& r : rows) printf("%-42s %14.3f %14.3f %10.3f %10.0f\n", r.name.c_str(), r.winst_per_clk_sm, r.tlaneops, r.ms, r.sm_mhz);
  if (argc > 1) {
    FILE* f = fopen(argv[1], "w");
    fprintf(f, "{\"device\": \"%s\", \"sms\": %d, \"rows\": [", p.name, nsm);
    for (size_t i = 0; i < rows.size(); ++i)
      fprintf(f, "%s{\"inst\": \"%s\", \"warp_inst_per_clk_per_sm\": %.4f, \"tera_lane_ops_per_s\": %.4f, \"ms\": %.4f, \"sm_mhz_implied\": %.0f}",
              i ? ", " : "", rows[i].name.c_str(), rows[i].winst_per_clk_sm, rows[i].tlaneops, rows[i].ms, rows[i].sm_mhz);
    fprintf(f, "]}\n"); fclose(f);
  }
  return 0;
}
