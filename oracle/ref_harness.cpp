// ref_harness.cpp — C-ABI wrapper around the UNMODIFIED reference aligner classes.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; the reference sources it wraps are compiled in
// place from /root/reference/src/aligner/*.cpp by oracle/build_ref.py and are never copied into
// this repository.  The resulting oracle/_ref/libref_aligner.so is used (a) to pin the C
// restatement in oracle/sw_oracle.c, (b) to generate the golden vectors under tests/golden/
// (tests/golden/make_golden.py) and (c) as the timed CPU baseline of bench.py
// (cpu_baseline.kind == "reference").  Nothing in the product path links or loads it.
//
// Wrapped reference entry points:
//   SWAligner<SMT>                       src/aligner/smithwaterman.{h,cpp}
//   OMPParallelLocalAligner<SMT, LAT>    src/aligner/plocalaligner.{h,cpp}   (serial build)
//   Similarity_Matrix / _Skewed          src/aligner/similaritymatrix.{h,cpp}
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <string_view>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "smithwaterman.h"
#include "plocalaligner.h"

namespace {

using ScoreFn = std::function<float(const char&, const char&)>;

// scoring spec: kind 0 = reference default ctor (+3/-3, gap 2);
//               kind 1 = match/mismatch/gap through the callback ctor;
//               kind 2 = 256x256 float table through the callback ctor (fn(a,b) = table[a][b]).
struct Scoring {
  int kind; float match; float mismatch; float gap; const float* table;
  ScoreFn fn() const {
    if (kind == 2) { const float* t = table; return [t](const char& a, const char& b) { return t[(unsigned char)a * 256 + (unsigned char)b]; }; }
    float ma = match, mi = mismatch;
    return [ma, mi](const char& a, const char& b) { return a == b ? ma : mi; };
  }
};

template <class Aligner>
int finish(Aligner& al, float* score, unsigned* pos, char* cx, char* cy, int cap, int* cons_len, float* iterate_us) {
  float s = al.calculateScore();
  if (score) *score = s;
  if (pos) *pos = al.getPos();
  auto vx = al.getConsensus_x(); auto vy = al.getConsensus_y();
  if (cons_len) *cons_len = (int)vx.size();
  if (cx && cy) {
    if ((int)vx.size() > cap || (int)vy.size() > cap) return -2;
    std::memcpy(cx, vx.data(), vx.size()); std::memcpy(cy, vy.data(), vy.size());
  }
  if (iterate_us) *iterate_us = al.getTimings()[0];
  return 0;
}

template <class SMT>
int run_sw(std::string_view x, std::string_view y, const Scoring& sc, float* score, unsigned* pos,
           char* cx, char* cy, int cap, int* cons_len, float* us) {
  if (sc.kind == 0) { SWAligner<SMT> al(x, y); return finish(al, score, pos, cx, cy, cap, cons_len, us); }
  SWAligner<SMT> al(x, y, sc.fn(), sc.gap);
  return finish(al, score, pos, cx, cy, cap, cons_len, us);
}

template <class SMT>
int run_omp(std::string_view x, std::string_view y, const Scoring& sc, int npiece, float ratio, float* score,
            unsigned* pos, char* cx, char* cy, int cap, int* cons_len, float* us) {
  using P = OMPParallelLocalAligner<SMT, SWAligner<SMT>>;
  if (sc.kind == 0) { P al(x, y, npiece, ratio); return finish(al, score, pos, cx, cy, cap, cons_len, us); }
  P al(x, y, npiece, ratio, sc.fn(), sc.gap);
  return finish(al, score, pos, cx, cy, cap, cons_len, us);
}

}  // namespace

extern "C" {

// smt: 0 = Similarity_Matrix_Skewed (u8 AVX2), 1 = Similarity_Matrix (f32).
// npiece <= 0: SWAligner<SMT>(x, y, ...); npiece >= 1: OMPParallelLocalAligner<SMT, SWAligner<SMT>>(x, y, npiece, ratio, ...).
// cx/cy receive the consensus strings exactly as the reference stores them (end -> start), not NUL-terminated.
int ref_align(int smt, const char* x, int64_t m, const char* y, int64_t n,
              int scoring_kind, float match, float mismatch, float gap, const float* table,
              int npiece, float ratio,
              float* score, unsigned* pos, char* cx, char* cy, int cap, int* cons_len, float* iterate_us) {
  std::string_view sx(x, (size_t)m), sy(y, (size_t)n);
  Scoring sc{scoring_kind, match, mismatch, gap, table};
  if (npiece <= 0) {
    return smt == 0 ? run_sw<Similarity_Matrix_Skewed>(sx, sy, sc, score, pos, cx, cy, cap, cons_len, iterate_us)
                    : run_sw<Similarity_Matrix>(sx, sy, sc, score, pos, cx, cy, cap, cons_len, iterate_us);
  }
  return smt == 0 ? run_omp<Similarity_Matrix_Skewed>(sx, sy, sc, npiece, ratio, score, pos, cx, cy, cap, cons_len, iterate_us)
                  : run_omp<Similarity_Matrix>(sx, sy, sc, npiece, ratio, score, pos, cx, cy, cap, cons_len, iterate_us);
}

// Dense H matrix through the reference's operator()(row, col): out is row-major (m+1) x (n+1) floats.
int ref_matrix(int smt, const char* x, int64_t m, const char* y, int64_t n,
               int scoring_kind, float match, float mismatch, float gap, const float* table, float* out) {
  std::string_view sx(x, (size_t)m), sy(y, (size_t)n);
  Scoring sc{scoring_kind, match, mismatch, gap, table};
  ScoreFn fn = sc.kind == 0 ? ScoreFn([](const char& a, const char& b) { return a == b ? 3.0f : -3.0f; }) : sc.fn();
  float g = sc.kind == 0 ? 2.0f : sc.gap;
  auto dump = [&](auto& sm) {
    sm.iterate(fn, g);
    for (int64_t i = 0; i <= m; ++i) for (int64_t j = 0; j <= n; ++j) out[i * (n + 1) + j] = sm(i, j);
  };
  if (smt == 0) { Similarity_Matrix_Skewed sm(sx, sy); dump(sm); }
  else { Similarity_Matrix sm(sx, sy); dump(sm); }
  return 0;
}

// (index_x, index_y, max) exactly as find_index_of_maximum() returns them after iterate().
int ref_argmax(int smt, const char* x, int64_t m, const char* y, int64_t n,
               int scoring_kind, float match, float mismatch, float gap, const float* table,
               int64_t* index_x, int64_t* index_y, float* maxv) {
  std::string_view sx(x, (size_t)m), sy(y, (size_t)n);
  Scoring sc{scoring_kind, match, mismatch, gap, table};
  ScoreFn fn = sc.kind == 0 ? ScoreFn([](const char& a, const char& b) { return a == b ? 3.0f : -3.0f; }) : sc.fn();
  float g = sc.kind == 0 ? 2.0f : sc.gap;
  auto go = [&](auto& sm) { sm.iterate(fn, g); auto [ix, iy, mx] = sm.find_index_of_maximum(); *index_x = ix; *index_y = iy; *maxv = mx; };
  if (smt == 0) { Similarity_Matrix_Skewed sm(sx, sy); go(sm); } else { Similarity_Matrix sm(sx, sy); go(sm); }
  return 0;
}

// _make_string_range is a free function with external linkage in plocalaligner.cpp:44.
}  // extern "C"
std::vector<std::pair<Eigen::Index, Eigen::Index>> _make_string_range(int, Eigen::Index, Eigen::Index, float);
extern "C" {
int ref_make_string_range(int npiece, int64_t shortlen, int64_t longlen, float ratio, int64_t* left, int64_t* right) {
  auto v = _make_string_range(npiece, shortlen, longlen, ratio);
  for (size_t i = 0; i < v.size(); ++i) { left[i] = v[i].first; right[i] = v[i].second; }
  return (int)v.size();
}

// CPU baseline: the serial reference aligner per (read, reference) pair, with a
// harness-level "omp parallel for" over reads.  Returns wall seconds; *iterate_us_sum receives the
// sum of getTimings()[0] (the reference drivers' own GCUPS numerator, sw_solve_small.cpp:88-89).
double ref_bench_reads(int smt, const char* reads, const int64_t* offsets, int64_t n_reads,
                       const char* y, int64_t n, int npiece, float ratio, int nthreads,
                       double* iterate_us_sum, unsigned* pos_out, float* score_out) {
  std::string_view sy(y, (size_t)n);
  double us_sum = 0.0;
  auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : us_sum)
#endif
  for (int64_t r = 0; r < n_reads; ++r) {
    std::string_view sx(reads + offsets[r], (size_t)(offsets[r + 1] - offsets[r]));
    float s = 0, us = 0; unsigned p = 0;
    Scoring sc{0, 3, -3, 2, nullptr};
    if (npiece <= 0) {
      if (smt == 0) run_sw<Similarity_Matrix_Skewed>(sx, sy, sc, &s, &p, nullptr, nullptr, 0, nullptr, &us);
      else run_sw<Similarity_Matrix>(sx, sy, sc, &s, &p, nullptr, nullptr, 0, nullptr, &us);
    } else {
      if (smt == 0) run_omp<Similarity_Matrix_Skewed>(sx, sy, sc, npiece, ratio, &s, &p, nullptr, nullptr, 0, nullptr, &us);
      else run_omp<Similarity_Matrix>(sx, sy, sc, npiece, ratio, &s, &p, nullptr, nullptr, 0, nullptr, &us);
    }
    us_sum += us;
    if (pos_out) pos_out[r] = p;
    if (score_out) score_out[r] = s;
  }
  auto t1 = std::chrono::steady_clock::now();
  if (iterate_us_sum) *iterate_us_sum = us_sum;
  return std::chrono::duration<double>(t1 - t0).count();
}

#ifdef USEOMP
// ---- the reference's OWN OpenMP paths (second build: -DUSEOMP -fopenmp -> oracle/_ref/libref_aligner_omp.so) -------------
// TIMING ONLY: the USEOMP build of OMPParallelLocalAligner is racy, so nothing here
// is ever used as an oracle.

// B2: sw_solve_small.cpp:82 under USEOMP — per read, OMPParallelLocalAligner<Skewed, SWAligner<Skewed>>(read, ref, npiece,
// ratio).calculateScore(); the reads run one after the other, the reference's pragmas use npiece threads per read
// (plocalaligner.cpp:93,111,120).  Returns wall seconds; *iterate_us_sum = sum of getTimings()[0].
double ref_omp_chunked_bench(const char* reads, const int64_t* offsets, int64_t n_reads, const char* y, int64_t n,
                             int npiece, float ratio, double* iterate_us_sum) {
  std::string_view sy(y, (size_t)n);
  double us_sum = 0.0;
  auto t0 = std::chrono::steady_clock::now();
  for (int64_t r = 0; r < n_reads; ++r) {
    std::string_view sx(reads + offsets[r], (size_t)(offsets[r + 1] - offsets[r]));
    OMPParallelLocalAligner<Similarity_Matrix_Skewed, SWAligner<Similarity_Matrix_Skewed>> al(sx, sy, npiece, ratio);
    al.calculateScore();
    us_sum += al.getTimings()[0];
  }
  auto t1 = std::chrono::steady_clock::now();
  if (iterate_us_sum) *iterate_us_sum = us_sum;
  return std::chrono::duration<double>(t1 - t0).count();
}

// B5: omp_sw_solve_small.cpp:164-189 — SWAligner<Similarity_Matrix> with sw_finegrain_type / sw_nthreads set by the
// driver (one "omp parallel for" per anti-diagonal, similaritymatrix.cpp:118-245).  Returns wall seconds of
// calculateScore(); timings[0..1] = getTimings() (iterate us, sum of per-diagonal us).
double ref_omp_finegrain(const char* x, int64_t m, const char* y, int64_t n, int finegrain_type, int nthreads,
                         float* score, unsigned* pos, float* timings) {
  std::string_view sx(x, (size_t)m), sy(y, (size_t)n);
  SWAligner<Similarity_Matrix> al(sx, sy);
  al.sw_finegrain_type = finegrain_type;
  al.sw_nthreads = nthreads;
  omp_set_num_threads(nthreads);
  auto t0 = std::chrono::steady_clock::now();
  const float s = al.calculateScore();
  auto t1 = std::chrono::steady_clock::now();
  if (score) *score = s;
  if (pos) *pos = al.getPos();
  if (timings) { auto t = al.getTimings(); timings[0] = t[0]; timings[1] = t[1]; }
  return std::chrono::duration<double>(t1 - t0).count();
}
#endif  // USEOMP

int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}  // extern "C"
