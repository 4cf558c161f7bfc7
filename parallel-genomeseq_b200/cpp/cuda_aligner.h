// cuda_aligner.h — C++ shims that give libswb200.so (include/swb200.h) the reference's aligner API.
//
// Drop-in surface (reference paths relative to kosta777/parallel-genomeseq):
//   CUDASWAligner<SMT>                  <->  SWAligner<SMT>                    src/aligner/smithwaterman.h:12-58
//   CUDAParallelLocalAligner<SMT, LAT>  <->  OMPParallelLocalAligner<SMT, LAT> src/aligner/plocalaligner.h:6-33
//   both implement LocalAligner<SMT> / ParallelLocalAligner<SMT, LAT>          src/aligner/localaligner.h:7-28
//   CUDABatchAligner                    <->  the per-read driver loops         src/sw_solve_small.cpp:56-101,
//                                                                              src/mpi_sw_solve_uniprot.cpp:95-138
// SMT is a tag type selecting the arithmetic:
//   CUDA_Similarity_Matrix_Skewed  -> SWB_MODE_SAT_U8  (what Similarity_Matrix_Skewed computes)
//   CUDA_Similarity_Matrix         -> SWB_MODE_EXACT   (what Similarity_Matrix computes)
//
// Build modes:
//   * standalone (default): the abstract interfaces are mirrored here (swb::LocalAligner, ...) and
//     getTimings() returns swb::Timings (operator[] / operator() / size(), like the 2-vector the reference returns);
//   * -DSWB_WITH_REFERENCE_HEADERS with the reference's src/aligner and Eigen on the include path: the
//     classes derive from the reference's own LocalAligner / ParallelLocalAligner and getTimings() returns
//     Eigen::VectorXf, so sw_solve_small.cpp:82-88 compiles unchanged after a type swap.
//
// Semantics kept from the reference: inputs are borrowed string_views that must outlive the object
// (smithwaterman.h:48-49); consensus views point into object-owned strings; objects are not thread-safe;
// getTimings()[0] is the time spent computing the matrix in microseconds — here DEVICE time from CUDA events.
// Errors: the reference has none in-band (it asserts/aborts); the shims throw swb::Error.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "../../include/swb200.h"

#ifdef SWB_WITH_REFERENCE_HEADERS
#include "localaligner.h"
#endif

namespace swb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#ifndef SWB_WITH_REFERENCE_HEADERS
struct Timings {
  float v[2] = {0.f, 0.f};
  float operator[](int i) const { return v[i]; }
  float& operator[](int i) { return v[i]; }
  float operator()(int i) const { return v[i]; }
  float& operator()(int i) { return v[i]; }
  int size() const { return 2; }
};
using TimingsVec = Timings;
inline TimingsVec make_timings(float a, float b) { Timings t; t.v[0] = a; t.v[1] = b; return t; }
// mirrors of src/aligner/localaligner.h:7-28
template <class SMT>
class LocalAligner {
 public:
  virtual ~LocalAligner() = default;
  virtual float calculateScore() = 0;
  virtual float getScore() const = 0;
  virtual unsigned int getPos() const = 0;
  virtual std::string_view getConsensus_x() const = 0;
  virtual std::string_view getConsensus_y() const = 0;
  virtual const SMT& getSimilarity_matrix() const = 0;
  virtual TimingsVec getTimings() const = 0;
};
template <class SMT, class LAT>
class ParallelLocalAligner {
 public:
  virtual ~ParallelLocalAligner() = default;
  virtual float calculateScore() = 0;
  virtual float getScore() const = 0;
  virtual unsigned int getPos() const = 0;
  virtual std::string_view getConsensus_x() const = 0;
  virtual std::string_view getConsensus_y() const = 0;
  virtual TimingsVec getTimings() const = 0;
};
#else
using TimingsVec = Eigen::VectorXf;
inline TimingsVec make_timings(float a, float b) { Eigen::VectorXf t(2); t(0) = a; t(1) = b; return t; }
template <class SMT> using LocalAligner = ::LocalAligner<SMT>;
template <class SMT, class LAT> using ParallelLocalAligner = ::ParallelLocalAligner<SMT, LAT>;
#endif

using ScoringFn = std::function<float(const char&, const char&)>;

// One swb_ctx per host thread, shared by all shim objects of that thread (the reference constructs one
// aligner per read; re-creating a CUDA context per read would dominate).  The reference sequence and the
// scoring are re-sent only when they change.
class Context {
 public:
  static Context& instance() { static thread_local Context c; return c; }
  swb_ctx* raw() { ensure(); return ctx_; }
  void check(int rc) { if (rc != SWB_OK) throw Error(rc, swb_last_error(ctx_)); }

  void set_reference(std::string_view y) {
    ensure();
    if (y.size() == ref_.size() && (y.empty() || std::memcmp(y.data(), ref_.data(), y.size()) == 0)) return;
    check(swb_set_reference(ctx_, y.data(), y.size()));
    ref_.assign(y.data(), y.size());
  }
  void set_scoring_match(int mode, float match, float mismatch, float gap) {
    ensure();
    if (have_match_ && mode == mode_ && match == match_ && mismatch == mismatch_ && gap == gap_) return;
    check(swb_set_scoring_match(ctx_, mode, match, mismatch, gap));
    have_match_ = true; mode_ = mode; match_ = match; mismatch_ = mismatch; gap_ = gap;
  }
  // Tabulate the callback over all byte pairs: table[a*256+b] = fn(a, b)  (argument order of similaritymatrix.cpp:252-254)
  void set_scoring_fn(int mode, const ScoringFn& fn, float gap) {
    ensure();
    std::vector<float> t(65536);
    if (mode == SWB_MODE_SAT_U8) {
      // Similarity_Matrix_Skewed only ever probes fn('A','A') and fn('A','T') (similaritymatrix.cpp:389-390)
      const char A = 'A', T = 'T';
      t[(unsigned char)A * 256 + (unsigned char)A] = fn(A, A);
      t[(unsigned char)A * 256 + (unsigned char)T] = fn(A, T);
    } else {
      for (int a = 0; a < 256; ++a)
        for (int b = 0; b < 256; ++b) { const char ca = (char)a, cb = (char)b; t[a * 256 + b] = fn(ca, cb); }
    }
    check(swb_set_scoring(ctx_, mode, t.data(), gap));
    have_match_ = false;
  }
  ~Context() { if (ctx_) swb_destroy(ctx_); }

 private:
  void ensure() {
    if (ctx_) return;
    int dev = 0;
    if (const char* e = std::getenv("SWB_DEVICE")) dev = std::atoi(e);
    int rc = swb_create(dev, &ctx_);
    if (rc != SWB_OK) throw Error(rc, "swb_create failed: no usable CUDA device (libswb200 has no CPU fallback)");
  }
  swb_ctx* ctx_ = nullptr;
  std::string ref_;
  bool have_match_ = false;
  int mode_ = -1;
  float match_ = 0, mismatch_ = 0, gap_ = 0;
};

// SMT tags.  operator()(row, col) is the Abstract_Similarity_Matrix accessor (similaritymatrix.h:13-24); the
// matrix is never materialised by the aligners, so the accessor fills a dense copy on first use (small inputs).
template <int MODE>
class CUDA_Similarity_Matrix_Tag {
 public:
  static constexpr int mode = MODE;
  CUDA_Similarity_Matrix_Tag() = default;
  CUDA_Similarity_Matrix_Tag(std::string_view x, std::string_view y) : x_(x), y_(y) {}
  float operator()(int64_t row, int64_t col) const {
    if (dense_.empty()) {
      dense_.resize((x_.size() + 1) * (y_.size() + 1));
      Context& c = Context::instance();
      if (apply_scoring_) apply_scoring_(c); else c.set_scoring_match(MODE, 3.0f, -3.0f, 2.0f);
      c.set_reference(y_);
      c.check(swb_matrix(c.raw(), x_.data(), x_.size(), dense_.data()));
    }
    return (float)dense_[(size_t)row * (y_.size() + 1) + (size_t)col];
  }
 private:
  template <class> friend class CUDASWAligner;
  std::string_view x_, y_;
  std::function<void(Context&)> apply_scoring_;   // set by the owning aligner: its callback + gap
  mutable std::vector<int32_t> dense_;
};
using CUDA_Similarity_Matrix_Skewed = CUDA_Similarity_Matrix_Tag<SWB_MODE_SAT_U8>;
using CUDA_Similarity_Matrix = CUDA_Similarity_Matrix_Tag<SWB_MODE_EXACT>;

namespace detail {
struct Scoring {
  bool is_fn = false;
  float match = 3.0f, mismatch = -3.0f, gap = 2.0f;   // smithwaterman.cpp:8: a == b ? 3.0 : -3.0, gap 2.0
  ScoringFn fn;
  void apply(Context& c, int mode) const { if (is_fn) c.set_scoring_fn(mode, fn, gap); else c.set_scoring_match(mode, match, mismatch, gap); }
};
struct Result {
  float score = -1.f;          // max_score(-1), smithwaterman.cpp:27
  unsigned pos = 0;
  std::string cx, cy;
  float device_us = 0.f;
};
inline Result run_one(int mode, const Scoring& sc, std::string_view x, std::string_view y, int npiece, float ratio) {
  Context& c = Context::instance();
  sc.apply(c, mode);
  c.set_reference(y);
  const uint64_t offs[2] = {0, x.size()};
  size_t cap = 2 * x.size() + 64;
  Result r;
  for (;;) {
    int32_t score = 0; uint32_t pos = 0, len = 0, flags = 0;
    r.cx.assign(cap, '\0'); r.cy.assign(cap, '\0');
    c.check(swb_align_batch(c.raw(), x.data(), offs, 1, npiece, ratio, SWB_FLAG_CONSENSUS, &score, &pos, nullptr,
                            r.cx.data(), r.cy.data(), &len, cap, &flags, &r.device_us));
    if (flags & SWB_RES_CONS_TRUNCATED) { cap = x.size() + y.size() + 2; continue; }   // grow and retry (rare)
    r.score = (float)score; r.pos = pos; r.cx.resize(len); r.cy.resize(len);
    return r;
  }
}
}  // namespace detail

template <class SMT>
class CUDASWAligner : public LocalAligner<SMT> {
 public:
  // the four constructors of smithwaterman.h:14-17
  CUDASWAligner(std::string_view x, std::string_view y) : x_(x), y_(y), sm_(x, y) {}
  CUDASWAligner(std::string_view x, std::string_view y, float gap) : x_(x), y_(y), sm_(x, y) { sc_.gap = gap; }
  CUDASWAligner(std::string_view x, std::string_view y, ScoringFn&& fn) : x_(x), y_(y), sm_(x, y) { sc_.is_fn = true; sc_.fn = std::move(fn); }
  CUDASWAligner(std::string_view x, std::string_view y, ScoringFn&& fn, float gap) : x_(x), y_(y), sm_(x, y) { sc_.is_fn = true; sc_.fn = std::move(fn); sc_.gap = gap; }

  float calculateScore() override { r_ = detail::run_one(SMT::mode, sc_, x_, y_, 0, 0.f); return r_.score; }
  float getScore() const override { return r_.score; }
  unsigned int getPos() const override { return r_.pos; }
  std::string_view getConsensus_x() const override { return r_.cx; }
  std::string_view getConsensus_y() const override { return r_.cy; }
  const SMT& getSimilarity_matrix() const override {
    const detail::Scoring* sc = &sc_;
    sm_.apply_scoring_ = [sc](Context& c) { sc->apply(c, SMT::mode); };
    return sm_;
  }
  TimingsVec getTimings() const override { return make_timings(r_.device_us, r_.device_us); }

 private:
  std::string_view x_, y_;
  mutable SMT sm_;
  detail::Scoring sc_;
  detail::Result r_;
};

template <class SMT, class LAT = CUDASWAligner<SMT>>
class CUDAParallelLocalAligner : public ParallelLocalAligner<SMT, LAT> {
 public:
  // the four constructors of plocalaligner.h:9-12
  CUDAParallelLocalAligner(std::string_view x, std::string_view y, int npiece, float ratio) : x_(x), y_(y), npiece_(npiece), ratio_(ratio) {}
  CUDAParallelLocalAligner(std::string_view x, std::string_view y, int npiece, float ratio, float gap) : x_(x), y_(y), npiece_(npiece), ratio_(ratio) { sc_.gap = gap; }
  CUDAParallelLocalAligner(std::string_view x, std::string_view y, int npiece, float ratio, ScoringFn&& fn) : x_(x), y_(y), npiece_(npiece), ratio_(ratio) { sc_.is_fn = true; sc_.fn = std::move(fn); }
  CUDAParallelLocalAligner(std::string_view x, std::string_view y, int npiece, float ratio, ScoringFn&& fn, float gap) : x_(x), y_(y), npiece_(npiece), ratio_(ratio) { sc_.is_fn = true; sc_.fn = std::move(fn); sc_.gap = gap; }

  float calculateScore() override { r_ = detail::run_one(SMT::mode, sc_, x_, y_, npiece_, ratio_); return r_.score; }
  float getScore() const override { return r_.score; }
  unsigned int getPos() const override { return r_.pos; }
  std::string_view getConsensus_x() const override { return r_.cx; }
  std::string_view getConsensus_y() const override { return r_.cy; }
  TimingsVec getTimings() const override { return make_timings(r_.device_us, r_.device_us); }

 private:
  std::string_view x_, y_;
  int npiece_;
  float ratio_;
  detail::Scoring sc_;
  detail::Result r_;
};

// Batched entry point: what the rewritten driver loops call (one launch for all reads / DB entries).
class CUDABatchAligner {
 public:
  struct Out {
    std::vector<int32_t> score;
    std::vector<uint32_t> pos, len, flags;
    std::vector<char> cx, cy;
    size_t stride = 0;
    float device_us = 0.f;
    std::string_view consensus_x(size_t i) const { return std::string_view(cx.data() + i * stride, len[i]); }
    std::string_view consensus_y(size_t i) const { return std::string_view(cy.data() + i * stride, len[i]); }
  };
  explicit CUDABatchAligner(int mode) : mode_(mode) {}
  void set_scoring(float match, float mismatch, float gap) { sc_.is_fn = false; sc_.match = match; sc_.mismatch = mismatch; sc_.gap = gap; }
  void set_scoring(ScoringFn fn, float gap) { sc_.is_fn = true; sc_.fn = std::move(fn); sc_.gap = gap; }
  void set_reference(std::string_view y) { y_ = y; }
  // npiece <= 0: SWAligner semantics; npiece >= 1: OMPParallelLocalAligner(x, y, npiece, ratio) semantics
  Out align(const std::vector<std::string_view>& xs, int npiece = 0, float ratio = 0.f, bool consensus = true) {
    Context& c = Context::instance();
    sc_.apply(c, mode_);
    c.set_reference(y_);
    std::string blob;
    std::vector<uint64_t> offs(xs.size() + 1, 0);
    size_t maxlen = 0;
    for (size_t i = 0; i < xs.size(); ++i) { blob.append(xs[i]); offs[i + 1] = blob.size(); maxlen = std::max(maxlen, xs[i].size()); }
    Out o;
    o.stride = consensus ? 2 * maxlen + 64 : 0;
    o.score.resize(xs.size()); o.pos.resize(xs.size()); o.len.resize(xs.size()); o.flags.resize(xs.size());
    if (consensus) { o.cx.resize(xs.size() * o.stride); o.cy.resize(xs.size() * o.stride); }
    c.check(swb_align_batch(c.raw(), blob.data(), offs.data(), xs.size(), npiece, ratio, consensus ? SWB_FLAG_CONSENSUS : 0u,
                            o.score.data(), o.pos.data(), nullptr, consensus ? o.cx.data() : nullptr, consensus ? o.cy.data() : nullptr,
                            o.len.data(), o.stride, o.flags.data(), &o.device_us));
    return o;
  }
 private:
  int mode_;
  detail::Scoring sc_;
  std::string_view y_;
};

}  // namespace swb
