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
// Lazy batching: the reference API is one object per alignment, and one kernel launch per object would be hopeless.
// Every shim object registers itself with its thread's Context when it is constructed; the first calculateScore() of any
// of them aligns ALL pending objects that share its arithmetic mode, scoring, reference and chunking in ONE
// swb_align_batch call and hands every object its own result, so a driver that constructs its aligners first and then
// queries them (the per-read loop of sw_solve_small.cpp:56-101 split in two) runs at batch speed.  A loop that constructs,
// queries and destroys one object at a time still gets one launch per object — use CUDABatchAligner there.
//
// Semantics kept from the reference: inputs are borrowed string_views that must outlive the object
// (smithwaterman.h:48-49); consensus views point into object-owned strings; objects are not thread-safe;
// getTimings()[0] is the time spent computing the matrix in microseconds — here DEVICE time from CUDA events.
// Errors: the reference has none in-band (it asserts/aborts); the shims throw swb::Error.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <type_traits>
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
  // device of this thread's context (before its first use); default: SWB_DEVICE or 0
  void set_device(int dev) { if (!ctx_) device_ = dev; else if (dev != device_) throw Error(SWB_ERR_STATE, "Context::set_device after first use"); }
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

  // lazy batching (see the header comment): objects waiting for their first calculateScore()
  void enqueue(struct Pending* p) { pending_.push_back(p); }
  void remove(struct Pending* p) { for (size_t i = 0; i < pending_.size(); ++i) if (pending_[i] == p) { pending_.erase(pending_.begin() + i); return; } }
  inline void flush(struct Pending* trigger);

 private:
  std::vector<struct Pending*> pending_;
  void ensure() {
    if (ctx_) return;
    int dev = device_;
    if (dev < 0) { dev = 0; if (const char* e = std::getenv("SWB_DEVICE")) dev = std::atoi(e); }
    device_ = dev;
    int rc = swb_create(dev, &ctx_);
    if (rc != SWB_OK) throw Error(rc, "swb_create failed: no usable CUDA device (libswb200 has no CPU fallback)");
  }
  swb_ctx* ctx_ = nullptr;
  int device_ = -1;
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
}  // namespace detail

// State every lazily batched aligner object shares (one alignment: x against y, optional chunking).
struct Pending {
  int mode = SWB_MODE_SAT_U8;
  detail::Scoring sc;
  std::string_view x, y;
  int npiece = 0;
  float ratio = 0.f;
  detail::Result r;
  float pass1_us = 0.f;
  bool ready = false, queued = false;
  std::vector<float> table;      // tabulated callback (filled when the queue is flushed)
  Pending() = default;
  Pending(const Pending&) = delete;
  Pending& operator=(const Pending&) = delete;
  void init(int m, std::string_view xs, std::string_view ys, int np, float rt) {
    mode = m; x = xs; y = ys; npiece = np; ratio = rt;
    Context::instance().enqueue(this); queued = true;
  }
  ~Pending() { if (queued) Context::instance().remove(this); }
  // the callback's table: the two probes of similaritymatrix.cpp:389-390 in SAT_U8 mode, every byte pair in EXACT
  void tabulate() {
    if (!sc.is_fn || !table.empty()) return;
    if (mode == SWB_MODE_SAT_U8) { const char A = 'A', T = 'T'; table = {sc.fn(A, A), sc.fn(A, T)}; return; }
    table.resize(65536);
    for (int a = 0; a < 256; ++a)
      for (int b = 0; b < 256; ++b) { const char ca = (char)a, cb = (char)b; table[a * 256 + b] = sc.fn(ca, cb); }
  }
  bool same_job(Pending& o) {
    if (mode != o.mode || npiece != o.npiece || ratio != o.ratio || sc.is_fn != o.sc.is_fn || sc.gap != o.sc.gap) return false;
    if (y.size() != o.y.size() || (y.data() != o.y.data() && std::memcmp(y.data(), o.y.data(), y.size()) != 0)) return false;
    if (!sc.is_fn) return sc.match == o.sc.match && sc.mismatch == o.sc.mismatch;
    tabulate(); o.tabulate();
    return table == o.table;
  }
  float calculate() {
    if (!ready) Context::instance().flush(this);
    return r.score;
  }
};

namespace detail {
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

inline void Context::flush(Pending* trigger) {
  std::vector<Pending*> job;
  for (Pending* p : pending_) if (p == trigger || (!p->ready && trigger->same_job(*p))) job.push_back(p);
  if (job.empty()) job.push_back(trigger);
  for (Pending* p : job) { remove(p); p->queued = false; }
  trigger->sc.apply(*this, trigger->mode);
  set_reference(trigger->y);
  std::string blob;
  std::vector<uint64_t> offs(job.size() + 1, 0);
  size_t maxlen = 0;
  for (size_t i = 0; i < job.size(); ++i) { blob.append(job[i]->x); offs[i + 1] = blob.size(); maxlen = std::max(maxlen, job[i]->x.size()); }
  const size_t stride = 2 * maxlen + 64;
  std::vector<int32_t> score(job.size());
  std::vector<uint32_t> pos(job.size()), len(job.size()), flags(job.size());
  std::vector<char> cx(job.size() * stride), cy(job.size() * stride);
  float device_us = 0.f;
  check(swb_align_batch(ctx_, blob.data(), offs.data(), job.size(), trigger->npiece, trigger->ratio, SWB_FLAG_CONSENSUS, score.data(), pos.data(), nullptr,
                        cx.data(), cy.data(), len.data(), stride, flags.data(), &device_us));
  swb_stats st{};
  swb_last_stats(ctx_, &st);
  for (size_t i = 0; i < job.size(); ++i) {
    Pending* p = job[i];
    if (flags[i] & SWB_RES_CONS_TRUNCATED) p->r = detail::run_one(p->mode, p->sc, p->x, p->y, p->npiece, p->ratio);   // a consensus longer than 2 * len + 64: alone, with room
    else {
      p->r.score = (float)score[i]; p->r.pos = pos[i];
      p->r.cx.assign(cx.data() + i * stride, len[i]); p->r.cy.assign(cy.data() + i * stride, len[i]);
    }
    // the drivers SUM getTimings()[0] over their reads (sw_solve_small.cpp:88-89): every object gets its share of the batch
    p->r.device_us = device_us / (float)job.size();
    p->pass1_us = st.pass1_us / (float)job.size();
    p->ready = true;
  }
}

template <class SMT>
class CUDASWAligner : public LocalAligner<SMT> {
 public:
  // the four constructors of smithwaterman.h:14-17
  CUDASWAligner(std::string_view x, std::string_view y) : sm_(x, y) { job_.init(SMT::mode, x, y, 0, 0.f); }
  CUDASWAligner(std::string_view x, std::string_view y, float gap) : sm_(x, y) { job_.sc.gap = gap; job_.init(SMT::mode, x, y, 0, 0.f); }
  CUDASWAligner(std::string_view x, std::string_view y, ScoringFn&& fn) : sm_(x, y) { job_.sc.is_fn = true; job_.sc.fn = std::move(fn); job_.init(SMT::mode, x, y, 0, 0.f); }
  CUDASWAligner(std::string_view x, std::string_view y, ScoringFn&& fn, float gap) : sm_(x, y) { job_.sc.is_fn = true; job_.sc.fn = std::move(fn); job_.sc.gap = gap; job_.init(SMT::mode, x, y, 0, 0.f); }

  float calculateScore() override { return job_.calculate(); }
  float getScore() const override { return job_.r.score; }
  unsigned int getPos() const override { return job_.r.pos; }
  std::string_view getConsensus_x() const override { return job_.r.cx; }
  std::string_view getConsensus_y() const override { return job_.r.cy; }
  const SMT& getSimilarity_matrix() const override {
    const detail::Scoring* sc = &job_.sc;
    sm_.apply_scoring_ = [sc](Context& c) { sc->apply(c, SMT::mode); };
    return sm_;
  }
  // [0] = microseconds spent computing the matrix (the drivers' GCUPS numerator, sw_solve_small.cpp:88-89): device time of
  // this object's share of its batch; [1] = the score-pass part of it (the reference: sum of its per-diagonal timers)
  TimingsVec getTimings() const override { return make_timings(job_.r.device_us, job_.pass1_us); }
  // public knobs of the reference's -DUSEOMP build that omp_sw_solve_small.cpp:165-171 sets and prints
  // (smithwaterman.h:37-41); the CUDA path has no use for them
  int sw_nthreads = 1, sw_finegrain_type = -1, sw_mt_simd = 0;

 private:
  mutable SMT sm_;
  Pending job_;
};

template <class SMT, class LAT = CUDASWAligner<SMT>>
class CUDAParallelLocalAligner : public ParallelLocalAligner<SMT, LAT> {
 public:
  // the four constructors of plocalaligner.h:9-12
  // The final alignment of the winning piece runs in SMT's arithmetic (LAT is only a type parameter here): the reference
  // instantiates <Skewed, SWAligner<Skewed>> and <plain, SWAligner<plain>> (plocalaligner.cpp:145-152 also lists mixed
  // pairs, which no driver uses); a mixed pair would silently change the score's range, so it does not compile.
  static_assert(std::is_same<LAT, CUDASWAligner<SMT>>::value, "CUDAParallelLocalAligner: LAT must be CUDASWAligner<SMT>");
  CUDAParallelLocalAligner(std::string_view x, std::string_view y, int npiece, float ratio) { job_.init(SMT::mode, x, y, npiece, ratio); }
  CUDAParallelLocalAligner(std::string_view x, std::string_view y, int npiece, float ratio, float gap) { job_.sc.gap = gap; job_.init(SMT::mode, x, y, npiece, ratio); }
  CUDAParallelLocalAligner(std::string_view x, std::string_view y, int npiece, float ratio, ScoringFn&& fn) { job_.sc.is_fn = true; job_.sc.fn = std::move(fn); job_.init(SMT::mode, x, y, npiece, ratio); }
  CUDAParallelLocalAligner(std::string_view x, std::string_view y, int npiece, float ratio, ScoringFn&& fn, float gap) { job_.sc.is_fn = true; job_.sc.fn = std::move(fn); job_.sc.gap = gap; job_.init(SMT::mode, x, y, npiece, ratio); }

  float calculateScore() override { return job_.calculate(); }
  float getScore() const override { return job_.r.score; }
  unsigned int getPos() const override { return job_.r.pos; }
  std::string_view getConsensus_x() const override { return job_.r.cx; }
  std::string_view getConsensus_y() const override { return job_.r.cy; }
  // plocalaligner.cpp:109-130: [0] = wall time of the piece loop, [1] = sum of the pieces' iterate times; here device
  // time of this object's share of its batch and its score-pass part
  TimingsVec getTimings() const override { return make_timings(job_.r.device_us, job_.pass1_us); }

 private:
  Pending job_;
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

// The same batch over several GPUs of one box behind the C ABI: one host thread and one swb_ctx per device, the batch
// partitioned over the devices, results merged on the host in input order.  This is the decomposition of the reference's
// MPI drivers — reads / database files divided over worker ranks, results funnelled to one writer
// (mpi_sw_solve_small.cpp:52-55,109-143; mpi_sw_solve_uniprot.cpp:65-72,95-138) — with host threads for ranks and no
// message passing: alignments are independent and the drivers end in a CSV.
class CUDAMultiGpuBatchAligner {
 public:
  enum Partition { BLOCK, BALANCED };   // BLOCK: contiguous floor(n / g) per device, the last takes the remainder
                                        // (mpi_sw_solve_small.cpp:52-55); BALANCED: by residues, longest first to the
                                        // least loaded device (ragged protein databases)
  CUDAMultiGpuBatchAligner(int mode, int n_gpus) : mode_(mode) {
    const int have = swb_device_count();
    if (have <= 0) throw Error(SWB_ERR_CUDA, "no CUDA device (libswb200 has no CPU fallback)");
    n_gpus_ = n_gpus <= 0 ? have : n_gpus;
    if (n_gpus_ > have) throw Error(SWB_ERR_ARG, "more GPUs requested than visible");
  }
  int gpus() const { return n_gpus_; }
  void set_scoring(float match, float mismatch, float gap) { sc_.is_fn = false; sc_.match = match; sc_.mismatch = mismatch; sc_.gap = gap; }
  void set_scoring(ScoringFn fn, float gap) { sc_.is_fn = true; sc_.fn = std::move(fn); sc_.gap = gap; }
  void set_reference(std::string_view y) { y_ = y; }

  static std::vector<std::vector<size_t>> partition(const std::vector<std::string_view>& xs, int g, Partition how) {
    std::vector<std::vector<size_t>> parts((size_t)g);
    if (how == BLOCK) {
      const size_t q = xs.size() / (size_t)g;
      for (int r = 0; r < g; ++r) for (size_t i = (size_t)r * q; i < (r + 1 == g ? xs.size() : (size_t)(r + 1) * q); ++i) parts[(size_t)r].push_back(i);
      return parts;
    }
    std::vector<size_t> order(xs.size());
    for (size_t i = 0; i < xs.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return xs[a].size() > xs[b].size(); });
    std::vector<unsigned long long> load((size_t)g, 0ull);
    for (size_t i : order) {
      size_t r = 0;
      for (size_t k = 1; k < (size_t)g; ++k) if (load[k] < load[r]) r = k;
      parts[r].push_back(i); load[r] += xs[i].size();
    }
    for (auto& p : parts) std::sort(p.begin(), p.end());
    return parts;
  }

  CUDABatchAligner::Out align(const std::vector<std::string_view>& xs, int npiece = 0, float ratio = 0.f, bool consensus = true, Partition how = BLOCK) {
    const auto parts = partition(xs, n_gpus_, how);
    size_t maxlen = 0;
    for (auto& x : xs) maxlen = std::max(maxlen, x.size());
    CUDABatchAligner::Out o;
    o.stride = consensus ? 2 * maxlen + 64 : 0;
    o.score.resize(xs.size()); o.pos.resize(xs.size()); o.len.resize(xs.size()); o.flags.resize(xs.size());
    if (consensus) { o.cx.resize(xs.size() * o.stride); o.cy.resize(xs.size() * o.stride); }
    std::vector<float> dev_us((size_t)n_gpus_, 0.f);
    std::vector<std::string> errors((size_t)n_gpus_);
    std::vector<std::thread> th;
    for (int r = 0; r < n_gpus_; ++r)
      th.emplace_back([&, r]() {
        try {
          const auto& idx = parts[(size_t)r];
          if (idx.empty()) return;
          Context::instance().set_device(r);
          CUDABatchAligner ba(mode_);
          if (sc_.is_fn) ba.set_scoring(sc_.fn, sc_.gap); else ba.set_scoring(sc_.match, sc_.mismatch, sc_.gap);
          ba.set_reference(y_);
          std::vector<std::string_view> mine;
          for (size_t i : idx) mine.push_back(xs[i]);
          CUDABatchAligner::Out part = ba.align(mine, npiece, ratio, consensus);
          for (size_t k = 0; k < idx.size(); ++k) {
            const size_t i = idx[k];
            o.score[i] = part.score[k]; o.pos[i] = part.pos[k]; o.len[i] = part.len[k]; o.flags[i] = part.flags[k];
            if (consensus) {
              const size_t nb = std::min<size_t>(part.len[k], std::min(part.stride, o.stride));
              std::memcpy(o.cx.data() + i * o.stride, part.cx.data() + k * part.stride, nb);
              std::memcpy(o.cy.data() + i * o.stride, part.cy.data() + k * part.stride, nb);
            }
          }
          dev_us[(size_t)r] = part.device_us;
        } catch (const std::exception& e) { errors[(size_t)r] = e.what(); }
      });
    for (auto& t : th) t.join();
    for (int r = 0; r < n_gpus_; ++r) if (!errors[(size_t)r].empty()) throw Error(SWB_ERR_CUDA, "GPU " + std::to_string(r) + ": " + errors[(size_t)r]);
    o.device_us = *std::max_element(dev_us.begin(), dev_us.end());   // the devices run concurrently
    return o;
  }

 private:
  int mode_, n_gpus_ = 1;
  detail::Scoring sc_;
  std::string_view y_;
};

}  // namespace swb
