// GPU test of the C++ shims: the reference's own unit tests (test/test_localaligner.cpp:8-27,53-59) with the
// aligner type swapped, plus the chunked aligner and the batched entry point.  Prints "SHIM OK" on success.
#include <chrono>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../parallel-genomeseq_b200/cpp/cuda_aligner.h"

#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main() {
  using namespace swb;
  const std::string sequence_x = "GGTTGACTA", sequence_y = "TGTTACGG";
  {
    CUDASWAligner<CUDA_Similarity_Matrix_Skewed> la(sequence_x, sequence_y);
    EXPECT(la.getScore() == -1.f);
    EXPECT(la.calculateScore() == 13.f);
    EXPECT(la.getScore() == 13.f);
    EXPECT(la.getPos() == 2u);
    EXPECT(la.getConsensus_x() == "CAGTTG");
    EXPECT(la.getConsensus_y() == "CA-TTG");
    EXPECT(la.getTimings()[0] > 0.f);
    swb::LocalAligner<CUDA_Similarity_Matrix_Skewed>* base = &la;     // usable through the abstract interface
    EXPECT(base->getPos() == 2u);
  }
  {
    CUDASWAligner<CUDA_Similarity_Matrix> la(sequence_x, sequence_y, [](const char& a, const char& b) { return a == b ? 3.0f : -3.0f; }, 2.0f);
    EXPECT(la.calculateScore() == 13.f);
    EXPECT(la.getPos() == 2u && la.getConsensus_x() == "CAGTTG" && la.getConsensus_y() == "CA-TTG");
  }
  {
    // chunked: 2 pieces over a reference made of filler + the Wikipedia pair
    std::string ref = std::string(300, 'C') + sequence_y + std::string(300, 'C');
    CUDAParallelLocalAligner<CUDA_Similarity_Matrix_Skewed> pa(sequence_x, ref, 2, 2.0f);
    EXPECT(pa.calculateScore() == 13.f);
    EXPECT(pa.getPos() == 302u);
    EXPECT(pa.getConsensus_x() == "CAGTTG");
  }
  {
    CUDABatchAligner ba(SWB_MODE_SAT_U8);
    ba.set_reference(sequence_y);
    std::vector<std::string_view> xs = {sequence_x, "TTAC", sequence_x};
    auto out = ba.align(xs);
    EXPECT(out.score[0] == 13 && out.score[2] == 13 && out.pos[0] == 2 && out.score[1] == 12);
    EXPECT(out.consensus_x(0) == "CAGTTG" && out.consensus_y(2) == "CA-TTG");
    EXPECT(out.device_us > 0.f);
  }
  {
    bool threw = false;
    try { CUDAParallelLocalAligner<CUDA_Similarity_Matrix_Skewed> pa(std::string(100, 'A'), std::string(160, 'A'), 4, 2.0f); pa.calculateScore(); }
    catch (const swb::Error& e) { threw = e.code == SWB_ERR_RANGE; }
    EXPECT(threw);    // the reference aborts on this assert (plocalaligner.cpp:52)
  }
  {
    // test/test_skewedmatrix.cpp:39-66: every cell of the two SMTs agrees (through getSimilarity_matrix())
    const std::string a = "GGTTGACTA", b = "TGTTACG";
    CUDASWAligner<CUDA_Similarity_Matrix_Skewed> s1(a, b);
    CUDASWAligner<CUDA_Similarity_Matrix> s2(a, b);
    s1.calculateScore(); s2.calculateScore();
    for (size_t i = 0; i <= a.size(); ++i)
      for (size_t j = 0; j <= b.size(); ++j) EXPECT(s1.getSimilarity_matrix()(i, j) == s2.getSimilarity_matrix()(i, j));
    EXPECT(s1.getSimilarity_matrix()(7, 6) == 13.f);
  }
  {
    // Lazy batching: the reference's per-read loop shape split in two (construct every aligner, then query them) runs
    // as ONE batch; results equal the batched entry point's, and it must not be slower than 2x that entry point.
    std::string ref;
    unsigned long long z = 88172645463325252ull;
    auto rnd = [&]() { z ^= z << 13; z ^= z >> 7; z ^= z << 17; return z; };
    for (int i = 0; i < 5000; ++i) ref += "ACGT"[rnd() & 3];
    std::vector<std::string> reads;
    for (int r = 0; r < 600; ++r) {
      std::string x = ref.substr(rnd() % (ref.size() - 125), 125);
      for (auto& ch : x) if (rnd() % 50 == 0) ch = "ACGT"[rnd() & 3];
      reads.push_back(x);
    }
    std::vector<std::string_view> views(reads.begin(), reads.end());
    CUDABatchAligner ba(SWB_MODE_SAT_U8);
    ba.set_reference(ref);
    ba.align(views);                                           // warm-up (buffers, kernels)
    auto t0 = std::chrono::steady_clock::now();
    auto out = ba.align(views);
    auto t1 = std::chrono::steady_clock::now();
    std::vector<std::unique_ptr<CUDASWAligner<CUDA_Similarity_Matrix_Skewed>>> las;
    for (auto& x : reads) las.push_back(std::make_unique<CUDASWAligner<CUDA_Similarity_Matrix_Skewed>>(x, ref));
    float us_sum = 0.f;
    for (size_t i = 0; i < las.size(); ++i) {
      EXPECT(las[i]->calculateScore() == (float)out.score[i]);
      EXPECT(las[i]->getPos() == out.pos[i]);
      EXPECT(las[i]->getConsensus_x() == out.consensus_x(i) && las[i]->getConsensus_y() == out.consensus_y(i));
      us_sum += las[i]->getTimings()[0];
      EXPECT(las[i]->getTimings()[1] > 0.f && las[i]->getTimings()[1] <= las[i]->getTimings()[0]);
    }
    auto t2 = std::chrono::steady_clock::now();
    const double batch_ms = std::chrono::duration<double, std::milli>(t1 - t0).count(), lazy_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    std::printf("LAZY n=%zu batch_ms=%.3f lazy_ms=%.3f device_us_sum=%.1f\n", reads.size(), batch_ms, lazy_ms, us_sum);
    EXPECT(lazy_ms < 2.0 * batch_ms + 2.0);
    // objects with another scoring or chunking are separate jobs; destroying a pending object is fine
    CUDASWAligner<CUDA_Similarity_Matrix> other(reads[0], ref);
    { CUDASWAligner<CUDA_Similarity_Matrix_Skewed> dropped(reads[1], ref); }
    CUDAParallelLocalAligner<CUDA_Similarity_Matrix_Skewed> chunked(reads[2], ref, 4, 2.0f);
    CUDASWAligner<CUDA_Similarity_Matrix_Skewed> plain(reads[2], ref);
    EXPECT(plain.calculateScore() == (float)out.score[2] && other.getScore() == -1.f && chunked.getScore() == -1.f);
    EXPECT(chunked.calculateScore() == 255.f && other.calculateScore() > 255.f);
  }
  {
    // several GPUs behind the C ABI (one host thread and one context per device): same results in input order
    const std::string ref = std::string(40, 'C') + sequence_y + std::string(40, 'G') + sequence_y + std::string(40, 'C');
    std::vector<std::string> xs_s = {sequence_x, "TTAC", sequence_x, "GACTA", "GGTTGACTAGGTTGACTA", "TGTTACGG", "A"};
    std::vector<std::string_view> xs(xs_s.begin(), xs_s.end());
    CUDABatchAligner one(SWB_MODE_SAT_U8);
    one.set_reference(ref);
    auto want = one.align(xs);
    for (int g = 1; g <= swb_device_count() && g <= 4; ++g)
      for (auto how : {CUDAMultiGpuBatchAligner::BLOCK, CUDAMultiGpuBatchAligner::BALANCED}) {
        CUDAMultiGpuBatchAligner many(SWB_MODE_SAT_U8, g);
        many.set_reference(ref);
        auto got = many.align(xs, 0, 0.f, true, how);
        for (size_t i = 0; i < xs.size(); ++i) {
          EXPECT(got.score[i] == want.score[i] && got.pos[i] == want.pos[i]);
          EXPECT(got.consensus_x(i) == want.consensus_x(i) && got.consensus_y(i) == want.consensus_y(i));
        }
      }
    auto parts = CUDAMultiGpuBatchAligner::partition(xs, 3, CUDAMultiGpuBatchAligner::BLOCK);
    EXPECT(parts[0].size() == 2 && parts[1].size() == 2 && parts[2].size() == 3 && parts[2].back() == 6);   // mpi_sw_solve_small.cpp:52-55
    std::printf("MULTI-GPU OK (%d device(s) visible)\n", swb_device_count());
  }
  std::printf("SHIM OK\n");
  return 0;
}
