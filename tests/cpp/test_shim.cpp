// GPU test of the C++ shims: the reference's own unit tests (test/test_localaligner.cpp:8-27,53-59) with the
// aligner type swapped, plus the chunked aligner and the batched entry point.  Prints "SHIM OK" on success.
#include <cstdio>
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
  std::printf("SHIM OK\n");
  return 0;
}
