// The reference's OWN unit test file (test/test_localaligner.cpp: score 13 / pos 2 / consensus "CAGTTG" / "CA-TTG"),
// compiled UNMODIFIED against the reference's own headers and its vendored googletest, with only the aligner type
// swapped: SWAligner<Similarity_Matrix_Skewed> -> swb::CUDASWAligner<swb::CUDA_Similarity_Matrix_Skewed>.
// This is the -DSWB_WITH_REFERENCE_HEADERS mode INTEGRATION.md §1 describes: the shims derive from the reference's
// LocalAligner (src/aligner/localaligner.h:7-17) and getTimings() returns Eigen::VectorXf.
//
// Built by tests/cpp/build_ref_tests.py in the build container only (it needs /root/reference: the test file, the
// headers, Eigen and googletest from the reference's cmake/*.zip); the binary travels to the GPU box.
#define SWB_WITH_REFERENCE_HEADERS 1
#include <gtest/gtest.h>
#include <memory>
#include "localaligner.h"
#include "smithwaterman.h"
#include "similaritymatrix.h"
#include "../../parallel-genomeseq_b200/cpp/cuda_aligner.h"

// the type swap (the reference headers above are already included, their include guards keep them out of the rest)
#define SWAligner swb::CUDASWAligner
#define Similarity_Matrix_Skewed swb::CUDA_Similarity_Matrix_Skewed
#include REF_TEST_FILE

int main(int argc, char** argv) {
  testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
