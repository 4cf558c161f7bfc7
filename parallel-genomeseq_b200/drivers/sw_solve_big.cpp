// sw_solve_big — batched B200 replacement of the reference benchmark driver src/sw_solve_big.cpp.
//
// Same inputs and the same [INFO] report lines:
//   * reference: ONE raw line, no FASTA header                                  (sw_solve_big.cpp:33-37)
//   * reads CSV "index,QNAME,SEQ,POS" with one header line, field 2 = read      (sw_solve_big.cpp:53-67)
//   * nrepeat runs, the minimum time is kept                                    (sw_solve_big.cpp:82-87)
//   * "[INFO] Average SW iter_ad_read times: <s>s, GCUPS:<g>, GCPUS per iteration: <g>" and
//     "[INFO] GCUPS avg:<mean>, GCUPS std:<std>"                                (sw_solve_big.cpp:99-106)
// Differences: all reads are aligned by ONE batched call per repeat (inputs stay resident in HBM between
// repeats: swb_batch_stage once, swb_batch_run nrepeat times), the time is DEVICE time of all kernels of the
// batch, and mean/std are taken over the repeats (the reference takes them over reads).
//   sw_solve_big [npiece nrepeat] [--fa FILE] [--reads FILE]
//     npiece > 0 reproduces the -DUSEOMP build: OMPParallelLocalAligner(read, ref, npiece*2, 2.0)  (:78)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/swb200.h"

static int die(swb_ctx* ctx, const char* what) {
  std::cerr << what << ": " << (ctx ? swb_last_error(ctx) : "no context") << std::endl;
  return 1;
}

int main(int argc, char** argv) {
  int npiece = 0, nrepeat = 1;
  std::string fa_file_path = "data/custom_ref_1.fa", input_file_path = "data/custom_reads_1.csv";
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--fa") && i + 1 < argc) fa_file_path = argv[++i];
    else if (!std::strcmp(argv[i], "--reads") && i + 1 < argc) input_file_path = argv[++i];
    else pos.push_back(argv[i]);
  }
  if (pos.size() >= 2) { npiece = std::atoi(pos[0].c_str()); nrepeat = std::max(1, std::atoi(pos[1].c_str())); }
  std::cout << "[INFO] npiece: " << npiece << ", nrepeat:" << nrepeat << std::endl;

  std::ifstream fa(fa_file_path);
  if (!fa) { std::cerr << "cannot open " << fa_file_path << std::endl; return 2; }
  std::string fa_string;
  std::getline(fa, fa_string);
  std::ifstream in(input_file_path);
  if (!in) { std::cerr << "cannot open " << input_file_path << std::endl; return 2; }
  std::string line, blob;
  std::vector<uint64_t> offs(1, 0);
  unsigned long long num_cells = 0;
  for (int i = 0; std::getline(in, line); ++i) {
    if (i == 0 || line.empty()) continue;
    size_t a = line.find(','), b = a == std::string::npos ? a : line.find(',', a + 1), c = b == std::string::npos ? b : line.find(',', b + 1);
    if (b == std::string::npos) { std::cerr << "malformed line " << i << std::endl; return 2; }
    const std::string read = line.substr(b + 1, c == std::string::npos ? std::string::npos : c - b - 1);
    blob += read; offs.push_back(blob.size());
    num_cells += (unsigned long long)read.size() * fa_string.size();
  }
  const size_t n_reads = offs.size() - 1;
  if (!n_reads) { std::cerr << "no reads" << std::endl; return 2; }
  std::cout << "[INFO] Estimated Memory consumption of the reference's matrices " << (double)num_cells * 1e-9 << "GB (never materialised here)" << std::endl;

  swb_ctx* ctx = nullptr;
  if (swb_create(0, &ctx) != SWB_OK) return die(nullptr, "swb_create (a CUDA device is required)");
  if (swb_set_scoring_match(ctx, SWB_MODE_SAT_U8, 3.f, -3.f, 2.f)) return die(ctx, "scoring");
  if (swb_set_reference(ctx, fa_string.data(), fa_string.size())) return die(ctx, "reference");
  if (swb_batch_stage(ctx, blob.data(), offs.data(), n_reads, npiece > 0 ? npiece * 2 : 0, 2.0f, 0, 0)) return die(ctx, "stage");
  std::vector<double> gcups;
  double time_min = 9e20;
  for (int j = 0; j < nrepeat; ++j) {
    float us = 0.f;
    if (swb_batch_run(ctx, &us)) return die(ctx, "run");
    time_min = std::min(time_min, (double)us);
    gcups.push_back(num_cells / (double)us * 1e-3);
  }
  std::vector<int32_t> score(n_reads);
  std::vector<uint32_t> posv(n_reads);
  if (swb_batch_fetch(ctx, score.data(), posv.data(), nullptr, nullptr, nullptr, nullptr, nullptr)) return die(ctx, "fetch");
  double mean = 0, var = 0;
  for (double g : gcups) mean += g;
  mean /= gcups.size();
  for (double g : gcups) var += (g - mean) * (g - mean);
  const double GCUPS = num_cells / time_min * 1e-3;
  std::cout << "[INFO] Average SW iter_ad_read times: " << time_min * 1e-6 / (double)n_reads << "s, GCUPS:" << GCUPS
            << ", GCPUS per iteration: " << GCUPS << std::endl;
  std::cout << "[INFO] GCUPS avg:" << mean << ", GCUPS std:" << std::sqrt(var / gcups.size()) << std::endl;
  for (double g : gcups) std::cout << g << " ";
  std::cout << std::endl;
  std::cout << "[INFO] first read: pos " << posv[0] << " score " << score[0] << std::endl;
  swb_destroy(ctx);
  return 0;
}
