// sw_solve_small — batched B200 replacement of the reference driver src/sw_solve_small.cpp.
//
// Same inputs, same output file format, same final GCUPS line:
//   * reference FASTA: header line skipped, remaining lines concatenated      (sw_solve_small.cpp:25-30)
//   * reads CSV "index,QNAME,SEQ,POS" with one header line, field 2 = read    (sw_solve_small.cpp:56-67)
//   * output: header + ",pos_pred,score", rows "<input line>, <pos_pred>, <score>"  (:72-74,91-93)
//   * "Average SW iter_ad_read times: <us>us, GCUP:<gcups>" with cells = sum len(read)*len(ref) (:88-89,102-106)
// Difference: the per-read loop "construct aligner -> calculateScore -> getPos" becomes ONE batched call.
//   sw_solve_small [fa] [reads.csv] [out.csv] [--npiece N --ratio R] [--float] [--gpus G]
//     --npiece 17 --ratio 2.0 reproduces the reference's -DUSEOMP build (OMPParallelLocalAligner, :82);
//     --float selects Similarity_Matrix (EXACT) arithmetic instead of Similarity_Matrix_Skewed (SAT_U8);
//     --gpus G divides the reads over G GPUs (0 = all) like mpi_sw_solve_small.cpp:52-55 divides them over ranks.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../cpp/cuda_aligner.h"

int main(int argc, char** argv) {
  std::string fa_file_path = "data/data_small/genome.chr22.5K.fa";
  std::string input_file_path = "data/data_small_ground_truth.csv";
  std::string output_file_path = "data/align_output.csv";
  int npiece = 0; float ratio = 2.0f; int mode = SWB_MODE_SAT_U8; int gpus = 1;
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--npiece") && i + 1 < argc) npiece = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--ratio") && i + 1 < argc) ratio = (float)std::atof(argv[++i]);
    else if (!std::strcmp(argv[i], "--float")) mode = SWB_MODE_EXACT;
    else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else pos.push_back(argv[i]);
  }
  if (pos.size() > 0) fa_file_path = pos[0];
  if (pos.size() > 1) input_file_path = pos[1];
  if (pos.size() > 2) output_file_path = pos[2];
  std::cout << "Hello sw_solve_small" << std::endl;

  std::ifstream fa(fa_file_path);
  if (!fa) { std::cerr << "cannot open " << fa_file_path << std::endl; return 2; }
  std::string fa_string, line;
  for (int i = 0; std::getline(fa, line); ++i) if (i > 0) fa_string += line;

  std::ifstream in(input_file_path);
  if (!in) { std::cerr << "cannot open " << input_file_path << std::endl; return 2; }
  std::vector<std::string> lines, reads;
  std::string header;
  for (int i = 0; std::getline(in, line); ++i) {
    if (i == 0) { header = line; continue; }
    if (line.empty()) continue;
    size_t a = line.find(','), b = a == std::string::npos ? a : line.find(',', a + 1), c = b == std::string::npos ? b : line.find(',', b + 1);
    if (b == std::string::npos) { std::cerr << "malformed line " << i << std::endl; return 2; }
    lines.push_back(line);
    reads.push_back(line.substr(b + 1, c == std::string::npos ? std::string::npos : c - b - 1));
  }
  std::vector<std::string_view> views(reads.begin(), reads.end());

  swb::CUDABatchAligner::Out out;
  try {
    if (gpus == 1) {
      swb::CUDABatchAligner aligner(mode);
      aligner.set_reference(fa_string);
      out = aligner.align(views, npiece, ratio, /*consensus=*/false);
    } else {
      swb::CUDAMultiGpuBatchAligner aligner(mode, gpus);
      aligner.set_reference(fa_string);
      out = aligner.align(views, npiece, ratio, /*consensus=*/false, swb::CUDAMultiGpuBatchAligner::BLOCK);
      std::cout << "reads divided over " << aligner.gpus() << " GPUs" << std::endl;
    }
  } catch (const swb::Error& e) { std::cerr << "alignment failed: " << e.what() << std::endl; return 1; }

  std::ofstream align_output(output_file_path);
  align_output << header << ",pos_pred,score\n";
  unsigned long long num_cells = 0;
  for (size_t i = 0; i < lines.size(); ++i) {
    align_output << lines[i] << ", " << out.pos[i] << ", " << (float)out.score[i] << "\n";
    num_cells += (unsigned long long)reads[i].size() * fa_string.size();
  }
  const double time_us = out.device_us;                     // device time of the whole batch (all kernels)
  const double GCUPs = num_cells / time_us * 1e-3;
  std::cout << "Average SW iter_ad_read times: " << time_us / (double)(lines.size() + 1) << "us, GCUP:" << GCUPs << std::endl;
  std::cout << "Done, output file see: " << output_file_path << std::endl;
  return 0;
}
