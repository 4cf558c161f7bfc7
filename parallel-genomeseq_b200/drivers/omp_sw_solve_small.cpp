// omp_sw_solve_small — B200 replacement of the reference's fine-grain benchmark driver src/omp_sw_solve_small.cpp
// (the long-pair workload: 10 kbp reads against a long reference).
//
// Same command line and the same timing-CSV schema, so py/eval.py --option ompfg can plot GPU rows beside the
// Leonhard data (omp_sw_solve_small.cpp:66-73,233-239):
//   omp_sw_solve_small solve_small <n_reads> <n_threads> <finegrain_type> <timing.csv> <ref file> <reads.csv> [align_out.csv]
//   timing.csv (appended; header written when the file is new):
//       n_reads,n_threads,finegrain_type,avg_t_calcscore,avg_t_adread,avg_t_adisum         (microseconds per read)
//   align_out.csv (default data/ompfg_align_output.csv): "<input line>, <pos_pred>, <score>"   (:150-156,199-201)
// The reference file is read whole, every line concatenated (fa_file_has_header = 0, :88-104); the first n_reads
// reads of the CSV (field 2) are aligned.  n_threads is only echoed (the wavefront parallelism is the GPU's);
// finegrain_type selects the arithmetic like the reference's build flavours: -1 = Similarity_Matrix_Skewed
// (the -DMTSIMD build, :164), anything else = Similarity_Matrix (EXACT, :167).
// Difference: all reads go through one batched call (long reads are cut into row strips that run concurrently);
// avg_t_calcscore is the whole call's wall time per read, avg_t_adread / avg_t_adisum the device time per read.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../cpp/cuda_aligner.h"

int main(int argc, char** argv) {
  if (argc < 8) { std::cerr << "usage: omp_sw_solve_small solve_small n_reads n_threads finegrain_type timing.csv ref reads.csv [align_out.csv]" << std::endl; return 255; }
  const std::string argv1 = argv[1];
  const int arg_nreads = std::atoi(argv[2]), arg_nthreads = std::atoi(argv[3]), arg_finegrain_type = std::atoi(argv[4]);
  const std::string timing_file_path = argv[5], fa_file_path = argv[6], input_file_path = argv[7];
  const std::string align_output_file_path = argc > 8 ? argv[8] : "data/ompfg_align_output.csv";
  if (argv1 != "solve_small") { std::cout << "Hello omp (nothing to do for section '" << argv1 << "')" << std::endl; return 0; }
  std::cout << "Hello sw_solve_small" << std::endl;

  std::ifstream fa(fa_file_path);
  if (!fa) { std::cerr << "cannot open " << fa_file_path << std::endl; return 2; }
  std::string fa_string, line;
  while (std::getline(fa, line)) fa_string += line;
  std::ifstream in(input_file_path);
  if (!in) { std::cerr << "cannot open " << input_file_path << std::endl; return 2; }
  std::string header;
  std::vector<std::string> lines, reads;
  for (int i = 0; std::getline(in, line) && i <= arg_nreads; ++i) {
    if (i == 0) { header = line; continue; }
    if (line.empty()) continue;
    size_t a = line.find(','), b = a == std::string::npos ? a : line.find(',', a + 1), c = b == std::string::npos ? b : line.find(',', b + 1);
    if (b == std::string::npos) { std::cerr << "malformed line " << i << std::endl; return 2; }
    lines.push_back(line);
    reads.push_back(line.substr(b + 1, c == std::string::npos ? std::string::npos : c - b - 1));
  }
  if (reads.empty()) { std::cerr << "no reads" << std::endl; return 2; }
  std::vector<std::string_view> views(reads.begin(), reads.end());

  swb::CUDABatchAligner aligner(arg_finegrain_type == -1 ? SWB_MODE_SAT_U8 : SWB_MODE_EXACT);
  aligner.set_reference(fa_string);
  swb::CUDABatchAligner::Out out;
  const auto start = std::chrono::high_resolution_clock::now();
  try { out = aligner.align(views, 0, 0.f, /*consensus=*/false); }
  catch (const swb::Error& e) { std::cerr << "alignment failed: " << e.what() << std::endl; return 1; }
  const auto end = std::chrono::high_resolution_clock::now();
  const double wall_us = (double)std::chrono::duration_cast<std::chrono::microseconds>(end - start).count();

  std::ofstream align_output(align_output_file_path);
  align_output << header << ",pos_pred,score\n";
  for (size_t i = 0; i < lines.size(); ++i) align_output << lines[i] << ", " << out.pos[i] << ", " << (float)out.score[i] << "\n";
  std::cout << "Done, align output file see: " << align_output_file_path << std::endl;

  const double n = (double)reads.size();
  const bool exists = std::ifstream(timing_file_path).good();
  std::ofstream timing(timing_file_path, std::ios::out | std::ios::app);
  if (!exists) timing << "n_reads,n_threads,finegrain_type,avg_t_calcscore,avg_t_adread,avg_t_adisum\n";
  timing << (float)arg_nreads << "," << (float)arg_nthreads << "," << (float)arg_finegrain_type << "," << (float)(wall_us / n) << ","
         << (float)(out.device_us / n) << "," << (float)(out.device_us / n) << "\n";
  unsigned long long cells = 0;
  for (auto& r : reads) cells += (unsigned long long)r.size() * fa_string.size();
  std::cout << "device time per read: " << out.device_us / n << "us, GCUPS: " << cells / (double)out.device_us * 1e-3 << std::endl;
  return 0;
}
