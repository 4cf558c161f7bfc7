// sw_search_uniprot — batched single-GPU replacement of the reference driver src/mpi_sw_solve_uniprot.cpp.
//
// What the reference does per worker rank (mpi_sw_solve_uniprot.cpp:95-138): for every database protein file,
// SWAligner<Similarity_Matrix>(protein, query).calculateScore() — x = DB protein, y = query, exact (f32)
// arithmetic, default scoring — and a record {char read[126], int pos_pred, double score} sent to the writer
// rank, which prints "read,pos_pred,score" rows ("%.126s" of the protein, :132; header :151-156).
// Here: the database is ONE multi-FASTA file (the 561 356 one-protein files become one blob +
// offsets), all proteins go through ONE batched call, and the same CSV is written by the same process.
// Several GPUs (--gpus G, 0 = all): the database is partitioned by residues over G GPUs, one host thread each, and the
// rows are written in database order (mpi_sw_solve_uniprot.cpp:65-72 gives every rank a block of files).
//   sw_search_uniprot QUERY.fasta DB.fasta|DB.swbdb OUT.csv [--blosum62 GAP] [--first N --count M] [--gpus G]
//   sw_search_uniprot --pack DB.fasta DB.swbdb        convert a multi-FASTA database into the packed 5-bit blob once
//     A database whose name ends in .swbdb is the packed blob (cpp/packed_db.h): no FASTA parsing per run; rows are written
//     in the ORIGINAL database order.
//     default scoring = the reference's (a == b ? 3 : -3, gap 2); --blosum62 10 tabulates BLOSUM62 through the
//     callback constructor surface (smithwaterman.h:16-17) with linear gap 10.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include <chrono>

#include "../cpp/cuda_aligner.h"
#include "../cpp/packed_db.h"

static const char* kOrder = "ARNDCQEGHILKMFPSTWYVBZX*";
static const int kBlosum62[24][24] = {
  { 4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0,-2,-1, 0,-4}, {-1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3,-1, 0,-1,-4},
  {-2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3, 3, 0,-1,-4}, {-2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3, 4, 1,-1,-4},
  { 0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1,-3,-3,-2,-4}, {-1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2, 0, 3,-1,-4},
  {-1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1,-4}, { 0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3,-1,-2,-1,-4},
  {-2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3, 0, 0,-1,-4}, {-1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3,-3,-3,-1,-4},
  {-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1,-4,-3,-1,-4}, {-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2, 0, 1,-1,-4},
  {-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1,-3,-1,-1,-4}, {-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1,-3,-3,-1,-4},
  {-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2,-2,-1,-2,-4}, { 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2, 0, 0, 0,-4},
  { 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0,-1,-1, 0,-4}, {-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3,-4,-3,-2,-4},
  {-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1,-3,-2,-1,-4}, { 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4,-3,-2,-1,-4},
  {-2,-1, 3, 4,-3, 0, 1,-1, 0,-3,-4, 0,-3,-3,-2, 0,-1,-4,-3,-3, 4, 1,-1,-4}, {-1, 0, 0, 1,-3, 3, 4,-2, 0,-3,-3, 1,-1,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1,-4},
  { 0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2, 0, 0,-2,-1,-1,-1,-1,-1,-4}, {-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4,-4, 1}};

static bool read_fasta_records(const std::string& path, std::vector<std::string>* seqs) {
  std::ifstream f(path);
  if (!f) return false;
  std::string line, cur;
  bool open = false;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (!line.empty() && line[0] == '>') { if (open) seqs->push_back(cur); cur.clear(); open = true; continue; }
    if (!open) { open = true; }         // header-less file: one record
    cur += line;
  }
  if (open) seqs->push_back(cur);
  return true;
}

int main(int argc, char** argv) {
  std::vector<std::string> pos;
  bool blosum = false, pack = false; float gap = 2.f; size_t first = 0, count = (size_t)-1; int gpus = 1;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--blosum62") && i + 1 < argc) { blosum = true; gap = (float)std::atof(argv[++i]); }
    else if (!std::strcmp(argv[i], "--first") && i + 1 < argc) first = (size_t)std::atoll(argv[++i]);
    else if (!std::strcmp(argv[i], "--count") && i + 1 < argc) count = (size_t)std::atoll(argv[++i]);
    else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--pack")) pack = true;
    else pos.push_back(argv[i]);
  }
  if (pack) {
    if (pos.size() < 2) { std::cerr << "usage: sw_search_uniprot --pack DB.fasta DB.swbdb" << std::endl; return 2; }
    std::vector<std::string> db;
    if (!read_fasta_records(pos[0], &db) || db.empty()) { std::cerr << "cannot read database " << pos[0] << std::endl; return 2; }
    if (!swb::write_packed_db(db, pos[1])) { std::cerr << "cannot write " << pos[1] << std::endl; return 2; }
    std::cout << "Packed " << db.size() << " entries into " << pos[1] << std::endl;
    return 0;
  }
  if (pos.size() < 3) { std::cerr << "usage: sw_search_uniprot QUERY.fasta DB.fasta|DB.swbdb OUT.csv [--blosum62 GAP] [--first N --count M] [--gpus G]" << std::endl; return 2; }
  std::vector<std::string> q, db;
  if (!read_fasta_records(pos[0], &q) || q.empty()) { std::cerr << "cannot read query " << pos[0] << std::endl; return 2; }
  const auto t_stage0 = std::chrono::steady_clock::now();
  const bool packed = pos[1].size() > 6 && pos[1].compare(pos[1].size() - 6, 6, ".swbdb") == 0;
  swb::PackedDb pdb;
  std::vector<std::string_view> all;           // every database entry, in the order it is stored
  std::vector<size_t> orig;                    // its index in the original database
  if (packed) {
    std::string err;
    if (!swb::load_packed_db(pos[1], &pdb, &err)) { std::cerr << err << std::endl; return 2; }
    for (size_t i = 0; i < pdb.size(); ++i) { all.emplace_back(pdb.blob.data() + pdb.offsets[i], pdb.offsets[i + 1] - pdb.offsets[i]); orig.push_back(pdb.orig[i]); }
  } else {
    if (!read_fasta_records(pos[1], &db) || db.empty()) { std::cerr << "cannot read database " << pos[1] << std::endl; return 2; }
    for (size_t i = 0; i < db.size(); ++i) { all.emplace_back(db[i]); orig.push_back(i); }
  }
  const double stage_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_stage0).count();
  const std::string& fa_string = q[0];
  // --first / --count select by ORIGINAL index
  const size_t ndb = all.size();
  if (first > ndb) first = ndb;
  if (count > ndb - first) count = ndb - first;
  std::vector<std::string_view> xs;
  std::vector<size_t> row_of;                  // output row (original order) of every aligned entry
  for (size_t i = 0; i < ndb; ++i) if (orig[i] >= first && orig[i] < first + count) { xs.push_back(all[i]); row_of.push_back(orig[i] - first); }

  auto blosum_fn = [](const char& a, const char& b) -> float {
    const char* pa = std::strchr(kOrder, a); const char* pb = std::strchr(kOrder, b);
    if (!pa || !pb || !a || !b) return -4.f;
    return (float)kBlosum62[pa - kOrder][pb - kOrder];
  };
  swb::CUDABatchAligner::Out out;
  try {
    if (gpus == 1) {
      swb::CUDABatchAligner aligner(SWB_MODE_EXACT);
      aligner.set_reference(fa_string);
      if (blosum) aligner.set_scoring(blosum_fn, gap); else aligner.set_scoring(3.f, -3.f, 2.f);
      out = aligner.align(xs, 0, 0.f, /*consensus=*/false);
    } else {
      swb::CUDAMultiGpuBatchAligner aligner(SWB_MODE_EXACT, gpus);
      aligner.set_reference(fa_string);
      if (blosum) aligner.set_scoring(blosum_fn, gap); else aligner.set_scoring(3.f, -3.f, 2.f);
      out = aligner.align(xs, 0, 0.f, /*consensus=*/false, swb::CUDAMultiGpuBatchAligner::BALANCED);
      std::cout << "database partitioned by residues over " << aligner.gpus() << " GPUs" << std::endl;
    }
  } catch (const swb::Error& e) { std::cerr << "search failed: " << e.what() << std::endl; return 1; }

  std::ofstream csv(pos[2]);
  csv << "read,pos_pred,score\n";
  unsigned long long cells = 0;
  char buff[127];
  std::vector<size_t> at(xs.size());           // aligned entry of every output row
  for (size_t i = 0; i < xs.size(); ++i) at[row_of[i]] = i;
  for (size_t r = 0; r < xs.size(); ++r) {
    const size_t i = at[r];
    std::snprintf(buff, sizeof buff, "%.126s", std::string(xs[i]).c_str());
    csv << buff << ", " << (int)out.pos[i] << ", " << (double)out.score[i] << "\n";
    cells += (unsigned long long)xs[i].size() * fa_string.size();
  }
  std::cout << "Searched " << xs.size() << " proteins against a " << fa_string.size() << "-residue query: device time "
            << out.device_us * 1e-3 << " ms, GCUPS " << cells / (double)out.device_us * 1e-3 << " (database " << (packed ? "packed blob" : "FASTA")
            << " read in " << stage_s << " s)" << std::endl;
  return 0;
}
