// packed_db.h — the packed sequence database of the search driver.
//
// The reference keeps UniProt as 561 356 one-protein FASTA files (py/reader.py:52-73) that every MPI rank opens, skips the
// header of and concatenates one by one (mpi_sw_solve_uniprot.cpp:97-110).  Here the database is ONE file: residues as 5-bit
// codes (8 residues in 5 bytes) + offsets, entries sorted by decreasing length (the order the kernels pair and schedule
// them in), the original index of every entry kept so that results are written in input order.  Same layout as
// parallel-genomeseq_b200/dataprep.py (pack_database / load_database):
//   MAGIC "SWBDB001" | n_entries u64 | n_residues u64 | alphabet_len u32 | alphabet (padded to 4 bytes) |
//   orig_index u32[n] | offsets u64[n+1] | packed codes ceil(n_residues / 8) * 5 bytes          (little endian)
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <string>
#include <vector>

namespace swb {

static const char kPackedDbMagic[9] = "SWBDB001";
static const char kPackedDbAlphabet[] = "ARNDCQEGHILKMFPSTWYVBZX*UOJ-";   // code = index, anything else -> 'X'

struct PackedDb {
  std::string blob;                  // residue BYTES, entries back to back in stored (length-sorted) order
  std::vector<uint64_t> offsets;     // n + 1
  std::vector<uint32_t> orig;        // original index of every stored entry
  size_t size() const { return orig.size(); }
};

inline bool write_packed_db(const std::vector<std::string>& seqs, const std::string& path) {
  const uint32_t alen = (uint32_t)std::strlen(kPackedDbAlphabet);
  uint8_t lut[256];
  const uint8_t xcode = (uint8_t)(std::strchr(kPackedDbAlphabet, 'X') - kPackedDbAlphabet);
  std::memset(lut, xcode, sizeof lut);
  for (uint32_t i = 0; i < alen; ++i) { lut[(unsigned char)kPackedDbAlphabet[i]] = (uint8_t)i; lut[(unsigned char)std::tolower(kPackedDbAlphabet[i])] = (uint8_t)i; }
  std::vector<uint32_t> order(seqs.size());
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return seqs[a].size() > seqs[b].size(); });
  std::vector<uint64_t> offs(seqs.size() + 1, 0);
  for (size_t i = 0; i < order.size(); ++i) offs[i + 1] = offs[i] + seqs[order[i]].size();
  const uint64_t n = offs.back(), nw = (n + 7) / 8;
  std::vector<uint8_t> packed(nw * 5, 0);
  uint64_t r = 0;
  for (uint32_t e : order)
    for (unsigned char ch : seqs[e]) {
      const uint64_t w = r >> 3, k = r & 7;
      uint64_t word = 0;
      std::memcpy(&word, packed.data() + w * 5, 5);
      word |= (uint64_t)lut[ch] << (5 * k);
      std::memcpy(packed.data() + w * 5, &word, 5);
      ++r;
    }
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  const uint64_t ne = seqs.size();
  f.write(kPackedDbMagic, 8);
  f.write((const char*)&ne, 8); f.write((const char*)&n, 8); f.write((const char*)&alen, 4);
  std::string alpha(kPackedDbAlphabet, alen);
  alpha.append((4 - alen % 4) % 4, '\0');
  f.write(alpha.data(), (std::streamsize)alpha.size());
  f.write((const char*)order.data(), (std::streamsize)(order.size() * 4));
  f.write((const char*)offs.data(), (std::streamsize)(offs.size() * 8));
  f.write((const char*)packed.data(), (std::streamsize)packed.size());
  return (bool)f;
}

inline bool load_packed_db(const std::string& path, PackedDb* db, std::string* err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { if (err) *err = "cannot open " + path; return false; }
  char magic[8];
  uint64_t ne = 0, n = 0; uint32_t alen = 0;
  f.read(magic, 8); f.read((char*)&ne, 8); f.read((char*)&n, 8); f.read((char*)&alen, 4);
  if (!f || std::memcmp(magic, kPackedDbMagic, 8) != 0 || alen == 0 || alen > 32) { if (err) *err = path + " is not a packed database"; return false; }
  std::string alpha(alen + (4 - alen % 4) % 4, '\0');
  f.read(alpha.data(), (std::streamsize)alpha.size());
  db->orig.resize(ne); db->offsets.resize(ne + 1);
  f.read((char*)db->orig.data(), (std::streamsize)(ne * 4));
  f.read((char*)db->offsets.data(), (std::streamsize)((ne + 1) * 8));
  const uint64_t nw = (n + 7) / 8;
  std::vector<uint8_t> packed(nw * 5);
  f.read((char*)packed.data(), (std::streamsize)packed.size());
  if (!f || db->offsets.back() != n) { if (err) *err = path + " is truncated"; return false; }
  db->blob.resize(n);
  for (uint64_t w = 0; w < nw; ++w) {
    uint64_t word = 0;
    std::memcpy(&word, packed.data() + w * 5, 5);
    for (uint64_t k = 0; k < 8 && w * 8 + k < n; ++k) db->blob[w * 8 + k] = alpha[(word >> (5 * k)) & 31];
  }
  return true;
}

}  // namespace swb
