// GLF v3 streams (one per person) and their N-way merge into packed sites.
//
// Record layout and parsing follow core/glfHandler.h:21-42 and core/glfHandler.cpp:194-261; the merge
// follows PedigreeGLF::Move2NextSection / Move2NextBaseEntry (src/PedigreeGLF.cpp:196-220, 282-324).
// Files may be plain or gzip (the reference's IFILE is gz-transparent; so is zlib's gzread).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "polymutt_b200.h"

namespace pmh {

class GlfStream {
 public:
  ~GlfStream();
  bool open(const std::string &path, std::string *err);
  bool is_open() const { return fh_ != nullptr; }
  bool next_section();
  bool next_entry();
  bool next_base_entry();

  std::string label;
  int max_position = 0;
  int position = 0;
  int record_type = 0;
  int ref_base = 0;  // 0..4 after translateBase
  uint8_t lk[10] = {0};
  uint32_t depth = 0;
  uint8_t map_quality = 0;

 private:
  void *fh_ = nullptr;  // gzFile
  bool end_of_section_ = true;
  size_t read(void *buf, size_t n);
  bool eof();
};

// Writer used by the synthetic generator and the tests (glfHandler::WriteHeader/BeginSection/WriteEntry,
// core/glfHandler.cpp:319-375).
class GlfWriter {
 public:
  ~GlfWriter();
  bool create(const std::string &path, bool gzip);
  void begin_section(const std::string &label, int length);
  void write_entry(int position, int ref_base, uint32_t depth, uint8_t map_quality, const uint8_t lk[10]);
  void end_section();
  void close();

 private:
  void *gz_ = nullptr;
  void *fp_ = nullptr;
  int position_ = 0;
  void put(const void *p, size_t n);
};

class GlfSet {
 public:
  // paths[c] is the GLF of VCF column c, empty string = no GLF for that person.
  bool open(const std::vector<std::string> &paths, std::string *err);
  // PedigreeGLF::Move2NextSection; false at end of input. Throws std::runtime_error on incompatible sections.
  bool next_section();
  // PedigreeGLF::Move2NextBaseEntry; fills one header and n_person records.
  bool next_site(pm_site_hdr *hdr, pm_person_site *out);
  const std::string &label() const;
  int max_position() const;
  int n_person() const { return (int)streams_.size(); }

 private:
  std::vector<GlfStream> streams_;
  int lead_ = -1;  // nonNULLglf: first stream that opened
  int current_pos_ = 0;
};

}  // namespace pmh
