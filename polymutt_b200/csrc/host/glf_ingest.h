// Batched, multi-threaded GLF ingest: the N-way merge of PedigreeGLF::Move2NextBaseEntry
// (src/PedigreeGLF.cpp:282-324) restated for throughput.  The reference advances N cursors one site at a
// time with two tiny reads per person-site; with thousands of people that merge, not the likelihood engine,
// bounds the executable.  Here every stream is block-decoded (gz-transparent) into arrays of
// (position, ref base, 16-byte packed record) by a pool of threads, the site list of a window is the union of
// the streams' positions (a bitmap), and every stream scatters its own records into the site-major batch.
//
// Semantics kept bit for bit (checked against GlfSet on ragged fixtures, tests/test_host.py):
//   * a site exists wherever at least one stream has a base record; people without a record there are zeros;
//   * the reference base comes from the lead stream (first opened) if it has a record at the site, else from the
//     first column that has one (strict `<` in the min scan);
//   * a chromosome ends one site after the first stream runs out (the recordType==0 check at the top of
//     Move2NextBaseEntry), or when the position passes the lead stream's section length;
//   * indel records (type 2) are skipped; sections must agree in label and length across files.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <string>
#include <vector>

#include "polymutt_b200.h"

namespace pmh {

class GlfBatchReader {
 public:
  ~GlfBatchReader();
  // paths[c] is the GLF of VCF column c, empty = no GLF.  threads <= 0: hardware concurrency (capped at 32).
  bool open(const std::vector<std::string> &paths, int threads, std::string *err);
  // Move2NextSection for every stream; false at end of input; throws std::runtime_error on incompatible sections.
  bool next_section();
  // Up to max_sites merged sites of the current section, in position order.  Returns the number written
  // (0 = the section is finished).  out must hold max_sites * n_person records.
  size_t next_batch(pm_site_hdr *hdr, pm_person_site *out, size_t max_sites);
  const std::string &label() const { return label_; }
  int max_position() const { return max_position_; }
  int n_person() const { return (int)streams_.size(); }

 private:
  struct Stream {
    gzFile f = nullptr;
    std::vector<unsigned char> raw;   // undecoded bytes
    size_t raw_beg = 0, raw_end = 0;
    bool file_eof = false;
    // decoded, not yet consumed base records of the current section
    std::vector<int32_t> pos;
    std::vector<uint8_t> ref;
    std::vector<pm_person_site> rec;
    size_t head = 0;
    int position = 0;                 // running position of the decoder
    bool ended = true;                // end-of-section marker (or end of file) reached by the decoder
    int last_pos = -1;                // position of the last base record of the section (valid once ended)
    std::string label;
    int max_position = 0;
    bool fill(size_t need);           // makes `need` raw bytes available at raw_beg
    void decode(size_t want_records); // decodes until want_records are pending, or the section ends
    size_t pending() const { return pos.size() - head; }
    void compact();
  };
  std::vector<Stream> streams_;
  int lead_ = -1;
  int threads_ = 1;
  std::string label_;
  int max_position_ = 0;
  bool section_done_ = true;
  long long prev1_ = -1, prev2_ = -1; // positions of the last two sites handed out (termination rule)
  std::vector<uint8_t> mark_;         // per position of the window: some stream has a record
  std::vector<int32_t> row_;          // per position of the window: row in the batch, -1 = none
  std::vector<uint32_t> owner_;       // per row: (priority << 8) | ref base of the stream that names the reference base
  template <typename F>
  void parallel_streams(F fn);
};

}  // namespace pmh
