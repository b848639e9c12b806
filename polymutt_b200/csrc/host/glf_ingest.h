// Batched, multi-threaded GLF ingest: the N-way merge of PedigreeGLF::Move2NextBaseEntry
// (src/PedigreeGLF.cpp:282-324) restated for throughput.  The reference advances N cursors one site at a
// time with two tiny reads per person-site; with thousands of people that merge, not the likelihood engine,
// bounds the executable.  Here a pool of threads scans every stream's (gz-transparent) bytes for record boundaries
// and positions only, the site list of a window is the union of the streams' positions (a bitmap), and the batch is
// filled row by row: a thread owns a contiguous range of columns and writes its stretch of every site's row in one go,
// straight from the 20-byte file records into the 16-byte packed ones.
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
  // The record conversion has an SSSE3 form (chosen at run time where the CPU has it) and a portable one; this switches
  // every reader of the process to the portable form (pm-tools pack --portable: the tests compare the two).
  static void use_portable_convert(bool on);

 private:
  struct Stream {
    gzFile f = nullptr;               // gzip / BGZF input; nullptr for an uncompressed file, which is read through fd
    int fd = -1;
    bool live = false;                // the column has a GLF
    // Bytes of the (inflated) stream: [raw_keep, raw_dec) holds the records that are decoded but not consumed yet,
    // [raw_dec, raw_end) what has been read but not looked at.  Base records are never copied out: a pending record is
    // (pos[k], off[k]) = its position and the offset of its 20 bytes in `raw`.
    std::vector<unsigned char> raw;
    size_t raw_keep = 0, raw_dec = 0, raw_end = 0;
    bool file_eof = false;
    std::vector<int32_t> pos;
    std::vector<uint32_t> off;
    size_t head = 0;                  // first pending record
    int position = 0;                 // running position of the decoder
    bool ended = true;                // end-of-section marker (or end of file) reached by the decoder
    int last_pos = -1;                // position of the last base record of the section (valid once ended)
    int last_rank = 0;                // how many base records right before it sit at that same position (a repeated position)
    bool dup_wait = false;            // the next record repeats last_pos: not decoded before everything pending is consumed
    std::string label;
    int max_position = 0;
    bool fill(size_t need);           // makes `need` undecoded bytes available at raw_dec
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
  long long prev1_ = -1, prev2_ = -1; // the last two sites handed out, as (position << 20 | rank) (termination rule)
  std::vector<uint8_t> mark_;         // per position of the window: some stream has a record
  std::vector<int32_t> rowpos_;       // position of every row of the batch
  std::vector<uint32_t> owner_;       // per (thread, row): (priority << 8) | ref base of the best stream of that thread's column range
  template <typename F>
  void parallel_streams(F fn);
};

}  // namespace pmh
