#include "glf.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace pmh {

static const uint8_t kTranslateBase[16] = {0, 1, 2, 0, 3, 0, 0, 0, 4, 0, 0, 0, 0, 0, 0, 0};  // glfHandler.cpp:4
static const uint8_t kBackTranslateBase[5] = {15, 1, 2, 4, 8};                               // glfHandler.cpp:5

GlfStream::~GlfStream() {
  if (fh_) gzclose((gzFile)fh_);
}
size_t GlfStream::read(void *buf, size_t n) {
  int r = gzread((gzFile)fh_, buf, (unsigned)n);
  return r < 0 ? 0 : (size_t)r;
}
bool GlfStream::eof() { return gzeof((gzFile)fh_) != 0; }

bool GlfStream::open(const std::string &path, std::string *err) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) { if (err) *err = "GLF file " + path + " can  not be opened!"; return false; }
  gzbuffer(f, 1 << 18);
  fh_ = f;
  char magic[4];
  uint32_t header_len = 0;
  if (read(magic, 4) != 4 || magic[0] != 'G' || magic[1] != 'L' || magic[2] != 'F' || magic[3] != 3 ||
      read(&header_len, 4) != 4 || header_len > 1024 * 1024) {
    if (err) *err = "GLF file " + path + ": invalid format or unsupported version";
    gzclose(f); fh_ = nullptr;
    return false;
  }
  std::string text(header_len, '\0');
  if (header_len && read(&text[0], header_len) != header_len) {
    if (err) *err = "GLF file " + path + ": unexpected end of file";
    gzclose(f); fh_ = nullptr;
    return false;
  }
  end_of_section_ = true;
  return true;
}

bool GlfStream::next_section() {  // glfHandler.cpp:139-171
  while (!end_of_section_ && !eof()) next_entry();
  end_of_section_ = false;
  int32_t label_len = 0;
  position = 0;
  if (read(&label_len, 4) == 4) {
    std::string buf((size_t)(label_len > 0 ? label_len : 0), '\0');
    if (label_len > 0) read(&buf[0], (size_t)label_len);
    label = std::string(buf.c_str());  // stored length includes the NUL
    max_position = 0;
    read(&max_position, 4);
    return max_position > 0 && !eof();
  }
  return false;
}

bool GlfStream::next_base_entry() {  // glfHandler.cpp:173-183
  bool result;
  do { result = next_entry(); } while (result && record_type == 2);
  return result;
}

bool GlfStream::next_entry() {  // glfHandler.cpp:186-261
  uint8_t rec[20];
  if (end_of_section_ || read(rec, 1) != 1) {
    end_of_section_ = true; record_type = 0; position = max_position + 1;
    return false;
  }
  record_type = rec[0] >> 4;
  ref_base = rec[0] & 0xf;
  switch (record_type) {
    case 0:
      end_of_section_ = true; position = max_position + 1;
      return true;
    case 1:
      if (read(rec + 1, 19) == 19) {
        ref_base = kTranslateBase[ref_base];
        uint32_t offset; memcpy(&offset, rec + 1, 4);
        uint32_t dm; memcpy(&dm, rec + 5, 4);
        depth = dm & 0xffffff;
        map_quality = rec[9];
        memcpy(lk, rec + 10, 10);
        position = position + (int)offset;
        return true;
      }
      record_type = 0; position = max_position + 1;
      return false;
    case 2:
      // indel record: 16 more fixed bytes, then two allele strings (glfHandler.cpp:233-255)
      if (read(rec + 1, 16) == 16) {
        ref_base = kTranslateBase[ref_base];
        uint32_t offset; memcpy(&offset, rec + 1, 4);
        position = position + (int)offset;
        int16_t len[2]; memcpy(len, rec + 13, 4);
        char skip[65536];
        bool ok = true;
        for (int a = 0; a < 2 && ok; a++) {
          size_t n = (size_t)(len[a] < 0 ? -len[a] : len[a]);
          if (n && read(skip, n) != n) ok = false;
        }
        if (ok) return true;
      }
      record_type = 0; position = max_position + 1;
      return false;
  }
  return false;
}

// ---- writer ---------------------------------------------------------------------------------
GlfWriter::~GlfWriter() { close(); }
void GlfWriter::put(const void *p, size_t n) {
  if (gz_) gzwrite((gzFile)gz_, p, (unsigned)n);
  else if (fp_) fwrite(p, 1, n, (FILE *)fp_);
}
bool GlfWriter::create(const std::string &path, bool gzip) {
  if (gzip) gz_ = gzopen(path.c_str(), "wb1");
  else fp_ = fopen(path.c_str(), "wb");
  if (!gz_ && !fp_) return false;
  const char magic[4] = {'G', 'L', 'F', 3};
  uint32_t zero = 0;
  put(magic, 4);
  put(&zero, 4);
  return true;
}
void GlfWriter::begin_section(const std::string &label, int length) {
  int32_t n = (int32_t)label.size() + 1;
  put(&n, 4);
  put(label.c_str(), (size_t)n);
  int32_t len = length;
  put(&len, 4);
  position_ = 0;
}
void GlfWriter::write_entry(int position, int ref_base, uint32_t depth, uint8_t map_quality, const uint8_t lk[10]) {
  uint8_t rec[20];
  rec[0] = (uint8_t)((1 << 4) | kBackTranslateBase[ref_base]);
  uint32_t offset = (uint32_t)(position - position_);
  position_ = position;
  memcpy(rec + 1, &offset, 4);
  uint32_t dm = depth & 0xffffff;  // minLLK = 0 (ignored by the reader)
  memcpy(rec + 5, &dm, 4);
  rec[9] = map_quality;
  memcpy(rec + 10, lk, 10);
  put(rec, 20);
}
void GlfWriter::end_section() { uint8_t z = 0; put(&z, 1); }
void GlfWriter::close() {
  if (gz_) { gzclose((gzFile)gz_); gz_ = nullptr; }
  if (fp_) { fclose((FILE *)fp_); fp_ = nullptr; }
}

// ---- merge ----------------------------------------------------------------------------------
bool GlfSet::open(const std::vector<std::string> &paths, std::string *err) {
  streams_ = std::vector<GlfStream>(paths.size());
  lead_ = -1;
  for (size_t i = 0; i < paths.size(); i++) {
    if (paths[i].empty()) continue;
    if (!streams_[i].open(paths[i], err)) return false;
    if (lead_ < 0) lead_ = (int)i;
  }
  if (lead_ < 0) { if (err) *err = "no GLF file could be opened"; return false; }
  return true;
}
const std::string &GlfSet::label() const { return streams_[lead_].label; }
int GlfSet::max_position() const { return streams_[lead_].max_position; }

bool GlfSet::next_section() {
  bool flag = false;
  for (auto &s : streams_) {
    if (!s.is_open()) continue;
    flag = s.next_section();
    const GlfStream &lead = streams_[lead_];
    if (s.max_position != lead.max_position || s.label != lead.label)
      throw std::runtime_error("GLF files are not compatible:\n\tsection " + lead.label + " with " +
                               std::to_string(lead.max_position) + " entries vs section " + s.label + " with " +
                               std::to_string(s.max_position) + " entries");
    if (!flag) return false;
  }
  current_pos_ = 0;
  return flag;
}

bool GlfSet::next_site(pm_site_hdr *hdr, pm_person_site *out) {
  if (current_pos_ > 0)
    for (auto &s : streams_)
      if (s.is_open() && s.record_type == 0) return false;
  for (auto &s : streams_)
    if (s.is_open() && s.position == current_pos_) s.next_base_entry();
  current_pos_ = streams_[lead_].position;
  int ref = streams_[lead_].ref_base;
  for (auto &s : streams_)
    if (s.is_open() && s.position < current_pos_) { current_pos_ = s.position; ref = s.ref_base; }
  if (current_pos_ > streams_[lead_].max_position) return false;
  hdr->pos = (uint32_t)current_pos_;
  hdr->ref_base = (uint8_t)ref;
  hdr->chr_class = PM_CHR_AUTO;
  hdr->reserved = 0;
  for (size_t i = 0; i < streams_.size(); i++) {
    pm_person_site &p = out[i];
    memset(&p, 0, sizeof p);
    const GlfStream &s = streams_[i];
    if (!s.is_open() || s.position != current_pos_) continue;
    memcpy(p.lk, s.lk, 10);
    p.depth[0] = (uint8_t)(s.depth & 0xff); p.depth[1] = (uint8_t)((s.depth >> 8) & 0xff); p.depth[2] = (uint8_t)((s.depth >> 16) & 0xff);
    p.map_quality = s.map_quality;
  }
  return true;
}

}  // namespace pmh
