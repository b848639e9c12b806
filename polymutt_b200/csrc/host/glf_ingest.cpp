#include "glf_ingest.h"

#include <sys/resource.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace pmh {

static const uint8_t kTranslateBase[16] = {0, 1, 2, 0, 3, 0, 0, 0, 4, 0, 0, 0, 0, 0, 0, 0};  // core/glfHandler.cpp:4

GlfBatchReader::~GlfBatchReader() {
  for (auto &s : streams_) if (s.f) gzclose(s.f);
}

template <typename F>
void GlfBatchReader::parallel_streams(F fn) {
  const int n = (int)streams_.size();
  const int T = std::max(1, std::min(threads_, n));
  if (T == 1) { fn(0, n, 0); return; }
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> errs((size_t)T);
  for (int t = 0; t < T; t++) {
    const int lo = (int)((long long)n * t / T), hi = (int)((long long)n * (t + 1) / T);
    pool.emplace_back([&, lo, hi, t]() {
      try { fn(lo, hi, t); } catch (...) { errs[(size_t)t] = std::current_exception(); }
    });
  }
  for (auto &th : pool) th.join();
  for (auto &e : errs) if (e) std::rethrow_exception(e);
}

bool GlfBatchReader::Stream::fill(size_t need) {
  if (raw_end - raw_beg >= need) return true;
  if (raw_beg > 0) {
    memmove(raw.data(), raw.data() + raw_beg, raw_end - raw_beg);
    raw_end -= raw_beg;
    raw_beg = 0;
  }
  if (need > raw.size()) raw.resize(need + (1 << 16));  // an indel record can carry up to 2 x 32 KiB of allele text
  while (raw_end < need && !file_eof) {
    int got = gzread(f, raw.data() + raw_end, (unsigned)(raw.size() - raw_end));
    if (got < 0) {  // a damaged .gz is an error, not the end of the chromosome (the run would exit 0 with a partial VCF)
      int errnum = 0;
      const char *msg = gzerror(f, &errnum);
      if (errnum == Z_BUF_ERROR) { file_eof = true; break; }  // truncated last block: premature end of file, as the reference
      throw std::runtime_error(std::string("GLF stream: read error (") + (msg ? msg : "?") + ")");
    }
    if (got == 0) { file_eof = true; break; }
    raw_end += (size_t)got;
  }
  return raw_end - raw_beg >= need;
}

void GlfBatchReader::Stream::compact() {
  if (head == 0) return;
  if (head == pos.size()) { pos.clear(); ref.clear(); rec.clear(); head = 0; return; }
  if (head < pos.size() / 2) return;  // amortise
  pos.erase(pos.begin(), pos.begin() + (long)head);
  ref.erase(ref.begin(), ref.begin() + (long)head);
  rec.erase(rec.begin(), rec.begin() + (long)head);
  head = 0;
}

// glfHandler::NextEntry (core/glfHandler.cpp:186-261) on a byte buffer
void GlfBatchReader::Stream::decode(size_t want_records) {
  compact();
  while (!ended && pending() < want_records) {
    // fast path: a run of base records that are completely inside the buffer goes into pre-sized arrays without the
    // per-record refill / capacity checks of the general path below
    {
      const size_t avail = (raw_end - raw_beg) / 20, room = want_records - pending();
      size_t n = avail < room ? avail : room;
      if (n > 0) {
        const size_t old = pos.size();
        pos.resize(old + n); ref.resize(old + n); rec.resize(old + n);
        const unsigned char *r = raw.data() + raw_beg;
        size_t k = 0;
        int p = position, lp = last_pos;
        for (; k < n && (r[0] >> 4) == 1; k++, r += 20) {
          uint32_t offset, dm;
          memcpy(&offset, r + 1, 4);
          memcpy(&dm, r + 5, 4);
          p += (int)offset;
          if (lp == p && lp >= 0 && offset == 0) {
            pos.resize(old + k); ref.resize(old + k); rec.resize(old + k);
            throw std::runtime_error("GLF stream repeats a position (offset 0): not supported by the batched reader");
          }
          pm_person_site &o = rec[old + k];
          memcpy(o.lk, r + 10, 10);
          o.depth[0] = (uint8_t)(dm & 0xff); o.depth[1] = (uint8_t)((dm >> 8) & 0xff); o.depth[2] = (uint8_t)((dm >> 16) & 0xff);
          o.map_quality = r[9];
          o.pad[0] = o.pad[1] = 0;
          pos[old + k] = p;
          ref[old + k] = kTranslateBase[r[0] & 0xf];
          lp = p;
        }
        position = p; last_pos = lp;
        raw_beg += 20 * k;
        if (k < n) { pos.resize(old + k); ref.resize(old + k); rec.resize(old + k); }
        if (k > 0) continue;
      }
    }
    if (!fill(1)) { ended = true; break; }  // premature end of file = end of section
    const unsigned char b0 = raw[raw_beg];
    const int type = b0 >> 4;
    if (type == 0) { raw_beg += 1; ended = true; break; }
    if (type == 1) {
      if (!fill(20)) { raw_beg = raw_end; ended = true; break; }
      const unsigned char *r = raw.data() + raw_beg;
      uint32_t offset, dm;
      memcpy(&offset, r + 1, 4);
      memcpy(&dm, r + 5, 4);
      position += (int)offset;
      if (last_pos == position && last_pos >= 0 && offset == 0)
        throw std::runtime_error("GLF stream repeats a position (offset 0): not supported by the batched reader");
      pm_person_site p;
      memset(&p, 0, sizeof p);
      memcpy(p.lk, r + 10, 10);
      p.depth[0] = (uint8_t)(dm & 0xff); p.depth[1] = (uint8_t)((dm >> 8) & 0xff); p.depth[2] = (uint8_t)((dm >> 16) & 0xff);
      p.map_quality = r[9];
      pos.push_back(position);
      ref.push_back(kTranslateBase[b0 & 0xf]);
      rec.push_back(p);
      last_pos = position;
      raw_beg += 20;
      continue;
    }
    if (type == 2) {  // indel: 17 fixed bytes + two allele strings, skipped (NextBaseEntry)
      if (!fill(17)) { raw_beg = raw_end; ended = true; break; }
      const unsigned char *r = raw.data() + raw_beg;
      uint32_t offset;
      memcpy(&offset, r + 1, 4);
      int16_t len[2];
      memcpy(len, r + 13, 4);
      const size_t extra = (size_t)(len[0] < 0 ? -len[0] : len[0]) + (size_t)(len[1] < 0 ? -len[1] : len[1]);
      if (!fill(17 + extra)) { raw_beg = raw_end; ended = true; break; }
      position += (int)offset;
      raw_beg += 17 + extra;
      continue;
    }
    ended = true;  // unknown record type: the reference's NextEntry returns false
    break;
  }
}

bool GlfBatchReader::open(const std::vector<std::string> &paths, int threads, std::string *err) {
  streams_ = std::vector<Stream>(paths.size());
  lead_ = -1;
  threads_ = threads > 0 ? threads : (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
  {  // one descriptor per person stays open for the whole run: lift the soft limit as far as the hard one allows
    struct rlimit rl;
    const rlim_t want = (rlim_t)paths.size() + 256;
    if (getrlimit(RLIMIT_NOFILE, &rl) == 0 && rl.rlim_cur < want && rl.rlim_cur < rl.rlim_max) {
      rl.rlim_cur = rl.rlim_max == RLIM_INFINITY ? want : std::min(want, rl.rlim_max);
      setrlimit(RLIMIT_NOFILE, &rl);
    }
  }
  for (size_t i = 0; i < paths.size(); i++) {
    if (paths[i].empty()) continue;
    Stream &s = streams_[i];
    s.f = gzopen(paths[i].c_str(), "rb");
    if (!s.f) { if (err) *err = "GLF file " + paths[i] + " can  not be opened!"; return false; }
    gzbuffer(s.f, 1 << 16);
    s.raw.resize(1 << 16);
    if (!s.fill(8) || memcmp(s.raw.data(), "GLF\3", 4) != 0) { if (err) *err = "GLF file " + paths[i] + ": invalid format or unsupported version"; return false; }
    uint32_t hl;
    memcpy(&hl, s.raw.data() + 4, 4);
    if (hl > 1024 * 1024) { if (err) *err = "GLF file " + paths[i] + ": header too large -- bailing"; return false; }
    s.raw_beg = 8;
    size_t left = hl;  // skip the header text
    while (left) {
      if (!s.fill(1)) { if (err) *err = "GLF file " + paths[i] + ": unexpected end of file"; return false; }
      size_t take = std::min(left, s.raw_end - s.raw_beg);
      s.raw_beg += take; left -= take;
    }
    if (lead_ < 0) lead_ = (int)i;
  }
  if (lead_ < 0) { if (err) *err = "no GLF file could be opened"; return false; }
  section_done_ = true;
  return true;
}

bool GlfBatchReader::next_section() {
  // glfHandler::NextSection (core/glfHandler.cpp:139-171) for every stream, in parallel
  std::vector<char> ok(streams_.size(), 1);
  parallel_streams([&](int lo, int hi, int) {
    for (int i = lo; i < hi; i++) {
      Stream &s = streams_[(size_t)i];
      if (!s.f) continue;
      while (!s.ended) { s.pos.clear(); s.ref.clear(); s.rec.clear(); s.head = 0; s.decode(4096); }  // drain the old section
      s.pos.clear(); s.ref.clear(); s.rec.clear(); s.head = 0;
      s.position = 0; s.last_pos = -1;
      int32_t label_len = 0;
      if (!s.fill(4)) { ok[(size_t)i] = 0; continue; }
      memcpy(&label_len, s.raw.data() + s.raw_beg, 4);
      s.raw_beg += 4;
      const size_t ll = (size_t)(label_len > 0 ? label_len : 0);
      if (!s.fill(ll + 4)) { ok[(size_t)i] = 0; continue; }
      s.label = std::string(std::string((const char *)s.raw.data() + s.raw_beg, ll).c_str());
      s.raw_beg += ll;
      memcpy(&s.max_position, s.raw.data() + s.raw_beg, 4);
      s.raw_beg += 4;
      s.ended = false;
      if (s.max_position <= 0) ok[(size_t)i] = 0;
    }
  });
  const Stream &lead = streams_[(size_t)lead_];
  // the reference walks the streams in order and stops at the first that has no further section
  for (size_t i = 0; i < streams_.size(); i++) {
    const Stream &s = streams_[i];
    if (!s.f) continue;
    if (ok[i] && ok[(size_t)lead_] && (s.max_position != lead.max_position || s.label != lead.label))
      throw std::runtime_error("GLF files are not compatible:\n\tsection " + lead.label + " with " + std::to_string(lead.max_position) +
                               " entries vs section " + s.label + " with " + std::to_string(s.max_position) + " entries");
    if (!ok[i]) return false;
  }
  label_ = lead.label;
  max_position_ = lead.max_position;
  section_done_ = false;
  prev1_ = -1; prev2_ = -1;
  return true;
}

size_t GlfBatchReader::next_batch(pm_site_hdr *hdr, pm_person_site *out, size_t max_sites) {
  if (section_done_ || max_sites == 0) return 0;
  const size_t np = streams_.size();
  // 1. every live stream gets at least `want` pending records (or reaches its end)
  size_t want = std::max<size_t>(64, std::min<size_t>(max_sites, ((size_t)96 << 20) / (np * 24)));
  parallel_streams([&](int lo, int hi, int) {
    for (int i = lo; i < hi; i++) {
      Stream &s = streams_[(size_t)i];
      if (s.f && !s.ended && s.pending() < want) s.decode(want);
    }
  });
  // 2. window: every position <= wend is completely known
  long long base = LLONG_MAX, wend = LLONG_MAX;
  long long T = LLONG_MAX;  // min over ended streams of their last base-record position (-1: none at all)
  for (const Stream &s : streams_) {
    if (!s.f) continue;
    if (s.pending()) base = std::min<long long>(base, s.pos[s.head]);
    if (s.ended) T = std::min<long long>(T, s.last_pos);
    else wend = std::min<long long>(wend, s.pos.back());
  }
  if (base == LLONG_MAX) { section_done_ = true; return 0; }  // nothing left anywhere
  const long long cap = (long long)std::max<size_t>(1 << 16, 4 * max_sites);
  if (wend == LLONG_MAX) {  // every stream has been decoded to its end: the rest of the section is known
    wend = base;
    for (const Stream &s : streams_) if (s.f && s.pending()) wend = std::max<long long>(wend, s.pos.back());
  }
  wend = std::max(base, std::min(wend, base + cap - 1));
  const size_t width = (size_t)(wend - base + 1);
  mark_.assign(width, 0);
  parallel_streams([&](int lo, int hi, int) {
    for (int i = lo; i < hi; i++) {
      const Stream &s = streams_[(size_t)i];
      if (!s.f) continue;
      for (size_t k = s.head; k < s.pos.size() && s.pos[k] <= wend; k++) mark_[(size_t)(s.pos[k] - base)] = 1;
    }
  });
  // 3. rows, with the reference's termination rules
  row_.assign(width, -1);
  size_t n = 0;
  long long limit = base - 1;
  for (size_t w = 0; w < width && n < max_sites; w++) {
    if (!mark_[w]) continue;
    const long long p = base + (long long)w;
    // Move2NextBaseEntry top check: some stream read its end marker in an earlier call
    if (prev1_ > 0 && T <= prev2_) { section_done_ = true; break; }
    if (p > max_position_) { section_done_ = true; break; }
    row_[w] = (int32_t)n++;
    limit = p;
    prev2_ = prev1_; prev1_ = p;
  }
  if (n == 0) { section_done_ = true; return 0; }
  // 4. scatter (parallel over streams), reference base = lead stream's if present, else the lowest column's
  memset(out, 0, n * np * sizeof(pm_person_site));
  owner_.assign(n, UINT32_MAX);
  std::atomic<uint32_t> *own = reinterpret_cast<std::atomic<uint32_t> *>(owner_.data());
  parallel_streams([&](int lo, int hi, int) {
    for (int i = lo; i < hi; i++) {
      Stream &s = streams_[(size_t)i];
      if (!s.f) continue;
      const uint32_t key = (i == lead_) ? 0u : (uint32_t)i + 1u;
      size_t k = s.head;
      for (; k < s.pos.size() && s.pos[k] <= limit; k++) {
        const int32_t r = row_[(size_t)(s.pos[k] - base)];
        if (r < 0) continue;
        out[(size_t)r * np + (size_t)i] = s.rec[k];
        const uint32_t v = (key << 8) | s.ref[k];
        uint32_t cur = own[r].load(std::memory_order_relaxed);
        while (v < cur && !own[r].compare_exchange_weak(cur, v, std::memory_order_relaxed)) {}
      }
      s.head = k;
    }
  });
  size_t w = 0;
  for (size_t r = 0; r < n; r++) {
    while (row_[w] != (int32_t)r) w++;
    hdr[r].pos = (uint32_t)(base + (long long)w);
    hdr[r].ref_base = (uint8_t)(owner_[r] & 0xff);
    hdr[r].chr_class = PM_CHR_AUTO;
    hdr[r].reserved = 0;
  }
  return n;
}

}  // namespace pmh
