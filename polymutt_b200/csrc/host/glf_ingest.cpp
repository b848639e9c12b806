#include "glf_ingest.h"

#include <fcntl.h>
#include <sys/resource.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>
#if defined(__x86_64__) && defined(__GNUC__)
#include <tmmintrin.h>
#endif

namespace pmh {

static const uint8_t kTranslateBase[16] = {0, 1, 2, 0, 3, 0, 0, 0, 4, 0, 0, 0, 0, 0, 0, 0};  // core/glfHandler.cpp:4

// 20-byte glfEntry (core/glfHandler.h:21-42: type/ref, offset u32, depth:24 | minLLK:8, mapQ, lk[10]) ->
// 16-byte pm_person_site (lk[10], depth[3], mapQ, pad[2]) for a run of records of one stream that sit at consecutive
// rows of the batch (no position compares: the caller has established the run).  dst advances by one row per record.
static void convert_run_scalar(const unsigned char *raw, const uint32_t *op, size_t count, unsigned char *dst, size_t stride) {
  for (size_t i = 0; i < count; i++, dst += stride) {
    const unsigned char *rec = raw + op[i];
    uint64_t lo8, hi8;
    uint32_t dm;
    uint16_t lk89;
    memcpy(&lo8, rec + 10, 8);
    memcpy(&lk89, rec + 18, 2);
    memcpy(&dm, rec + 5, 4);
    hi8 = (uint64_t)lk89 | ((uint64_t)(dm & 0xffffffu) << 16) | ((uint64_t)rec[9] << 40);
    memcpy(dst, &lo8, 8);
    memcpy(dst + 8, &hi8, 8);
  }
}
#if defined(__x86_64__) && defined(__GNUC__)
// the same with one 16-byte load of record bytes 4..19, one byte shuffle and one store per record
__attribute__((target("ssse3"))) static void convert_run_ssse3(const unsigned char *raw, const uint32_t *op, size_t count, unsigned char *dst, size_t stride) {
  const __m128i pick = _mm_setr_epi8(6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 1, 2, 3, 5, (char)0x80, (char)0x80);
  for (size_t i = 0; i < count; i++, dst += stride) {
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(raw + op[i] + 4));
    _mm_storeu_si128(reinterpret_cast<__m128i *>(dst), _mm_shuffle_epi8(v, pick));
  }
}
#endif
typedef void (*ConvertRun)(const unsigned char *, const uint32_t *, size_t, unsigned char *, size_t);
static bool g_portable_convert = false;
void GlfBatchReader::use_portable_convert(bool on) { g_portable_convert = on; }
static ConvertRun pick_convert_run() {
#if defined(__x86_64__) && defined(__GNUC__)
  static const bool have_ssse3 = __builtin_cpu_supports("ssse3");
  if (have_ssse3 && !g_portable_convert) return convert_run_ssse3;
#endif
  return convert_run_scalar;
}

static double g_t[4];
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
GlfBatchReader::~GlfBatchReader() {
  for (auto &s : streams_) { if (s.f) gzclose(s.f); else if (s.fd >= 0) close(s.fd); }
  if (getenv("PM_TIMING")) fprintf(stderr, "[pm timing] ingest: decode %.3f s, window+mark %.3f s, rows %.3f s, fill %.3f s\n", g_t[0], g_t[1], g_t[2], g_t[3]);
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

// Makes `need` undecoded bytes available at raw_dec.  Bytes of records that are decoded but not consumed yet
// (from raw_keep on) stay where the record offsets point: the buffer is compacted relative to raw_keep.
bool GlfBatchReader::Stream::fill(size_t need) {
  if (raw_end - raw_dec >= need) return true;
  if (raw_keep > 0 && (raw_keep >= raw.size() / 2 || raw_dec - raw_keep + need > raw.size() - raw_keep)) {
    memmove(raw.data(), raw.data() + raw_keep, raw_end - raw_keep);
    for (size_t k = head; k < off.size(); k++) off[k] -= (uint32_t)raw_keep;
    raw_dec -= raw_keep; raw_end -= raw_keep; raw_keep = 0;
  }
  if (raw_dec + need > raw.size()) raw.resize(std::max(raw.size() * 2, raw_dec + need + (1 << 16)));  // an indel record can carry 2 x 32 KiB of allele text
  while (raw_end - raw_dec < need && !file_eof) {
    if (!f) {  // an uncompressed GLF: straight from the page cache into the record buffer (no second copy through zlib)
      const ssize_t n = read(fd, raw.data() + raw_end, std::min<size_t>(raw.size() - raw_end, (size_t)1 << 30));
      if (n < 0) { if (errno == EINTR) continue; throw std::runtime_error(std::string("GLF stream: read error (") + strerror(errno) + ")"); }
      if (n == 0) { file_eof = true; break; }
      raw_end += (size_t)n;
      continue;
    }
    int got = gzread(f, raw.data() + raw_end, (unsigned)std::min<size_t>(raw.size() - raw_end, 1u << 30));
    if (got < 0) {  // a damaged .gz is an error, not the end of the chromosome (the run would exit 0 with a partial VCF)
      int errnum = 0;
      const char *msg = gzerror(f, &errnum);
      if (errnum == Z_BUF_ERROR) { file_eof = true; break; }  // truncated last block: premature end of file, as the reference
      throw std::runtime_error(std::string("GLF stream: read error (") + (msg ? msg : "?") + ")");
    }
    if (got == 0) { file_eof = true; break; }
    raw_end += (size_t)got;
  }
  return raw_end - raw_dec >= need;
}

void GlfBatchReader::Stream::compact() {
  if (head == 0) return;
  if (head == pos.size()) { pos.clear(); off.clear(); head = 0; raw_keep = raw_dec; return; }
  raw_keep = off[head];
  if (head < pos.size() / 2) return;  // amortise
  pos.erase(pos.begin(), pos.begin() + (long)head);
  off.erase(off.begin(), off.begin() + (long)head);
  head = 0;
}

// glfHandler::NextEntry (core/glfHandler.cpp:186-261) on a byte buffer.  Base records are NOT copied: a pending record
// is its position and the offset of its 20 bytes in `raw`.
//
// A base record with offset 0 right after a base record repeats a position.  The reference's cursor walks through it
// (src/PedigreeGLF.cpp:282-324: the stream advances to the repeat, stays at currentPos, and the minimum over the streams
// gives that position again): a second site at the same position, made of the second records of the streams that have one.
// Here a window holds every position once, so the decoder stops in front of a repeat (dup_wait) until the records
// before it have been consumed; the repeat then opens the next window, at the same position.
void GlfBatchReader::Stream::decode(size_t want_records) {
  compact();
  if (dup_wait) {
    if (pending() > 0) return;
    dup_wait = false;
  }
  while (!ended && !dup_wait && pending() < want_records) {
    // refill in big steps: the window of undecoded bytes should cover what is still wanted
    if (raw_end - raw_dec < 20) {
      const size_t want_bytes = std::min<size_t>((want_records - pending()) * 20 + 1, (size_t)4 << 20);
      fill(want_bytes);
      if (raw_end == raw_dec) { ended = true; break; }  // premature end of file = end of section
    }
    {  // fast path: a run of base records that are completely inside the buffer
      const size_t avail = (raw_end - raw_dec) / 20, room = want_records - pending();
      const size_t n = avail < room ? avail : room;
      if (n > 0) {
        const size_t old = pos.size();
        pos.resize(old + n); off.resize(old + n);
        const unsigned char *r = raw.data() + raw_dec;
        size_t k = 0;
        int p = position, lp = last_pos, rank = last_rank;
        for (; k < n && (r[0] >> 4) == 1; k++, r += 20) {
          uint32_t offset;
          memcpy(&offset, r + 1, 4);
          if (offset == 0 && lp == p && lp >= 0) {
            if (old + k != head) { dup_wait = true; break; }  // records before the repeat are still pending
            rank++;
          } else {
            rank = 0;
          }
          p += (int)offset;
          pos[old + k] = p;
          off[old + k] = (uint32_t)(raw_dec + 20 * k);
          lp = p;
        }
        position = p; last_pos = lp; last_rank = rank;
        raw_dec += 20 * k;
        if (k < n) { pos.resize(old + k); off.resize(old + k); }
        if (k > 0 || dup_wait) continue;
      }
    }
    if (!fill(1)) { ended = true; break; }
    const unsigned char b0 = raw[raw_dec];
    const int type = b0 >> 4;
    if (type == 0) { raw_dec += 1; ended = true; break; }
    if (type == 1) {  // a base record that straddles the end of what has been read
      if (!fill(20)) { raw_dec = raw_end; ended = true; break; }
      continue;
    }
    if (type == 2) {  // indel: 17 fixed bytes + two allele strings, skipped (NextBaseEntry)
      if (!fill(17)) { raw_dec = raw_end; ended = true; break; }
      uint32_t offset;
      memcpy(&offset, raw.data() + raw_dec + 1, 4);
      int16_t len[2];
      memcpy(len, raw.data() + raw_dec + 13, 4);
      const size_t extra = (size_t)(len[0] < 0 ? -len[0] : len[0]) + (size_t)(len[1] < 0 ? -len[1] : len[1]);
      if (!fill(17 + extra)) { raw_dec = raw_end; ended = true; break; }
      position += (int)offset;
      raw_dec += 17 + extra;
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
  // thousands of files: opened and their headers read by the thread pool; the first failure in column order is reported
  std::vector<std::string> errs(paths.size());
  parallel_streams([&](int lo, int hi, int) {
    for (int i = lo; i < hi; i++) {
      if (paths[(size_t)i].empty()) continue;
      Stream &s = streams_[(size_t)i];
      const std::string &path = paths[(size_t)i];
      s.fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
      if (s.fd < 0) { errs[(size_t)i] = "GLF file " + path + " can  not be opened!"; continue; }
      unsigned char magic[2] = {0, 0};
      const bool gz = pread(s.fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
      if (gz) {  // gzip / BGZF: through zlib; anything else is read as it is (what gzread's transparent mode would do)
        s.f = gzdopen(s.fd, "rb");
        if (!s.f) { close(s.fd); s.fd = -1; errs[(size_t)i] = "GLF file " + path + " can  not be opened!"; continue; }
        gzbuffer(s.f, 1 << 17);
      }
      s.live = true;
      s.raw.resize(1 << 16);
      if (!s.fill(8) || memcmp(s.raw.data(), "GLF\3", 4) != 0) { errs[(size_t)i] = "GLF file " + path + ": invalid format or unsupported version"; continue; }
      uint32_t hl;
      memcpy(&hl, s.raw.data() + 4, 4);
      if (hl > 1024 * 1024) { errs[(size_t)i] = "GLF file " + path + ": header too large -- bailing"; continue; }
      s.raw_dec = 8;
      size_t left = hl;  // skip the header text
      while (left) {
        if (!s.fill(1)) { errs[(size_t)i] = "GLF file " + path + ": unexpected end of file"; break; }
        size_t take = std::min(left, s.raw_end - s.raw_dec);
        s.raw_dec += take; left -= take;
        s.raw_keep = s.raw_dec;
      }
      s.raw_keep = s.raw_dec;
    }
  });
  for (size_t i = 0; i < paths.size(); i++) {
    if (!errs[i].empty()) { if (err) *err = errs[i]; return false; }
    if (lead_ < 0 && streams_[i].live) lead_ = (int)i;
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
      if (!s.live) continue;
      while (!s.ended) { s.head = s.pos.size(); s.decode(4096); }  // drain the old section
      s.pos.clear(); s.off.clear(); s.head = 0; s.raw_keep = s.raw_dec;
      s.position = 0; s.last_pos = -1; s.last_rank = 0; s.dup_wait = false;
      int32_t label_len = 0;
      if (!s.fill(4)) { ok[(size_t)i] = 0; continue; }
      memcpy(&label_len, s.raw.data() + s.raw_dec, 4);
      s.raw_dec += 4;
      s.raw_keep = s.raw_dec;
      const size_t ll = (size_t)(label_len > 0 ? label_len : 0);
      if (!s.fill(ll + 4)) { ok[(size_t)i] = 0; continue; }
      s.label = std::string(std::string((const char *)s.raw.data() + s.raw_dec, ll).c_str());
      s.raw_dec += ll;
      memcpy(&s.max_position, s.raw.data() + s.raw_dec, 4);
      s.raw_dec += 4;
      s.raw_keep = s.raw_dec;
      s.ended = false;
      if (s.max_position <= 0) ok[(size_t)i] = 0;
    }
  });
  const Stream &lead = streams_[(size_t)lead_];
  // the reference walks the streams in order and stops at the first that has no further section
  for (size_t i = 0; i < streams_.size(); i++) {
    const Stream &s = streams_[i];
    if (!s.live) continue;
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
  const double t0 = now_s();
  // 1. every live stream gets at least `want` pending records (or reaches its end)
  size_t want = std::max<size_t>(64, std::min<size_t>(max_sites, ((size_t)96 << 20) / (np * 24)));
  parallel_streams([&](int lo, int hi, int) {
    for (int i = lo; i < hi; i++) {
      Stream &s = streams_[(size_t)i];
      if (s.live && !s.ended && s.pending() < want) s.decode(want);
    }
  });
  const double t1 = now_s();
  // 2. window: every position <= wend is completely known
  long long base = LLONG_MAX, wend = LLONG_MAX;
  // Sites are ordered by (position, rank), rank > 0 only for the further sites of a repeated position; kept as one number.
  auto site_key = [](long long p, long long rank) { return p < 0 ? -1LL : (p << 20) | std::min<long long>(rank, (1 << 20) - 1); };
  long long T = LLONG_MAX;  // min over ended streams of the site of their last base record (-1: none at all)
  for (const Stream &s : streams_) {
    if (!s.live) continue;
    if (s.pending()) base = std::min<long long>(base, s.pos[s.head]);
    if (s.ended) T = std::min<long long>(T, site_key(s.last_pos, s.last_rank));
    else wend = std::min<long long>(wend, s.pos.back());
  }
  if (base == LLONG_MAX) { section_done_ = true; return 0; }  // nothing left anywhere
  const long long cap = (long long)std::max<size_t>(1 << 16, 4 * max_sites);
  if (wend == LLONG_MAX) {  // every stream has been decoded to its end: the rest of the section is known
    wend = base;
    for (const Stream &s : streams_) if (s.live && s.pending()) wend = std::max<long long>(wend, s.pos.back());
  }
  wend = std::max(base, std::min(wend, base + cap - 1));
  const size_t width = (size_t)(wend - base + 1);
  mark_.assign(width, 0);
  parallel_streams([&](int lo, int hi, int) {
    for (int i = lo; i < hi; i++) {
      const Stream &s = streams_[(size_t)i];
      if (!s.live) continue;
      for (size_t k = s.head; k < s.pos.size() && s.pos[k] <= wend; k++) mark_[(size_t)(s.pos[k] - base)] = 1;
    }
  });
  const double t2 = now_s();
  // 3. rows, with the reference's termination rules
  rowpos_.clear();
  for (size_t w = 0; w < width && rowpos_.size() < max_sites; w++) {
    if (!mark_[w]) continue;
    const long long p = base + (long long)w;
    // Move2NextBaseEntry top check: some stream read its end marker in an earlier call
    if (prev1_ >= 0 && (prev1_ >> 20) > 0 && T <= prev2_) { section_done_ = true; break; }
    if (p > max_position_) { section_done_ = true; break; }
    rowpos_.push_back((int32_t)p);
    const long long rank = (prev1_ >= 0 && (prev1_ >> 20) == p) ? (prev1_ & ((1 << 20) - 1)) + 1 : 0;
    prev2_ = prev1_; prev1_ = site_key(p, rank);
  }
  const size_t n = rowpos_.size();
  if (n == 0) { section_done_ = true; return 0; }
  // 4. fill, row by row: a thread owns a contiguous range of columns and writes its stretch of every site's row in
  // one go (records and the zeros of people without a record alike), walking its streams' cursors in step.  The
  // reference base of a site comes from the lead stream if it has a record there, else from the lowest column that has
  // one: every thread reports its best (priority, base) per row, merged below.
  const double t3 = now_s();
  const int Tn = std::max(1, std::min(threads_, (int)np));
  owner_.assign(n * (size_t)Tn, UINT32_MAX);
  parallel_streams([&](int lo, int hi, int t) {
    // Blocks of 64 rows x tiles of 4 columns (one 64-byte line of a site's row): inside a block of rows the thread goes
    // through all its columns, so the lines it writes there — its stretch of 64 rows, 64 pages — stay in reach of the
    // TLB while every stream is still read sequentially (64 records at a time).  Measured on the GPU box's 16 cores, 3,000
    // streams: 180 k sites/s against 165 k with the columns outermost; non-temporal stores were slower (148 k).
    uint32_t *own = owner_.data() + (size_t)t * n;
    constexpr int CW = 4;
    constexpr size_t RB = 64;
    struct Cur { const int32_t *pp, *pe; const uint32_t *op; const unsigned char *raw; uint32_t prio; };
    std::vector<Cur> cur((size_t)(hi - lo));
    for (int c = lo; c < hi; c++) {
      const Stream &s = streams_[(size_t)c];
      Cur &k = cur[(size_t)(c - lo)];
      k.pp = s.pos.data() + s.head; k.pe = s.pos.data() + s.pos.size();
      k.op = s.off.data() + s.head; k.raw = s.raw.data();
      k.prio = ((c == lead_) ? 0u : (uint32_t)c + 1u) << 8;
    }
    const ConvertRun convert_run = pick_convert_run();
    const int32_t *const rowpos = rowpos_.data();
    const size_t stride = np * sizeof(pm_person_site);
    for (size_t r0 = 0; r0 < n; r0 += RB) {
      const size_t r1 = std::min(n, r0 + RB), m = r1 - r0;
      const int32_t *const rp = rowpos + r0;
      uint32_t *const ownr = own + r0;
      // The reference base of a row comes from the lowest column that has a record there (the lead stream is the lowest
      // live column).  The thread goes through its columns in increasing order, so the first record it sees at a row
      // settles the row; once every row of the block is settled the columns that follow skip the bookkeeping.
      size_t unowned = m;
      for (int c0 = lo; c0 < hi; c0 += CW) {
        const int cw = std::min(CW, hi - c0);
        // One column (stream) at a time through the block's rows, its cursor in registers; the four columns of a tile write
        // the same 64 cache lines one after the other.
        for (int j = 0; j < cw; j++) {
          Cur &k = cur[(size_t)(c0 + j - lo)];
          const int32_t *pp = k.pp, *const pe = k.pe;
          const uint32_t *op = k.op;
          const unsigned char *const raw = k.raw;
          unsigned char *dst = reinterpret_cast<unsigned char *>(out + r0 * np + (size_t)(c0 + j));
          // Runs of rows at which the stream has a record, separated by single rows at which it has none.  Both position
          // lists increase strictly and the stream's positions are among the rows', so "the stream's L-th pending record
          // sits at the L-th row from here" holds for a prefix of L and fails from there on: the end of a run is found by
          // bisection, and a stream with a record at every row left (the common case) by one compare.
          size_t i = 0;
          while (i < m) {
            const size_t lim = std::min(m - i, (size_t)(pe - pp));
            size_t run = lim;
            if (lim > 0 && pp[lim - 1] != rp[i + lim - 1]) {
              size_t ok = 0, bad = lim - 1;  // records [0, ok) match; record `bad` does not
              while (ok < bad) {
                const size_t mid = (ok + bad) / 2;
                if (pp[mid] == rp[i + mid]) ok = mid + 1; else bad = mid;
              }
              run = ok;
            }
            if (run) {
              convert_run(raw, op, run, dst, stride);
              if (unowned) {
                for (size_t q = 0; q < run; q++)
                  if (ownr[i + q] == UINT32_MAX) { ownr[i + q] = k.prio | kTranslateBase[raw[op[q]] & 0xf]; unowned--; }
              }
              pp += run; op += run; i += run; dst += run * stride;
            }
            if (i < m) { memset(dst, 0, sizeof(pm_person_site)); i++; dst += stride; }  // no record of this stream at this row
          }
          k.pp = pp; k.op = op;
        }
      }
    }
    for (int c = lo; c < hi; c++) {
      Stream &s = streams_[(size_t)c];
      s.head = (size_t)(cur[(size_t)(c - lo)].pp - s.pos.data());
    }
  });
  for (size_t r = 0; r < n; r++) {
    uint32_t best = UINT32_MAX;
    for (int t = 0; t < Tn; t++) best = std::min(best, owner_[(size_t)t * n + r]);
    hdr[r].pos = (uint32_t)rowpos_[r];
    hdr[r].ref_base = (uint8_t)(best & 0xff);
    hdr[r].chr_class = PM_CHR_AUTO;
    hdr[r].reserved = 0;
  }
  g_t[0] += t1 - t0; g_t[1] += t2 - t1; g_t[2] += t3 - t2; g_t[3] += now_s() - t3;
  return n;
}

}  // namespace pmh
