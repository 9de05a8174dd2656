// cfr_cli_bulk.hpp -- part of the `centrifuger-b200` command line (host I/O only, see cfr_main.cpp).
//
// Block-parallel ingest for the common case: plain (not gzip'ed) FASTQ files of four-line records.  The reference parses
// on one thread (ReadFiles::NextBatch over kseq.h, ReadFiles.hpp:337); once the classification runs on the GPU that parser
// is the whole run time.  Here a file is mmap'ed and cut into super-blocks; T threads parse one super-block together
// (each takes a byte range, finds the first record that STARTS in its range and parses until a record starts past it),
// and the batch assembler copies record ranges out of the parsed sub-blocks.
//
// The fast path accepts only what it can parse exactly like kseq: records of exactly four lines ('@' header, one
// sequence line that does not start with '+', '>' or '@', a '+' line, a quality line as long as the sequence), no '\r',
// no blank lines, no empty or very long reads, and neighbouring threads must agree on where each other's records begin
// and end.  Anything else -- FASTA, multi-line records, CRLF, a cut-off last record -- sends that file, from the start of
// the super-block in question, to the serial SeqReader, which restates kseq's behaviour for every irregular input
// (tests/fuzz/fuzz_cli_parser.py); so does any file that is gzip'ed, a pipe, or stdin.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "cfr_cli_reads.hpp"

struct BulkSink {  // where records are appended (a Batch's arrays; ids / qual may be null)
  std::string *ids;
  std::vector<uint32_t> *id_off;
  std::string *seq;
  std::vector<uint64_t> *off;
  std::string *qual;
  std::vector<uint64_t> *qoff;
};

class BulkReader {
 public:
  // files are read back to back (ReadFiles::AddReadFile); `threads` parse a super-block together
  void init(const std::vector<std::string> &files, unsigned threads, bool allow_fast) {
    files_ = files;
    threads_ = std::max(1u, threads);
    allow_fast_ = allow_fast;
  }
  ~BulkReader() { unmap(); }

  // appends up to `want` records to the sink, stopping early once the sink's sequence bytes reach max_bases;
  // returns how many (fewer than `want` without the byte limit being hit: every file has ended)
  size_t take(size_t want, size_t max_bases, const BulkSink &out, size_t *max_len) {
    size_t got = 0;
    while (got < want && out.seq->size() < max_bases) {
      if (parked_) {  // the record exhausted() looked at
        take_parked(out, max_len);
        ++got;
        continue;
      }
      if (!ensure()) break;
      if (serial_) {
        name_.clear();
        const size_t before = out.seq->size();
        if (!rd_.next(name_, *out.seq, out.qual)) {
          rd_.close();
          serial_ = false;
          opened_ = false;
          ++cur_;
          continue;
        }
        RemoveReadIdSuffix(name_);
        if (out.ids) {
          *out.ids += name_;
          out.id_off->push_back((uint32_t)out.ids->size());
        }
        out.off->push_back(out.seq->size());
        if (out.qoff) out.qoff->push_back(out.qual->size());
        if (max_len) *max_len = std::max(*max_len, out.seq->size() - before);
        ++got;
        continue;
      }
      // Fast mode: plan which record ranges of the parsed sub-blocks go into the sink (as many as wanted and as the byte limit
      // lets through, at least one record), size the sink's arrays once without touching the new bytes, and let the
      // threads copy the ranges side by side -- a serial memcpy of the batch costs more than parsing it did.
      struct Seg {
        Sub *sb;
        size_t from, n, seq_dst, id_dst, rec_dst, qual_dst;
      };
      std::vector<Seg> segs;
      size_t seq_at = out.seq->size(), id_at = out.ids ? out.ids->size() : 0, rec_at = out.off->size();
      size_t qual_at = out.qual ? out.qual->size() : 0;  // (FASTA records served by the serial reader have no qualities)
      const size_t qrec0 = out.qoff ? out.qoff->size() : 0, rec0 = rec_at;
      size_t planned = 0;
      for (size_t si = sub_at_; si < subs_.size() && got + planned < want && seq_at < max_bases; ++si) {
        Sub &sb = subs_[si];
        if (sb.taken >= sb.n) continue;
        size_t n = std::min(want - got - planned, sb.n - sb.taken);
        const uint64_t s0 = sb.off[sb.taken];
        const size_t room = max_bases - seq_at;
        bool cut = false;  // the byte limit ends the plan inside this sub-block: nothing behind it may be taken
        if (sb.off[sb.taken + n] - s0 > room) {
          const auto it = std::upper_bound(sb.off.begin() + (long)sb.taken, sb.off.begin() + (long)(sb.taken + n) + 1, s0 + room);
          n = std::max<size_t>(1, (size_t)(it - (sb.off.begin() + (long)sb.taken)) - 1);
          cut = true;
        }
        segs.push_back(Seg{&sb, sb.taken, n, seq_at, id_at, rec_at, qual_at});
        seq_at += (size_t)(sb.off[sb.taken + n] - s0);
        qual_at += (size_t)(sb.off[sb.taken + n] - s0);
        id_at += sb.id_off[sb.taken + n] - sb.id_off[sb.taken];
        rec_at += n;
        planned += n;
        if (max_len) *max_len = std::max(*max_len, sb.max_len);
        if (cut) break;
      }
      grow_uninitialized(*out.seq, seq_at);
      if (out.ids) grow_uninitialized(*out.ids, id_at);
      if (out.qual) grow_uninitialized(*out.qual, qual_at);
      out.off->resize(rec_at);
      if (out.ids) out.id_off->resize(rec_at);
      if (out.qoff) out.qoff->resize(qrec0 + (rec_at - rec0));
      auto copy = [&](const Seg &g) {
        const Sub &sb = *g.sb;
        const uint64_t s0 = sb.off[g.from], s1 = sb.off[g.from + g.n];
        memcpy(&(*out.seq)[g.seq_dst], sb.seq.data() + s0, (size_t)(s1 - s0));
        const uint64_t base = g.seq_dst - s0;
        for (size_t i = 0; i < g.n; ++i) (*out.off)[g.rec_dst + i] = sb.off[g.from + 1 + i] + base;
        if (out.ids) {
          const uint32_t i0 = sb.id_off[g.from], i1 = sb.id_off[g.from + g.n];
          memcpy(&(*out.ids)[g.id_dst], sb.ids.data() + i0, i1 - i0);
          const uint32_t ibase = (uint32_t)g.id_dst - i0;
          for (size_t i = 0; i < g.n; ++i) (*out.id_off)[g.rec_dst + i] = sb.id_off[g.from + 1 + i] + ibase;
        }
        if (out.qual) {  // a quality string is as long as its read: the same offsets
          memcpy(&(*out.qual)[g.qual_dst], sb.qual.data() + s0, (size_t)(s1 - s0));
          const uint64_t qbase = g.qual_dst - s0;
          for (size_t i = 0; i < g.n; ++i) (*out.qoff)[qrec0 + (g.rec_dst - rec0) + i] = sb.off[g.from + 1 + i] + qbase;
        }
      };
      if (segs.size() > 1 && threads_ > 1 && seq_at - segs[0].seq_dst > (1u << 20)) {
        std::vector<std::thread> pool;
        const size_t T = std::min<size_t>(threads_, segs.size());
        for (size_t t = 1; t < T; ++t)
          pool.emplace_back([&, t] {
            for (size_t k = t; k < segs.size(); k += T) copy(segs[k]);
          });
        for (size_t k = 0; k < segs.size(); k += T) copy(segs[k]);
        for (auto &th : pool) th.join();
      } else {
        for (const Seg &g : segs) copy(g);
      }
      for (const Seg &g : segs) g.sb->taken += g.n;
      got += planned;
    }
    return got;
  }

  // std::string::resize writes zeros over the new bytes -- a pass over a 150 MB batch that the copy repeats at once
  static void grow_uninitialized(std::string &str, size_t n) {
#if defined(__cpp_lib_string_resize_and_overwrite)
    str.resize_and_overwrite(n, [n](char *, size_t) { return n; });
#else
    str.resize(n);
#endif
  }

  // true when no record is left in any file
  bool exhausted() {
    if (!ensure()) return true;
    if (!serial_) return false;
    // serial mode has no look-ahead: park one record
    if (parked_) return false;
    for (;;) {
      park_name_.clear();
      park_seq_.clear();
      park_qual_.clear();
      if (rd_.next(park_name_, park_seq_, &park_qual_)) {
        parked_ = true;
        return false;
      }
      rd_.close();
      serial_ = false;
      opened_ = false;
      ++cur_;
      if (!ensure()) return true;
      if (!serial_) return false;
    }
  }
  bool used_fast_path() const { return fast_blocks_ > 0; }
  // records are being served from parsed sub-blocks (reads of at most 1000 bases): a caller may ask for many at once
  bool in_fast_mode() const { return !serial_ && !parked_ && map_ != nullptr; }
  double parse_seconds() const { return parse_seconds_; }  // wall time inside the parallel parse

 private:
  struct Sub {
    std::string ids, seq, qual;
    std::vector<uint32_t> id_off;
    std::vector<uint64_t> off;
    size_t n = 0, taken = 0, begin = 0, end = 0, max_len = 0;
    bool ok = true;
  };

  // makes a record available: a sub-block with records left (fast mode) or an open serial reader; false = all files done
  bool ensure() {
    for (;;) {
      if (serial_) return true;
      while (sub_at_ < subs_.size() && subs_[sub_at_].taken >= subs_[sub_at_].n) ++sub_at_;
      if (sub_at_ < subs_.size()) return true;
      if (opened_ && map_ && pos_ < size_) {
        if (parse_super_block()) continue;
        // irregular input: the serial reader takes over this file from the start of the super-block
        start_serial(pos_);
        continue;
      }
      if (opened_) {  // the mapped file is done
        unmap();
        opened_ = false;
        ++cur_;
      }
      if (cur_ >= files_.size()) return false;
      opened_ = true;
      if (!(allow_fast_ && map_file(files_[cur_]))) start_serial(0);
    }
  }

  void start_serial(size_t offset) {
    unmap();
    subs_.clear();
    sub_at_ = 0;
    if (!rd_.open(files_[cur_]) || !rd_.seek(offset)) {
      PrintLog("ERROR: cannot open read file %s", files_[cur_].c_str());
      exit(EXIT_FAILURE);
    }
    serial_ = true;
  }

  bool map_file(const std::string &path) {
    if (path == "-") return false;
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;  // the serial reader reports the error
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 4) {
      close(fd);
      return false;
    }
    void *p = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return false;
    const unsigned char *b = (const unsigned char *)p;
    if ((b[0] == 0x1f && b[1] == 0x8b) || b[0] != '@') {  // gzip, FASTA, anything unusual: serial
      munmap(p, (size_t)st.st_size);
      return false;
    }
    madvise(p, (size_t)st.st_size, MADV_SEQUENTIAL);
    map_ = (const char *)p;
    size_ = (size_t)st.st_size;
    pos_ = 0;
    return true;
  }
  void unmap() {
    if (map_) munmap((void *)map_, size_);
    map_ = nullptr;
    size_ = pos_ = 0;
  }

  static const char *find_nl(const char *p, const char *e) { return p < e ? (const char *)memchr(p, '\n', (size_t)(e - p)) : nullptr; }

  // does a four-line record start at p (p = a line start)?  Quality lines may begin with '@', so the test looks two
  // lines ahead for the '+' line: a quality line is followed by a header and a sequence line, never by a '+' line there.
  bool record_starts_at(const char *p) const {
    const char *e = map_ + size_;
    if (p >= e || *p != '@') return false;
    const char *n1 = find_nl(p, e);
    if (!n1) return false;
    const char *n2 = find_nl(n1 + 1, e);
    if (!n2 || n2 + 1 >= e) return false;
    return n2[1] == '+';
  }

  // records that START in [c0, c1) (the first one exactly at c0 when `at_start`)
  void parse_range(size_t c0, size_t c1, bool at_start, bool want_qual, Sub &sb) const {
    const char *e = map_ + size_;
    const char *p = map_ + c0;
    sb.ok = true;
    sb.n = sb.taken = 0;
    sb.max_len = 0;
    sb.ids.clear();
    sb.seq.clear();
    sb.qual.clear();
    sb.id_off.assign(1, 0);
    sb.off.assign(1, 0);
    if (!at_start) {
      if (p[-1] != '\n') {
        const char *nl = find_nl(p, e);
        p = nl ? nl + 1 : e;
      }
      int tries = 0;
      while (p < e && !record_starts_at(p)) {
        const char *nl = find_nl(p, e);
        p = nl ? nl + 1 : e;
        if (++tries > 8) {  // four-line records: one of any four consecutive lines starts a record
          sb.ok = false;
          break;
        }
      }
    }
    sb.begin = (size_t)(p - map_);
    const size_t est = c1 > c0 ? c1 - c0 : 0;
    sb.seq.reserve(est / 2 + 4096);
    if (want_qual) sb.qual.reserve(est / 2 + 4096);
    sb.ids.reserve(est / 8 + 1024);
    while (sb.ok && p < e && (size_t)(p - map_) < c1) {
      if (*p != '@') { sb.ok = false; break; }
      const char *h1 = find_nl(p, e);
      if (!h1) { sb.ok = false; break; }
      const char *s0 = h1 + 1;
      const char *s1 = find_nl(s0, e);
      if (!s1) { sb.ok = false; break; }
      const char *pl = s1 + 1;
      const char *p1 = find_nl(pl, e);
      if (!p1 || *pl != '+') { sb.ok = false; break; }
      const char *q0 = p1 + 1;
      const char *q1 = find_nl(q0, e);
      const char *next = q1 ? q1 + 1 : e;
      if (!q1) q1 = e;  // kseq takes a last quality line without a terminator
      const size_t sl = (size_t)(s1 - s0);
      if (sl == 0 || sl > 1000 || (size_t)(q1 - q0) != sl || s0[0] == '+' || s0[0] == '>' || s0[0] == '@' || h1[-1] == '\r' ||
          s1[-1] == '\r' || q1[-1] == '\r' || (next < e && *next != '@')) {
        sb.ok = false;
        break;
      }
      // read id: the first token of the header, a trailing /1 or /2 removed (ReadFiles.hpp:82-90)
      const char *t = p + 1;
      while (t < h1 && *t != ' ' && *t != '\t') ++t;
      size_t idn = (size_t)(t - (p + 1));
      if (idn >= 2 && (p[idn] == '1' || p[idn] == '2') && p[idn - 1] == '/') idn -= 2;
      sb.ids.append(p + 1, idn);
      sb.id_off.push_back((uint32_t)sb.ids.size());
      sb.seq.append(s0, sl);
      sb.off.push_back(sb.seq.size());
      if (want_qual) sb.qual.append(q0, sl);
      if (sl > sb.max_len) sb.max_len = sl;
      ++sb.n;
      p = next;
    }
    sb.end = (size_t)(p - map_);
  }

  bool parse_super_block() {
    struct Timer {
      double &acc;
      std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
      ~Timer() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
    } timer{parse_seconds_};
    // bytes per thread and super-block (CFR_B200_INGEST_BLOCK: tests make the ranges a few hundred bytes long)
    static const size_t per = getenv("CFR_B200_INGEST_BLOCK") ? (size_t)std::max(16ll, atoll(getenv("CFR_B200_INGEST_BLOCK"))) : (8u << 20);
    const size_t left = size_ - pos_;
    const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(threads_, (left + per - 1) / per));
    const size_t span = std::min(left, per * T);
    const size_t stop = pos_ + span;
    subs_.resize(T);
    sub_at_ = 0;
    const bool wq = want_qual_;
    auto run = [&](unsigned t) {
      const size_t c0 = pos_ + span * t / T, c1 = t + 1 == T ? stop : pos_ + span * (t + 1) / T;
      parse_range(c0, c1, t == 0, wq, subs_[t]);
    };
    if (T == 1) {
      run(0);
    } else {
      std::vector<std::thread> pool;
      for (unsigned t = 1; t < T; ++t) pool.emplace_back(run, t);
      run(0);
      for (auto &th : pool) th.join();
    }
    if (subs_[0].begin != pos_) return false;
    for (unsigned t = 0; t < T; ++t) {
      if (!subs_[t].ok) return false;
      if (t + 1 < T && subs_[t].end != subs_[t + 1].begin) return false;
    }
    pos_ = subs_[T - 1].end;
    ++fast_blocks_;
    return true;
  }

 public:
  void set_want_qual(bool w) { want_qual_ = w; }
  // serial mode only: the record exhausted() parked is handed out first
  bool take_parked(const BulkSink &out, size_t *max_len) {
    if (!parked_) return false;
    parked_ = false;
    RemoveReadIdSuffix(park_name_);
    if (out.ids) {
      *out.ids += park_name_;
      out.id_off->push_back((uint32_t)out.ids->size());
    }
    *out.seq += park_seq_;
    out.off->push_back(out.seq->size());
    if (out.qual) {
      *out.qual += park_qual_;
      out.qoff->push_back(out.qual->size());
    }
    if (max_len) *max_len = std::max(*max_len, park_seq_.size());
    return true;
  }

 private:
  std::vector<std::string> files_;
  size_t cur_ = 0;
  unsigned threads_ = 1;
  bool allow_fast_ = true, opened_ = false, serial_ = false, want_qual_ = false, parked_ = false;
  const char *map_ = nullptr;
  size_t size_ = 0, pos_ = 0;
  std::vector<Sub> subs_;
  size_t sub_at_ = 0, fast_blocks_ = 0;
  double parse_seconds_ = 0;
  SeqReader rd_;
  std::string name_, park_name_, park_seq_, park_qual_;
};
