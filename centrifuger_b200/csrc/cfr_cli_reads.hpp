// cfr_cli_reads.hpp -- part of the `centrifuger-b200` command line (host I/O only, see cfr_main.cpp).
#pragma once
#include <glob.h>
#include <zlib.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

// Utils::PrintLog (compactds/Utils.hpp:369-381)
static void PrintLog(const char *fmt, ...) {
  va_list args;
  va_start(args, fmt);
  char buffer[1000];
  vsnprintf(buffer, sizeof(buffer), fmt, args);
  va_end(args);
  time_t mytime = time(NULL);
  struct tm *localT = localtime(&mytime);
  char stime[500];
  strftime(stime, sizeof(stime), "%c", localT);
  fprintf(stderr, "[%s] %s\n", stime, buffer);
}

// FASTA/FASTQ reader with kseq.h semantics (name = first token after '>'/'@', sequence
// lines concatenated, FASTQ quality skipped by length; '-' = stdin; gz ok), block-buffered
// with memchr line splitting.
class SeqReader {
 public:
  bool open(const std::string &path) {
    fp_ = path == "-" ? gzdopen(fileno(stdin), "r") : gzopen(path.c_str(), "r");
    if (!fp_) return false;
    gzbuffer(fp_, 1 << 20);
    buf_.resize(4 << 20);
    len_ = pos_ = 0;
    eof_ = false;
    dead_ = false;
    return true;
  }
  void close() {
    if (fp_) gzclose(fp_);
    fp_ = nullptr;
  }
  // continue at byte `offset` of the (uncompressed) file: the block-parallel reader hands over here
  bool seek(size_t offset) {
    if (offset == 0) return true;
    if (gzseek(fp_, (z_off_t)offset, SEEK_SET) < 0) return false;
    len_ = pos_ = 0;
    eof_ = false;
    return true;
  }
  // appends the record's sequence to `seq` (and, for FASTQ records, its quality string to `qual` when
  // given: a FASTA record appends nothing there); returns false at end of file
  bool next(std::string &name, std::string &seq, std::string *qual = nullptr, std::string *comment = nullptr) {
    const char *ln;
    size_t n;
    if (dead_) return false;
    // header line
    for (;;) {
      if (!line(ln, n)) return false;
      if (n > 0 && (ln[0] == '>' || ln[0] == '@')) break;
    }
    const bool fastq = ln[0] == '@';
    size_t e = 1;
    while (e < n && ln[e] != ' ' && ln[e] != '\t') ++e;
    name.assign(ln + 1, e - 1);
    if (comment) comment->assign(e < n ? ln + e + 1 : ln + n, e < n ? n - e - 1 : 0);  // kseq: the rest of the line
    // sequence lines: until a line starting with '+' (FASTQ), '>' or '@' (next record)
    const size_t start = seq.size();
    for (;;) {
      if (!peek_line(ln, n)) return true;  // EOF ends the record
      if (n > 0 && (ln[0] == '+' || ln[0] == '>' || ln[0] == '@')) break;
      seq.append(ln, n);
      // kseq drops a line's trailing '\r' only once the record holds more than one character: an otherwise
      // empty CRLF line at the start of a record leaves a one-character sequence "\r" (kseq.h:146)
      if (n == 0 && cr_ && seq.size() == start) seq += '\r';
      consume();
    }
    if (ln[0] != '+') return true;  // FASTA: next header stays in the buffer
    (void)fastq;
    consume();  // the '+' line
    size_t q = 0;
    const size_t want = seq.size() - start, qstart = qual ? qual->size() : 0;
    // kseq_read returns an error -- which ends the FILE for ReadFiles::Next -- when the stream stops inside
    // the '+' line, when the quality string is cut short, or when it comes out longer than the sequence
    // (kseq.h:212-218); the record is dropped in all three cases
    bool broken = !nl_;
    while (!broken && q < want) {  // quality lines (may start with '@' or '+'): by length
      if (!line(ln, n)) {
        broken = true;
        break;
      }
      if (qual) qual->append(ln, n);
      q += n;
      if (n == 0 && cr_ && q == 0) {  // the same rule for the quality string
        if (qual) *qual += '\r';
        ++q;
      }
    }
    if (broken || q != want) {
      seq.resize(start);
      if (qual) qual->resize(qstart);
      dead_ = true;
      return false;
    }
    return true;
  }

 private:
  // returns the next line without its terminator ('\r' stripped) and consumes it
  bool line(const char *&p, size_t &n) {
    if (!peek_line(p, n)) return false;
    consume();
    return true;
  }
  bool peek_line(const char *&p, size_t &n) {
    for (;;) {
      const char *nl = (const char *)memchr(buf_.data() + pos_, '\n', len_ - pos_);
      if (nl) {
        p = buf_.data() + pos_;
        n = (size_t)(nl - p);
        next_ = pos_ + n + 1;
        nl_ = true;
        cr_ = n > 0 && p[n - 1] == '\r';
        if (cr_) --n;
        return true;
      }
      if (eof_) {
        if (pos_ >= len_) return false;
        p = buf_.data() + pos_;  // last line without '\n'
        n = len_ - pos_;
        next_ = len_;
        nl_ = false;
        cr_ = n > 0 && p[n - 1] == '\r';
        if (cr_) --n;
        return true;
      }
      // refill: keep the partial line at the front
      if (pos_ > 0) {
        memmove(&buf_[0], buf_.data() + pos_, len_ - pos_);
        len_ -= pos_;
        pos_ = 0;
      }
      if (len_ == buf_.size()) buf_.resize(buf_.size() * 2);
      const int got = gzread(fp_, &buf_[len_], (unsigned)std::min<size_t>(buf_.size() - len_, 1u << 30));
      if (got <= 0) eof_ = true; else len_ += (size_t)got;
    }
  }
  void consume() { pos_ = next_; }
  gzFile fp_ = nullptr;
  std::string buf_;
  size_t len_ = 0, pos_ = 0, next_ = 0;
  bool eof_ = false;
  bool cr_ = false;  // the line peek_line() returned last ended in "\r\n"
  bool nl_ = true;   // ... and had a line terminator at all (false: the stream ended inside it)
  bool dead_ = false;  // a broken FASTQ record ended this file
};

// ReadFiles::RemoveReadIdSuffix (ReadFiles.hpp:82-90)
static void RemoveReadIdSuffix(std::string &id) {
  const size_t len = id.size();
  if (len >= 2 && (id[len - 1] == '1' || id[len - 1] == '2') && id[len - 2] == '/') id.resize(len - 2);
}

struct ReadSource {  // a list of files read back to back (ReadFiles::AddReadFile)
  std::vector<std::string> files;
  size_t cur = 0;
  bool opened = false;
  bool markFileEnds = false;  // --sample-sheet: every file end is reported (ReadFiles::SetSpecialReadToMarkFileEnd)
  SeqReader rd;
  // a name with '*' stands for the files it matches, in glob(3) order (ReadFiles.hpp:135-172)
  void add(const char *file) {
    if (!strchr(file, '*')) {
      files.push_back(file);
      return;
    }
    glob_t g;
    memset(&g, 0, sizeof(g));
    const int rc = glob(file, GLOB_TILDE, NULL, &g);
    if (rc != 0) fprintf(stderr, "glob() failed with return value %d.\n", rc);
    for (size_t i = 0; rc == 0 && i < g.gl_pathc; ++i) files.push_back(g.gl_pathv[i]);
    globfree(&g);
  }
  enum { END = 0, RECORD = 1, FILE_END = 2 };
  // RECORD, END (no file left) or -- with markFileEnds -- FILE_END once per file, the last one included
  int step(std::string &name, std::string &seq, std::string *qual = nullptr, std::string *comment = nullptr) {
    for (;;) {
      if (!opened) {
        if (cur >= files.size()) return END;
        if (!rd.open(files[cur])) {
          PrintLog("ERROR: cannot open read file %s", files[cur].c_str());
          exit(EXIT_FAILURE);
        }
        opened = true;
      }
      if (rd.next(name, seq, qual, comment)) return RECORD;
      rd.close();
      opened = false;
      ++cur;
      if (markFileEnds) return FILE_END;
    }
  }
  bool next(std::string &name, std::string &seq, std::string *qual = nullptr, std::string *comment = nullptr) {
    int r;
    while ((r = step(name, seq, qual, comment)) == FILE_END) {
    }
    return r == RECORD;
  }
};
