// cfr_main.cpp -- `centrifuger-b200`: drop-in for the reference's classification
// binary (CentrifugerClass.cpp) on the paths this repo covers: same -x/-1/-2/-u/-i/
// -t/-k/--min-hitlen/--hitk-factor/--no-dust/--consider-secondary/-h/-v options, the
// reference's own *.cfr index files, the identical TSV on stdout and the same log
// lines on stderr.  All classification work is done by libcfrb200.so on the GPU
// (include/centrifuger_b200.h); this file is host I/O only: gz FASTA/FASTQ parsing
// (ReadFiles.hpp + kseq.h behaviour), batching and ResultWriter-style output.
//
// Not supported (the reference's single-cell / output-side extras, SURVEY.md 8 "out
// of scope"): --un/--cl, --merge-readpair, --expand-taxid, barcode/UMI/read-format
// options, --sample-sheet.  They are rejected with a log line and EXIT_FAILURE.
#include <getopt.h>
#include <zlib.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/centrifuger_b200.h"

#define CENTRIFUGER_VERSION "1.1.3-r347"  // defs.h:8 of the reference this build mirrors

static const char usage[] =
    "./centrifuger-b200 [OPTIONS] > output.tsv:\n"
    "Required:\n"
    "\t-x FILE: index prefix\n"
    "\t-1 FILE -2 FILE: paired-end read\n"
    "\t\tor\n"
    "\t-u FILE: single-end read\n"
    "\t\tor\n"
    "\t-i FILE: interleaved read file\n"
    "Optional:\n"
    "\t-t INT: number of threads [1] (accepted for compatibility; the work runs on the GPU)\n"
    "\t-k INT: report upto <int> distinct, primary assignments for each read pair [1]\n"
    "\t--no-dust: do not DUST-mask low-complexity regions of reads [mask]\n"
    "\t--min-hitlen INT: minimum length of partial hits [auto]\n"
    "\t--hitk-factor INT: resolve at most <int>*k entries for each hit [40; use 0 for no restriction]\n"
    "\t--consider-secondary STR: in the format INT,FLOAT consider the secondary hit if its hitlen>=INT,score>=FLOAT*best_score [2000,0.995]\n"
    "\t--gpu INT: CUDA device ordinal [0]\n"
    "\t--batch INT: reads per GPU batch [1048576]\n"
    "\t--layout STR: in-HBM BWT layout, occ|runblock|auto [auto]\n"
    "\t-h: print this usage message\n"
    "\t-v: print the version information and quit\n";

enum {
  ARGV_NO_DUST = 256, ARGV_MIN_HITLEN, ARGV_HITK, ARGV_SECONDARY, ARGV_GPU, ARGV_BATCH, ARGV_LAYOUT,
  ARGV_UNSUPPORTED
};

static const char *short_options = "x:1:2:u:i:o:t:k:hv";
static struct option long_options[] = {
    {"no-dust", no_argument, 0, ARGV_NO_DUST},
    {"min-hitlen", required_argument, 0, ARGV_MIN_HITLEN},
    {"hitk-factor", required_argument, 0, ARGV_HITK},
    {"consider-secondary", required_argument, 0, ARGV_SECONDARY},
    {"gpu", required_argument, 0, ARGV_GPU},
    {"batch", required_argument, 0, ARGV_BATCH},
    {"layout", required_argument, 0, ARGV_LAYOUT},
    {"un", required_argument, 0, ARGV_UNSUPPORTED},
    {"cl", required_argument, 0, ARGV_UNSUPPORTED},
    {"merge-readpair", no_argument, 0, ARGV_UNSUPPORTED},
    {"expand-taxid", no_argument, 0, ARGV_UNSUPPORTED},
    {"read-format", required_argument, 0, ARGV_UNSUPPORTED},
    {"barcode", required_argument, 0, ARGV_UNSUPPORTED},
    {"UMI", required_argument, 0, ARGV_UNSUPPORTED},
    {"barcode-whitelist", required_argument, 0, ARGV_UNSUPPORTED},
    {"barcode-translate", required_argument, 0, ARGV_UNSUPPORTED},
    {"sample-sheet", required_argument, 0, ARGV_UNSUPPORTED},
    {(char *)0, 0, 0, 0}};

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

// Minimal FASTA/FASTQ reader with kseq.h semantics: name = first token after '>'/'@',
// sequence lines concatenated, FASTQ quality skipped by length; '-' = stdin; gz ok.
class SeqReader {
 public:
  bool open(const std::string &path) {
    fp_ = path == "-" ? gzdopen(fileno(stdin), "r") : gzopen(path.c_str(), "r");
    if (!fp_) return false;
    gzbuffer(fp_, 1 << 20);
    len_ = pos_ = 0;
    eof_ = false;
    last_ = 0;
    return true;
  }
  void close() {
    if (fp_) gzclose(fp_);
    fp_ = nullptr;
  }
  // returns false at end of file
  bool next(std::string &name, std::string &seq) {
    int c;
    if (last_ == 0) {
      while ((c = getc_()) != -1 && c != '>' && c != '@') {
      }
      if (c == -1) return false;
      last_ = c;
    }
    name.clear();
    seq.clear();
    // name up to the first space; the rest of the line is the comment
    while ((c = getc_()) != -1 && c != '\n' && c != ' ' && c != '\t' && c != '\r') name.push_back((char)c);
    while (c != -1 && c != '\n') c = getc_();
    // sequence lines
    while ((c = getc_()) != -1 && c != '>' && c != '+' && c != '@') {
      if (c == '\n' || c == '\r') continue;
      seq.push_back((char)c);
      while ((c = getc_()) != -1 && c != '\n')
        if (c != '\r') seq.push_back((char)c);
    }
    if (c == '>' || c == '@') {
      last_ = c;
      return true;
    }
    last_ = 0;
    if (c != '+') return true;  // FASTA at EOF
    // FASTQ: skip the '+' line, then read quality of the same length
    while ((c = getc_()) != -1 && c != '\n') {
    }
    size_t q = 0;
    while (q < seq.size() && (c = getc_()) != -1)
      if (c != '\n' && c != '\r') ++q;
    return true;
  }

 private:
  int getc_() {
    if (pos_ >= len_) {
      if (eof_) return -1;
      len_ = gzread(fp_, buf_, sizeof(buf_));
      pos_ = 0;
      if (len_ <= 0) {
        eof_ = true;
        len_ = 0;
        return -1;
      }
    }
    return (unsigned char)buf_[pos_++];
  }
  gzFile fp_ = nullptr;
  char buf_[1 << 16];
  int len_ = 0, pos_ = 0;
  bool eof_ = false;
  int last_ = 0;
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
  SeqReader rd;
  bool next(std::string &name, std::string &seq) {
    for (;;) {
      if (!opened) {
        if (cur >= files.size()) return false;
        if (!rd.open(files[cur])) {
          PrintLog("ERROR: cannot open read file %s", files[cur].c_str());
          exit(EXIT_FAILURE);
        }
        opened = true;
      }
      if (rd.next(name, seq)) return true;
      rd.close();
      opened = false;
      ++cur;
    }
  }
};

int main(int argc, char *argv[]) {
  if (argc <= 1) {  // CentrifugerClass.cpp:347-351: usage on stderr, exit 0
    fprintf(stderr, "%s", usage);
    return 0;
  }
  cfr_params params;
  cfr_default_params(&params);
  const char *idxPrefix = NULL;
  ReadSource reads, mates;
  bool hasMate = false, interleaved = false;
  int device = 0;
  long batchReads = 1 << 20;
  int c, option_index = 0;
  while ((c = getopt_long(argc, argv, short_options, long_options, &option_index)) != -1) {
    if (c == 'x') idxPrefix = optarg;
    else if (c == 'u') reads.files.push_back(optarg);
    else if (c == '1') { reads.files.push_back(optarg); hasMate = true; }
    else if (c == '2') { mates.files.push_back(optarg); hasMate = true; }
    else if (c == 'i') { reads.files.push_back(optarg); hasMate = true; interleaved = true; }
    else if (c == 'o') { /* parsed but unused by the reference as well (CentrifugerClass.cpp:416) */ }
    else if (c == 't') { /* host thread count of the reference; nothing to do */ }
    else if (c == 'k') params.max_result = atoi(optarg);
    else if (c == 'h') { fprintf(stdout, "%s", usage); return 0; }  // the reference prints -h to stdout (CentrifugerClass.cpp:987-991)
    else if (c == 'v') { printf("Centrifuger v" CENTRIFUGER_VERSION "\n"); return 0; }  // :428-432
    else if (c == ARGV_NO_DUST) params.dust = 0;
    else if (c == ARGV_MIN_HITLEN) params.min_hit_len = atoi(optarg);
    else if (c == ARGV_HITK) params.max_result_per_hit_factor = atoi(optarg);
    else if (c == ARGV_SECONDARY) {
      unsigned long len = 0;
      double f = 0;
      if (sscanf(optarg, "%lu,%lf", &len, &f) != 2) {
        PrintLog("Invalid format for --consider-secondary option. It should be in the format of INT,FLOAT");
        return EXIT_FAILURE;
      }
      params.consider_secondary_hit_len = len;
      params.consider_secondary_score_factor = f;
    } else if (c == ARGV_GPU) device = atoi(optarg);
    else if (c == ARGV_BATCH) batchReads = atol(optarg);
    else if (c == ARGV_LAYOUT) {
      if (!strcmp(optarg, "occ")) params.layout = CFR_LAYOUT_OCCLINE;
      else if (!strcmp(optarg, "runblock")) params.layout = CFR_LAYOUT_RUNBLOCK;
      else params.layout = CFR_LAYOUT_AUTO;
    } else if (c == ARGV_UNSUPPORTED) {
      PrintLog("Option --%s is not supported by the B200 classification path.", long_options[option_index].name);
      return EXIT_FAILURE;
    } else {
      fprintf(stderr, "%s", usage);
      return EXIT_FAILURE;
    }
  }
  PrintLog("Centrifuger v" CENTRIFUGER_VERSION " starts.");
  if (idxPrefix == NULL) {
    PrintLog("Need to use -x to specify index prefix.");
    return EXIT_FAILURE;
  }
  if (batchReads < 1) batchReads = 1;
  params.max_batch_reads = (int32_t)(batchReads > (1 << 22) ? (1 << 22) : batchReads);

  cfr_handle *h = NULL;
  int st = cfr_open(idxPrefix, &params, device, &h);
  if (st != CFR_OK) {
    PrintLog("ERROR: %s", cfr_last_error());
    return EXIT_FAILURE;
  }
  PrintLog("Finishes loading index.");
  if (params.min_hit_len <= 0) PrintLog("Inferred --min-hitlen: %d", (int)cfr_index_info(h, 4));

  // ResultWriter::OutputHeader (ResultWriter.hpp:186-197)
  std::string out;
  out.reserve(64 << 20);
  out += "readID\tseqID\ttaxID\tscore\t2ndBestScore\thitLength\tqueryLength\tnumMatches\n";

  const int k = params.max_result;
  std::vector<std::string> ids;
  std::string seq1, seq2, name, s, name2;
  std::vector<uint64_t> off1, off2, assign;
  std::vector<cfr_result> results;
  unsigned long totalCnt = 0, classifiedCnt = 0;
  char line[1 << 16];
  for (;;) {
    ids.clear();
    seq1.clear();
    seq2.clear();
    off1.assign(1, 0);
    off2.assign(1, 0);
    while ((long)ids.size() < batchReads) {
      if (!reads.next(name, s)) break;
      RemoveReadIdSuffix(name);
      ids.push_back(name);
      seq1 += s;
      off1.push_back(seq1.size());
      if (hasMate) {
        bool ok = interleaved ? reads.next(name2, s) : mates.next(name2, s);
        if (!ok) {
          PrintLog("ERROR: The two mate-pair read files have different number of reads.");
          return EXIT_FAILURE;
        }
        seq2 += s;
        off2.push_back(seq2.size());
      }
    }
    if (hasMate && !interleaved && (long)ids.size() < batchReads) {  // mate 1 ended: mate 2 must end too
      if (mates.next(name2, s)) {
        PrintLog("ERROR: The two mate-pair read files have different number of reads.");
        return EXIT_FAILURE;
      }
    }
    if (ids.empty()) break;
    const size_t n = ids.size();
    results.resize(n);
    assign.resize(n * (size_t)k);
    cfr_read_batch b;
    b.n_reads = n;
    b.seq1 = seq1.data();
    b.off1 = off1.data();
    b.seq2 = hasMate ? seq2.data() : NULL;
    b.off2 = hasMate ? off2.data() : NULL;
    st = cfr_classify_batch(h, &b, results.data(), assign.data(), NULL);
    if (st != CFR_OK) {
      PrintLog("ERROR: %s", cfr_last_error());
      return EXIT_FAILURE;
    }
    for (size_t i = 0; i < n; ++i) {  // ResultWriter::Output (ResultWriter.hpp:199-236)
      const int w = cfr_format_tsv(h, ids[i].c_str(), &results[i], &assign[i * (size_t)k], line, sizeof(line));
      if (w < 0) {
        PrintLog("ERROR: output row too long for read %s", ids[i].c_str());
        return EXIT_FAILURE;
      }
      out.append(line, (size_t)w);
      ++totalCnt;
      if (results[i].n_assign > 0) ++classifiedCnt;
      if (out.size() > (48u << 20)) {
        fwrite(out.data(), 1, out.size(), stdout);
        out.clear();
      }
    }
  }
  fwrite(out.data(), 1, out.size(), stdout);
  fflush(stdout);
  // ResultWriter::Finalize (ResultWriter.hpp:279-283)
  PrintLog("Processed %lu read fragments, and %lu (%.2lf%%) can be classified.", totalCnt, classifiedCnt,
           (double)classifiedCnt / (double)totalCnt * 100.0);
  cfr_close(h);
  PrintLog("Centrifuger finishes.");
  return 0;
}
