// cfr_main.cpp -- `centrifuger-b200`: drop-in for the reference's classification
// binary (CentrifugerClass.cpp): the same options (-x/-1/-2/-u/-i/--sample-sheet/-t/-k/--min-hitlen/
// --hitk-factor/--no-dust/--consider-secondary/--un/--cl/--merge-readpair/--expand-taxid/--read-format/
// --barcode/--UMI/--barcode-whitelist/--barcode-translate/-h/-v), the reference's own *.cfr index files,
// the identical TSV on stdout and the same log lines on stderr.  All classification work is done by
// libcfrb200.so on the GPU (include/centrifuger_b200.h); this file is host I/O only: gz FASTA/FASTQ
// parsing (ReadFiles.hpp + kseq.h behaviour) on an ingest thread (mate 2 on a second one), batching, read
// stretches / barcodes / UMIs (ReadFormatter.hpp, BarcodeCorrector.hpp, BarcodeTranslator.hpp behaviour),
// the read-pair merger (ReadPairMerger.hpp behaviour) and ResultWriter-style output on an output thread.
// Protein indexes are refused by the library at load.
#include <getopt.h>
#include <glob.h>
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <chrono>
#include <thread>
#include <unordered_map>
#include <deque>
#include <iterator>
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
    "\t\tor\n"
    "\t--sample-sheet FILE: list of sample files, each row: \"read1 read2 barcode UMI output\". Use dot(.) to represent no such file\n"
    "Optional:\n"
    "\t-t INT: number of threads [1] (the classification runs on the GPU; given, it caps the host threads that parse the reads)\n"
    "\t-k INT: report upto <int> distinct, primary assignments for each read pair [1]\n"
    "\t--un STR: output unclassified reads to files with the prefix of <str>\n"
    "\t--cl STR: output classified reads to files with the prefix of <str>\n"
    "\t--merge-readpair: merge overlapped paired-end reads and trim adapters [no merge]\n"
    "\t--barcode STR: path to the barcode file\n"
    "\t--UMI STR: path to the UMI file\n"
    "\t--read-format STR: format for read, barcode and UMI files, e.g. r1:0:-1,r2:0:-1,bc:0:15,um:16:-1 for paired-end files with barcode and UMI\n"
    "\t--barcode-whitelist STR: path to the barcode whitelist file\n"
    "\t--barcode-translate STR: path to the barcode translation file\n"
    "\t--expand-taxid: output the tax IDs that are promoted to the final report tax ID [no]\n"
    "\t--no-dust: do not DUST-mask low-complexity regions of reads [mask]\n"
    "\t--min-hitlen INT: minimum length of partial hits [auto]\n"
    "\t--hitk-factor INT: resolve at most <int>*k entries for each hit [40; use 0 for no restriction]\n"
    "\t--consider-secondary STR: in the format INT,FLOAT consider the secondary hit if its hitlen>=INT,score>=FLOAT*best_score [2000,0.995]\n"
    "\t--gpu INT: CUDA device ordinal [0]\n"
    "\t--quant-report FILE: also write the abundance report `centrifuger-quant` computes from this output (- = stderr is not\n"
    "\t            touched; the report goes to FILE); --quant-format INT (0:centrifuge 1:metaphlan 2:CAMI 3:kraken-report) [0],\n"
    "\t            --quant-min-score INT, --quant-min-length INT as centrifuger-quant's --min-score / --min-length [0]\n"
    "\t--gpus INT: classify on INT GPUs (devices --gpu .. --gpu + INT - 1): the index is replicated per GPU, batch j goes\n"
    "\t            to GPU j mod INT, rows stay in input order; the per-taxon counters are summed over NCCL at the end [1]\n"
    "\t--batch INT: reads per GPU batch [1048576]\n"
    "\t--layout STR: in-HBM BWT layout, occ|runblock|auto [auto]\n"
    "\t-h: print this usage message\n"
    "\t-v: print the version information and quit\n";

enum {
  ARGV_NO_DUST = 256, ARGV_MIN_HITLEN, ARGV_HITK, ARGV_SECONDARY, ARGV_GPU, ARGV_GPUS, ARGV_QUANT_REPORT, ARGV_QUANT_FORMAT, ARGV_QUANT_MIN_SCORE, ARGV_QUANT_MIN_LENGTH, ARGV_BATCH, ARGV_LAYOUT,
  ARGV_UNSUPPORTED, ARGV_DRY_RUN, ARGV_DRY_PIPE, ARGV_UN, ARGV_CL, ARGV_MERGE, ARGV_EXPAND, ARGV_SAMPLE_SHEET, ARGV_DRY_OUT, ARGV_READ_FORMAT, ARGV_BARCODE, ARGV_UMI, ARGV_BC_WHITELIST, ARGV_BC_TRANSLATE
};

static const char *short_options = "x:1:2:u:i:o:t:k:hv";
static struct option long_options[] = {
    {"no-dust", no_argument, 0, ARGV_NO_DUST},
    {"min-hitlen", required_argument, 0, ARGV_MIN_HITLEN},
    {"hitk-factor", required_argument, 0, ARGV_HITK},
    {"consider-secondary", required_argument, 0, ARGV_SECONDARY},
    {"gpu", required_argument, 0, ARGV_GPU},
    {"gpus", required_argument, 0, ARGV_GPUS},
    {"quant-report", required_argument, 0, ARGV_QUANT_REPORT},
    {"quant-format", required_argument, 0, ARGV_QUANT_FORMAT},
    {"quant-min-score", required_argument, 0, ARGV_QUANT_MIN_SCORE},
    {"quant-min-length", required_argument, 0, ARGV_QUANT_MIN_LENGTH},
    {"batch", required_argument, 0, ARGV_BATCH},
    {"layout", required_argument, 0, ARGV_LAYOUT},
    {"dry-run", no_argument, 0, ARGV_DRY_RUN},
    {"dry-run-pipeline", no_argument, 0, ARGV_DRY_PIPE},
    {"dry-run-output", no_argument, 0, ARGV_DRY_OUT},
    {"un", required_argument, 0, ARGV_UN},
    {"cl", required_argument, 0, ARGV_CL},
    {"merge-readpair", no_argument, 0, ARGV_MERGE},
    {"expand-taxid", no_argument, 0, ARGV_EXPAND},
    {"read-format", required_argument, 0, ARGV_READ_FORMAT},
    {"barcode", required_argument, 0, ARGV_BARCODE},
    {"UMI", required_argument, 0, ARGV_UMI},
    {"barcode-whitelist", required_argument, 0, ARGV_BC_WHITELIST},
    {"barcode-translate", required_argument, 0, ARGV_BC_TRANSLATE},
    {"sample-sheet", required_argument, 0, ARGV_SAMPLE_SHEET},
    {(char *)0, 0, 0, 0}};
#include "cfr_cli_reads.hpp"   // PrintLog, SeqReader, ReadSource: FASTA/FASTQ ingest
#include "cfr_cli_bulk.hpp"    // BulkReader: block-parallel ingest of plain four-line FASTQ files
#include "cfr_cli_format.hpp"  // ReadFormat, BarcodeWhitelist, BarcodeTranslation
#include "cfr_cli_merge.hpp"   // PairMerger

// One batch travelling through the ingest -> classify -> output pipeline.
struct Batch {
  std::string ids;              // read ids back to back
  std::vector<uint32_t> id_off;  // n + 1
  std::string seq1, seq2;
  std::vector<uint64_t> off1, off2;
  std::vector<cfr_result> results;
  std::vector<uint64_t> assign;
  // the bases as the device wants them (cfr_pack_reads, written by the ingest stage): 2.25 bits per base cross the host link
  std::vector<uint64_t> pcodes, poff1, poff2;
  std::vector<uint32_t> pmask;
  cfr_packed_batch packed;
  bool isPacked = false;
  // --barcode / --UMI / bc,um stretches of --read-format: the strings printed for each read
  std::string bc, um;
  std::vector<uint32_t> bc_off, um_off;
  // --expand-taxid only: the ids promoted into each assignment (cfr_fetch_expanded layout)
  std::vector<uint32_t> exp_cnt;
  std::vector<uint64_t> exp_off, exp_ids;
  // --un / --cl only: quality strings (empty for a FASTA record) and the reads as classified
  // (DUST intervals as 'N'), which is what the reference writes (ResultWriter.hpp:244-262)
  std::string qual1, qual2, masked1, masked2;
  std::vector<uint64_t> qoff1, qoff2;
  // --merge-readpair: merged[i] != 0 when pair i is classified as one merged read (seq1 holds it, its
  // mate-2 slot is empty); orig1 / orig2 keep the pairs as read, which is what --un / --cl write for them
  std::vector<uint8_t> merged;
  std::string orig1, orig2;
  std::vector<uint64_t> ooff1, ooff2;
  size_t n = 0;
  bool last = false;
  bool fileEnd = false;  // --sample-sheet: an input file ended with this batch, the TSV moves to the next output
  void clear() {
    fileEnd = false;
    bc.clear();
    um.clear();
    bc_off.assign(1, 0);
    um_off.assign(1, 0);
    merged.clear();
    orig1.clear();
    orig2.clear();
    qual1.clear();
    qual2.clear();
    qoff1.assign(1, 0);
    qoff2.assign(1, 0);
    ids.clear();
    id_off.assign(1, 0);
    seq1.clear();
    seq2.clear();
    off1.assign(1, 0);
    off2.assign(1, 0);
    n = 0;
    last = false;
  }
};

// a tiny blocking hand-off slot between two pipeline stages
template <class T>
class Slot {
 public:
  void put(T v) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return !full_; });
    v_ = v;
    full_ = true;
    cv_.notify_all();
  }
  T take() {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return full_; });
    T v = v_;
    full_ = false;
    cv_.notify_all();
    return v;
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  T v_{};
  bool full_ = false;
};

static inline char *put_u64(char *p, uint64_t v) {
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = (char)('0' + v % 10);
    v /= 10;
  } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}
static inline char *put_i32(char *p, int v) {
  if (v < 0) {
    *p++ = '-';
    return put_u64(p, (uint64_t)(-(int64_t)v));
  }
  return put_u64(p, (uint64_t)v);
}
static inline char *put_str(char *p, const char *s, size_t n) {
  memcpy(p, s, n);
  return p + n;
}

// --merge-readpair over one parsed batch (host threads): pair i becomes (merged read, empty mate) when
// ReadPairMerger::Merge succeeds.  The originals move to orig1 / orig2.
static void MergeBatch(Batch &bt, bool haveQual, unsigned nthreads) {
  const size_t n = bt.n;
  std::vector<std::string> rms(n), qms(n);
  bt.merged.assign(n, 0);
  auto work = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
      const int l1 = (int)(bt.off1[i + 1] - bt.off1[i]), l2 = (int)(bt.off2[i + 1] - bt.off2[i]);
      const bool q = haveQual && bt.qoff1[i + 1] - bt.qoff1[i] == (uint64_t)l1 && bt.qoff2[i + 1] - bt.qoff2[i] == (uint64_t)l2 &&
                     l1 > 0 && l2 > 0;
      bt.merged[i] = (uint8_t)PairMerger::Merge(bt.seq1.data() + bt.off1[i], q ? bt.qual1.data() + bt.qoff1[i] : NULL, l1,
                                                bt.seq2.data() + bt.off2[i], q ? bt.qual2.data() + bt.qoff2[i] : NULL, l2,
                                                rms[i], qms[i]);
    }
  };
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> pool;
  const size_t per = (n + nthreads - 1) / nthreads;
  for (unsigned t = 0; t < nthreads; ++t) {
    const size_t lo = t * per, hi = std::min(n, lo + per);
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto &th : pool) th.join();
  bt.orig1.swap(bt.seq1);
  bt.orig2.swap(bt.seq2);
  bt.ooff1.swap(bt.off1);
  bt.ooff2.swap(bt.off2);
  bt.seq1.clear();
  bt.seq2.clear();
  bt.off1.assign(1, 0);
  bt.off2.assign(1, 0);
  for (size_t i = 0; i < n; ++i) {
    if (bt.merged[i]) {
      bt.seq1 += rms[i];
    } else {
      bt.seq1.append(bt.orig1, (size_t)bt.ooff1[i], (size_t)(bt.ooff1[i + 1] - bt.ooff1[i]));
      bt.seq2.append(bt.orig2, (size_t)bt.ooff2[i], (size_t)(bt.ooff2[i + 1] - bt.ooff2[i]));
    }
    bt.off1.push_back(bt.seq1.size());
    bt.off2.push_back(bt.seq2.size());
  }
}

int main(int argc, char *argv[]) {
  if (argc <= 1) {  // CentrifugerClass.cpp:347-351: usage on stderr, exit 0
    fprintf(stderr, "%s", usage);
    return 0;
  }
  cfr_params params;
  cfr_default_params(&params);
  const char *idxPrefix = NULL;
  ReadSource reads, mates, barcodes, umis;  // --barcode / --UMI: one record per read, in step with the read files
  ReadFormat fmt;                            // --read-format
  bool hasBarcode = false, hasUmi = false;
  BarcodeWhitelist whitelist;      // --barcode-whitelist
  BarcodeTranslation translation;  // --barcode-translate
  bool hasWhitelist = false;
  bool hasMate = false, interleaved = false;
  int device = 0;
  int nGpus = 1;
  const char *quantReport = NULL;
  int quantFormat = 0;
  unsigned long quantMinScore = 0, quantMinLength = 0;
  int hostThreads = 0;  // -t: 0 = not given (the ingest stage then uses every core)
  long batchReads = 1 << 20;
  const char *unPrefix = NULL, *clPrefix = NULL;  // --un / --cl
  bool mergePairs = false;                        // --merge-readpair
  bool useSheet = false;                          // --sample-sheet
  std::vector<std::string> sheetOutputs;          // TSV file of every input file, in input order
  bool dryPipe = false;  // diagnostics: the batches the threaded ingest stage hands to the GPU stage, no GPU work
  bool dryOut = false;  // diagnostics: ingest and output stages as in a real run, every read reported unclassified, no GPU work
  bool dryRun = false;  // diagnostics: parse the inputs and print id<TAB>mate1<TAB>mate2, no GPU work
  int c, option_index = 0;
  while ((c = getopt_long(argc, argv, short_options, long_options, &option_index)) != -1) {
    if (c == 'x') idxPrefix = optarg;
    else if (c == 'u') reads.add(optarg);
    else if (c == '1') { reads.add(optarg); hasMate = true; }
    else if (c == '2') { mates.add(optarg); hasMate = true; }
    else if (c == 'i') { reads.add(optarg); hasMate = true; interleaved = true; }
    else if (c == ARGV_SAMPLE_SHEET) {
      // rows "read1 read2 barcode UMI output", '.' = none (CentrifugerClass.cpp:467-516); every file end
      // moves the TSV to the next row's output file (ResultWriter.hpp:75-107)
      FILE *fs = fopen(optarg, "r");
      if (!fs) {
        PrintLog("Cannot open the sample sheet %s", optarg);
        return EXIT_FAILURE;
      }
      char line[8192], f1[2048], f2[2048], bc[2048], um[2048], of[2048];
      while (fgets(line, sizeof(line), fs)) {
        f1[0] = f2[0] = bc[0] = um[0] = of[0] = 0;
        if (sscanf(line, "%2047s %2047s %2047s %2047s %2047s", f1, f2, bc, um, of) < 1) continue;
        if (strcmp(bc, ".")) {  // CentrifugerClass.cpp:495-505
          barcodes.add(bc);
          hasBarcode = true;
        }
        if (strcmp(um, ".")) {
          umis.add(um);
          hasUmi = true;
        }
        const bool paired = strcmp(f2, ".") != 0;
        if (!sheetOutputs.empty() && paired != hasMate) {
          PrintLog("ERROR: a sample sheet mixing single-end and paired-end rows is not supported by centrifuger-b200.");
          return EXIT_FAILURE;
        }
        const size_t before = reads.files.size();
        reads.add(f1);
        if (paired) {
          mates.add(f2);
          hasMate = true;
        }
        // one output entry per input FILE (a row whose name globs to several files moves on at each of
        // them, as the reference's per-file end marker does)
        for (size_t q = before; q < reads.files.size(); ++q) sheetOutputs.push_back(of);
      }
      fclose(fs);
      useSheet = true;
    }
    else if (c == 'o') { /* parsed but unused by the reference as well (CentrifugerClass.cpp:416) */ }
    else if (c == 't') hostThreads = atoi(optarg);  // the reference's worker threads: here the cap on the ingest stage's threads
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
    else if (c == ARGV_GPUS) nGpus = atoi(optarg);
    else if (c == ARGV_QUANT_REPORT) quantReport = optarg;
    else if (c == ARGV_QUANT_FORMAT) quantFormat = atoi(optarg);
    else if (c == ARGV_QUANT_MIN_SCORE) quantMinScore = strtoul(optarg, NULL, 10);
    else if (c == ARGV_QUANT_MIN_LENGTH) quantMinLength = strtoul(optarg, NULL, 10);
    else if (c == ARGV_BATCH) batchReads = atol(optarg);
    else if (c == ARGV_LAYOUT) {
      if (!strcmp(optarg, "occ")) params.layout = CFR_LAYOUT_OCCLINE;
      else if (!strcmp(optarg, "runblock")) params.layout = CFR_LAYOUT_RUNBLOCK;
      else params.layout = CFR_LAYOUT_AUTO;
    } else if (c == ARGV_DRY_RUN) dryRun = true;
    else if (c == ARGV_READ_FORMAT) fmt.Init(optarg);
    else if (c == ARGV_BARCODE) { barcodes.add(optarg); hasBarcode = true; }
    else if (c == ARGV_UMI) { umis.add(optarg); hasUmi = true; }
    else if (c == ARGV_BC_WHITELIST) {
      if (!whitelist.Load(optarg)) {
        PrintLog("ERROR: cannot open the barcode whitelist %s", optarg);
        return EXIT_FAILURE;
      }
      hasWhitelist = true;
    } else if (c == ARGV_BC_TRANSLATE) {
      if (!translation.Load(optarg)) {
        PrintLog("ERROR: cannot open the barcode translation table %s", optarg);
        return EXIT_FAILURE;
      }
    }
    else if (c == ARGV_DRY_PIPE) dryPipe = true;
    else if (c == ARGV_DRY_OUT) dryOut = true;
    else if (c == ARGV_MERGE) mergePairs = true;
    else if (c == ARGV_EXPAND) params.expand_taxid = 1;  // classifierParam.outputExpandedResult, CentrifugerClass.cpp:453-456
    else if (c == ARGV_UN) unPrefix = optarg;
    else if (c == ARGV_CL) clPrefix = optarg;
    else if (c == ARGV_UNSUPPORTED) {
      PrintLog("Option --%s is not supported by the B200 classification path.", long_options[option_index].name);
      return EXIT_FAILURE;
    } else {
      fprintf(stderr, "%s", usage);
      return EXIT_FAILURE;
    }
  }
  // a barcode / UMI stretch in the description without a file of its own is cut out of read 1
  // (CentrifugerClass.cpp:560-563, :141-146)
  if (fmt.segs[ReadFormat::BARCODE].size() > 0) hasBarcode = true;
  if (fmt.segs[ReadFormat::UMI].size() > 0) hasUmi = true;
  if (hasBarcode && hasWhitelist) {  // CentrifugerClass.cpp:565-574
    if (barcodes.files.empty()) {
      PrintLog("Barcode whitelist has to be used with --barcode option, so cases like piping input is not supported.");
      return EXIT_FAILURE;
    }
    // BarcodeCorrector::CollectBackgroundDistribution: how often each listed barcode occurs among the first
    // two million records, then the barcode files are read again from the start
    std::string nm, rec;
    for (int seen = 0; seen < 2000000; ++seen) {
      rec.clear();
      if (!barcodes.next(nm, rec)) break;
      whitelist.Find(fmt.Extract(rec, ReadFormat::BARCODE, true), 1);
    }
    if (barcodes.opened) barcodes.rd.close();
    barcodes.opened = false;
    barcodes.cur = 0;
  }
  if (useSheet) {
    if (sheetOutputs.empty()) {
      PrintLog("ERROR: the sample sheet lists no files.");
      return EXIT_FAILURE;
    }
    reads.markFileEnds = mates.markFileEnds = true;
  }
  if (dryRun) {
    std::string name, name2, s1, s2, q1, q2;
    for (;;) {
      name.clear();
      s1.clear();
      s2.clear();
      q1.clear();
      q2.clear();
      if (!reads.next(name, s1, &q1)) break;
      RemoveReadIdSuffix(name);
      if (hasMate && !(interleaved ? reads.next(name2, s2, &q2) : mates.next(name2, s2, &q2))) {
        PrintLog("ERROR: The two mate-pair read files have different number of reads.");
        return EXIT_FAILURE;
      }
      if (mergePairs && hasMate) {  // diagnostics for --merge-readpair: code, merged read, merged qualities
        std::string rm, qm;
        const bool q = q1.size() == s1.size() && q2.size() == s2.size() && !s1.empty() && !s2.empty();
        const int code = PairMerger::Merge(s1.data(), q ? q1.data() : NULL, (int)s1.size(), s2.data(), q ? q2.data() : NULL,
                                           (int)s2.size(), rm, qm);
        printf("%d\t%s\t%s\n", code, rm.c_str(), qm.c_str());
      } else {
        printf("%s\t%s\t%s\n", name.c_str(), s1.c_str(), s2.c_str());
      }
    }
    return 0;
  }
  if (!dryPipe) PrintLog("Centrifuger v" CENTRIFUGER_VERSION " starts.");
  if (idxPrefix == NULL && !dryPipe && !dryOut) {
    PrintLog("Need to use -x to specify index prefix.");
    return EXIT_FAILURE;
  }
  if (batchReads < 1) batchReads = 1;
  if (batchReads > (1 << 22)) batchReads = 1 << 22;  // the ingest loop cuts batches by this too
  params.max_batch_reads = (int32_t)batchReads;

  {
    // The library's load-time tables (dense locate table, wide lookup table: DESIGN.md section 2) take
    // about a second per Gbp of index to build and pay off over tens of millions of reads.  For a small
    // input -- estimated from the file sizes -- they are left out, unless the environment already says
    // otherwise.  Results are identical either way.
    double bases = 0;
    bool known = true;
    for (int m = 0; m < 2; ++m)
      for (const std::string &f : (m ? mates : reads).files) {
        struct stat sb;
        if (f == "-" || stat(f.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) {
          known = false;
          continue;
        }
        const bool gz = f.size() > 3 && f.compare(f.size() - 3, 3, ".gz") == 0;
        bases += (double)sb.st_size * (gz ? 2.0 : 0.45);  // FASTQ: sequence + qualities + header
      }
    if (known && bases < 3e9) {
      setenv("CFR_B200_DENSE_LOCATE", "-1", 0);
      setenv("CFR_B200_WIDE_LOOKUP", "0", 0);
      setenv("CFR_B200_PAIRS", "0", 0);
    }
  }
  if (nGpus < 1) nGpus = 1;
  if (nGpus > 64) nGpus = 64;
  std::vector<cfr_handle *> handles((size_t)nGpus, (cfr_handle *)NULL);  // one replica of the index per GPU
  cfr_handle *h = NULL;
  int st = CFR_OK;
  if (!dryPipe && !dryOut) {
    // the replicas load side by side (CentrifugerClass.cpp loads once; here each GPU reads the files from the page cache)
    std::vector<int> rcs((size_t)nGpus, CFR_OK);
    std::vector<std::string> errs((size_t)nGpus);
    std::vector<std::thread> openers;
    for (int g = 0; g < nGpus; ++g)
      openers.emplace_back([&, g] {
        rcs[g] = cfr_open(idxPrefix, &params, device + g, &handles[g]);
        if (rcs[g] != CFR_OK) errs[g] = cfr_last_error();
      });
    for (auto &t : openers) t.join();
    for (int g = 0; g < nGpus; ++g)
      if (rcs[g] != CFR_OK) {
        PrintLog("ERROR: %s", errs[g].c_str());
        return EXIT_FAILURE;
      }
    h = handles[0];
    if (quantReport)
      for (cfr_handle *hh : handles)
        if (cfr_quant_enable(hh, quantMinScore, quantMinLength) != CFR_OK) {
          PrintLog("ERROR: %s", cfr_last_error());
          return EXIT_FAILURE;
        }
    PrintLog("Finishes loading index.");
    if (params.min_hit_len <= 0) PrintLog("Inferred --min-hitlen: %d", (int)cfr_index_info(h, 4));
  }

  // Three-stage pipeline over three rotating batches: the ingest thread parses batch i+1 while
  // the GPU classifies batch i and the output thread formats batch i-1 (ResultWriter::Output,
  // ResultWriter.hpp:199-236; rows in input order as in CentrifugerClass.cpp:690).
  const int k = h ? (int)cfr_index_info(h, 24) : (params.max_result > 0 ? params.max_result : 64);  // id slots per read
  const bool expandTaxid = params.expand_taxid != 0;  // --expand-taxid: one more TSV column
  const int NBATCH = 2 + 3 * nGpus;  // ingest (1) + in flight on each GPU (3) + output (1)
  std::vector<Batch> batches((size_t)NBATCH);
  std::vector<Slot<Batch *>> free_slots((size_t)NBATCH);
  Slot<Batch *> to_gpu, to_out;
  for (auto &bt : batches) bt.clear();
  unsigned long totalCnt = 0, classifiedCnt = 0;
  bool mate_mismatch = false;
  // --un / --cl (ResultWriter::SetOutputReads, ResultWriter.hpp:118-172): <prefix>_1.fq.gz / _2.fq.gz with
  // mates, <prefix>.fq.gz without, gzip level 1
  const bool writeReads = unPrefix != NULL || clPrefix != NULL;
  const bool keepReads = writeReads || mergePairs;  // qualities are parsed: --un / --cl print them, the merger reads them
  gzFile readOut[2][4] = {{NULL, NULL, NULL, NULL}, {NULL, NULL, NULL, NULL}};  // [0 = unclassified, 1 = classified][mate 1, mate 2, barcode, UMI]
  for (int cat = 0; cat < 2; ++cat) {
    const char *prefix = cat ? clPrefix : unPrefix;
    if (!prefix) continue;
    const std::string p(prefix);
    readOut[cat][0] = gzopen((hasMate ? p + "_1.fq.gz" : p + ".fq.gz").c_str(), "w1");
    if (hasMate) readOut[cat][1] = gzopen((p + "_2.fq.gz").c_str(), "w1");
    if (hasBarcode) readOut[cat][2] = gzopen((p + "_bc.fa.gz").c_str(), "w1");  // ResultWriter.hpp:158-168
    if (hasUmi) readOut[cat][3] = gzopen((p + "_um.fa.gz").c_str(), "w1");
    if (!readOut[cat][0] || (hasMate && !readOut[cat][1]) || (hasBarcode && !readOut[cat][2]) || (hasUmi && !readOut[cat][3])) {
      PrintLog("ERROR: cannot open the read output files with prefix %s", prefix);
      return EXIT_FAILURE;
    }
  }

  // The reads cross the host link packed (cfr_submit_packed) unless the masked reads must come back for --un / --cl
  // (CFR_B200_PACK_INPUT=0 sends the bytes instead)
  const bool packInput = !writeReads && !dryOut && !dryPipe && !(getenv("CFR_B200_PACK_INPUT") && atoi(getenv("CFR_B200_PACK_INPUT")) == 0);
  const unsigned packThreads = hostThreads > 0 ? (unsigned)std::min(hostThreads, 64) : std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  // CFR_B200_STAGE_REPORT=1: seconds each pipeline stage was busy (waits for the neighbouring stages excluded)
  const bool stageReport = getenv("CFR_B200_STAGE_REPORT") && atoi(getenv("CFR_B200_STAGE_REPORT")) != 0;
  typedef std::chrono::steady_clock StageClock;
  auto secondsSince = [](StageClock::time_point t0) { return std::chrono::duration<double>(StageClock::now() - t0).count(); };
  const StageClock::time_point tPipe0 = StageClock::now();
  double ingestWait = 0, ingestTotal = 0, outputWait = 0, outputTotal = 0, gpuWaitIn = 0, gpuWaitDev = 0;
  size_t batchCnt = 0;

  // Plain FASTQ files without read-format / barcode / UMI work are parsed block-parallel (cfr_cli_bulk.hpp); anything its
  // strict parser does not accept falls back, per file, to the serial reader below.  CFR_B200_BULK_INGEST=0 disables it.
  const bool bulkIngest = !useSheet && !interleaved && !hasBarcode && !hasUmi && !fmt.NeedExtract(ReadFormat::R1) &&
                          !fmt.NeedExtract(ReadFormat::R2) && !(getenv("CFR_B200_BULK_INGEST") && atoi(getenv("CFR_B200_BULK_INGEST")) == 0);
  const unsigned ingestThreads = getenv("CFR_B200_INGEST_THREADS") ? (unsigned)std::max(1, atoi(getenv("CFR_B200_INGEST_THREADS")))
                                 : hostThreads > 0          ? (unsigned)std::min(hostThreads, 64)
                                                            : std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  // Sizes a batch's arrays for batchReads records like its first `seen` ones: a string that grows to 150 MB by doubling is
  // copied and page-faulted several times over, which costs more than parsing the reads
  auto reserveBatch = [&](Batch &bt, size_t seen) {
    const size_t target = (size_t)batchReads;
    auto scaled = [&](size_t bytes) { return std::min<size_t>((size_t)((double)bytes / (double)seen * (double)target * 1.03) + 4096, 272u << 20); };
    bt.seq1.reserve(scaled(bt.seq1.size()));
    bt.ids.reserve(scaled(bt.ids.size()));
    bt.off1.reserve(target + 1);
    bt.id_off.reserve(target + 1);
    if (hasMate) {
      bt.seq2.reserve(scaled(bt.seq1.size()));
      bt.off2.reserve(target + 1);
    }
    if (keepReads) {
      bt.qual1.reserve(scaled(bt.seq1.size()));
      bt.qoff1.reserve(target + 1);
      if (hasMate) {
        bt.qual2.reserve(scaled(bt.seq1.size()));
        bt.qoff2.reserve(target + 1);
      }
    }
  };
  BulkReader bulk1, bulk2;
  if (bulkIngest) {
    bulk1.init(reads.files, ingestThreads, true);
    bulk2.init(mates.files, ingestThreads, true);
    bulk1.set_want_qual(keepReads);
    bulk2.set_want_qual(keepReads);
  }

  std::thread ingest([&] {
    std::string name, name2, tmp, comment1;
    int bi = 0;
    bool eof = false;
    const bool twoFiles = hasMate && !interleaved;
    const StageClock::time_point tStage0 = StageClock::now();
    for (;;) {
      const StageClock::time_point tw = StageClock::now();
      Batch *bt = free_slots[bi].take();
      ingestWait += secondsSince(tw);
      bi = (bi + 1) % NBATCH;
      bt->clear();
      if (bulkIngest) {
        // the same batch limits as the record-by-record loop below, applied per slice of records
        static const size_t maxBasesB = getenv("CFR_B200_MAX_BATCH_BASES") ? (size_t)std::max(1ll, atoll(getenv("CFR_B200_MAX_BATCH_BASES"))) : (256u << 20);  // (tests shrink it)
        static const size_t slotBudgetB = getenv("CFR_B200_SLOT_BUDGET") ? (size_t)atoll(getenv("CFR_B200_SLOT_BUDGET")) : (64u << 20);
        size_t maxLenB = 0;
        const BulkSink s1{&bt->ids, &bt->id_off, &bt->seq1, &bt->off1, keepReads ? &bt->qual1 : nullptr, keepReads ? &bt->qoff1 : nullptr};
        while ((long)bt->n < batchReads && bt->seq1.size() < maxBasesB &&
               (bt->n == 0 || (bt->n + 1) * (maxLenB / 24 + 1) <= slotBudgetB)) {
          // slices of 64 k records let the slot budget see a long read soon; the block-parallel reader serves short reads only
          // and copies large slices with all its threads
          size_t want = std::min<size_t>((size_t)batchReads - bt->n, bulk1.in_fast_mode() ? (size_t)batchReads : 65536);
          if (maxLenB > 0) want = std::max<size_t>(1, std::min(want, slotBudgetB / (maxLenB / 24 + 1) > bt->n ? slotBudgetB / (maxLenB / 24 + 1) - bt->n : 1));
          else want = std::min<size_t>(want, 4096);  // the first slice is short: it tells how long the reads are
          const bool firstSlice = bt->n == 0;
          const size_t got = bulk1.take(want, maxBasesB, s1, &maxLenB);
          bt->n += got;
          if (firstSlice && got > 0) reserveBatch(*bt, got);  // the batch's arrays are sized once, from the first records
          if (got < want && bt->seq1.size() < maxBasesB) {
            eof = true;
            break;
          }
        }
        if (twoFiles) {
          const BulkSink s2{nullptr, nullptr, &bt->seq2, &bt->off2, keepReads ? &bt->qual2 : nullptr, keepReads ? &bt->qoff2 : nullptr};
          const size_t n2 = bulk2.take(bt->n, ~(size_t)0, s2, &maxLenB);
          if (n2 < bt->n) {  // mate 2 ended first: keep the batch consistent for the stages behind
            mate_mismatch = true;
            bt->n = n2;
            bt->off1.resize(bt->n + 1);
            bt->id_off.resize(bt->n + 1);
          } else if (eof && !bulk2.exhausted()) {
            mate_mismatch = true;  // mate 1 ended: mate 2 must end here too
          }
        }
        bt->last = eof || mate_mismatch;
        if (mergePairs && hasMate && bt->n) MergeBatch(*bt, true, std::thread::hardware_concurrency());
        bt->isPacked = false;
        if (packInput && bt->n) {
          cfr_read_batch rb;
          rb.n_reads = bt->n;
          rb.seq1 = bt->seq1.data();
          rb.off1 = bt->off1.data();
          rb.seq2 = hasMate ? bt->seq2.data() : NULL;
          rb.off2 = hasMate ? bt->off2.data() : NULL;
          const uint64_t nw = cfr_packed_words(&rb);
          bt->pcodes.resize(nw);
          bt->pmask.resize(nw);
          bt->poff1.resize(bt->n + 1);
          if (hasMate) bt->poff2.resize(bt->n + 1);
          bt->isPacked = cfr_pack_reads(&rb, bt->pcodes.data(), bt->pmask.data(), bt->poff1.data(), hasMate ? bt->poff2.data() : NULL,
                                        (int)packThreads, &bt->packed) == CFR_OK;
        }
        to_gpu.put(bt);
        if (bt->last) break;
        continue;
      }
      // a batch ends after batchReads reads or 2^28 bases per mate, whichever comes first (long reads: the
      // device work areas grow with the bases and with the longest read of a batch)
      static const size_t maxBases = getenv("CFR_B200_MAX_BATCH_BASES") ? (size_t)std::max(1ll, atoll(getenv("CFR_B200_MAX_BATCH_BASES"))) : (256u << 20);
      // the device hit tables hold (longest read / (minHitLen + 1)) slots for every read of a batch, so one
      // very long read among short ones must not share a batch with a million of them: a batch also ends
      // when reads x (longest / 24 + 1) would pass 2^26 slots (CFR_B200_SLOT_BUDGET overrides, for tests)
      static const size_t slotBudget = getenv("CFR_B200_SLOT_BUDGET") ? (size_t)atoll(getenv("CFR_B200_SLOT_BUDGET")) : (64u << 20);
      size_t maxLen = 0;
      // mate 2 of separate files is parsed by a second thread that trails this one: it reads record j only
      // after mate 1's record j exists, so the two files advance by the same number of records per batch
      std::atomic<long> n1(0);       // mate-1 records parsed so far in this batch
      std::atomic<bool> done1(false);
      long n2 = 0;
      std::thread mate2;
      if (twoFiles) {
        mate2 = std::thread([&] {
          std::string nm;
          std::string *q2 = keepReads ? &bt->qual2 : nullptr;
          for (;;) {
            while (n2 >= n1.load(std::memory_order_acquire) && !done1.load(std::memory_order_acquire)) std::this_thread::yield();
            if (n2 >= n1.load(std::memory_order_acquire)) break;  // mate 1 is done and mate 2 has caught up
            if (mates.step(nm, bt->seq2, q2) != ReadSource::RECORD) break;  // mate 2 (or, with a sample sheet, its file) ended first
            fmt.ExtractTail(bt->seq2, (size_t)bt->off2.back(), ReadFormat::R2, true);
            if (keepReads && bt->qual2.size() > bt->qoff2.back()) fmt.ExtractTail(bt->qual2, (size_t)bt->qoff2.back(), ReadFormat::R2, false);
            bt->off2.push_back(bt->seq2.size());
            if (keepReads) bt->qoff2.push_back(bt->qual2.size());
            ++n2;
          }
        });
      }
      while ((long)bt->n < batchReads && bt->seq1.size() < maxBases &&
             (bt->n == 0 || (bt->n + 1) * (maxLen / 24 + 1) <= slotBudget)) {
        name.clear();
        const int got = reads.step(name, bt->seq1, keepReads ? &bt->qual1 : nullptr, (hasBarcode || hasUmi) ? &comment1 : nullptr);
        if (got == ReadSource::FILE_END) {  // sample sheet: the batch ends with the file
          bt->fileEnd = true;
          break;
        }
        if (got == ReadSource::END) {
          eof = true;
          break;
        }
        RemoveReadIdSuffix(name);
        bt->ids += name;
        bt->id_off.push_back((uint32_t)bt->ids.size());
        if (hasBarcode || hasUmi) {  // GetReadBatch, CentrifugerClass.cpp:128-227
          // each comes from its own file, or is a copy of read 1 as read (before read 1 is cut itself)
          const std::string raw1 = bt->seq1.substr((size_t)bt->off1.back());
          for (int which = 0; which < 2; ++which) {
            if (!(which ? hasUmi : hasBarcode)) continue;
            ReadSource &src = which ? umis : barcodes;
            const int cat = which ? ReadFormat::UMI : ReadFormat::BARCODE;
            std::string rec, com, nm, rq;
            if (!src.files.empty()) {
              if (!src.next(nm, rec, &rq, &com)) {
                PrintLog(which ? "ERROR: The UMI file and read file have different number of reads."
                               : "ERROR: The barcode file and read file have different number of reads.");
                exit(EXIT_FAILURE);
              }
            } else {
              rec = raw1;
              com = comment1;
            }
            std::string val = fmt.InComment(cat) ? fmt.Extract(com, cat, true) : fmt.Extract(rec, cat, true, fmt.inOrder[cat]);
            if (which == 0) {  // CentrifugerClass.cpp:186-206: whitelist correction, then translation; "N" when neither listed nor correctable
              int verdict = 0;
              if (whitelist.listed > 0) {
                const std::string qv = (!fmt.InComment(cat) && !rq.empty()) ? fmt.Extract(rq, cat, false, fmt.inOrder[cat]) : std::string();
                verdict = whitelist.Correct(val, qv);
              }
              if (verdict < 0) val = "N";
              else if (translation.set) val = translation.Translate(val);
            }
            std::string &dst = which ? bt->um : bt->bc;
            dst += val;
            (which ? bt->um_off : bt->bc_off).push_back((uint32_t)dst.size());
          }
        }
        fmt.ExtractTail(bt->seq1, (size_t)bt->off1.back(), ReadFormat::R1, true);
        if (keepReads && bt->qual1.size() > bt->qoff1.back()) fmt.ExtractTail(bt->qual1, (size_t)bt->qoff1.back(), ReadFormat::R1, false);
        bt->off1.push_back(bt->seq1.size());
        maxLen = std::max(maxLen, (size_t)(bt->off1[bt->n + 1] - bt->off1[bt->n]));
        if (keepReads) bt->qoff1.push_back(bt->qual1.size());
        if (interleaved) {
          std::string *q2 = keepReads ? &bt->qual2 : nullptr;
          if (!reads.next(name2, bt->seq2, q2)) {
            mate_mismatch = true;
            break;
          }
          fmt.ExtractTail(bt->seq2, (size_t)bt->off2.back(), ReadFormat::R2, true);
          if (keepReads && bt->qual2.size() > bt->qoff2.back()) fmt.ExtractTail(bt->qual2, (size_t)bt->qoff2.back(), ReadFormat::R2, false);
          bt->off2.push_back(bt->seq2.size());
          maxLen = std::max(maxLen, (size_t)(bt->off2[bt->n + 1] - bt->off2[bt->n]));
          if (keepReads) bt->qoff2.push_back(bt->qual2.size());
        }
        ++bt->n;
        if (bt->n == 4096 && !twoFiles) reserveBatch(*bt, 4096);  // (with a second parser thread appending, the arrays are left alone)
        n1.store((long)bt->n, std::memory_order_release);
      }
      if (twoFiles) {
        done1.store(true, std::memory_order_release);
        mate2.join();
        if (n2 < (long)bt->n) mate_mismatch = true;  // mate 2 ended first
        if (!mate_mismatch && (eof || bt->fileEnd)) {  // mate 1 (its file) ended: mate 2 must end here too
          tmp.clear();
          const int got2 = mates.step(name2, tmp);
          if (got2 == ReadSource::RECORD || (bt->fileEnd && got2 != ReadSource::FILE_END)) mate_mismatch = true;
        }
        if (mate_mismatch && n2 < (long)bt->n) {  // keep the batch consistent for the stages behind
          bt->n = (size_t)n2;
          bt->off1.resize(bt->n + 1);
          bt->id_off.resize(bt->n + 1);
        }
      }
      bt->last = eof || mate_mismatch;
      if (mergePairs && hasMate && bt->n) MergeBatch(*bt, true, std::thread::hardware_concurrency());
      bt->isPacked = false;
      if (packInput && bt->n) {  // 2-bit codes + N bits for the device, packed here while the GPU works on the batches before
        cfr_read_batch rb;
        rb.n_reads = bt->n;
        rb.seq1 = bt->seq1.data();
        rb.off1 = bt->off1.data();
        rb.seq2 = hasMate ? bt->seq2.data() : NULL;
        rb.off2 = hasMate ? bt->off2.data() : NULL;
        const uint64_t nw = cfr_packed_words(&rb);
        bt->pcodes.resize(nw);
        bt->pmask.resize(nw);
        bt->poff1.resize(bt->n + 1);
        if (hasMate) bt->poff2.resize(bt->n + 1);
        bt->isPacked = cfr_pack_reads(&rb, bt->pcodes.data(), bt->pmask.data(), bt->poff1.data(), hasMate ? bt->poff2.data() : NULL,
                                      (int)packThreads, &bt->packed) == CFR_OK;
      }
      to_gpu.put(bt);
      if (bt->last) break;
    }
    ingestTotal = secondsSince(tStage0);
  });

  for (int i = 0; i < NBATCH; ++i) free_slots[i].put(&batches[i]);
  if (dryPipe) {  // what the GPU stage would receive: id, mate 1, mate 2 (merged pairs: the merged read, empty mate)
    for (;;) {
      Batch *bt = to_gpu.take();
      for (size_t i = 0; i < bt->n; ++i) {
        fwrite(bt->ids.data() + bt->id_off[i], 1, bt->id_off[i + 1] - bt->id_off[i], stdout);
        fputc('\t', stdout);
        fwrite(bt->seq1.data() + bt->off1[i], 1, (size_t)(bt->off1[i + 1] - bt->off1[i]), stdout);
        fputc('\t', stdout);
        if (hasMate) fwrite(bt->seq2.data() + bt->off2[i], 1, (size_t)(bt->off2[i + 1] - bt->off2[i]), stdout);
        fputc('\n', stdout);
      }
      const bool last = bt->last;
      for (int q = 0; q < NBATCH; ++q)
        if (&batches[q] == bt) free_slots[q].put(bt);
      if (last) break;
    }
    ingest.join();
    if (mate_mismatch) {
      PrintLog("ERROR: The two mate-pair read files have different number of reads.");
      return EXIT_FAILURE;
    }
    return 0;
  }

  FILE *fpOut = stdout;
  if (useSheet) {  // ResultWriter::SetMultiOutputFileList: the first row's file, header included, instead of stdout
    fpOut = fopen(sheetOutputs[0].c_str(), "w");
    if (!fpOut) {
      PrintLog("ERROR: cannot open output file %s", sheetOutputs[0].c_str());
      return EXIT_FAILURE;
    }
  }
  std::thread output([&] {
    std::string out, rec;
    out.reserve(64 << 20);
    // ResultWriter::OutputHeader (ResultWriter.hpp:186-197)
    const std::string header = std::string("readID\tseqID\ttaxID\tscore\t2ndBestScore\thitLength\tqueryLength\tnumMatches") +
                               (hasBarcode ? "\tbarcode" : "") + (hasUmi ? "\tUMI" : "") +
                               (expandTaxid ? "\texpandedTaxIDs" : "") + "\n";
    // the barcode / UMI columns of read i (ResultWriter.hpp:222-225, :235-238)
    auto put_extra = [&](const Batch *bt, size_t i) {
      if (hasBarcode) {
        out += '\t';
        out.append(bt->bc, bt->bc_off[i], bt->bc_off[i + 1] - bt->bc_off[i]);
      }
      if (hasUmi) {
        out += '\t';
        out.append(bt->um, bt->um_off[i], bt->um_off[i + 1] - bt->um_off[i]);
      }
    };
    out += header;
    size_t sheetAt = 0;  // --sample-sheet: index of the input file being written (ResultWriter.hpp:75-107)
    std::vector<std::string> sheetSeen;
    if (useSheet) sheetSeen.push_back(sheetOutputs[0]);
    int bi = 0;
    const StageClock::time_point tStage0 = StageClock::now();
    for (;;) {
      const StageClock::time_point tw = StageClock::now();
      Batch *bt = to_out.take();
      outputWait += secondsSince(tw);
      for (size_t i = 0; i < bt->n; ++i) {
        const cfr_result &r = bt->results[i];
        const char *id = bt->ids.data() + bt->id_off[i];
        const size_t idn = bt->id_off[i + 1] - bt->id_off[i];
        const size_t base = out.size();
        ++totalCnt;
        if (r.n_assign > 0) {
          ++classifiedCnt;
          const int m = r.n_assign < k ? r.n_assign : k;
          uint64_t ex_at = expandTaxid ? bt->exp_off[i] : 0;
          for (int j = 0; j < m; ++j) {
            const uint64_t a = bt->assign[i * (size_t)k + j];
            const char *nm = r.by_rank ? cfr_rank_name(h, a) : cfr_seq_name(h, a);
            const uint64_t tax = r.by_rank ? cfr_orig_taxid(h, a) : cfr_orig_taxid(h, cfr_seq_taxid(h, a));
            const size_t nn = strlen(nm);
            const size_t at = out.size();
            out.resize(at + idn + nn + 160);
            char *p = &out[at];
            p = put_str(p, id, idn); *p++ = '\t';
            p = put_str(p, nm, nn); *p++ = '\t';
            p = put_u64(p, tax); *p++ = '\t';
            p = put_u64(p, r.score); *p++ = '\t';
            p = put_u64(p, r.secondary_score); *p++ = '\t';
            p = put_i32(p, r.hit_length); *p++ = '\t';
            p = put_i32(p, r.query_length); *p++ = '\t';
            p = put_i32(p, r.n_assign);
            if (hasBarcode || hasUmi) {
              out.resize((size_t)(p - out.data()));
              put_extra(bt, i);
              const size_t used = out.size();
              out.resize(used + 8);
              p = &out[used];
            }
            if (expandTaxid) {  // ResultWriter.hpp:226-227: the original ids of the promoted nodes, comma separated
              *p++ = '\t';
              const uint32_t cnt = bt->exp_cnt[i * (size_t)k + j];
              const size_t used = (size_t)(p - out.data());
              out.resize(used + (size_t)cnt * 21 + 8);
              p = &out[used];
              for (uint32_t q = 0; q < cnt; ++q) {
                if (q) *p++ = ',';
                p = put_u64(p, cfr_orig_taxid(h, bt->exp_ids[ex_at + q]));
              }
              ex_at += cnt;
            }
            *p++ = '\n';
            out.resize((size_t)(p - out.data()));
          }
        } else {
          out.resize(base + idn + 64);
          char *p = &out[base];
          p = put_str(p, id, idn);
          p = put_str(p, "\tunclassified\t0\t0\t0\t0\t", 22);
          p = put_i32(p, r.query_length);
          p = put_str(p, "\t1", 2);
          if (hasBarcode || hasUmi) {
            out.resize((size_t)(p - out.data()));
            put_extra(bt, i);
            const size_t used = out.size();
            out.resize(used + 8);
            p = &out[used];
          }
          if (expandTaxid) *p++ = '\t';  // PrintExtraCol(""), ResultWriter.hpp:239-240
          *p++ = '\n';
          out.resize((size_t)(p - out.data()));
        }
        if (out.size() > (48u << 20)) {
          fwrite(out.data(), 1, out.size(), fpOut);
          out.clear();
        }
        if (writeReads) {  // ResultWriter::Output, ResultWriter.hpp:244-262
          const int cat = r.n_assign > 0 ? 1 : 0;
          if (readOut[cat][0]) {
            // a merged pair was masked and classified as the merged read: its two reads are written as read
            const bool asRead = !bt->merged.empty() && bt->merged[i];
            for (int m = 0; m < (hasMate ? 2 : 1); ++m) {
              const std::string &ms = asRead ? (m ? bt->orig2 : bt->orig1) : (m ? bt->masked2 : bt->masked1);
              const std::vector<uint64_t> &so = asRead ? (m ? bt->ooff2 : bt->ooff1) : (m ? bt->off2 : bt->off1);
              const std::string &qs = m ? bt->qual2 : bt->qual1;
              const std::vector<uint64_t> &qo = m ? bt->qoff2 : bt->qoff1;
              const size_t sl = (size_t)(so[i + 1] - so[i]), ql = (size_t)(qo[i + 1] - qo[i]);
              const bool fq = ql > 0;  // qual == NULL iff kseq saw no quality string (ReadFiles.hpp:326-329)
              rec.clear();
              rec += fq ? '@' : '>';
              rec.append(id, idn);
              rec += '\n';
              rec.append(ms, (size_t)so[i], sl);
              rec += '\n';
              if (fq) {
                rec += "+\n";
                rec.append(qs, (size_t)qo[i], ql);
                rec += '\n';
              }
              gzwrite(readOut[cat][m], rec.data(), (unsigned)rec.size());
            }
            for (int w = 0; w < 2; ++w) {  // ResultWriter.hpp:265-273: ">id\n<barcode>\n", ">id\n<UMI>\n"
              if (!readOut[cat][2 + w]) continue;
              const std::string &vs = w ? bt->um : bt->bc;
              const std::vector<uint32_t> &vo = w ? bt->um_off : bt->bc_off;
              rec.clear();
              rec += '>';
              rec.append(id, idn);
              rec += '\n';
              rec.append(vs, vo[i], vo[i + 1] - vo[i]);
              rec += '\n';
              gzwrite(readOut[cat][2 + w], rec.data(), (unsigned)rec.size());
            }
          }
        }
      }
      if (useSheet && bt->fileEnd) {  // ResultWriter::NextMultiOutputFile: a file seen before is appended to, without a header
        if (fpOut) {
          fwrite(out.data(), 1, out.size(), fpOut);
          fclose(fpOut);
          fpOut = NULL;
        }
        out.clear();
        if (++sheetAt < sheetOutputs.size()) {
          const std::string &nm = sheetOutputs[sheetAt];
          const bool seen = std::find(sheetSeen.begin(), sheetSeen.end(), nm) != sheetSeen.end();
          fpOut = fopen(nm.c_str(), seen ? "a" : "w");
          if (!fpOut) {
            PrintLog("ERROR: cannot open output file %s", nm.c_str());
            exit(EXIT_FAILURE);
          }
          if (!seen) {
            sheetSeen.push_back(nm);
            out += header;
          }
        }
      }
      const bool last = bt->last;
      free_slots[bi].put(bt);
      bi = (bi + 1) % NBATCH;
      if (last) break;
    }
    if (fpOut) {
      fwrite(out.data(), 1, out.size(), fpOut);
      fflush(fpOut);
      if (fpOut != stdout) fclose(fpOut);
    }
    outputTotal = secondsSince(tStage0);
  });

  int rc = 0;
  // submitted, not yet waited for: up to two stay behind the batch just submitted (three in flight)
  struct InFlight {
    Batch *bt;
    int ticket;
    cfr_handle *h;  // the replica the batch went to
  };
  std::deque<InFlight> pending;
  size_t gpuBatchNo = 0;  // batch j is classified on GPU j mod nGpus (CentrifugerClass.cpp:674-691: batches in input order)
  auto retire = [&]() {
    Batch *pb = pending.front().bt;
    const int tk = pending.front().ticket;
    cfr_handle *h = pending.front().h;
    pending.pop_front();
    const StageClock::time_point tw = StageClock::now();
    const int wst = tk >= 0 ? cfr_wait_batch(h, tk) : CFR_OK;
    gpuWaitDev += secondsSince(tw);
    if (wst != CFR_OK) {
      PrintLog("ERROR: %s", cfr_last_error());
      rc = EXIT_FAILURE;
      pb->n = 0;
    }
    if (expandTaxid && tk >= 0 && pb->n > 0) {  // the lists leave the device before this slot's next batch
      pb->exp_cnt.resize(pb->n * (size_t)k);
      pb->exp_off.resize(pb->n);
      if (pb->exp_ids.size() < 4096) pb->exp_ids.resize(4096);
      uint64_t need = 0;
      int est = cfr_fetch_expanded(h, tk, pb->exp_cnt.data(), pb->exp_off.data(), pb->exp_ids.data(), pb->exp_ids.size(), &need);
      if (est != CFR_OK && need > pb->exp_ids.size()) {
        pb->exp_ids.resize(need);
        est = cfr_fetch_expanded(h, tk, pb->exp_cnt.data(), pb->exp_off.data(), pb->exp_ids.data(), pb->exp_ids.size(), &need);
      }
      if (est != CFR_OK) {
        PrintLog("ERROR: %s", cfr_last_error());
        rc = EXIT_FAILURE;
        pb->n = 0;
      }
    }
    to_out.put(pb);
  };
  for (;;) {
    const StageClock::time_point tw = StageClock::now();
    Batch *bt = to_gpu.take();
    gpuWaitIn += secondsSince(tw);
    ++batchCnt;
    int ticket = -1;
    cfr_handle *hb = h;
    if (dryOut) {  // no device: the output stage sees every read as unclassified
      bt->results.assign(bt->n, cfr_result());
      bt->assign.assign(bt->n * (size_t)k, 0);
      for (size_t i = 0; i < bt->n; ++i)
        bt->results[i].query_length = (int32_t)(bt->off1[i + 1] - bt->off1[i] + (hasMate ? bt->off2[i + 1] - bt->off2[i] : 0));
      bt->masked1 = bt->seq1;
      bt->masked2 = bt->seq2;
    } else if (rc == 0 && bt->n > 0) {
      cfr_handle *h = handles[gpuBatchNo++ % handles.size()];
      hb = h;
      bt->results.resize(bt->n);
      bt->assign.resize(bt->n * (size_t)k);
      cfr_read_batch b;
      b.n_reads = bt->n;
      b.seq1 = bt->seq1.data();
      b.off1 = bt->off1.data();
      b.seq2 = hasMate ? bt->seq2.data() : NULL;
      b.off2 = hasMate ? bt->off2.data() : NULL;
      // streaming form: this batch's upload overlaps the previous batches' kernels
      if (writeReads) {
        bt->masked1.resize(bt->seq1.size());
        bt->masked2.resize(bt->seq2.size());
        st = cfr_submit_batch_masked(h, &b, bt->results.data(), bt->assign.data(), &bt->masked1[0],
                                     hasMate ? &bt->masked2[0] : NULL, NULL, &ticket);
      } else if (bt->isPacked) {
        st = cfr_submit_packed(h, &bt->packed, bt->results.data(), bt->assign.data(), NULL, &ticket);
      } else {
        st = cfr_submit_batch(h, &b, bt->results.data(), bt->assign.data(), NULL, &ticket);
      }
      if (st != CFR_OK) {
        PrintLog("ERROR: %s", cfr_last_error());
        rc = EXIT_FAILURE;
        bt->n = 0;
      }
    } else if (rc != 0) {
      bt->n = 0;  // after a device error no batch reaches the output stage unclassified (its result arrays are stale)
    }
    pending.push_back(InFlight{bt, ticket, hb});
    if (pending.size() > 3 * handles.size() - 1) retire();  // three batches in flight per GPU
    if (bt->last) {
      while (!pending.empty()) retire();
      break;
    }
  }
  ingest.join();
  output.join();
  for (int cat = 0; cat < 2; ++cat)
    for (int m = 0; m < 4; ++m)
      if (readOut[cat][m]) gzclose(readOut[cat][m]);
  if (mate_mismatch) {
    PrintLog("ERROR: The two mate-pair read files have different number of reads.");  // CentrifugerClass.cpp:121-125
    return EXIT_FAILURE;
  }
  if (rc) return rc;
  if (stageReport) {
    // ingest / output: seconds spent parsing / formatting (total minus waits for a free or finished batch); gpu: seconds
    // the submitting thread waited for the device, and for the ingest stage (a large share = the parser is the bottleneck)
    fprintf(stderr, "[cfr-stages] {\"batches\": %zu, \"pipeline_s\": %.3f, \"ingest_busy_s\": %.3f, \"ingest_wait_s\": %.3f, "
                    "\"gpu_wait_device_s\": %.3f, \"gpu_wait_ingest_s\": %.3f, \"output_busy_s\": %.3f, \"output_wait_s\": %.3f, "
                    "\"ingest_parse_s\": %.3f, \"ingest_threads\": %u, \"block_parallel_ingest\": %s}\n",
            batchCnt, secondsSince(tPipe0), ingestTotal - ingestWait, ingestWait, gpuWaitDev, gpuWaitIn, outputTotal - outputWait,
            outputWait, bulk1.parse_seconds() + bulk2.parse_seconds(), ingestThreads, bulk1.used_fast_path() ? "true" : "false");
  }
  // ResultWriter::Finalize (ResultWriter.hpp:279-283)
  PrintLog("Processed %lu read fragments, and %lu (%.2lf%%) can be classified.", totalCnt, classifiedCnt,
           (double)classifiedCnt / (double)totalCnt * 100.0);
  if (h && nGpus > 1) {
    // the path's one collective: the per-taxon counters of the replicas are summed over NCCL (NVLink); the totals
    // the log line above printed are checked against the reduced vector's {reads, classified} entries
    const uint64_t nodeCnt = cfr_index_info(h, 5);
    std::vector<uint64_t> red(nodeCnt + 3, 0);
    if (cfr_counts_allreduce_local(handles.data(), nGpus, red.data(), red.size()) != CFR_OK) {
      PrintLog("ERROR: %s", cfr_last_error());
      return EXIT_FAILURE;
    }
    if (red[nodeCnt + 1] != totalCnt || red[nodeCnt + 2] != classifiedCnt) {
      PrintLog("WARNING: the counters reduced over %d GPUs (%lu reads, %lu classified) disagree with the rows written.",
               nGpus, (unsigned long)red[nodeCnt + 1], (unsigned long)red[nodeCnt + 2]);
    } else {
      PrintLog("Reduced the per-taxon counters of %d GPUs over NCCL: %lu reads, %lu classified.", nGpus,
               (unsigned long)red[nodeCnt + 1], (unsigned long)red[nodeCnt + 2]);
    }
  }
  if (h && quantReport) {  // Quantifier::Quantification + Output over the records the GPUs coalesced
    if (cfr_quant_report(handles.data(), nGpus, idxPrefix, quantFormat, quantReport) != CFR_OK) {
      PrintLog("ERROR: %s", cfr_last_error());
      return EXIT_FAILURE;
    }
    PrintLog("Wrote the abundance report to %s.", quantReport);
  }
  for (cfr_handle *hh : handles)
    if (hh) cfr_close(hh);
  PrintLog("Centrifuger finishes.");
  return 0;
}
