// cfr_quant_main.cpp -- `centrifuger-b200-quant`: drop-in for the reference's `centrifuger-quant`
// (CentrifugerQuant.cpp): classification TSV in, abundance report out, same options and bytes.
// Host-only: the quantifier works on coalesced assignments (csrc/cfr_quant.hpp); when the reads are
// classified in the same process (`centrifuger-b200 --quant-report`), the coalescing runs on the GPU and no
// TSV is written and read back.
#include <getopt.h>
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>

#include "cfr_quant.hpp"

static const char usage[] =
    "./centrifuger-b200-quant [OPTIONS]:\n"
    "Required:\n"
    "\t-x FILE: index prefix\n"
    "\t-c FILE: classification result file (- for stdin, may be gzip'ed)\n"
    "Optional:\n"
    "\t--min-score INT: only consider reads with score at least <int> \n"
    "\t--min-length INT: only consider reads with classified length at least <int>\n"
    "\t--output-format INT: output format. (0:centrifuge,default, 1:metaphlan, 2:CAMI, 3:kraken-report)\n"
    "\t-h: print this usage message\n";

static void log_line(const char *msg) {  // Utils::PrintLog (Utils.hpp:369-381)
  time_t t = time(NULL);
  struct tm *lt = localtime(&t);
  char buf[64];
  strftime(buf, sizeof(buf), "%c", lt);
  fprintf(stderr, "[%s] %s\n", buf, msg);
}

int main(int argc, char *argv[]) {
  if (argc <= 1) {
    fprintf(stderr, "%s", usage);
    return 0;
  }
  enum { OPT_MIN_SCORE = 256, OPT_MIN_LENGTH, OPT_FORMAT };
  static struct option longopts[] = {{"min-score", required_argument, 0, OPT_MIN_SCORE},
                                     {"min-length", required_argument, 0, OPT_MIN_LENGTH},
                                     {"output-format", required_argument, 0, OPT_FORMAT},
                                     {0, 0, 0, 0}};
  const char *idx = NULL, *cls = NULL;
  unsigned long min_score = 0, min_len = 0;
  int format = 0, c, oi;
  while ((c = getopt_long(argc, argv, "x:c:h", longopts, &oi)) != -1) {
    if (c == 'x') idx = optarg;
    else if (c == 'c') cls = optarg;
    else if (c == OPT_MIN_SCORE) min_score = strtoul(optarg, NULL, 10);
    else if (c == OPT_MIN_LENGTH) min_len = strtoul(optarg, NULL, 10);
    else if (c == OPT_FORMAT) format = atoi(optarg);
    else if (c == 'h') {
      fprintf(stdout, "%s", usage);
      return 0;
    } else {
      fprintf(stderr, "%s", usage);
      return EXIT_FAILURE;
    }
  }
  log_line("Centrifuger-quant v1.1.3-r347 starts.");
  if (!idx || !cls) {
    log_line("Need to use -x to specify index prefix and -c to specify the classification result.");
    return EXIT_FAILURE;
  }
  cfrb200::Quantifier q;
  std::string err;
  if (q.init(idx, err) != 0) {
    log_line(("ERROR: " + err).c_str());
    return EXIT_FAILURE;
  }
  // the TSV through zlib (plain or gzip'ed, like the reference's gzopen), handed to the parser as a stream
  gzFile gz = strcmp(cls, "-") ? gzopen(cls, "r") : gzdopen(fileno(stdin), "r");
  if (!gz) {
    log_line("ERROR: cannot open the classification file.");
    return EXIT_FAILURE;
  }
  FILE *tmp = tmpfile();
  char buf[1 << 16];
  int got;
  while ((got = gzread(gz, buf, sizeof(buf))) > 0) fwrite(buf, 1, (size_t)got, tmp);
  gzclose(gz);
  rewind(tmp);
  q.load_tsv(tmp, min_score, min_len);
  fclose(tmp);
  log_line("Finish loading the read classification result.");
  q.quantify();
  q.output(stdout, format);
  log_line("Centrifuger-quant finishes.");
  return 0;
}
