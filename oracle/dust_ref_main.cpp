// Test infrastructure ONLY: a driver around the UNMODIFIED reference header Dustmasker.hpp (included from
// the read-only reference tree at build time, never copied).  Built into oracle/_ref/dust_ref.
// stdin: one read per line; stdout: the read with the masked intervals replaced by 'N', exactly as
// ClassifyReads_Thread applies them (CentrifugerClass.cpp:276-290).
#include <stdio.h>
#include <string.h>

#include <vector>

#include "Dustmasker.hpp"

int main() {
  Dustmasker dustmasker;
  dustmasker.Init("ACGT");
  std::vector<struct _dustmasker_perfect_interval> intervals, windowIntervals;
  static char line[1 << 20];
  while (fgets(line, sizeof(line), stdin)) {
    size_t n = strlen(line);
    while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
    dustmasker.MaskWithBuffer(line, n, windowIntervals, intervals);
    for (size_t j = 0; j < intervals.size(); ++j)
      for (int k = intervals[j].start; k <= (int)intervals[j].end; ++k) line[k] = 'N';
    printf("%s\n", line);
  }
  return 0;
}
