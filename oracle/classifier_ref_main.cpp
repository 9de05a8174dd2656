// Test infrastructure ONLY: a driver around the UNMODIFIED reference header Classifier.hpp (included from
// the read-only reference tree at build time, never copied; its private search stage is reached by
// redefining `private` for this translation unit only).  Built into oracle/_ref/classifier_ref.
// argv: <index prefix> [minHitLen [k [hitkFactor [secondaryHitLen secondaryScoreFactor]]]].  stdin: "r1<TAB>r2" per line ("-" = no mate); stdout: the hits of
// Classifier::SearchForwardAndReverse as "sp,ep,l,offset,strand" separated by ';' -- or, when k is given,
// the result of Classifier::Query (no DUST): "score 2ndBest hitLength queryLength n name:taxID;...".
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// every standard header the reference includes, BEFORE the redefinition (their include guards keep
// them from being read again under it)
#include <assert.h>
#include <glob.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <time.h>
#include <zlib.h>

#include <algorithm>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#define private public
#include "Classifier.hpp"
#undef private

int main(int argc, char *argv[]) {
  if (argc < 2) return 2;
  struct _classifierParam param;
  if (argc > 2) param.minHitLen = atoi(argv[2]);
  const bool query = argc > 3;
  if (query) param.maxResult = atoi(argv[3]);
  if (argc > 4) param.maxResultPerHitFactor = atoi(argv[4]);
  if (argc > 6) {
    param.considerSecondaryHitLen = (size_t)atol(argv[5]);
    param.considerSecondaryScoreFactor = atof(argv[6]);
  }
  Classifier<Sequence_RunBlock> classifier;
  classifier.Init(argv[1], param);
  static char line[1 << 20];
  while (fgets(line, sizeof(line), stdin)) {
    char *r1 = strtok(line, "\t\n");
    char *r2 = strtok(NULL, "\t\n");
    if (!r1) {
      printf("\n");
      continue;
    }
    if (r2 && !strcmp(r2, "-")) r2 = NULL;
    if (query) {
      struct _classifierResult res;
      classifier.Query(r1, r2, res);
      printf("%lu %lu %d %d %d ", (unsigned long)res.score, (unsigned long)res.secondaryScore, res.hitLength,
             res.queryLength, (int)res.taxIds.size());
      for (size_t i = 0; i < res.taxIds.size(); ++i)
        printf("%s%s:%lu", i ? ";" : "", res.seqStrNames[i].c_str(), (unsigned long)res.taxIds[i]);
      printf("\n");
      continue;
    }
    SimpleVector<struct _BWTHit> hits;
    classifier.SearchForwardAndReverse(r1, r2, hits);
    for (int i = 0; i < (int)hits.Size(); ++i)
      printf("%s%lu,%lu,%d,%d,%d", i ? ";" : "", (unsigned long)hits[i].sp, (unsigned long)hits[i].ep, hits[i].l,
             hits[i].offset, hits[i].strand);
    printf("\n");
  }
  return 0;
}
