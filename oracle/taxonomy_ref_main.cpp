// Test infrastructure ONLY: a driver around the UNMODIFIED reference header Taxonomy.hpp (included from
// the read-only reference tree at build time, never copied).  Built into oracle/_ref/taxonomy_ref.
// argv[1] = <prefix>.2.cfr.  stdin: one query per line "k id id id ..." (compact tax ids); stdout:
// "<promoted ids, space separated>|<child list 0, comma separated>;<child list 1>;..." -- the lists
// exactly as Taxonomy::ReduceTaxIds returns them in promotedChildTaxIds (possibly fewer than ids).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "Taxonomy.hpp"

using namespace compactds;

int main(int argc, char *argv[]) {
  if (argc < 2) return 2;
  FILE *fp = fopen(argv[1], "rb");
  if (!fp) return 2;
  Taxonomy tax;
  tax.Load(fp);
  fclose(fp);
  static char line[1 << 16];
  while (fgets(line, sizeof(line), stdin)) {
    char *p = strtok(line, " \n");
    if (!p) continue;
    const int k = atoi(p);
    SimpleVector<size_t> ids, promoted;
    while ((p = strtok(NULL, " \n"))) ids.PushBack((size_t)strtoull(p, NULL, 10));
    std::vector<std::vector<size_t> > lists;
    tax.ReduceTaxIds(ids, promoted, k, &lists);
    for (int i = 0; i < promoted.Size(); ++i) printf(i ? " %lu" : "%lu", (unsigned long)promoted[i]);
    printf("|");
    for (size_t i = 0; i < lists.size(); ++i) {
      if (i) printf(";");
      for (size_t j = 0; j < lists[i].size(); ++j) printf(j ? ",%lu" : "%lu", (unsigned long)lists[i][j]);
    }
    printf("\n");
  }
  return 0;
}
