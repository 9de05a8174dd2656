// Test infrastructure ONLY: a driver around the UNMODIFIED reference headers compactds/FMIndex.hpp and
// Sequence_RunBlock.hpp (included from the read-only reference tree at build time, never copied).
// Built into oracle/_ref/fm_ref.  argv[1] = <prefix>.1.cfr.  stdin, one query per line:
//   R <c> <pos> <incl>   Sequence_RunBlock::Rank          -> count
//   A <pos>              Sequence_RunBlock::Access        -> symbol
//   F <c> <pos> <incl>   FMIndex::Rank                    -> count
//   S <string>           FMIndex::BackwardSearch          -> l sp ep
//   L <row>              FMIndex::BackwardToSampledSA     -> value steps
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>  // FMBuilder.hpp uses std::string without including it

#include "compactds/FMIndex.hpp"
#include "compactds/Sequence_RunBlock.hpp"

using namespace compactds;

int main(int argc, char *argv[]) {
  if (argc < 2) return 2;
  FILE *fp = fopen(argv[1], "rb");
  if (!fp) return 2;
  FMIndex<Sequence_RunBlock> fm;
  fm.Load(fp);
  fclose(fp);
  Sequence_RunBlock *bwt = fm.GetBWT();
  static char line[1 << 16], arg[1 << 16];
  while (fgets(line, sizeof(line), stdin)) {
    char c;
    unsigned long long pos;
    int incl;
    if (line[0] == 'R' && sscanf(line + 1, " %c %llu %d", &c, &pos, &incl) == 3) {
      printf("%llu\n", (unsigned long long)bwt->Rank(c, (size_t)pos, incl));
    } else if (line[0] == 'A' && sscanf(line + 1, " %llu", &pos) == 1) {
      printf("%c\n", (char)bwt->Access((size_t)pos));
    } else if (line[0] == 'F' && sscanf(line + 1, " %c %llu %d", &c, &pos, &incl) == 3) {
      printf("%llu\n", (unsigned long long)fm.Rank(c, (size_t)pos, incl));
    } else if (line[0] == 'S' && sscanf(line + 1, " %65000s", arg) == 1) {
      size_t sp = 0, ep = 0;
      const size_t l = fm.BackwardSearch(arg, strlen(arg), sp, ep);
      printf("%llu %llu %llu\n", (unsigned long long)l, (unsigned long long)sp, (unsigned long long)ep);
    } else if (line[0] == 'L' && sscanf(line + 1, " %llu", &pos) == 1) {
      size_t steps = 0;
      const size_t v = fm.BackwardToSampledSA((size_t)pos, steps);
      printf("%llu %llu\n", (unsigned long long)v, (unsigned long long)steps);
    } else {
      printf("?\n");
    }
  }
  return 0;
}
