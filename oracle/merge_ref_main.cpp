// Test infrastructure ONLY: a driver around the UNMODIFIED reference header ReadPairMerger.hpp (included
// from the read-only reference tree at build time, never copied).  Built into oracle/_ref/merge_ref.
// stdin: one pair per line "r1<TAB>q1<TAB>r2<TAB>q2" ("-" = no qualities); stdout: "code<TAB>rm<TAB>qm".
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "ReadPairMerger.hpp"

int main() {
  ReadPairMerger merger;
  static char line[1 << 20];
  while (fgets(line, sizeof(line), stdin)) {
    char *f[4] = {NULL, NULL, NULL, NULL};
    int n = 0;
    for (char *p = strtok(line, "\t\n"); p && n < 4; p = strtok(NULL, "\t\n")) f[n++] = p;
    if (n < 4) continue;
    char *q1 = strcmp(f[1], "-") ? f[1] : NULL, *q2 = strcmp(f[3], "-") ? f[3] : NULL;
    char *rm = NULL, *qm = NULL;
    const int code = merger.Merge(f[0], q1, f[2], q2, &rm, &qm);
    printf("%d\t%s\t%s\n", code, rm ? rm : "", qm ? qm : "");
    free(rm);
    free(qm);
  }
  return 0;
}
