// cfr_cli_merge.hpp -- part of the `centrifuger-b200` command line (host I/O only, see cfr_main.cpp).
#pragma once
#include <algorithm>
#include <string>

// --merge-readpair (ReadPairMerger.hpp; applied in CentrifugerClass.cpp:271-272): before a pair is
// classified, mate 2 is reverse-complemented and laid over mate 1.  If the fragment was shorter than a
// read ("read-through": mate 1 starts inside rc(mate 2)) the pair is trimmed to the fragment; if the
// mates overlap at their ends they are joined; either way the result is classified as ONE read
// (Query(rm, NULL), :323-333).  Same decisions as the reference, bit for bit (tests/test_cli_ingest.py
// runs this against the unmodified reference header):
//   * placing `b` at offset j of `a` is accepted when the matches can still reach
//     int((|a| - j) * t), t = 0.95 below 50 overlapping bases, rising linearly to 0.85 at 100 and above;
//   * exactly one offset j < |a| - minOverlap may be accepted, minOverlap = min(31, (|a| + |b|) / 10);
//   * an end overlap of at most 2 * minOverlap bases is refused when `b` starts with a tandem repeat
//     of period <= overlap / 2 (every complete repeat unit inside the overlap equals the first).
struct PairMerger {
  struct Placement {
    int offset = -1, length = -1;  // where b sits on a, and how many bases were compared
  };

  static double RequiredIdentity(int span) {
    if (span >= 100) return 0.85;
    if (span >= 50) return 0.85 + (span - 50) / 50.0 * 0.1;
    return 0.95;
  }

  // true when b laid at a[j..] stays above the identity bound; `compared` = bases looked at
  static bool Fits(const char *a, int alen, const char *b, int blen, int j, int &compared) {
    const int span = alen - j;
    const int need = int(span * RequiredIdentity(span));
    int same = 0, k = 0;
    for (; j + k < alen && k < blen; ++k) {
      same += a[j + k] == b[k];
      if (same + (span - k - 1) < need) return false;  // even a perfect rest cannot reach the bound
    }
    compared = k;
    return true;
  }

  static bool StartsWithTandem(const char *b, int n) {
    for (int period = 1; period <= n / 2; ++period) {
      const int whole = (n / period) * period;  // only complete repeat units are compared
      int k = period;
      while (k < whole && b[k] == b[k % period]) ++k;
      if (k == whole) return true;
    }
    return false;
  }

  static Placement UniquePlacement(const char *a, int alen, const char *b, int blen, int minOverlap, bool refuseTandem) {
    Placement found;
    int accepted = 0;
    for (int j = 0; j < alen - minOverlap; ++j) {
      int compared;
      if (Fits(a, alen, b, blen, j, compared)) {
        ++accepted;
        found.offset = j;
        found.length = compared;
      }
    }
    if (accepted != 1) return Placement();
    if (refuseTandem && found.length <= 2 * minOverlap && StartsWithTandem(b, found.length)) return Placement();
    return found;
  }

  static char Complement(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; }

  // 0 = left as a pair, 1 = joined at an end overlap, 2 = trimmed to a read-through fragment.
  // q1 / q2 may be NULL (FASTA input); rm / qm receive the single read and its qualities.
  static int Merge(const char *r1, const char *q1, int len1, const char *r2, const char *q2, int len2, std::string &rm,
                   std::string &qm) {
    rm.clear();
    qm.clear();
    std::string m2(len2, 'N'), m2q;  // mate 2 on mate 1's strand
    for (int i = 0; i < len2; ++i) m2[i] = Complement(r2[len2 - 1 - i]);
    if (q2) m2q.assign(std::reverse_iterator<const char *>(q2 + len2), std::reverse_iterator<const char *>(q2));
    const int minOverlap = std::min(31, (len1 + len2) / 10);

    // read-through: mate 1 begins somewhere inside m2; the fragment is what they share
    Placement p = UniquePlacement(m2.data(), len2, r1, len1, minOverlap, false);
    if (p.length >= 0) {
      rm.assign(r1, p.length);
      if (q1) {
        qm.assign(q1, p.length);
        for (int i = 0; i < p.length; ++i)
          if (m2q[i + p.offset] > q1[i] || rm[i] == 'N') {  // the better-called base wins
            rm[i] = m2[i + p.offset];
            qm[i] = m2q[i + p.offset];
          }
      }
      return 2;
    }

    // end overlap: m2 begins inside mate 1
    p = UniquePlacement(r1, len1, m2.data(), len2, minOverlap, true);
    if (p.length < 0) return 0;
    const int total = p.offset + len2;  // (mate 2 may end before mate 1 does)
    rm.assign(total, 'N');
    rm.replace(p.offset, len2, m2);
    if (q2) {
      qm.assign(total, '!');
      qm.replace(p.offset, len2, m2q);
    }
    for (int i = 0; i < std::min(len1, total); ++i) {
      // mate 1 keeps its bases in front of the overlap, where its call is at most 14 below mate 2's, and
      // where mate 2 has no call
      if (i < p.offset || (q1 != NULL && q2 != NULL && q1[i] >= qm[i] - 14) || rm[i] == 'N') {
        rm[i] = r1[i];
        if (q1) qm[i] = q1[i];
      }
    }
    return 1;
  }
};
