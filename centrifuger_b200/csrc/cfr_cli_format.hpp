// cfr_cli_format.hpp -- part of the `centrifuger-b200` command line (host I/O only, see cfr_main.cpp).
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

// --read-format (ReadFormatter.hpp): which stretches of read 1 / read 2 / the barcode record / the UMI
// record are used.  A description is a ',' or ';' separated list of  <r1|r2|bc|um>:START:END[:STRAND]
// (0-based, END inclusive, negative = counted from the end, STRAND '-' reverse-complements the
// assembled stretch) or  <bc|um>:hd:FIELD:START:END[:STRAND]  for a stretch of the header comment
// (FIELD = number of the whitespace separated field, or a prefix to search for).  Stretches of one
// category are concatenated in the order given (ReadFormatter.hpp:275-391).
struct ReadFormat {
  enum { R1 = 0, R2 = 1, BARCODE = 2, UMI = 3, NCAT = 4 };
  struct Seg {
    int start = 0, end = -1, strand = 1;
    bool inComment = false;
    int field = -1;
    std::string prefix;
  };
  std::vector<Seg> segs[NCAT];

  // one item of the description (ReadFormatter.hpp:50-139)
  bool ParseItem(const char *s, int len) {
    if (len < 3 || s[2] != ':') return false;
    int cat;
    if (s[0] == 'r' && s[1] == '1') cat = R1;
    else if (s[0] == 'r' && s[1] == '2') cat = R2;
    else if (s[0] == 'b' && s[1] == 'c') cat = BARCODE;
    else if (s[0] == 'u' && s[1] == 'm') cat = UMI;
    else return false;
    Seg seg;
    int at = 3;
    if (len >= 6 && s[3] == 'h' && s[4] == 'd' && s[5] == ':') {
      seg.inComment = true;
      int e = 6;
      while (e < len && s[e] != ':') ++e;
      const std::string tok(s + 6, (size_t)(e - 6));
      const bool digits = tok.find_first_not_of("0123456789") == std::string::npos;
      if (digits) seg.field = atoi(tok.c_str());
      else seg.prefix = tok;
      at = e + 1;
    }
    int part = 0;
    std::string tok;
    for (int i = at; i <= len; ++i) {
      if (i >= len || s[i] == ':') {
        if (part == 0) seg.start = atoi(tok.c_str());
        else if (part == 1) seg.end = atoi(tok.c_str());
        else seg.strand = (!tok.empty() && tok[0] == '+') ? 1 : -1;
        tok.clear();
        if (i < len && s[i] == ':') ++part;
      } else {
        tok += s[i];
      }
    }
    if (part >= 3 || part < 1) return false;
    segs[cat].push_back(seg);
    return true;
  }
  void Init(const char *desc) {  // ReadFormatter.hpp:198-225
    for (int i = 0; desc[i];) {
      int j = i;
      while (desc[j] && desc[j] != ';' && desc[j] != ',') ++j;
      if (!ParseItem(desc + i, j - i)) {
        fprintf(stderr, "Format description error in %s\n", desc);
        exit(1);
      }
      i = desc[j] ? j + 1 : j;
    }
    for (int c = 0; c < NCAT; ++c) {  // ReadFormatter::AreSegmentsSorted (:140-149); an END of -1 always passes
      inOrder[c] = true;
      for (size_t q = 1; q < segs[c].size(); ++q)
        if (segs[c][q].start <= segs[c][q - 1].end) inOrder[c] = false;
    }
  }
  bool inOrder[NCAT] = {true, true, true, true};
  bool InComment(int cat) const { return !segs[cat].empty() && segs[cat][0].inComment; }
  bool NeedExtract(int cat) const {  // ReadFormatter.hpp:259-273
    if (segs[cat].empty()) return false;
    if (segs[cat].size() == 1) {
      const Seg &g = segs[cat][0];
      if (g.start == 0 && g.end == -1 && g.strand == 1 && !g.inComment) return false;
    }
    return true;
  }
  // the stretches of `in` (a sequence, a quality string or a header comment); complement = false for qualities.
  // overwrite = true restates ReadFormatter::InplaceExtractSeqAndQual for stretches it considers in order:
  // the reference then assembles the result inside the record itself, so a stretch that lies in front of an
  // earlier one (possible when the earlier one ends at -1) is read after it was overwritten.
  std::string Extract(const std::string &given, int cat, bool complement, bool overwrite = false) const {
    if (!NeedExtract(cat)) return given;
    const int len = (int)given.size();
    std::string in = given, out;
    int strand = 1;
    for (const Seg &g : segs[cat]) {
      int start = g.start, end = g.end, lenk = len;
      if (InComment(cat)) {  // find the field, then count inside it (ReadFormatter.hpp:318-366)
        int fstart = 0, fend = 0;
        if (g.field >= 0) {
          int f = 0;
          for (int j = 0; j <= len; ++j) {
            const char ch = j < len ? in[j] : '\0';
            if (ch == ' ' || ch == '\t' || ch == '\0') {
              ++f;
              if (f == g.field) fstart = j + 1;
              else if (f == g.field + 1) {
                fend = j - 1;
                break;
              }
            }
          }
          if (f <= g.field) {
            fstart = len;
            fend = len - 1;
          }
        } else {
          const size_t p = in.find(g.prefix);
          if (p != std::string::npos) {
            fstart = (int)p;
            size_t q = p;
            while (q < in.size() && in[q] != ' ' && in[q] != '\t') ++q;
            fend = (int)q - 1;
          } else {
            fstart = len;
            fend = len - 1;
          }
        }
        if (start >= 0) start += fstart;
        if (end >= 0) end += fstart;
        lenk = fend + 1;
      }
      if (start < 0) start = lenk + start;
      if (end >= lenk) end = lenk - 1;
      else if (end < 0) end = lenk + end;
      if (start < 0) start = 0;  // the reference would read in front of its buffer here
      for (int j = start; j <= end && j < len; ++j) {
        out += in[j];
        if (overwrite && out.size() <= in.size()) in[out.size() - 1] = in[j];
      }
      if (g.strand == -1) strand = -1;
    }
    if (strand == -1) {
      std::reverse(out.begin(), out.end());
      if (complement)
        for (char &c : out) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
    return out;
  }
  // replaces the record just appended to `buf` (from `from` on) by its stretches
  void ExtractTail(std::string &buf, size_t from, int cat, bool complement) const {
    if (!NeedExtract(cat)) return;
    const std::string rec = buf.substr(from);
    buf.resize(from);
    buf += Extract(rec, cat, complement, inOrder[cat]);
  }
};

// --barcode-whitelist (BarcodeCorrector.hpp): a barcode that is not on the list is replaced by the listed
// barcode one substitution away that was seen most often among the first two million barcodes (ties: the
// one whose changed base has the lowest quality, then the first in position / base order); none -> "N".
// The list lives in a 4-ary trie like the reference's, because its look-up also "finds" a proper prefix
// of a listed barcode (with whatever count that inner node has), and that decides what gets corrected.
struct BarcodeWhitelist {
  struct Node {
    int next[4] = {-1, -1, -1, -1};
    int count = 0;
  };
  std::vector<Node> nodes;
  int listed = 0;
  BarcodeWhitelist() : nodes(1) {}
  static int Code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }
  void Insert(const std::string &s, int weight) {  // Trie::Insert (:69-94)
    for (char c : s)
      if (Code(c) < 0) return;
    int p = 0;
    bool grew = false;
    for (char c : s) {
      const int t = Code(c);
      if (nodes[p].next[t] < 0) {
        nodes[p].next[t] = (int)nodes.size();
        nodes.push_back(Node());
        grew = true;
      }
      p = nodes[p].next[t];
    }
    nodes[p].count += weight;
    if (grew) ++listed;
  }
  int Find(const std::string &s, int weight) {  // Trie::SearchAndUpdate (:96-113): count after the update, -1 = absent
    for (char c : s)
      if (Code(c) < 0) return -1;
    int p = 0;
    for (char c : s) {
      p = nodes[p].next[Code(c)];
      if (p < 0) return -1;
    }
    nodes[p].count += weight;
    return nodes[p].count;
  }
  bool Load(const char *file) {  // SetWhitelist (:123-143)
    gzFile fp = gzopen(file, "r");
    if (!fp) return false;
    char buffer[256];
    while (gzgets(fp, buffer, sizeof(buffer)) != NULL) {
      size_t len = strlen(buffer);
      if (len && buffer[len - 1] == '\n') buffer[--len] = 0;
      Insert(buffer, 1);
    }
    gzclose(fp);
    return true;
  }
  // Correct (:164-233): -1 could not correct, 0 listed, 1 corrected in place; qual may be empty
  int Correct(std::string &bc, const std::string &qual) {
    if (Find(bc, 0) != -1) return 0;
    int bestCnt = -1, bestPos = -1, bestBase = -1, bestLowQual = 255;
    std::string probe = bc;
    for (size_t i = 0; i < bc.size(); ++i)
      for (int j = 0; j < 4; ++j) {
        if ("ACGT"[j] == bc[i]) continue;
        probe[i] = "ACGT"[j];
        const int cnt = Find(probe, 0);
        probe[i] = bc[i];
        if (cnt == -1) continue;
        const bool haveQual = i < qual.size();
        if (cnt > bestCnt) {
          bestCnt = cnt;
          bestPos = (int)i;
          bestBase = j;
          if (!qual.empty()) bestLowQual = haveQual ? qual[i] : 0;
        } else if (cnt == bestCnt && !qual.empty() && (haveQual ? qual[i] : 0) < bestLowQual) {
          bestLowQual = haveQual ? qual[i] : 0;
          bestPos = (int)i;
          bestBase = j;
        }
      }
    if (bestPos < 0) return -1;
    bc[(size_t)bestPos] = "ACGT"[bestBase];
    return 1;
  }
};

// --barcode-translate (BarcodeTranslator.hpp): lines "<to><sep><from>"; a barcode is cut into pieces as long
// as the last line's <from>, each piece is replaced, the results are joined with '-'
struct BarcodeTranslation {
  std::unordered_map<std::string, std::string> table;  // from -> to, a later line replaces an earlier one
  int fromLen = -1;
  bool set = false;
  bool Load(const char *file) {
    gzFile fp = gzopen(file, "r");
    if (!fp) return false;
    set = true;
    char line[512];
    while (gzgets(fp, line, sizeof(line)) != NULL) {
      size_t len = strlen(line);
      if (len && line[len - 1] == '\n') line[--len] = 0;
      size_t i = 0;
      while (i < len && line[i] != ',' && line[i] != '\t' && line[i] != ' ') ++i;
      const std::string to(line, i), from(i < len ? line + i + 1 : "");
      fromLen = i < len ? (int)(len - i - 1) : -1;
      table[from] = to;
    }
    gzclose(fp);
    return true;
  }
  std::string Translate(const std::string &bc) const {
    std::string ret;
    if (fromLen <= 0) return ret;
    for (size_t i = 0; i < bc.size() / (size_t)fromLen; ++i) {
      const std::string piece = bc.substr(i * (size_t)fromLen, (size_t)fromLen);
      const auto hit = table.find(piece);
      const std::string *to = hit == table.end() ? nullptr : &hit->second;
      if (!to) {
        fprintf(stderr, "Barcode %s does not exist in the translation table.\n", piece.c_str());
        exit(-1);
      }
      ret += i ? "-" + *to : *to;
    }
    return ret;
  }
};
