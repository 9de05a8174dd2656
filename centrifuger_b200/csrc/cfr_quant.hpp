// cfr_quant.hpp -- taxonomic quantification, host side.
//
// Restates what the reference's `centrifuger-quant` computes from the classification of a read set
// (Quantifier.hpp): coalesced read assignments -> per-taxon read counts and their subtree sums ->
// abundances by EM over the covered part of the taxonomy -> one of four report formats.
//
// How the work is split here.  Everything that touches every READ is on the device: the assignment of a read
// is reduced to a key (target taxa, weight class, unique flag) by k_quant_keys and the keys of a batch are
// coalesced by radix sorts (csrc/cfr_api.cu: quant_coalesce_batch), so only the distinct keys with their
// multiplicities -- a few hundred to a few thousand entries per million reads -- reach the host, and no
// classification TSV is written and read back.  What is left (this file) works on those few entries and on
// the taxonomy tree: a 1000-iteration EM over a tree of some thousand nodes is latency-bound scalar
// work, not GPU work.  The arithmetic follows the reference operation for operation, in its order, so
// the doubles -- and the printed report -- are identical, not merely close:
//   weights are sums of 4^-d, exact in binary floating point whatever the order of the reads
//   (CalculateAssignmentWeight, Quantifier.hpp:283-293); everything after coalescing runs in the
//   reference's order over assignments sorted by its operator< (:50-63).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "cfr_format.hpp"

namespace cfrb200 {

enum { QUANT_FORMAT_CENTRIFUGER = 0, QUANT_FORMAT_METAPHLAN = 1, QUANT_FORMAT_CAMI = 2, QUANT_FORMAT_KREPORT = 3 };

// Tree_Plain as the quantifier uses it (compactds/Tree_Plain.hpp:58-206): children in insertion order, the
// root's id doubles as the "none" mark of the child / sibling links.
struct QuantTree {
  struct Node {
    size_t parent, sibling, child, last_child;
  };
  size_t root = 0;
  std::vector<Node> nodes;
  void init(size_t n) { nodes.assign(n, Node{root, root, root, root}); }
  void add_edge(size_t c, size_t parent) {
    nodes[c].parent = parent;
    const size_t last = nodes[parent].last_child;
    if (last == root) nodes[parent].child = c; else nodes[last].sibling = c;
    nodes[parent].last_child = c;
  }
  std::vector<size_t> children(size_t v) const {
    std::vector<size_t> r;
    for (size_t c = nodes[v].child; c != root; c = nodes[c].sibling) r.push_back(c);
    return r;
  }
  size_t size() const { return nodes.size(); }
};

struct QuantAssignment {  // _readAssignment (Quantifier.hpp:25-78)
  std::vector<uint64_t> targets;
  double weight = 0, count = 0, uniq = 0;
  bool operator<(const QuantAssignment &b) const {
    if (targets.size() != b.targets.size()) return targets.size() < b.targets.size();
    for (size_t i = 0; i < targets.size(); ++i)
      if (targets[i] != b.targets[i]) return targets[i] < b.targets[i];
    return false;
  }
  bool same_targets(const QuantAssignment &b) const { return targets == b.targets; }
};

class Quantifier {
 public:
  // Quantifier::Init(indexPrefix) (Quantifier.hpp:432-458): taxonomy from .2.cfr, sequence lengths from .3.cfr
  int init(const std::string &prefix, std::string &err) {
    int st = load_taxonomy_file(prefix + ".2.cfr", tax_, err);
    if (st) return st;
    std::map<uint64_t, uint64_t> seq_len;
    FILE *fp = fopen((prefix + ".3.cfr").c_str(), "rb");
    if (!fp) {
      err = prefix + ".3.cfr: cannot open";
      return -2;
    }
    uint64_t two[2];
    while (fread(two, 8, 2, fp) == 2) seq_len[two[0]] = two[1];
    fclose(fp);
    init_from(tax_, seq_len);
    return 0;
  }
  void init_from(const TaxonomyHost &t, const std::map<uint64_t, uint64_t> &seq_len) {
    if (&t != &tax_) tax_ = t;
    const size_t n = tax_.node_cnt;
    abund_.assign(n + 1, 0.0);
    read_count_.assign(n + 1, 0.0);
    uniq_count_.assign(n + 1, 0.0);
    tax_len_.assign(n + 1, 0);
    seq_length_to_tax_length(seq_len);
    assignments_.clear();
    unclassified_ = 0;
  }
  uint64_t node_cnt() const { return tax_.node_cnt; }

  // Taxonomy::CompactTaxId (Taxonomy.hpp:646-652)
  uint64_t compact_taxid(uint64_t orig) const {
    if (orig_to_compact_.empty())
      for (uint64_t i = 0; i < tax_.orig_taxid.size(); ++i) orig_to_compact_[tax_.orig_taxid[i]] = i;
    auto it = orig_to_compact_.find(orig);
    return it == orig_to_compact_.end() ? tax_.node_cnt : it->second;
  }

  // CalculateAssignmentWeight (:283-293) as a class: weight = 4^-class, class in 0..11
  static int weight_class(uint64_t hit_length, uint64_t read_length) {
    int diff = (int)read_length - (int)hit_length;
    const int slack = (int)((double)read_length * 0.01);
    if (diff < slack) return 0;
    diff -= slack;
    if (diff > 10) diff = 11;
    return diff;
  }
  static double class_weight(int c) { return 1.0 / (double)(1 << (2 * c)); }

  // one coalesced entry: `count` reads with these targets (compact ids), this weight class and unique flag
  void add(const uint64_t *targets, size_t n_targets, int wclass, bool uniq, uint64_t count) {
    QuantAssignment a;
    a.targets.assign(targets, targets + n_targets);
    a.weight = class_weight(wclass) * (double)count;  // exact: a multiple of 4^-11 below 2^53
    a.count = (double)count;
    a.uniq = uniq ? (double)count : 0.0;
    assignments_.push_back(a);
  }
  void add_unclassified(uint64_t n) { unclassified_ += n; }

  // LoadReadAssignments (:515-622): the classification TSV of `centrifuger` / `centrifuger-b200`
  int load_tsv(FILE *fp, uint64_t min_score, uint64_t min_hit_length) {
    std::vector<char> line(1 << 16), rid(1 << 16), name(1 << 16);
    std::string prev;
    QuantAssignment cur;
    bool have = false;
    size_t line_no = 0;
    while (fgets(line.data(), (int)line.size(), fp)) {
      if (line_no++ == 0) continue;  // header
      unsigned long taxid = 0, score = 0, second = 0, hit = 0, rlen = 0;
      if (sscanf(line.data(), "%s\t%[^\t]\t%lu\t%lu\t%lu\t%lu\t%lu", rid.data(), name.data(), &taxid, &score, &second, &hit, &rlen) < 7)
        continue;
      if (hit < min_hit_length || score < min_score || taxid == 0) {
        ++unclassified_;
        continue;
      }
      if (prev != rid.data()) {
        if (have && !cur.targets.empty()) assignments_.push_back(cur);
        cur.targets.clear();
        cur.weight = class_weight(weight_class(hit, rlen));
        cur.count = 1;
        cur.uniq = score > second ? 1 : 0;
        prev = rid.data();
        have = true;
      }
      cur.targets.push_back(compact_taxid(taxid));
    }
    if (!cur.targets.empty()) assignments_.push_back(cur);
    return 0;
  }

  // Quantification (:640-744)
  void quantify() {
    coalesce();
    const size_t n_assign = assignments_.size();
    QuantTree all;
    general_tree(all);
    // the part of the taxonomy that some read reaches, ids in order of first appearance (root first)
    std::map<size_t, size_t> covered;
    std::vector<size_t> covered_inv;
    auto cover_add = [&](size_t v) {  // MapID::Add: existing id, or the next one
      auto it = covered.find(v);
      if (it != covered.end()) return it->second;
      const size_t id = covered_inv.size();
      covered[v] = id;
      covered_inv.push_back(v);
      return id;
    };
    size_t sub_size = 1;
    cover_add(all.root);
    std::vector<QuantAssignment> sub(assignments_);
    for (size_t i = 0; i < n_assign; ++i) {
      const size_t tc = assignments_[i].targets.size();
      for (size_t j = 0; j < tc; ++j) {
        const uint64_t ctid = sub[i].targets[j];
        if (ctid == tax_.node_cnt) {  // a taxon outside the tree counts for the root
          sub[i].targets[j] = 0;
          read_count_[all.root] += assignments_[i].count / (double)tc;
          uniq_count_[all.root] += assignments_[i].uniq;
          continue;
        }
        read_count_[assignments_[i].targets[j]] += assignments_[i].count / (double)tc;
        uniq_count_[assignments_[i].targets[j]] += assignments_[i].uniq;
        uint64_t p = ctid;
        while (cover_add(p) == sub_size) {
          ++sub_size;
          p = tax_.parent[p];
        }
        sub[i].targets[j] = covered[ctid];
      }
    }
    tree_sum(all.root, read_count_.data(), all);
    tree_sum(all.root, uniq_count_.data(), all);
    QuantTree st;
    st.init(sub_size);
    for (size_t i = 1; i < sub_size; ++i) st.add_edge(i, covered[tax_.parent[covered_inv[i]]]);
    std::vector<size_t> st_len(sub_size, 0);
    for (size_t i = 0; i < all.size(); ++i) {
      auto it = covered.find(i);
      if (it != covered.end()) st_len[it->second] = tax_len_[i] + tax_len_[tax_.root] / 10;  // baseline against tiny genomes
    }
    std::vector<double> st_abund(sub_size, 0.0), st_reads(sub_size, 0.0);
    em(sub, st, st_len.data(), st_reads.data(), st_abund.data());
    for (size_t i = 0; i < sub_size; ++i) abund_[covered_inv[i]] = st_abund[i];
  }

  // Output (:746-818)
  void output(FILE *fp, int format) const {
    const size_t n = tax_.node_cnt;
    if (format == QUANT_FORMAT_METAPHLAN) {
      fprintf(fp, "#clade_name\tNCBI_tax_id\trelative_abundance\tadditional_species\n");
      for (size_t i = 0; i < n; ++i) {
        if (read_count_[i] < 1e-6 || !canonical(i)) continue;
        fprintf(fp, "%s\t%s\t%.5lf\t\n", lineage(i, format, true).c_str(), lineage(i, format, false).c_str(), abund_[i] * 100.0);
      }
    } else if (format == QUANT_FORMAT_CAMI) {
      fprintf(fp, "@@TAXID\tRANK\tTAXPATH\tTAXPATHSN\tPERCENTAGE\n");
      for (size_t i = 0; i < n; ++i) {
        if (read_count_[i] < 1e-6 || !canonical(i)) continue;
        fprintf(fp, "%lu\t%s\t%s\t%s\t%.5lf\n", (unsigned long)tax_.orig_taxid[i], tax_rank_string(tax_.rank[i]),
                lineage(i, format, false).c_str(), lineage(i, format, true).c_str(), abund_[i] * 100.0);
      }
    } else if (format == QUANT_FORMAT_KREPORT) {
      QuantTree all;
      general_tree(all);
      kreport(all, all.root, 0, 0, '\0', fp);
    } else {
      fprintf(fp, "name\ttaxID\ttaxRank\tgenomeSize\tnumReads\tnumUniqueReads\tabundance\n");
      for (size_t i = 0; i < n; ++i) {
        if (read_count_[i] < 1e-6) continue;
        fprintf(fp, "%s\t%lu\t%s\t%lu\t%d\t%d\t%.7lf\n", tax_.tax_name[i].c_str(), (unsigned long)tax_.orig_taxid[i],
                tax_rank_string(tax_.rank[i]), (unsigned long)tax_len_[i], (int)(read_count_[i] + 1e-3),
                (int)(uniq_count_[i] + 1e-3), abund_[i]);
      }
    }
  }
  uint64_t unclassified() const { return unclassified_; }

 private:
  TaxonomyHost tax_;
  mutable std::map<uint64_t, uint64_t> orig_to_compact_;
  std::vector<double> abund_, read_count_, uniq_count_;
  std::vector<size_t> tax_len_;
  std::vector<QuantAssignment> assignments_;
  uint64_t unclassified_ = 0;

  bool canonical(size_t i) const { return tax_rank_is_canonical(tax_.rank[i]); }

  // CoalesceAssignments (:490-513)
  void coalesce() {
    std::sort(assignments_.begin(), assignments_.end());
    size_t k = assignments_.empty() ? 0 : 1;
    for (size_t i = 1; i < assignments_.size(); ++i) {
      if (assignments_[i].same_targets(assignments_[k - 1])) {
        assignments_[k - 1].weight += assignments_[i].weight;
        assignments_[k - 1].count += assignments_[i].count;
        assignments_[k - 1].uniq += assignments_[i].uniq;
      } else {
        assignments_[k] = assignments_[i];
        ++k;
      }
    }
    assignments_.resize(k);
  }

  // Taxonomy::ConvertToGeneralTree (Taxonomy.hpp:1086-1107)
  void general_tree(QuantTree &t) const {
    t.root = tax_.root;
    t.init(tax_.node_cnt);
    for (size_t i = 0; i < tax_.node_cnt; ++i)
      if (i != tax_.parent[i]) t.add_edge(i, tax_.parent[i]);
    std::vector<size_t> rc = t.children(t.root);
    std::map<size_t, int> rcm;
    for (size_t c : rc) rcm[c] = 1;
    for (size_t i = 0; i < tax_.node_cnt; ++i)
      if (t.nodes[i].parent == t.root && rcm.find(i) == rcm.end()) t.add_edge(i, t.root);
  }

  // Taxonomy::IsNextSeqNameFromTheSameGenome (Taxonomy.hpp:372-406)
  static bool next_seq_same_genome(const char *a, const char *b) {
    uint64_t id[2];
    for (int i = 0; i < 2; ++i) {
      const char *s = i ? b : a;
      id[i] = 0;
      int j = 0;
      for (; s[j]; ++j)
        if (s[j] >= '0' && s[j] <= '9') break;
      for (; s[j]; ++j) {
        if (s[j] >= '0' && s[j] <= '9') id[i] = id[i] * 10 + (uint64_t)(s[j] - '0'); else break;
      }
      if (j < 3 || s[2] != '_') return false;
    }
    return id[1] == id[0] + 1;
  }

  // Taxonomy::ConvertSeqLengthToTaxLength + InferAllTaxLength (Taxonomy.hpp:1111-1207)
  void seq_length_to_tax_length(std::map<uint64_t, uint64_t> seq_len) {
    const size_t n = tax_.node_cnt;
    std::vector<std::string> names(tax_.seq_name);
    std::map<std::string, uint64_t> id_of;
    for (uint64_t i = 0; i < names.size(); ++i)
      if (id_of.find(names[i]) == id_of.end()) id_of[names[i]] = i;
    std::sort(names.begin(), names.end());
    auto tax_of = [&](uint64_t sid) { return sid < tax_.seq_cnt ? tax_.seq_to_tax[sid] : (uint64_t)n; };
    for (size_t i = 0; i < n; ++i) tax_len_[i] = 0;
    for (size_t i = 0; i < names.size();) {
      const uint64_t sid = id_of[names[i]];
      size_t len = seq_len[sid];
      const uint64_t taxid = tax_of(sid);
      size_t j;
      for (j = i + 1; j < names.size(); ++j) {
        const uint64_t nsid = id_of[names[j]];
        if (tax_of(nsid) != taxid || !next_seq_same_genome(names[j - 1].c_str(), names[j].c_str())) break;
        len += seq_len[nsid];
      }
      if (taxid < n && len > tax_len_[taxid]) tax_len_[taxid] = len;
      i = j;
    }
    std::vector<size_t> cnt(n, 0), nlen(n, 0);
    std::vector<char> preset(n, 0);
    for (size_t i = 0; i < n; ++i)
      if (tax_len_[i] != 0) {
        preset[i] = 1;
        cnt[i] = 1;
      }
    for (size_t i = 0; i < n; ++i) {
      if (!preset[i]) continue;
      if (i == tax_.parent[i] || !tax_.leaf[i]) continue;
      size_t p = tax_.parent[i];
      for (;;) {
        ++cnt[p];
        nlen[p] += tax_len_[i];
        if (p == tax_.parent[p]) break;
        p = tax_.parent[p];
      }
    }
    for (size_t i = 0; i < n; ++i) {  // lengthFromSeqLength = true: every node is recomputed
      size_t sum = nlen[i];
      if (preset[i]) sum += tax_len_[i];
      tax_len_[i] = cnt[i] == 0 ? sum : sum / cnt[i];
    }
  }

  // GenerateTreeAbundance (:123-133)
  static double tree_sum(size_t v, double *a, const QuantTree &t) {
    double sum = a[v];
    for (size_t c : t.children(v)) sum += tree_sum(c, a, t);
    return a[v] = sum;
  }

  // RedistributeAbundToChildren (:136-182) without expanded-taxid edge weights (the reference never sets them:
  // _hasExpandedTaxIds stays false, :97,:536-540)
  static void redistribute(size_t v, double *a, const QuantTree &t, const size_t *len) {
    const std::vector<size_t> ch = t.children(v);
    const size_t cs = ch.size();
    double csum = 0, wsum = 0;
    for (size_t c : ch) csum += a[c];
    double excess = a[v] - csum;
    if (excess < 0) excess = 0;
    if (csum == 0) return;
    const double expanded = 0;
    for (size_t c : ch) wsum += a[c] / (double)(len ? len[c] : 1) * ((excess - expanded) / (double)cs + 0);
    if (wsum == 0) wsum = 1;
    for (size_t c : ch) {
      a[c] += excess * (a[c] / (double)(len ? len[c] : 1) * ((excess - expanded) / (double)cs + 0)) / wsum;
      redistribute(c, a, t, len);
    }
  }

  // EMupdate (:186-234)
  static double em_update(const double *a0, double *a1, double *reads, const std::vector<QuantAssignment> &as, const QuantTree &t,
                          const size_t *len) {
    const size_t ts = t.size();
    memset(reads, 0, sizeof(double) * ts);
    for (const QuantAssignment &q : as) {
      double sum = 0;
      for (uint64_t x : q.targets) sum += a0[x];
      for (uint64_t x : q.targets) reads[x] += q.weight * a0[x] / sum;
    }
    double sum = 0;
    for (size_t i = 0; i < ts; ++i) sum += reads[i] / (double)len[i];
    for (size_t i = 0; i < ts; ++i) a1[i] = reads[i] / (double)len[i] / sum;
    tree_sum(0, a1, t);
    redistribute(0, a1, t, nullptr);
    double diff = 0;
    for (size_t i = 0; i < ts; ++i) diff += std::fabs(a0[i] - a1[i]);
    return diff;
  }

  // EstimateAbundanceWithEM (:236-281)
  static void em(const std::vector<QuantAssignment> &as, const QuantTree &t, const size_t *len, double *reads, double *abund) {
    for (const QuantAssignment &q : as)
      for (uint64_t x : q.targets) reads[x] += q.weight / (double)q.targets.size();
    tree_sum(t.root, reads, t);
    redistribute(t.root, reads, t, len);
    const size_t ts = t.size();
    const double factor = reads[t.root];
    for (size_t i = 0; i < ts; ++i) abund[i] = reads[i] / factor;
    std::vector<double> next(ts);
    for (int it = 0; it < 1000; ++it) {
      const double delta = em_update(abund, next.data(), reads, as, t, len);
      memcpy(abund, next.data(), sizeof(double) * ts);
      if (delta < 1e-6 && delta < 0.1 / (double)ts) break;
    }
  }

  // GetTaxLineagePathString (:300-350), canonical ranks only
  std::string lineage(size_t ctid, int style, bool use_name) const {
    std::vector<size_t> path;  // Taxonomy::GetTaxLineagePath (Taxonomy.hpp:977-993): up to, not including, the top node
    if (ctid >= tax_.node_cnt) {
      path.push_back(tax_.root);
    } else {
      size_t c = ctid;
      do {
        path.push_back(c);
        c = tax_.parent[c];
      } while (c != tax_.parent[c]);
    }
    std::reverse(path.begin(), path.end());
    std::string out;
    for (size_t i = 0; i < path.size(); ++i) {
      if (!canonical(path[i])) continue;
      if (style == QUANT_FORMAT_METAPHLAN && use_name) {
        const uint8_t r = tax_.rank[path[i]];
        char pre[4] = {0, '_', '_', 0};
        pre[0] = (r == TAX_RANK_SUPER_KINGDOM || r == TAX_RANK_ACELLULAR_ROOT) ? 'd' : tax_rank_string(r)[0];
        out += pre;
      }
      if (use_name) {
        out += tax_.tax_name[path[i]];
      } else {
        char buf[32];
        snprintf(buf, sizeof(buf), "%lu", (unsigned long)tax_.orig_taxid[path[i]]);
        out += buf;
      }
      if (i + 1 < path.size()) out += "|";
    }
    return out;
  }

  // OutputKreportDFS (:353-399)
  void kreport(const QuantTree &t, size_t v, int depth, int dist, char prev_sym, FILE *fp) const {
    if (read_count_[v] < 1e-6) return;
    char r[32];
    const uint8_t rk = tax_.rank[v];
    if (canonical(v) && rk != TAX_RANK_STRAIN) {
      r[0] = (rk == TAX_RANK_SUPER_KINGDOM || rk == TAX_RANK_ACELLULAR_ROOT) ? 'D' : (char)(tax_rank_string(rk)[0] - 'a' + 'A');
      r[1] = 0;
      dist = 0;
    } else if (prev_sym == '\0') {
      r[0] = 'R';
      r[1] = 0;
    } else {
      snprintf(r, sizeof(r), "%c%d", prev_sym, dist);
    }
    double child_reads = 0;
    const std::vector<size_t> ch = t.children(v);
    for (size_t c : ch) child_reads += read_count_[c];
    fprintf(fp, "%.2lf\t%.0lf\t%.0lf\t%s\t%lu\t", abund_[v] * 100, read_count_[v], read_count_[v] - child_reads, r,
            (unsigned long)tax_.orig_taxid[v]);
    for (int i = 0; i < depth; ++i) fprintf(fp, "  ");
    fprintf(fp, "%s\n", tax_.tax_name[v].c_str());
    for (size_t c : ch) kreport(t, c, depth + 1, dist + 1, r[0], fp);
  }
};

}  // namespace cfrb200
