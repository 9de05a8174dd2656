/*
 * cfr_oracle.c -- TEST INFRASTRUCTURE ONLY.  See cfr_oracle.h.
 *
 * CPU restatement of the reference's classification path.  All file:line
 * citations are relative to the reference tree (mourisl/centrifuger
 * v1.1.3-r347).  This is deliberately a literal, scalar, single-threaded
 * restatement: its job is to be obviously equal to the reference, not fast.
 */
#define _GNU_SOURCE
#include "cfr_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* small utilities                                                           */
/* ------------------------------------------------------------------------ */

#define DIV_CEIL(x, y) (((x) % (y)) ? ((x) / (y) + 1) : ((x) / (y))) /* Utils.hpp:30 */

static int popc64(uint64_t x) { return __builtin_popcountll(x); } /* Utils.hpp:92 */

static void die(const char *msg) {
  fprintf(stderr, "cfr_oracle: %s\n", msg);
  exit(2);
}

static void xread(void *dst, size_t sz, size_t cnt, FILE *fp) {
  if (cnt == 0) return;
  if (fread(dst, sz, cnt, fp) != cnt) die("short read in .cfr file");
}

static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) die("out of memory");
  return p;
}

/* ------------------------------------------------------------------------ */
/* Alphabet  (compactds/Alphabet.hpp)                                        */
/* ------------------------------------------------------------------------ */

typedef struct {
  uint64_t n;
  int32_t method;
  char *list;
  int32_t code[256];
  int16_t code_len[256];
} alphabet_t;

/* Alphabet.hpp:207-221 */
static void alphabet_load(alphabet_t *a, FILE *fp) {
  uint64_t space;
  memset(a, 0, sizeof(*a));
  xread(&space, 8, 1, fp);
  xread(&a->method, 4, 1, fp);
  xread(&a->n, 8, 1, fp);
  if (a->n != 0) {
    a->list = (char *)xmalloc(a->n);
    xread(a->list, 1, a->n, fp);
    xread(a->code, 4, 256, fp);
    xread(a->code_len, 2, 256, fp);
  }
}

/* Alphabet.hpp:169-176 */
static int alphabet_is_in(const alphabet_t *a, char c) {
  uint64_t i;
  for (i = 0; i < a->n; ++i)
    if (a->list[i] == c) return 1;
  return 0;
}

/* Alphabet.hpp:145-167; callers only pass characters that are IsIn */
static uint64_t alphabet_encode(const alphabet_t *a, char c, int *l) {
  if (a->method == 0) { /* ALPHABET_CODE_NOCODE */
    if (l) *l = 0;
    return (uint64_t)(unsigned char)c;
  }
  if (l) *l = a->code_len[(unsigned char)c];
  return (uint64_t)a->code[(unsigned char)c];
}

/* Alphabet.hpp:127-143 (plain method only; Huffman is never used for .cfr) */
static char alphabet_decode(const alphabet_t *a, uint64_t code) {
  if (a->method == 0) return (char)code;
  if (a->method != 1) die("huffman alphabet not supported by the oracle");
  return a->list[code];
}

/* ------------------------------------------------------------------------ */
/* Bitvector_Plain + DS_Rank9                                                */
/* ------------------------------------------------------------------------ */

typedef struct {
  uint64_t n;      /* bits */
  uint64_t *B;     /* Bitvector_Plain::_B */
  uint64_t word_cnt;
  uint64_t *R;     /* DS_Rank9::_R */
} bv_t;

/* Bitvector_Plain.hpp:198-221, DS_Rank.hpp:284-296, DS_Select.hpp:679-686 */
static void bv_load(bv_t *v, FILE *fp) {
  uint64_t space;
  int32_t rb, sb, select_speed, select_type;
  memset(v, 0, sizeof(*v));
  xread(&space, 8, 1, fp); /* Bitvector::_space */
  xread(&v->n, 8, 1, fp);
  xread(&rb, 4, 1, fp);
  xread(&sb, 4, 1, fp);
  xread(&select_speed, 4, 1, fp);
  xread(&select_type, 4, 1, fp);
  if (v->n > 0) {
    uint64_t words = DIV_CEIL(v->n, 64);
    uint64_t rspace, block_cnt, sspace, sn;
    int32_t speed;
    v->B = (uint64_t *)xmalloc(words * 8);
    xread(v->B, 8, words, fp);
    xread(&rspace, 8, 1, fp);
    xread(&v->word_cnt, 8, 1, fp);
    block_cnt = DIV_CEIL(v->word_cnt, 8);
    v->R = (uint64_t *)xmalloc(block_cnt * 2 * 8);
    xread(v->R, 8, block_cnt * 2, fp);
    xread(&sspace, 8, 1, fp);
    xread(&sn, 8, 1, fp);
    xread(&speed, 4, 1, fp);
    if (!(speed == 0 || sn == 0)) die("DS_Select tables present (select speed != 0): not a centrifuger index");
  }
}

static void bv_free(bv_t *v) {
  free(v->B);
  free(v->R);
}

/* Utils.hpp:244 (BitRead) via Bitvector_Plain.hpp:128 */
static int bv_access(const bv_t *v, uint64_t i) {
  return (int)((v->B[i >> 6] >> (i & 63)) & 1ull);
}

/* DS_Rank.hpp:255-273 */
static uint64_t bv_rank1(const bv_t *v, uint64_t i, int inclusive) {
  if (i >= v->n) i = v->n - 1; /* :259-260 (re-enters with n-1) */
  {
    const uint64_t wi = i >> 6;
    const uint64_t ri = (wi >> 3) * 2;
    const uint64_t t = (wi & 7) - 1;
    const uint64_t mask = (((1ull << (i & 63)) - 1ull) << inclusive) + (uint64_t)inclusive;
    return v->R[ri] + ((v->R[ri + 1] >> ((t + ((t >> 60) & 8)) * 9)) & 0x1ff) +
           (uint64_t)popc64(v->B[wi] & mask);
  }
}

/* Bitvector.hpp:45-57 */
static uint64_t bv_rank(const bv_t *v, int type, uint64_t i, int inclusive) {
  if (type == 1) return bv_rank1(v, i, inclusive);
  return i + (uint64_t)inclusive - bv_rank1(v, i, inclusive);
}

/* ------------------------------------------------------------------------ */
/* Sequence_WaveletTree<Bitvector_Plain>                                     */
/* ------------------------------------------------------------------------ */

typedef struct {
  uint64_t n;
  alphabet_t alphabet;
  int32_t node_cnt;
  bv_t *v;
  int32_t (*children)[2];
} wt_t;

/* Sequence.hpp:31-36 + Sequence_WaveletTree.hpp:313-327 and :29-35 */
static void wt_load(wt_t *t, FILE *fp) {
  uint64_t space;
  int32_t select_speed, i;
  memset(t, 0, sizeof(*t));
  xread(&space, 8, 1, fp);
  xread(&t->n, 8, 1, fp);
  alphabet_load(&t->alphabet, fp);
  xread(&t->node_cnt, 4, 1, fp);
  xread(&select_speed, 4, 1, fp);
  if (t->alphabet.n == 0) { /* empty tree */
    t->node_cnt = 0;
    return;
  }
  t->v = (bv_t *)xmalloc(sizeof(bv_t) * (size_t)t->node_cnt);
  t->children = (int32_t(*)[2])xmalloc(sizeof(int32_t[2]) * (size_t)t->node_cnt);
  for (i = 0; i < t->node_cnt; ++i) {
    uint64_t prefix;
    int32_t prefix_len;
    xread(&prefix, 8, 1, fp);
    xread(&prefix_len, 4, 1, fp);
    xread(t->children[i], 4, 2, fp);
    bv_load(&t->v[i], fp);
  }
}

static void wt_free(wt_t *t) {
  int i;
  for (i = 0; i < t->node_cnt; ++i) bv_free(&t->v[i]);
  free(t->v);
  free(t->children);
  free(t->alphabet.list);
}

/* Sequence_WaveletTree.hpp:215-232 */
static char wt_access(const wt_t *t, uint64_t i) {
  int l;
  uint64_t code = 0;
  int ti = 0;
  for (l = 0; ti != -1; ++l) {
    int b = bv_access(&t->v[ti], i);
    code = (code << 1) | (uint64_t)b;
    i = bv_rank(&t->v[ti], b, i, 1) - 1;
    ti = t->children[ti][b];
  }
  return alphabet_decode(&t->alphabet, code);
}

/* Sequence_WaveletTree.hpp:235-264 */
static uint64_t wt_rank(const wt_t *t, char c, uint64_t i, int inclusive) {
  int l = 0;
  uint64_t code = alphabet_encode(&t->alphabet, c, &l);
  int depth, ti = 0;
  if (!inclusive) {
    if (i == 0) return 0;
    --i;
  }
  for (depth = 0; depth < l; ++depth) {
    int b = (int)((code >> (l - depth - 1)) & 1);
    i = bv_rank(&t->v[ti], b, i, 1);
    if (i == 0 || depth == l - 1) break;
    --i;
    ti = t->children[ti][b];
  }
  return i;
}

/* Sequence_WaveletTree.hpp:268-293 */
static uint64_t wt_rank_and_test(const wt_t *t, char c, uint64_t i, int *is_c) {
  int l = 0;
  uint64_t code = alphabet_encode(&t->alphabet, c, &l);
  int depth, ti = 0;
  *is_c = 1;
  for (depth = 0; depth < l; ++depth) {
    int b = (int)((code >> (l - depth - 1)) & 1);
    if (*is_c && b != bv_access(&t->v[ti], i)) *is_c = 0;
    i = bv_rank(&t->v[ti], b, i, 1);
    if (i == 0 || depth == l - 1) break;
    --i;
    ti = t->children[ti][b];
  }
  return i;
}

/* ------------------------------------------------------------------------ */
/* FixedSizeElemArray                                                        */
/* ------------------------------------------------------------------------ */

typedef struct {
  uint64_t size; /* words allocated */
  int32_t l;
  uint64_t n;
  uint64_t *W;
} fsea_t;

/* FixedSizeElemArray.hpp:396-403 */
static void fsea_load(fsea_t *a, FILE *fp) {
  uint64_t words;
  memset(a, 0, sizeof(*a));
  xread(&a->size, 8, 1, fp);
  xread(&a->l, 4, 1, fp);
  xread(&a->n, 8, 1, fp);
  words = DIV_CEIL(a->n * (uint64_t)a->l, 64);
  a->W = (uint64_t *)calloc((a->size > words ? a->size : words) + 1, 8);
  if (!a->W) die("out of memory");
  xread(a->W, 8, words, fp);
}

/* FixedSizeElemArray.hpp:102-105 + Utils.hpp:197-219 (BitsRead) */
static uint64_t fsea_read(const fsea_t *a, uint64_t i) {
  const uint64_t s = i * (uint64_t)a->l, e = (i + 1) * (uint64_t)a->l - 1;
  const uint64_t is = s >> 6, ie = e >> 6;
  const int rs = (int)(s & 63);
  if (is == ie) {
    const uint64_t len = e - s + 1;
    const uint64_t m = (len >= 64) ? 0xffffffffffffffffull : ((1ull << len) - 1ull);
    return (a->W[is] >> rs) & m;
  } else {
    const int re = (int)(e & 63);
    return (a->W[is] >> rs) | ((a->W[ie] & ((1ull << (re + 1)) - 1ull)) << (64 - rs));
  }
}

/* ------------------------------------------------------------------------ */
/* Taxonomy (query-side subset)                                              */
/* ------------------------------------------------------------------------ */

enum { /* Taxonomy.hpp:25-59 */
  RANK_UNKNOWN = 0, RANK_STRAIN, RANK_SPECIES, RANK_GENUS, RANK_FAMILY, RANK_ORDER,
  RANK_CLASS, RANK_PHYLUM, RANK_KINGDOM, RANK_DOMAIN, RANK_FORMA, RANK_INFRA_CLASS,
  RANK_INFRA_ORDER, RANK_PARV_ORDER, RANK_SUB_CLASS, RANK_SUB_FAMILY, RANK_SUB_GENUS,
  RANK_SUB_KINGDOM, RANK_SUB_ORDER, RANK_SUB_PHYLUM, RANK_SUB_SPECIES, RANK_SUB_TRIBE,
  RANK_SUPER_CLASS, RANK_SUPER_FAMILY, RANK_SUPER_KINGDOM, RANK_SUPER_ORDER,
  RANK_SUPER_PHYLUM, RANK_TRIBE, RANK_VARIETAS, RANK_LIFE, RANK_ACELLULAR_ROOT, RANK_MAX
};

typedef struct { /* Taxonomy.hpp:61-82 */
  uint64_t parent;
  uint8_t rank;
  uint8_t leaf;
  uint8_t pad[6];
} tax_node_t;

typedef struct {
  uint64_t node_cnt, seq_cnt, extra_seq_cnt;
  tax_node_t *tree;
  uint64_t *orig_tax_id; /* MapID<uint64_t>::_toOrigElem */
  uint64_t orig_cnt;
  char **tax_name;
  uint64_t *seq_to_tax;
  char **seq_name;
  uint8_t rank_num[RANK_MAX];
  uint64_t root;
} taxonomy_t;

/* Taxonomy.hpp:100-144 */
static void tax_init_rank_num(taxonomy_t *t) {
  uint8_t rank = 0;
  t->rank_num[RANK_SUB_SPECIES] = rank;
  t->rank_num[RANK_STRAIN] = rank++;
  t->rank_num[RANK_SPECIES] = rank++;
  t->rank_num[RANK_SUB_GENUS] = rank;
  t->rank_num[RANK_GENUS] = rank++;
  t->rank_num[RANK_SUB_FAMILY] = rank;
  t->rank_num[RANK_FAMILY] = rank;
  t->rank_num[RANK_SUPER_FAMILY] = rank++;
  t->rank_num[RANK_SUB_ORDER] = rank;
  t->rank_num[RANK_INFRA_ORDER] = rank;
  t->rank_num[RANK_PARV_ORDER] = rank;
  t->rank_num[RANK_ORDER] = rank;
  t->rank_num[RANK_SUPER_ORDER] = rank++;
  t->rank_num[RANK_INFRA_CLASS] = rank;
  t->rank_num[RANK_SUB_CLASS] = rank;
  t->rank_num[RANK_CLASS] = rank;
  t->rank_num[RANK_SUPER_CLASS] = rank++;
  t->rank_num[RANK_SUB_PHYLUM] = rank;
  t->rank_num[RANK_PHYLUM] = rank;
  t->rank_num[RANK_SUPER_PHYLUM] = rank++;
  t->rank_num[RANK_SUB_KINGDOM] = rank;
  t->rank_num[RANK_KINGDOM] = rank++;
  t->rank_num[RANK_SUPER_KINGDOM] = rank;
  t->rank_num[RANK_ACELLULAR_ROOT] = rank;
  t->rank_num[RANK_DOMAIN] = rank++;
  t->rank_num[RANK_FORMA] = rank;
  t->rank_num[RANK_SUB_TRIBE] = rank;
  t->rank_num[RANK_TRIBE] = rank;
  t->rank_num[RANK_VARIETAS] = rank;
  t->rank_num[RANK_LIFE] = rank;
  t->rank_num[RANK_UNKNOWN] = rank;
}

/* Taxonomy.hpp:415-425 */
static char *tax_load_string(FILE *fp) {
  uint64_t len;
  char *s;
  xread(&len, 8, 1, fp);
  s = (char *)xmalloc(len + 1);
  xread(s, 1, len, fp);
  s[len] = '\0';
  return s;
}

/* Taxonomy.hpp:1259-1287, MapID.hpp:83-99 */
static void tax_load(taxonomy_t *t, FILE *fp) {
  uint64_t i;
  memset(t, 0, sizeof(*t));
  tax_init_rank_num(t);
  xread(&t->node_cnt, 8, 1, fp);
  xread(&t->seq_cnt, 8, 1, fp);
  xread(&t->extra_seq_cnt, 8, 1, fp);
  t->tree = (tax_node_t *)xmalloc(sizeof(tax_node_t) * t->node_cnt);
  xread(t->tree, sizeof(tax_node_t), t->node_cnt, fp);
  xread(&t->orig_cnt, 8, 1, fp);
  t->orig_tax_id = (uint64_t *)xmalloc(8 * t->orig_cnt);
  xread(t->orig_tax_id, 8, t->orig_cnt, fp);
  t->tax_name = (char **)xmalloc(sizeof(char *) * t->node_cnt);
  for (i = 0; i < t->node_cnt; ++i) t->tax_name[i] = tax_load_string(fp);
  t->seq_to_tax = (uint64_t *)xmalloc(8 * t->seq_cnt);
  xread(t->seq_to_tax, 8, t->seq_cnt, fp);
  t->seq_name = (char **)xmalloc(sizeof(char *) * (t->seq_cnt + t->extra_seq_cnt));
  for (i = 0; i < t->seq_cnt + t->extra_seq_cnt; ++i) t->seq_name[i] = tax_load_string(fp);
  /* FindRoot, Taxonomy.hpp:426-433 */
  t->root = t->node_cnt;
  for (i = 0; i < t->node_cnt; ++i)
    if (t->tree[i].parent == i) {
      t->root = i;
      break;
    }
}

static void tax_free(taxonomy_t *t) {
  uint64_t i;
  for (i = 0; i < t->node_cnt; ++i) free(t->tax_name[i]);
  for (i = 0; i < t->seq_cnt + t->extra_seq_cnt; ++i) free(t->seq_name[i]);
  free(t->tax_name);
  free(t->seq_name);
  free(t->tree);
  free(t->orig_tax_id);
  free(t->seq_to_tax);
}

/* Taxonomy.hpp:497-532 */
static const char *tax_rank_string(uint8_t rank) {
  switch (rank) {
    case RANK_STRAIN: return "strain";
    case RANK_SPECIES: return "species";
    case RANK_GENUS: return "genus";
    case RANK_FAMILY: return "family";
    case RANK_ORDER: return "order";
    case RANK_CLASS: return "class";
    case RANK_PHYLUM: return "phylum";
    case RANK_KINGDOM: return "kingdom";
    case RANK_DOMAIN: return "domain";
    case RANK_ACELLULAR_ROOT: return "acellular root";
    case RANK_FORMA: return "forma";
    case RANK_INFRA_CLASS: return "infraclass";
    case RANK_INFRA_ORDER: return "infraorder";
    case RANK_PARV_ORDER: return "parvorder";
    case RANK_SUB_CLASS: return "subclass";
    case RANK_SUB_FAMILY: return "subfamily";
    case RANK_SUB_GENUS: return "subgenus";
    case RANK_SUB_KINGDOM: return "subkingdom";
    case RANK_SUB_ORDER: return "suborder";
    case RANK_SUB_PHYLUM: return "subphylum";
    case RANK_SUB_SPECIES: return "subspecies";
    case RANK_SUB_TRIBE: return "subtribe";
    case RANK_SUPER_CLASS: return "superclass";
    case RANK_SUPER_FAMILY: return "superfamily";
    case RANK_SUPER_KINGDOM: return "superkingdom";
    case RANK_SUPER_ORDER: return "superorder";
    case RANK_SUPER_PHYLUM: return "superphylum";
    case RANK_TRIBE: return "tribe";
    case RANK_VARIETAS: return "varietas";
    case RANK_LIFE: return "life";
    default: return "no rank";
  }
}

/* Taxonomy.hpp:633-639 */
static uint64_t tax_orig_id(const taxonomy_t *t, uint64_t ctid) {
  if (ctid >= t->node_cnt) return t->orig_tax_id[t->root];
  return t->orig_tax_id[ctid];
}

/* Taxonomy.hpp:659-665 */
static uint8_t tax_rank_of(const taxonomy_t *t, uint64_t ctid) {
  if (ctid >= t->node_cnt) return RANK_UNKNOWN;
  return t->tree[ctid].rank;
}

/* Taxonomy.hpp:718-724 */
static uint64_t tax_seq_to_tax(const taxonomy_t *t, uint64_t seq_id) {
  if (seq_id < t->seq_cnt) return t->seq_to_tax[seq_id];
  return t->node_cnt;
}

/* a tiny ordered set of uint64 (stands in for std::map<size_t,int> keys) */
typedef struct {
  uint64_t *a;
  int n, cap;
} u64set;

static int u64set_find(const u64set *s, uint64_t x, int *pos) {
  int lo = 0, hi = s->n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (s->a[mid] < x) lo = mid + 1; else hi = mid;
  }
  *pos = lo;
  return lo < s->n && s->a[lo] == x;
}

static int u64set_insert(u64set *s, uint64_t x) { /* returns 1 if newly inserted */
  int pos;
  if (u64set_find(s, x, &pos)) return 0;
  if (s->n == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 16;
    s->a = (uint64_t *)realloc(s->a, sizeof(uint64_t) * (size_t)s->cap);
    if (!s->a) die("out of memory");
  }
  memmove(s->a + pos + 1, s->a + pos, sizeof(uint64_t) * (size_t)(s->n - pos));
  s->a[pos] = x;
  ++s->n;
  return 1;
}

/* Taxonomy.hpp:733-836.  `children` (may be NULL) is lcaChildTaxIds: when the answer is a backbone
 * node it receives, ascending, the distinct nodes directly below the answer where the inputs'
 * lineages leave it (:767-773, :811-812, :825-831); it is left untouched when the answer is the
 * root by default (all-root input :756, no shared backbone node :819) */
static uint64_t tax_lca(const taxonomy_t *t, const uint64_t *tax_ids, int tax_cnt, u64set *children) {
  int i, j, k;
  u64set *below = NULL; /* backboneChildTaxIds */
  uint64_t cur;
  uint64_t *path = NULL, *tmp = NULL;
  int *path_cnt = NULL;
  int path_len = 0, path_cap = 0, tmp_len = 0, tmp_cap = 0;
  int root_count = 0;
  uint64_t ret;

  for (i = 0; i < tax_cnt; ++i)
    if (tax_ids[i] != t->root) break;
  if (i < tax_cnt) k = i; else return t->root;

  cur = tax_ids[k];
  do {
    if (path_len == path_cap) {
      path_cap = path_cap ? path_cap * 2 : 32;
      path = (uint64_t *)realloc(path, 8 * (size_t)path_cap);
      path_cnt = (int *)realloc(path_cnt, sizeof(int) * (size_t)path_cap);
    }
    path[path_len] = cur;
    path_cnt[path_len++] = 1;
    cur = t->tree[cur].parent;
  } while (cur != t->tree[cur].parent);
  if (path_len == path_cap) {
    path_cap += 1;
    path = (uint64_t *)realloc(path, 8 * (size_t)path_cap);
    path_cnt = (int *)realloc(path_cnt, sizeof(int) * (size_t)path_cap);
  }
  path[path_len] = t->root;
  path_cnt[path_len++] = 1;
  if (children) { /* :767-773 */
    below = (u64set *)calloc((size_t)path_len, sizeof(u64set));
    for (j = 1; j < path_len; ++j) u64set_insert(&below[j], path[j - 1]);
  }

  for (i = 0; i < tax_cnt; ++i) {
    int ib, it;
    if (i == k) continue;
    tmp_len = 0;
    cur = tax_ids[i];
    if (cur == t->tree[cur].parent) {
      ++root_count;
      continue;
    }
    do {
      if (tmp_len + 1 >= tmp_cap) {
        tmp_cap = tmp_cap ? tmp_cap * 2 : 32;
        tmp = (uint64_t *)realloc(tmp, 8 * (size_t)tmp_cap);
      }
      tmp[tmp_len++] = cur;
      cur = t->tree[cur].parent;
    } while (cur != t->tree[cur].parent);
    tmp[tmp_len++] = t->root;
    for (ib = path_len - 1, it = tmp_len - 1; ib >= 0 && it >= 0; --ib, --it) {
      if (tmp[it] != path[ib]) break;
      path_cnt[ib] += 1;
    }
    if (children && it >= 0 && ib + 1 < path_len) u64set_insert(&below[ib + 1], tmp[it]); /* :811-812 */
  }
  for (j = 0; j < path_len; ++j)
    if (path_cnt[j] == tax_cnt - root_count) break;
  ret = (j >= path_len) ? t->root : path[j];
  if (children) {
    if (j < path_len) { /* :825-831 */
      children->n = 0;
      for (i = 0; i < below[j].n; ++i) u64set_insert(children, below[j].a[i]);
    }
    for (j = 0; j < path_len; ++j) free(below[j].a);
    free(below);
  }
  free(path);
  free(path_cnt);
  free(tmp);
  return ret;
}

/* plain append (the lists of promotedChildTaxIds are vectors: order kept, duplicates allowed) */
static void u64vec_push(u64set *v, uint64_t x) {
  if (v->n == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 8;
    v->a = (uint64_t *)realloc(v->a, sizeof(uint64_t) * (size_t)v->cap);
    if (!v->a) die("out of memory");
  }
  v->a[v->n++] = x;
}

/* Taxonomy.hpp:839-973.  `lists`/`n_lists` (may be NULL) are promotedChildTaxIds: *n_lists vectors,
 * allocated here (caller frees each .a and the array); the caller prints them only when
 * *n_lists equals the number of ids returned (Classifier.hpp:823) */
static int tax_reduce(const taxonomy_t *t, const uint64_t *tax_ids, int tax_cnt, int k,
                      uint64_t *out, int cap, u64set **lists, int *n_lists) {
  int i, n_out = 0;
  u64set level[RANK_MAX];
  uint8_t ri;
#define PUSH(x) do { if (n_out < cap) out[n_out] = (x); ++n_out; } while (0)
  if (lists) {
    *lists = NULL;
    *n_lists = 0;
  }
  if (tax_cnt <= k) { /* :847-851 */
    for (i = 0; i < tax_cnt; ++i) PUSH(tax_ids[i]);
    return n_out;
  }
  for (i = 0; i < tax_cnt; ++i) /* :855-882 */
    if (tax_ids[i] >= t->node_cnt) {
      PUSH(t->node_cnt);
      if (lists) { /* :871-878: one list holding every input id as given */
        int j;
        *lists = (u64set *)calloc(1, sizeof(u64set));
        *n_lists = 1;
        for (j = 0; j < tax_cnt; ++j) u64vec_push(&(*lists)[0], tax_ids[j]);
      }
      return n_out;
    }
  if (k == 1) { /* :884-901 */
    if (lists) {
      *lists = (u64set *)calloc(1, sizeof(u64set));
      *n_lists = 1;
      PUSH(tax_lca(t, tax_ids, tax_cnt, &(*lists)[0]));
    } else {
      PUSH(tax_lca(t, tax_ids, tax_cnt, NULL));
    }
    return n_out;
  }
  memset(level, 0, sizeof(level));
  for (i = 0; i < tax_cnt; ++i) { /* :904-930 */
    uint64_t cur = tax_ids[i];
    uint8_t prev_rank_num = 0;
    u64set_insert(&level[prev_rank_num], cur);
    do {
      uint8_t rank_num = t->rank_num[t->tree[cur].rank];
      if (rank_num != t->rank_num[RANK_UNKNOWN] && rank_num > prev_rank_num) {
        int pos;
        for (ri = (uint8_t)(rank_num - 1); ri > prev_rank_num; --ri) u64set_insert(&level[ri], cur);
        if (!u64set_find(&level[rank_num], cur, &pos))
          u64set_insert(&level[rank_num], cur);
        else
          break;
        prev_rank_num = rank_num;
      }
      cur = t->tree[cur].parent;
    } while (cur != t->tree[cur].parent);
  }
  for (ri = 0; ri < t->rank_num[RANK_UNKNOWN]; ++ri) /* :933-936 */
    if (level[ri].n <= k) break;
  for (i = 0; i < level[ri].n; ++i) PUSH(level[ri].a[i]);
  if (n_out == 0) {
    PUSH(t->root);
  } else if (lists && ri > 0) { /* :941-971: the level below, grouped by the promoted node above it */
    *lists = (u64set *)calloc((size_t)level[ri].n, sizeof(u64set));
    *n_lists = level[ri].n;
    for (i = 0; i < level[ri - 1].n; ++i) {
      uint64_t cur = level[ri - 1].a[i];
      while (cur != t->tree[cur].parent) {
        uint8_t rank_num;
        cur = t->tree[cur].parent;
        rank_num = t->rank_num[t->tree[cur].rank];
        if (rank_num > ri) break;
        if (rank_num == ri) {
          int pos;
          if (u64set_find(&level[ri], cur, &pos)) u64vec_push(&(*lists)[pos], level[ri - 1].a[i]);
          break;
        }
      }
    }
  }
  for (i = 0; i < RANK_MAX; ++i) free(level[i].a);
#undef PUSH
  return n_out;
}

/* ------------------------------------------------------------------------ */
/* the index object                                                          */
/* ------------------------------------------------------------------------ */

typedef struct { uint64_t start, len; } range_t;
typedef struct { uint64_t row, val; } selsa_t;

struct cfr_oracle {
  /* FMIndex (FMIndex.hpp:191-199) */
  uint64_t n, alphabet_bits, first_isa;
  char last_chr;
  /* Sequence_RunBlock (Sequence_RunBlock.hpp:15-20) */
  uint64_t rb_n;
  alphabet_t rb_alphabet;
  uint64_t b, block_cnt;
  bv_t use_run_block;
  wt_t wavelet_seq, run_block_seq;
  alphabet_t alphabets, plain_coder;
  uint64_t *C;
  /* _FMIndexAuxData (FMIndex.hpp:13-41) */
  uint64_t aux_n;
  int32_t sample_strategy, sample_rate;
  uint64_t sample_size, precompute_width, precompute_size, adjusted_sa0;
  fsea_t sampled_sa;
  range_t *precomputed;
  uint64_t max_lcp;
  uint64_t sel_cnt;
  int32_t sel_filter_rate;
  selsa_t *sel; /* sorted by row (std::map order) */
  uint64_t *sel_filter;
  uint8_t has_end_marker;
  fsea_t end_marker_sa;
  taxonomy_t tax;
  cfr_oracle_counters cnt;
};

/* FMIndex.hpp:136-185 */
static void aux_load(cfr_oracle *o, FILE *fp) {
  uint64_t i;
  xread(&o->aux_n, 8, 1, fp);
  xread(&o->sample_strategy, 4, 1, fp);
  xread(&o->sample_rate, 4, 1, fp);
  xread(&o->sample_size, 8, 1, fp);
  xread(&o->precompute_width, 8, 1, fp);
  xread(&o->precompute_size, 8, 1, fp);
  xread(&o->adjusted_sa0, 8, 1, fp);
  fsea_load(&o->sampled_sa, fp);
  o->precomputed = (range_t *)xmalloc(sizeof(range_t) * o->precompute_size);
  xread(o->precomputed, sizeof(range_t), o->precompute_size, fp);
  xread(&o->max_lcp, 8, 1, fp);
  if (o->max_lcp > 0) { /* two bit arrays, unused by the classifier */
    uint64_t words = DIV_CEIL(o->aux_n, 64);
    if (fseek(fp, (long)(words * 8 * 2), SEEK_CUR) != 0) die("seek failed");
  }
  xread(&o->sel_cnt, 8, 1, fp);
  xread(&o->sel_filter_rate, 4, 1, fp);
  if (o->sel_cnt > 0) {
    uint64_t fbits = DIV_CEIL(o->aux_n, (uint64_t)o->sel_filter_rate);
    o->sel_filter = (uint64_t *)calloc(DIV_CEIL(fbits, 64) + 1, 8);
    o->sel = (selsa_t *)xmalloc(sizeof(selsa_t) * o->sel_cnt);
    for (i = 0; i < o->sel_cnt; ++i) {
      uint64_t pair[2], fb;
      xread(pair, 8, 2, fp);
      o->sel[i].row = pair[0];
      o->sel[i].val = pair[1];
      fb = pair[0] / (uint64_t)o->sel_filter_rate;
      o->sel_filter[fb >> 6] |= 1ull << (fb & 63);
    }
    /* the file stores std::map order (ascending row); keep a defensive check */
    for (i = 1; i < o->sel_cnt; ++i)
      if (o->sel[i - 1].row >= o->sel[i].row) die("selectedSA not strictly ascending");
  }
  if (fread(&o->has_end_marker, 1, 1, fp) != 1) o->has_end_marker = 0; /* :178-181 */
  if (o->has_end_marker) fsea_load(&o->end_marker_sa, fp);
}

cfr_oracle *cfr_oracle_open(const char *idx_prefix) {
  cfr_oracle *o = (cfr_oracle *)calloc(1, sizeof(*o));
  char *name = (char *)xmalloc(strlen(idx_prefix) + 32);
  FILE *fp;
  uint64_t space, sigma;
  if (!o) die("out of memory");

  /* Classifier.hpp:907-911 -> FMIndex.hpp:588-606 */
  sprintf(name, "%s.1.cfr", idx_prefix);
  fp = fopen(name, "rb");
  if (!fp) { free(name); free(o); return NULL; }
  xread(&o->n, 8, 1, fp);
  xread(&o->alphabet_bits, 8, 1, fp);
  xread(&o->first_isa, 8, 1, fp);
  xread(&o->last_chr, 1, 1, fp);
  /* Sequence_RunBlock.hpp:478-488 */
  xread(&space, 8, 1, fp);
  xread(&o->rb_n, 8, 1, fp);
  alphabet_load(&o->rb_alphabet, fp);
  xread(&o->b, 8, 1, fp);
  xread(&o->block_cnt, 8, 1, fp);
  bv_load(&o->use_run_block, fp);
  wt_load(&o->wavelet_seq, fp);
  wt_load(&o->run_block_seq, fp);
  alphabet_load(&o->alphabets, fp);
  alphabet_load(&o->plain_coder, fp);
  sigma = o->plain_coder.n;
  o->C = (uint64_t *)xmalloc(8 * (sigma + 1));
  xread(o->C, 8, sigma + 1, fp);
  aux_load(o, fp);
  {
    long here = ftell(fp), end;
    fseek(fp, 0, SEEK_END);
    end = ftell(fp);
    if (here != end) die(".1.cfr: trailing bytes after parse (grammar mismatch)");
  }
  fclose(fp);

  /* Classifier.hpp:913-918 */
  sprintf(name, "%s.2.cfr", idx_prefix);
  fp = fopen(name, "rb");
  if (!fp) die("cannot open .2.cfr");
  tax_load(&o->tax, fp);
  fclose(fp);
  free(name);
  return o;
}

void cfr_oracle_close(cfr_oracle *o) {
  if (!o) return;
  bv_free(&o->use_run_block);
  wt_free(&o->wavelet_seq);
  wt_free(&o->run_block_seq);
  free(o->rb_alphabet.list);
  free(o->alphabets.list);
  free(o->plain_coder.list);
  free(o->C);
  free(o->sampled_sa.W);
  free(o->precomputed);
  free(o->sel);
  free(o->sel_filter);
  free(o->end_marker_sa.W);
  tax_free(&o->tax);
  free(o);
}

void cfr_oracle_default_param(cfr_oracle_param *p) { /* Classifier.hpp:28-37 */
  p->max_result = 1;
  p->min_hit_len = 0;
  p->max_result_per_hit_factor = 40;
  p->pad_ = 0;
  p->consider_secondary_hit_len = 2000;
  p->consider_secondary_score_factor = 0.995;
}

uint64_t cfr_oracle_scalar(const cfr_oracle *o, int which) {
  switch (which) {
    case 0: return o->n;
    case 1: return o->b;
    case 2: return o->block_cnt;
    case 3: return o->first_isa;
    case 4: return (uint64_t)(unsigned char)o->last_chr;
    case 5: return (uint64_t)o->sample_rate;
    case 6: return (uint64_t)o->sampled_sa.l;
    case 7: return o->sample_size;
    case 8: return o->precompute_width;
    case 9: return o->sel_cnt;
    case 10: return o->tax.node_cnt;
    case 11: return o->tax.seq_cnt;
    case 12: return o->tax.extra_seq_cnt;
    case 13: return o->tax.root;
    case 14: return (uint64_t)cfr_oracle_infer_min_hit_len(o);
    case 15: return o->wavelet_seq.n;
    case 16: return o->run_block_seq.n;
    case 17: return o->adjusted_sa0;
    case 18: case 19: case 20: case 21: case 22:
      return ((uint64_t)(which - 18) <= o->plain_coder.n) ? o->C[which - 18] : 0;
    default: return 0;
  }
}

/* ------------------------------------------------------------------------ */
/* Sequence_RunBlock                                                         */
/* ------------------------------------------------------------------------ */

/* Sequence_RunBlock.hpp:360-376 */
static char rb_access(cfr_oracle *o, uint64_t i) {
  uint64_t bi = i / o->b;
  int type = bv_access(&o->use_run_block, bi);
  ++o->cnt.n_access;
  if (type == 0) {
    uint64_t r = bv_rank(&o->use_run_block, 1, bi, 1);
    i -= o->b * r;
    return wt_access(&o->wavelet_seq, i);
  } else {
    uint64_t r = bv_rank(&o->use_run_block, 0, bi, 1);
    i -= o->b * r;
    return wt_access(&o->run_block_seq, i / o->b);
  }
}

/* Sequence_RunBlock.hpp:378-416 */
static uint64_t rb_rank(cfr_oracle *o, char c, uint64_t i, int inclusive) {
  uint64_t bi, ranki, other_ranki, ret;
  int type;
  ++o->cnt.n_rank;
  if (!inclusive) {
    if (i == 0) return 0;
    --i;
  }
  bi = i / o->b;
  type = bv_access(&o->use_run_block, bi);
  ranki = (o->b < o->rb_n) ? bv_rank(&o->use_run_block, type, bi, 1) : 1;
  other_ranki = (bi + 1) - ranki;
  if (type == 0) {
    ret = wt_rank(&o->wavelet_seq, c, (ranki - 1) * o->b + i % o->b, 1);
  } else {
    int in_run = 1;
    uint64_t rb = wt_rank_and_test(&o->run_block_seq, c, ranki - 1, &in_run);
    if (in_run)
      ret = (rb - 1) * o->b + i % o->b + 1;
    else
      ret = rb * o->b;
  }
  if (other_ranki == 0) return ret;
  if (type == 0)
    ret += wt_rank(&o->run_block_seq, c, other_ranki - 1, 1) * o->b;
  else
    ret += wt_rank(&o->wavelet_seq, c, other_ranki * o->b - 1, 1);
  return ret;
}

uint64_t cfr_oracle_bwt_rank(cfr_oracle *o, char c, uint64_t i, int inclusive) {
  return rb_rank(o, c, i, inclusive);
}
char cfr_oracle_bwt_access(cfr_oracle *o, uint64_t i) { return rb_access(o, i); }

/* ------------------------------------------------------------------------ */
/* FMIndex                                                                   */
/* ------------------------------------------------------------------------ */

/* FMIndex.hpp:352-362 */
static uint64_t fm_rank(cfr_oracle *o, char c, uint64_t p, int inclusive) {
  uint64_t ret = rb_rank(o, c, p, inclusive);
  if (c == o->last_chr && (p < o->first_isa || (!inclusive && p == o->first_isa))) ++ret;
  return ret;
}
uint64_t cfr_oracle_fm_rank(cfr_oracle *o, char c, uint64_t p, int inclusive) {
  return fm_rank(o, c, p, inclusive);
}

/* FMIndex.hpp:364-379 */
static void fm_backward_extend(cfr_oracle *o, char c, uint64_t sp, uint64_t ep,
                               uint64_t *next_sp, uint64_t *next_ep) {
  uint64_t offset = o->C[alphabet_encode(&o->plain_coder, c, NULL)];
  ++o->cnt.n_extend;
  *next_sp = offset + fm_rank(o, c, sp, 0) + 1 - 1;
  if (sp != ep)
    *next_ep = offset + fm_rank(o, c, ep, 1) - 1;
  else
    *next_ep = *next_sp + ((rb_access(o, ep) == c) ? 0 : (uint64_t)-1);
}

/* FMIndex.hpp:382-386 */
static uint64_t fm_lf(cfr_oracle *o, char c, uint64_t p) {
  uint64_t offset = o->C[alphabet_encode(&o->plain_coder, c, NULL)];
  return offset + fm_rank(o, c, p, 1) - 1;
}

/* FMIndex.hpp:388-422 */
static uint64_t fm_initial_range(cfr_oracle *o, const char *s, uint64_t m, uint64_t *sp,
                                 uint64_t *ep) {
  uint64_t i;
  if (o->precompute_width > 0) {
    uint64_t w = 0;
    for (i = 0; i < o->precompute_width; ++i) {
      if (!alphabet_is_in(&o->alphabets, s[m - 1 - i])) {
        *sp = 1;
        *ep = 0;
        return i;
      }
      w = (w << o->alphabet_bits) | alphabet_encode(&o->plain_coder, s[m - 1 - i], NULL);
    }
    if (o->precomputed[w].len == 0) {
      *sp = 1;
      *ep = 0;
      return o->precompute_width - 1;
    }
    *sp = o->precomputed[w].start;
    *ep = *sp + o->precomputed[w].len - 1;
    return o->precompute_width;
  }
  *sp = 0;
  *ep = o->n - 1;
  return 0;
}

/* FMIndex.hpp:487-510 */
static uint64_t fm_backward_search(cfr_oracle *o, const char *s, uint64_t m, uint64_t *sp,
                                   uint64_t *ep) {
  uint64_t l, next_sp, next_ep;
  if (m < o->precompute_width) return 0;
  ++o->cnt.n_search;
  l = fm_initial_range(o, s, m, sp, ep);
  if (l < o->precompute_width) return l;
  next_sp = *sp;
  next_ep = *ep;
  while (l < m) {
    if (!alphabet_is_in(&o->alphabets, s[m - 1 - l])) break;
    fm_backward_extend(o, s[m - 1 - l], *sp, *ep, &next_sp, &next_ep);
    if (next_sp > next_ep || next_ep > o->n) break;
    *sp = next_sp;
    *ep = next_ep;
    ++l;
  }
  return l;
}

uint64_t cfr_oracle_backward_search(cfr_oracle *o, const char *s, uint64_t m, uint64_t *sp,
                                    uint64_t *ep) {
  return fm_backward_search(o, s, m, sp, ep);
}

/* FMIndex.hpp:203-231 */
static int fm_get_sampled_sa(cfr_oracle *o, uint64_t i, uint64_t *sa) {
  if (i == o->first_isa) {
    *sa = o->adjusted_sa0;
    return 1;
  } else if (i % (uint64_t)o->sample_rate == 0) {
    *sa = fsea_read(&o->sampled_sa, i / (uint64_t)o->sample_rate);
    return 1;
  } else if (o->sel_filter) {
    uint64_t fb = i / (uint64_t)o->sel_filter_rate;
    if ((o->sel_filter[fb >> 6] >> (fb & 63)) & 1ull) {
      uint64_t lo = 0, hi = o->sel_cnt;
      while (lo < hi) {
        uint64_t mid = (lo + hi) / 2;
        if (o->sel[mid].row < i) lo = mid + 1; else hi = mid;
      }
      if (lo < o->sel_cnt && o->sel[lo].row == i) {
        *sa = o->sel[lo].val;
        return 1;
      }
    }
  } else if (o->has_end_marker && i < o->end_marker_sa.n) {
    *sa = fsea_read(&o->end_marker_sa, i);
    return 1;
  }
  return 0;
}

/* FMIndex.hpp:514-524 */
static uint64_t fm_backward_to_sampled_sa(cfr_oracle *o, uint64_t i, uint64_t *l) {
  uint64_t ret = 0;
  *l = 0;
  while (!fm_get_sampled_sa(o, i, &ret)) {
    i = fm_lf(o, rb_access(o, i), i);
    ++*l;
    ++o->cnt.n_lf;
  }
  ++o->cnt.n_locate;
  return ret;
}

uint64_t cfr_oracle_locate(cfr_oracle *o, uint64_t row, uint64_t *steps) {
  uint64_t l;
  uint64_t r = fm_backward_to_sampled_sa(o, row, &l);
  if (steps) *steps = l;
  return r;
}

/* ------------------------------------------------------------------------ */
/* Classifier                                                                */
/* ------------------------------------------------------------------------ */

typedef struct {
  cfr_oracle_hit *a;
  int n, cap;
} hitvec;

static void hv_push(hitvec *v, cfr_oracle_hit h) {
  if (v->n == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 16;
    v->a = (cfr_oracle_hit *)realloc(v->a, sizeof(cfr_oracle_hit) * (size_t)v->cap);
    if (!v->a) die("out of memory");
  }
  v->a[v->n++] = h;
}
static void hv_append(hitvec *v, const hitvec *w) {
  int i;
  for (i = 0; i < w->n; ++i) hv_push(v, w->a[i]);
}

static cfr_oracle_hit mk_hit(uint64_t sp, uint64_t ep, int l, int offset, int strand) {
  cfr_oracle_hit h;
  h.sp = sp; h.ep = ep; h.l = l; h.offset = offset; h.strand = strand; h.pad_ = 0;
  return h;
}

/* Classifier.hpp:113-129 (nucleotide: mhl starts at 23) */
int cfr_oracle_infer_min_hit_len(const cfr_oracle *o) {
  int mhl = 23;
  uint64_t alphabet_size = o->alphabets.n;
  uint64_t kmerspace = 1, n = o->n;
  int i;
  for (i = 0; i < mhl; ++i) kmerspace *= alphabet_size; /* Utils::PowerInt */
  kmerspace /= 2;
  for (; mhl <= 32; ++mhl) {
    if (kmerspace >= 100 * n) break;
    kmerspace *= alphabet_size;
  }
  return mhl;
}

static int eff_min_hit_len(const cfr_oracle *o, const cfr_oracle_param *p) {
  return p->min_hit_len > 0 ? p->min_hit_len : cfr_oracle_infer_min_hit_len(o);
}

/* Classifier.hpp:243-252 (nucleotide: _scoreHitLenAdjust = 15) */
static uint64_t hit_score_len(int l, int min_hit_len) {
  if (l < min_hit_len) return 0;
  return (uint64_t)(l - 15) * (uint64_t)(l - 15);
}

/* Classifier.hpp:261-271 */
static uint64_t hits_score(const hitvec *v, int min_hit_len) {
  uint64_t s = 0;
  int i;
  for (i = 0; i < v->n; ++i) s += hit_score_len(v->a[i].l, min_hit_len);
  return s;
}

/* Classifier.hpp:99-111 with the table from :846-856 */
static char *reverse_complement_dup(const char *r, int len) {
  char *rc = (char *)xmalloc((size_t)len + 1);
  int i;
  for (i = 0; i < len; ++i) {
    char c = r[len - 1 - i], d = 'N';
    if (c == 'A') d = 'T';
    else if (c == 'C') d = 'G';
    else if (c == 'G') d = 'C';
    else if (c == 'T') d = 'A';
    rc[i] = d;
  }
  rc[len] = '\0';
  return rc;
}

/* Classifier.hpp:274-293 */
static void get_hits_from_read(cfr_oracle *o, int min_hit_len, const char *r, int len,
                               hitvec *hits) {
  uint64_t sp = 0, ep = 0;
  int l = 0;
  int remaining = len;
  while (remaining >= min_hit_len) {
    l = (int)fm_backward_search(o, r, (uint64_t)remaining, &sp, &ep);
    if (l >= min_hit_len && sp <= ep) hv_push(hits, mk_hit(sp, ep, l, len - remaining, 0));
    remaining -= (l + 1);
  }
}

/* Classifier.hpp:303-401 */
static void adjust_hit_boundary(cfr_oracle *o, const char *r, const char *rc, int len,
                                hitvec *strand_hits /* [2] */) {
  int i, j, k;
  int hit_size[2];
  uint64_t sp = 0, ep = 0;
  int l;
  int need_fix[2] = {0, 0};
  if (!strand_hits[0].n || !strand_hits[1].n) return;
  hit_size[0] = strand_hits[0].n;
  hit_size[1] = strand_hits[1].n;
  j = hit_size[0] - 1;
  for (i = 0; i < hit_size[1]; ++i) {
    int left, right;
    right = len - strand_hits[1].a[i].offset - 1;
    left = right - strand_hits[1].a[i].l + 1;
    for (; j >= 0; --j) {
      int rc_left, rc_right;
      rc_left = strand_hits[0].a[j].offset;
      rc_right = rc_left + strand_hits[0].a[j].l - 1;
      if (rc_left >= right) continue;
      if (left >= rc_right) break;
      if (left == rc_left && right == rc_right) break;
      if (left < rc_left && rc_right < right) break;
      if (rc_left < left && right < rc_right) break;
      if (rc_right > right) {
        l = (int)fm_backward_search(o, r, (uint64_t)(rc_right + 1), &sp, &ep);
        if (rc_right - l + 1 == left && sp <= ep) {
          strand_hits[1].a[i] = mk_hit(sp, ep, l, len - rc_right - 1, 1);
          need_fix[1] = 1;
        }
      }
      if (left < rc_left) {
        l = (int)fm_backward_search(o, rc, (uint64_t)(len - left), &sp, &ep);
        if (left + l - 1 == rc_right && sp <= ep) {
          strand_hits[0].a[j] = mk_hit(sp, ep, l, left, -1);
          need_fix[0] = 1;
        }
      }
    }
  }
  for (k = 0; k <= 1; ++k) { /* :361-400 */
    cfr_oracle_hit *h = strand_hits[k].a;
    if (!need_fix[k]) continue;
    for (i = 0; i < hit_size[k] - 1; ++i) {
      int starti = h[i].offset;
      int endi = starti + h[i].l - 1;
      for (j = i + 1; j < hit_size[k]; ++j) {
        int startj = h[j].offset, endj;
        if (startj > endi) break;
        endj = startj + h[j].l - 1;
        if (h[j].l >= h[i].l) {
          h[i].l = startj - starti;
          break;
        } else {
          if (endj <= endi)
            h[j].l = 0;
          else {
            h[j].offset = endi + 1;
            h[j].l = endj - (endi + 1) + 1;
            break;
          }
        }
      }
    }
  }
}

/* Classifier.hpp:509-583 (nucleotide branch) */
static void search_forward_and_reverse(cfr_oracle *o, int min_hit_len, const char *r1,
                                       const char *r2, hitvec *hits) {
  int i, k;
  int r1len = (int)strlen(r1);
  char *rc_r1 = reverse_complement_dup(r1, r1len);
  char *rc_r2 = NULL;
  hitvec strand_hits[2] = {{0, 0, 0}, {0, 0, 0}};
  uint64_t strand_score[2];

  get_hits_from_read(o, min_hit_len, r1, r1len, &strand_hits[1]);
  get_hits_from_read(o, min_hit_len, rc_r1, r1len, &strand_hits[0]);
  adjust_hit_boundary(o, r1, rc_r1, r1len, strand_hits);

  if (r2) {
    int r2len = (int)strlen(r2);
    hitvec r2_hits[2] = {{0, 0, 0}, {0, 0, 0}};
    rc_r2 = reverse_complement_dup(r2, r2len);
    get_hits_from_read(o, min_hit_len, r2, r2len, &r2_hits[1]);
    get_hits_from_read(o, min_hit_len, rc_r2, r2len, &r2_hits[0]);
    adjust_hit_boundary(o, r2, rc_r2, r2len, r2_hits);
    for (i = 0; i <= 1; ++i) hv_append(&strand_hits[i], &r2_hits[1 - i]);
    free(r2_hits[0].a);
    free(r2_hits[1].a);
  }
  for (k = 0; k < 2; ++k) {
    for (i = 0; i < strand_hits[k].n; ++i) strand_hits[k].a[i].strand = 2 * k - 1;
    strand_score[k] = hits_score(&strand_hits[k], min_hit_len);
  }
  hits->n = 0;
  if (strand_score[1] > strand_score[0] + strand_score[0] / 100)
    hv_append(hits, &strand_hits[1]);
  else if (strand_score[0] > strand_score[1] + strand_score[1] / 100)
    hv_append(hits, &strand_hits[0]);
  else {
    hv_append(hits, &strand_hits[1]);
    hv_append(hits, &strand_hits[0]);
  }
  free(strand_hits[0].a);
  free(strand_hits[1].a);
  free(rc_r1);
  free(rc_r2);
}

int cfr_oracle_search(cfr_oracle *o, const cfr_oracle_param *p, const char *r1,
                      const char *r2, cfr_oracle_hit *out, int cap) {
  hitvec hits = {0, 0, 0};
  int i, n;
  search_forward_and_reverse(o, eff_min_hit_len(o, p), r1, r2, &hits);
  n = hits.n;
  for (i = 0; i < n && i < cap; ++i) out[i] = hits.a[i];
  free(hits.a);
  return n;
}

/* per-strand std::map<size_t,_seqHitRecord> (Classifier.hpp:590) */
typedef struct { uint64_t seq_id, score; int hit_length; } seq_rec;
typedef struct { seq_rec *a; int n, cap; } recmap;

static seq_rec *recmap_find(recmap *m, uint64_t seq_id, int *pos) {
  int lo = 0, hi = m->n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (m->a[mid].seq_id < seq_id) lo = mid + 1; else hi = mid;
  }
  *pos = lo;
  if (lo < m->n && m->a[lo].seq_id == seq_id) return &m->a[lo];
  return NULL;
}

/* operator[]: default-constructs {0,0,0} when absent */
static seq_rec *recmap_get(recmap *m, uint64_t seq_id) {
  int pos;
  seq_rec *r = recmap_find(m, seq_id, &pos);
  if (r) return r;
  if (m->n == m->cap) {
    m->cap = m->cap ? m->cap * 2 : 16;
    m->a = (seq_rec *)realloc(m->a, sizeof(seq_rec) * (size_t)m->cap);
    if (!m->a) die("out of memory");
  }
  memmove(m->a + pos + 1, m->a + pos, sizeof(seq_rec) * (size_t)(m->n - pos));
  m->a[pos].seq_id = seq_id; /* the map key */
  m->a[pos].score = 0;
  m->a[pos].hit_length = 0;
  ++m->n;
  return &m->a[pos];
}

/* the child lists of one result (Classifier.hpp:807-838, outputExpandedResult) */
typedef struct {
  uint64_t *child; /* compact tax ids, list after list */
  int child_cap;
  int32_t *cnt;    /* entries per reported id (64 slots) */
  int total;
} expand_out;

/* Classifier.hpp:585-843 */
static void classification_from_hits(cfr_oracle *o, const cfr_oracle_param *p, int min_hit_len,
                                     const hitvec *hv, cfr_oracle_result *res, expand_out *ex) {
  int i, k;
  uint64_t j;
  const cfr_oracle_hit *hits = hv->a;
  int hit_cnt = hv->n;
  recmap rec[2] = {{0, 0, 0}, {0, 0, 0}};
  seq_rec prev_uniq = {0, 0, 0};
  int mix_strand = 0;
  uint64_t best = 0, second = 0, best_len = 0, second_len = 0;
  u64set used = {0, 0, 0};
  uint64_t *best_ids = NULL;
  int best_n = 0, best_cap = 0;

  for (i = 1; i < hit_cnt; ++i)
    if (hits[i].strand != hits[i - 1].strand) {
      mix_strand = 1;
      break;
    }

  for (i = 0; i < hit_cnt; ++i) {
    uint64_t score;
    u64set local = {0, 0, 0};
    uint64_t max_entries;
    int q;
    if (hits[i].l < min_hit_len) continue;
    score = hit_score_len(hits[i].l, min_hit_len);
    k = (hits[i].strand + 1) / 2;
    max_entries = (uint64_t)(int64_t)(p->max_result * p->max_result_per_hit_factor); /* :620 */
    if (hits[i].ep - hits[i].sp + 1 <= max_entries || p->max_result_per_hit_factor <= 0 ||
        p->max_result <= 0) {
      for (j = hits[i].sp; j <= hits[i].ep; ++j) {
        uint64_t l;
        u64set_insert(&local, fm_backward_to_sampled_sa(o, j, &l));
      }
    } else { /* :635-666 */
      uint64_t range_size = hits[i].ep - hits[i].sp + 1;
      uint64_t step = DIV_CEIL(range_size, max_entries);
      uint64_t resolved = 0, l;
      for (j = hits[i].sp; j <= hits[i].ep; j += step) {
        u64set_insert(&local, fm_backward_to_sampled_sa(o, j, &l));
        ++resolved;
      }
      for (j = hits[i].ep; j >= hits[i].sp && j <= hits[i].ep; j -= step) {
        u64set_insert(&local, fm_backward_to_sampled_sa(o, j, &l));
        ++resolved;
        if (resolved >= max_entries) break;
      }
    }
    for (q = 0; q < local.n; ++q) { /* :669-707 */
      uint64_t seq_id = local.a[q];
      if (!mix_strand && i > 0 && hits[i].ep == hits[i].sp && hits[i - 1].ep == hits[i - 1].sp &&
          hits[i - 1].strand == hits[i].strand &&
          hits[i - 1].offset + hits[i - 1].l + 1 == hits[i].offset &&
          seq_id == prev_uniq.seq_id) {
        seq_rec *r = recmap_get(&rec[k], seq_id);
        r->score -= prev_uniq.score;
        prev_uniq.hit_length += hits[i].l;
        prev_uniq.score = hit_score_len(prev_uniq.hit_length, min_hit_len);
        r->score += prev_uniq.score;
        r->hit_length += hits[i].l;
      } else {
        int pos;
        seq_rec *r = recmap_find(&rec[k], seq_id, &pos);
        if (!r) {
          r = recmap_get(&rec[k], seq_id);
          r->score = score;
          r->hit_length = hits[i].l;
        } else {
          r->score += score;
          r->hit_length += hits[i].l;
        }
        if (hits[i].ep == hits[i].sp) {
          prev_uniq.seq_id = seq_id;
          prev_uniq.score = score;
          prev_uniq.hit_length = hits[i].l;
        }
      }
    }
    free(local.a);
  }

  for (k = 0; k <= 1; ++k) /* :711-736 */
    for (i = 0; i < rec[k].n; ++i) {
      const seq_rec *r = &rec[k].a[i];
      if (r->score > best) {
        second = best;
        second_len = best_len;
        best = r->score;
        best_len = (uint64_t)(int64_t)r->hit_length;
      } else if (r->score > second) {
        second = r->score;
        second_len = (uint64_t)(int64_t)r->hit_length;
      }
    }
  res->score = best;
  res->secondary_score = second;
  res->hit_length = (int32_t)best_len;

#define BEST_PUSH(x) do { if (best_n == best_cap) { best_cap = best_cap ? best_cap * 2 : 16; \
    best_ids = (uint64_t *)realloc(best_ids, 8 * (size_t)best_cap); } best_ids[best_n++] = (x); } while (0)
  for (k = 0; k <= 1; ++k) /* :743-757 */
    for (i = 0; i < rec[k].n; ++i)
      if (rec[k].a[i].score == best && u64set_insert(&used, rec[k].a[i].seq_id))
        BEST_PUSH(rec[k].a[i].seq_id);
  if (best_n > 1) res->secondary_score = best;

  if (second_len >= p->consider_secondary_hit_len && second < best &&
      second >= (uint64_t)(p->consider_secondary_score_factor * (double)best)) { /* :763-781 */
    for (k = 0; k <= 1; ++k)
      for (i = 0; i < rec[k].n; ++i)
        if (rec[k].a[i].score == second && u64set_insert(&used, rec[k].a[i].seq_id))
          BEST_PUSH(rec[k].a[i].seq_id);
    res->secondary_score = second;
  }
#undef BEST_PUSH

  res->n = 0;
  res->by_rank = 0;
  if (best_n <= p->max_result || p->max_result <= 0) { /* :784-797 */
    for (i = 0; i < best_n; ++i) {
      if (i < 64) {
        res->ids[i] = best_ids[i];
        res->tax_ids[i] = tax_orig_id(&o->tax, tax_seq_to_tax(&o->tax, best_ids[i]));
      }
    }
    res->n = best_n;
  } else { /* :798-841 */
    uint64_t *tids = (uint64_t *)xmalloc(8 * (size_t)best_n);
    uint64_t *red = (uint64_t *)xmalloc(8 * (size_t)(best_n + 1));
    int nred;
    for (i = 0; i < best_n; ++i) tids[i] = tax_seq_to_tax(&o->tax, best_ids[i]);
    if (ex) { /* :807-838 */
      u64set *lists = NULL;
      int n_lists = 0, q;
      nred = tax_reduce(&o->tax, tids, best_n, p->max_result, red, best_n + 1, &lists, &n_lists);
      for (i = 0; i < n_lists; ++i) {
        if (n_lists == nred && i < 64) { /* :823 */
          ex->cnt[i] = lists[i].n;
          for (q = 0; q < lists[i].n; ++q) {
            if (ex->total < ex->child_cap) ex->child[ex->total] = lists[i].a[q];
            ++ex->total;
          }
        }
        free(lists[i].a);
      }
      free(lists);
    } else {
      nred = tax_reduce(&o->tax, tids, best_n, p->max_result, red, best_n + 1, NULL, NULL);
    }
    for (i = 0; i < nred && i < 64; ++i) {
      res->ids[i] = red[i];
      res->tax_ids[i] = tax_orig_id(&o->tax, red[i]);
    }
    res->n = nred;
    res->by_rank = 1;
    free(tids);
    free(red);
  }
  free(best_ids);
  free(used.a);
  free(rec[0].a);
  free(rec[1].a);
}

/* Classifier.hpp:950-961 */
void cfr_oracle_query(cfr_oracle *o, const cfr_oracle_param *p, const char *r1, const char *r2,
                      cfr_oracle_result *res) {
  hitvec hits = {0, 0, 0};
  int mhl = eff_min_hit_len(o, p);
  memset(res, 0, sizeof(*res));
  search_forward_and_reverse(o, mhl, r1, r2, &hits);
  classification_from_hits(o, p, mhl, &hits, res, NULL);
  res->query_length = (int32_t)strlen(r1);
  if (r2) res->query_length += (int32_t)strlen(r2);
  free(hits.a);
}

/* Query with _classifierParam.outputExpandedResult (Classifier.hpp:22, :792-838) */
int cfr_oracle_query_expanded(cfr_oracle *o, const cfr_oracle_param *p, const char *r1, const char *r2,
                              cfr_oracle_result *res, uint64_t *child, int child_cap, int32_t *child_cnt) {
  hitvec hits = {0, 0, 0};
  expand_out ex;
  int mhl = eff_min_hit_len(o, p);
  memset(res, 0, sizeof(*res));
  memset(child_cnt, 0, 64 * sizeof(int32_t));
  ex.child = child;
  ex.child_cap = child_cap;
  ex.cnt = child_cnt;
  ex.total = 0;
  search_forward_and_reverse(o, mhl, r1, r2, &hits);
  classification_from_hits(o, p, mhl, &hits, res, &ex);
  res->query_length = (int32_t)strlen(r1);
  if (r2) res->query_length += (int32_t)strlen(r2);
  free(hits.a);
  return ex.total;
}

uint64_t cfr_oracle_seqid_to_taxid(const cfr_oracle *o, uint64_t s) { return tax_seq_to_tax(&o->tax, s); }
uint64_t cfr_oracle_orig_taxid(const cfr_oracle *o, uint64_t c) { return tax_orig_id(&o->tax, c); }
const char *cfr_oracle_seq_name(const cfr_oracle *o, uint64_t s) {
  if (s >= o->tax.seq_cnt + o->tax.extra_seq_cnt) return "";
  return o->tax.seq_name[s];
}
const char *cfr_oracle_rank_name(const cfr_oracle *o, uint64_t c) {
  return tax_rank_string(tax_rank_of(&o->tax, c));
}
int cfr_oracle_reduce_taxids(const cfr_oracle *o, const uint64_t *t, int n, int k, uint64_t *out,
                             int cap) {
  return tax_reduce(&o->tax, t, n, k, out, cap, NULL, NULL);
}
int cfr_oracle_reduce_taxids_expanded(const cfr_oracle *o, const uint64_t *t, int n, int k, uint64_t *out,
                                      int cap, uint64_t *child, int child_cap, int32_t *child_cnt,
                                      int *n_lists) {
  u64set *lists = NULL;
  int i, q, total = 0;
  int nred = tax_reduce(&o->tax, t, n, k, out, cap, &lists, n_lists);
  for (i = 0; i < *n_lists; ++i) {
    child_cnt[i] = lists[i].n;
    for (q = 0; q < lists[i].n; ++q) {
      if (total < child_cap) child[total] = lists[i].a[q];
      ++total;
    }
    free(lists[i].a);
  }
  free(lists);
  return nred;
}

/* ResultWriter.hpp:199-236 */
int cfr_oracle_format_tsv(const cfr_oracle *o, const char *read_id, const cfr_oracle_result *r,
                          char *buf, size_t cap) {
  size_t off = 0;
  int i, w;
  if (r->n > 0) {
    for (i = 0; i < r->n && i < 64; ++i) {
      const char *name = r->by_rank ? cfr_oracle_rank_name(o, r->ids[i]) : cfr_oracle_seq_name(o, r->ids[i]);
      w = snprintf(buf + off, cap - off, "%s\t%s\t%lu\t%lu\t%lu\t%d\t%d\t%d\n", read_id, name,
                   (unsigned long)r->tax_ids[i], (unsigned long)r->score,
                   (unsigned long)r->secondary_score, r->hit_length, r->query_length, r->n);
      if (w < 0 || (size_t)w >= cap - off) return -1;
      off += (size_t)w;
    }
  } else {
    w = snprintf(buf + off, cap - off, "%s\tunclassified\t0\t0\t0\t0\t%d\t1\n", read_id, r->query_length);
    if (w < 0 || (size_t)w >= cap - off) return -1;
    off += (size_t)w;
  }
  return (int)off;
}

/* ResultWriter.hpp:186-241 with _outputExpandedTaxIds: one more column, the original ids of the
 * child list joined by ',' (Classifier.hpp:826-835); empty for unclassified reads */
int cfr_oracle_format_tsv_expanded(const cfr_oracle *o, const char *read_id, const cfr_oracle_result *r,
                                   const uint64_t *child, const int32_t *child_cnt, char *buf, size_t cap) {
  size_t off = 0;
  int i, q, w, at = 0;
  if (r->n > 0) {
    for (i = 0; i < r->n && i < 64; ++i) {
      const char *name = r->by_rank ? cfr_oracle_rank_name(o, r->ids[i]) : cfr_oracle_seq_name(o, r->ids[i]);
      w = snprintf(buf + off, cap - off, "%s\t%s\t%lu\t%lu\t%lu\t%d\t%d\t%d\t", read_id, name,
                   (unsigned long)r->tax_ids[i], (unsigned long)r->score,
                   (unsigned long)r->secondary_score, r->hit_length, r->query_length, r->n);
      if (w < 0 || (size_t)w >= cap - off) return -1;
      off += (size_t)w;
      for (q = 0; q < child_cnt[i]; ++q, ++at) {
        w = snprintf(buf + off, cap - off, "%s%lu", q ? "," : "", (unsigned long)tax_orig_id(&o->tax, child[at]));
        if (w < 0 || (size_t)w >= cap - off) return -1;
        off += (size_t)w;
      }
      if (cap - off < 2) return -1;
      buf[off++] = '\n';
      buf[off] = 0;
    }
  } else {
    w = snprintf(buf + off, cap - off, "%s\tunclassified\t0\t0\t0\t0\t%d\t1\t\n", read_id, r->query_length);
    if (w < 0 || (size_t)w >= cap - off) return -1;
    off += (size_t)w;
  }
  return (int)off;
}

void cfr_oracle_get_counters(const cfr_oracle *o, cfr_oracle_counters *c) { *c = o->cnt; }
void cfr_oracle_reset_counters(cfr_oracle *o) { memset(&o->cnt, 0, sizeof(o->cnt)); }

/* ------------------------------------------------------------------------ */
/* Dustmasker                                                                */
/* ------------------------------------------------------------------------ */

#define DUST_W 64 /* Dustmasker.hpp:247 */
#define DUST_T 20 /* :248 */
#define DUST_L 1  /* :249 */
#define DUST_ABITS 3 /* "ACGT" + 1 catch-all -> 5 symbols -> 3 bits (:282-309) */
#define DUST_NCODE 4

typedef struct { uint64_t start, end; int score; } dust_iv;
typedef struct { dust_iv *a; int n, cap; } dust_ivvec;

static void ivv_insert(dust_ivvec *v, int pos, dust_iv x) {
  if (v->n == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 16;
    v->a = (dust_iv *)realloc(v->a, sizeof(dust_iv) * (size_t)v->cap);
    if (!v->a) die("out of memory");
  }
  memmove(v->a + pos + 1, v->a + pos, sizeof(dust_iv) * (size_t)(v->n - pos));
  v->a[pos] = x;
  ++v->n;
}

static int dust_code(char c) { /* :291-302 */
  switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return DUST_NCODE;
  }
}

typedef struct { /* Dustmasker_Queue, :33-90, capacity 128 for w=64 */
  int head, tail, mask;
  int s[128];
} dust_queue;

static int dq_size(const dust_queue *q) { return (q->tail - q->head) & q->mask; }
static void dq_push(dust_queue *q, int t) { q->s[q->tail] = t; q->tail = (q->tail + 1) & q->mask; }
static int dq_pop(dust_queue *q) { int t = q->s[q->head]; q->head = (q->head + 1) & q->mask; return t; }
static int dq_at(const dust_queue *q, int i) { return q->s[(q->head + i) & q->mask]; }

static void dust_add(int t, int *count, int *r) { *r += count[t]; ++count[t]; }     /* :93-97 */
static void dust_remove(int t, int *count, int *r) { --count[t]; *r -= count[t]; }   /* :99-103 */

/* :106-136 */
static void dust_shift_window(int t, dust_queue *window, int *lv, int *rw, int *rv, int *cw, int *cv) {
  if (dq_size(window) >= DUST_W - 2) {
    int old = window->s[window->head];
    dust_remove(old, cw, rw);
    dq_pop(window);
    if (*lv > dq_size(window)) {
      dust_remove(old, cv, rv);
      --*lv;
    }
  }
  dq_push(window, t);
  ++*lv;
  dust_add(t, cw, rw);
  dust_add(t, cv, rv);
  if (cv[t] * 10 > 2 * DUST_T) {
    for (;;) {
      int s = dq_at(window, dq_size(window) - *lv);
      dust_remove(s, cv, rv);
      --*lv;
      if (s == t) break;
    }
  }
}

/* :139-168 */
static void dust_save_masked(dust_ivvec *result, dust_ivvec *P, uint64_t window_start) {
  if (P->n > 0 && P->a[P->n - 1].start < window_start) {
    dust_iv last = P->a[P->n - 1];
    int l = result->n;
    if (l > 0) {
      if (last.start <= result->a[l - 1].end + 1) {
        if (last.end > result->a[l - 1].end) result->a[l - 1].end = last.end;
      } else
        ivv_insert(result, result->n, last);
    } else
      ivv_insert(result, result->n, last);
    while (P->n > 0 && P->a[P->n - 1].start < window_start) --P->n;
  }
}

/* :173-242 */
static void dust_find_perfect(dust_ivvec *P, dust_queue *window, uint64_t window_start, int lv,
                              int rv, int *cv) {
  int i;
  int max_score = 0;
  int max_score_triplets = 1;
  int it = 0; /* index standing in for the std::vector iterator, re-set per i (:187) */
  for (i = dq_size(window) - lv - 1; i >= 0; --i) {
    int t = dq_at(window, i);
    dust_add(t, cv, &rv);
    it = 0; /* std::vector::iterator it = P.begin() (:187) */
    if (rv * 10 > DUST_T * (dq_size(window) - i - 1)) {
      while (it != P->n && P->a[it].start >= (uint64_t)i + window_start) {
        if ((uint64_t)(int64_t)P->a[it].score * (uint64_t)(int64_t)max_score_triplets >
            (uint64_t)(int64_t)max_score * (P->a[it].end - P->a[it].start - 2)) {
          max_score = P->a[it].score;
          max_score_triplets = (int)(P->a[it].end - P->a[it].start - 2);
        }
        ++it;
      }
      if (rv * max_score_triplets >= max_score * (dq_size(window) - i - 1)) {
        dust_iv np;
        max_score = rv;
        max_score_triplets = dq_size(window) - i - 1;
        np.start = (uint64_t)i + window_start;
        np.end = window_start + (uint64_t)dq_size(window) + 1;
        np.score = rv;
        ivv_insert(P, it, np);
      }
    }
  }
  for (i = dq_size(window) - lv - 1; i >= 0; --i) dust_remove(dq_at(window, i), cv, &rv);
}

/* :312-354 */
static void dust_sdust(const char *S, uint64_t n, dust_ivvec *result) {
  uint64_t wstart, wfinish;
  int triplet;
  const int triplet_mask = (1 << (3 * DUST_ABITS)) - 1;
  int count_v[512], count_w[512];
  int rv = 0, rw = 0, lv = 0;
  dust_queue window;
  dust_ivvec P = {0, 0, 0};
  if (n < 3) return;
  memset(count_v, 0, sizeof(count_v));
  memset(count_w, 0, sizeof(count_w));
  window.head = window.tail = 0;
  window.mask = 127; /* (1<<capacityBits)-1 with capacityBits = 7 for sz = 64 (:47-55) */
  triplet = (dust_code(S[0]) << DUST_ABITS) + dust_code(S[1]);
  for (wfinish = 2; wfinish < n; ++wfinish) {
    wstart = 0;
    if (wfinish + 1 > (uint64_t)DUST_W) wstart = wfinish + 1 - DUST_W;
    dust_save_masked(result, &P, wstart);
    triplet = ((triplet << DUST_ABITS) & triplet_mask) + dust_code(S[wfinish]);
    dust_shift_window(triplet, &window, &lv, &rw, &rv, count_w, count_v);
    if (rw * 10 > lv * DUST_T) dust_find_perfect(&P, &window, wstart, lv, rv, count_v);
  }
  wstart = 0;
  if (wfinish + 1 > (uint64_t)DUST_W) wstart = wfinish + 1 - DUST_W;
  while (P.n > 0) {
    dust_save_masked(result, &P, wstart);
    ++wstart;
  }
  free(P.a);
}

/* Dustmasker.hpp:357-421, then the in-place masking of CentrifugerClass.cpp:281-289 */
int cfr_oracle_dust_mask(char *seq, size_t n) {
  uint64_t i, j;
  dust_ivvec result = {0, 0, 0}, wres = {0, 0, 0};
  int q, nres;
  if (n < 3) return 0;
  for (i = 0; i < n && dust_code(seq[i]) == DUST_NCODE; ++i)
    ;
  for (; i < n;) {
    uint64_t n_count = 0;
    uint64_t last_valid = i;
    for (j = i; j < n; ++j) {
      if (dust_code(seq[j]) == DUST_NCODE)
        ++n_count;
      else {
        if (n_count > (uint64_t)DUST_W) break;
        last_valid = j;
        n_count = 0;
      }
    }
    if (last_valid > i) {
      wres.n = 0;
      dust_sdust(seq + i, last_valid - i + 1, &wres);
      for (q = 0; q < wres.n; ++q) {
        dust_iv x = wres.a[q];
        x.start += i;
        x.end += i;
        ivv_insert(&result, result.n, x);
      }
    }
    i = j;
  }
  /* _l == 1: the linker merge (:404-420) is disabled */
  nres = result.n;
  for (q = 0; q < nres; ++q) {
    uint64_t k2;
    for (k2 = result.a[q].start; k2 <= result.a[q].end; ++k2) seq[k2] = 'N';
  }
  free(result.a);
  free(wres.a);
  return nres;
}
