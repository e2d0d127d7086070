/* TEST INFRASTRUCTURE — CPU restatement of the reference decoder (the parity oracle).
 *
 * Restates, in plain C99 float arithmetic, the reference's lattice-generating
 * token-passing Viterbi beam search:
 *   OnlineLatticeDecoderBase<Fst,StdToken>   src/my-decoder/online-decoder-base-inl.h
 *   HashList<StateId,Token*>                 src/util/hash-list-inl.h
 *   LatticeToVector                          src/newfst/lattice-functions.cc:179-217
 * Each function cites the reference lines it follows.  Parity status: PINNED — in
 * ORC_MODE_REFERENCE this file is checked bit-for-bit (one-best words, alignment,
 * cost bits, per-frame token counts and cutoffs) against the compiled reference
 * (oracle/_ref/ref_decode) by tests/test_oracle_vs_ref.py and against the committed
 * fixtures under tests/golden/ that the compiled reference generated.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path never does.
 */
#include "wfst_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_INF (INFINITY)
#define NO_ARC 0xFFFFFFFFu
#define NO_BUCKET ((size_t)-1)

/* ------------------------------------------------------------------ graph */

struct OrcGraph {
  OrcArc *arcs;
  int64_t *row_off;
  uint32_t *n_ieps;
  int32_t n_states;
  int64_t n_arcs;
  int32_t start, final_state;
  /* CLG graphs (orc_graph_set_clg): the reference's CLG decoder compares differently and adds the two
   * weights of an arc that leaves a CLG state through an HMM one after the other in its best-token
   * pre-pass (src/my-decoder/online-clg-decoder-mempool-base.h:66-112, :128, :150) */
  int clg;
  float *w_clg, *w_hmm;
  unsigned char *from_clg;
};

OrcGraph *orc_graph_create(const OrcArc *arcs, const int64_t *row_off, const uint32_t *n_ieps,
                           int32_t n_states, int64_t n_arcs, int32_t start, int32_t final_state) {
  OrcGraph *g = (OrcGraph *)calloc(1, sizeof(OrcGraph));
  g->arcs = (OrcArc *)malloc(sizeof(OrcArc) * (size_t)(n_arcs ? n_arcs : 1));
  memcpy(g->arcs, arcs, sizeof(OrcArc) * (size_t)n_arcs);
  g->row_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_states + 1));
  memcpy(g->row_off, row_off, sizeof(int64_t) * (size_t)(n_states + 1));
  g->n_ieps = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n_states);
  memcpy(g->n_ieps, n_ieps, sizeof(uint32_t) * (size_t)n_states);
  g->n_states = n_states;
  g->n_arcs = n_arcs;
  g->start = start;
  g->final_state = final_state;
  return g;
}

void orc_graph_set_clg(OrcGraph *g, const float *w_clg, const float *w_hmm, const unsigned char *from_clg) {
  g->clg = 1;
  g->w_clg = (float *)malloc(sizeof(float) * (size_t)(g->n_arcs ? g->n_arcs : 1));
  g->w_hmm = (float *)malloc(sizeof(float) * (size_t)(g->n_arcs ? g->n_arcs : 1));
  g->from_clg = (unsigned char *)malloc((size_t)(g->n_arcs ? g->n_arcs : 1));
  memcpy(g->w_clg, w_clg, sizeof(float) * (size_t)g->n_arcs);
  memcpy(g->w_hmm, w_hmm, sizeof(float) * (size_t)g->n_arcs);
  memcpy(g->from_clg, from_clg, (size_t)g->n_arcs);
}

void orc_graph_destroy(OrcGraph *g) {
  if (!g) return;
  free(g->w_clg);
  free(g->w_hmm);
  free(g->from_clg);
  free(g->arcs);
  free(g->row_off);
  free(g->n_ieps);
  free(g);
}

/* ------------------------------------------------------------------ LM FSAs (biglm) */

typedef struct OrcLm {
  int32_t bos, eos;
  int32_t n_states;
  int32_t *arc_num;
  float *backoff_prob;
  int32_t *backoff_id;
  int64_t *arc_off;
  OrcLmArc *arcs;
} OrcLm;

OrcLm *orc_lm_create(int32_t bos, int32_t eos, int32_t n_states, const int32_t *arc_num,
                     const float *backoff_prob, const int32_t *backoff_id, const OrcLmArc *arcs,
                     int64_t n_arcs) {
  OrcLm *lm = (OrcLm *)calloc(1, sizeof(OrcLm));
  lm->bos = bos;
  lm->eos = eos;
  lm->n_states = n_states;
  lm->arc_num = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_states);
  lm->backoff_prob = (float *)malloc(sizeof(float) * (size_t)n_states);
  lm->backoff_id = (int32_t *)malloc(sizeof(int32_t) * (size_t)n_states);
  lm->arc_off = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n_states + 1));
  lm->arcs = (OrcLmArc *)malloc(sizeof(OrcLmArc) * (size_t)(n_arcs ? n_arcs : 1));
  memcpy(lm->arc_num, arc_num, sizeof(int32_t) * (size_t)n_states);
  memcpy(lm->backoff_prob, backoff_prob, sizeof(float) * (size_t)n_states);
  memcpy(lm->backoff_id, backoff_id, sizeof(int32_t) * (size_t)n_states);
  memcpy(lm->arcs, arcs, sizeof(OrcLmArc) * (size_t)n_arcs);
  lm->arc_off[0] = 0;
  for (int32_t i = 0; i < n_states; ++i) lm->arc_off[i + 1] = lm->arc_off[i] + arc_num[i];
  return lm;
}

void orc_lm_destroy(OrcLm *lm) {
  if (!lm) return;
  free(lm->arc_num);
  free(lm->backoff_prob);
  free(lm->backoff_id);
  free(lm->arc_off);
  free(lm->arcs);
  free(lm);
}

/* Fsa::GetArc, src/newlm/arpa2fsa.cc:244-262 (start state 0 is direct-indexed by word id,
 * FsaState::SearchStartArc arpa2fsa.h:211-214; other states binary search, arpa2fsa.h:194-210) */
static int fsa_get_arc(const OrcLm *lm, int32_t id, int32_t word, float *weight, int32_t *to) {
  if (word == 0) {
    *weight = lm->backoff_prob[id];
    *to = lm->backoff_id[id];
    return 1;
  }
  const OrcLmArc *arc = lm->arcs + lm->arc_off[id];
  const OrcLmArc *hit = NULL;
  if (id == 0) {
    hit = &arc[word];
  } else {
    int start = 0, end = lm->arc_num[id] - 1, mid = (start + end) / 2;
    while (start <= end) {
      if (arc[mid].wordid > word) end = mid - 1;
      else if (arc[mid].wordid < word) start = mid + 1;
      else { hit = &arc[mid]; break; }
      mid = (start + end) / 2;
    }
  }
  if (!hit) return 0;
  *weight = hit->weight;
  *to = hit->tostateid;
  return 1;
}

/* ComposeArpaLm::Start, src/newlm/compose-arpalm.cc:5-13 */
static int32_t clm_start(const OrcLm *lm) {
  float w = 0.0f;
  int32_t to = 0;
  fsa_get_arc(lm, 0, lm->bos, &w, &to);
  return to;
}

/* ComposeArpaLm::GetArc, compose-arpalm.cc:52-70: walk the back-off chain; Value1 = -(sum) */
static void clm_get_arc(const OrcLm *lm, int32_t s, int32_t word, int32_t *next, float *value1) {
  float weight = 0.0f, w_arc = 0.0f;
  int32_t to = 0;
  while (!fsa_get_arc(lm, s, word, &w_arc, &to)) {
    fsa_get_arc(lm, s, 0, &w_arc, &to);
    s = to;
    weight += w_arc;
  }
  weight += w_arc;
  *value1 = -1 * weight;
  *next = to;
}

/* ComposeArpaLm::Final, compose-arpalm.cc:15-29 */
static float clm_final(const OrcLm *lm, int32_t s) {
  float weight = 0.0f, w_arc = 0.0f;
  int32_t to = 0;
  while (!fsa_get_arc(lm, s, lm->eos, &w_arc, &to)) {
    fsa_get_arc(lm, s, 0, &w_arc, &to);
    s = to;
    weight += w_arc;
  }
  weight += w_arc;
  return (-1.0 * weight);
}

/* -------------------------------------------------------- tokens and links */

typedef struct Lnk {
  struct Tok *next_tok;
  int32_t ilabel, olabel;
  float graph, ac;
  struct Lnk *next;
} Lnk; /* ForwardLink, src/my-decoder/online-decoder-base.h:28-48 */

typedef struct Tok {
  float tot, extra;
  Lnk *links;
  struct Tok *next;
  struct Tok *back;
  int32_t state; /* convenience only (the reference keeps it in the hash Elem) */
  int32_t frame;
  uint32_t arc;  /* global index of the arc that set `tot` (canonical tie-break) */
  int32_t lm_state; /* biglm: DiffArpaLm pair-state id */
  int has_final;    /* member of the reference's _final_costs map */
  float final_cost;
} Tok; /* StdToken, src/my-decoder/online-decoder-base.h:54-84 */

typedef struct Elem {
  uint64_t key; /* StateId, or for biglm PairId = fst_state + (lm_state << 32) (…-biglm.h:77-80) */
  Tok *val;
  struct Elem *tail;
} Elem; /* HashList::Elem, src/util/hash-list.h:16-21 */

typedef struct {
  size_t prev_bucket;
  Elem *last_elem;
} Bucket; /* src/util/hash-list.h:83-89 */

typedef struct {
  Tok *toks;
  int must_prune_links, must_prune_toks;
} TokList; /* src/my-decoder/online-decoder-base.h:252-260 */

struct OrcDecoder {
  const OrcGraph *g;
  OrcConfig cfg;
  int mode;
  /* HashList state (src/util/hash-list.h:91-97) */
  Elem *list_head;
  size_t bucket_list_tail;
  size_t hash_size;
  Bucket *buckets;
  size_t n_buckets;
  Elem *freed_elems;
  /* pools (MemPool<T>, src/util/mem-pool.h:16-65: free-list allocators) */
  Tok *freed_toks;
  Lnk *freed_lnks;
  void **blocks;
  size_t n_blocks, cap_blocks;
  /* decoder state */
  TokList *active;
  size_t n_active, cap_active;
  Elem **queue;
  size_t n_queue, cap_queue;
  float *tmp;
  size_t n_tmp, cap_tmp;
  int64_t num_toks, num_links;
  int32_t num_frames_decoded;
  int warned, finalized;
  /* final costs: there is a single super-final state, so the reference's
     unordered_map<Token*,BaseFloat> (inl.h:691-697) has at most one entry */
  int final_costs_nonempty;
  /* biglm (…-biglm.h): LM-difference composition */
  const struct OrcLm *lm1, *lm2;
  int32_t (*pairs)[2];   /* DiffArpaLm::_state_vec */
  int32_t n_pairs, cap_pairs;
  int32_t *pair_map;     /* open-addressing map pair -> id (DiffArpaLm::_state_map) */
  int32_t pair_map_size;
  float final_relative_cost, final_best_cost;
  /* stats */
  OrcFrameStat *stats;
  size_t n_stats, cap_stats;
  OrcFrameStat pending;
};

static void *block_alloc(OrcDecoder *d, size_t bytes) {
  void *p = malloc(bytes);
  if (d->n_blocks == d->cap_blocks) {
    d->cap_blocks = d->cap_blocks ? d->cap_blocks * 2 : 64;
    d->blocks = (void **)realloc(d->blocks, d->cap_blocks * sizeof(void *));
  }
  d->blocks[d->n_blocks++] = p;
  return p;
}

static Tok *new_token(OrcDecoder *d, float tot, float extra, Lnk *links, Tok *next, Tok *back) {
  if (!d->freed_toks) {
    size_t n = 4096;
    Tok *b = (Tok *)block_alloc(d, n * sizeof(Tok));
    for (size_t i = 0; i + 1 < n; ++i) b[i].next = &b[i + 1];
    b[n - 1].next = NULL;
    d->freed_toks = b;
  }
  Tok *t = d->freed_toks;
  d->freed_toks = t->next;
  t->tot = tot;
  t->extra = extra;
  t->links = links;
  t->next = next;
  t->back = back;
  t->state = -1;
  t->frame = -1;
  t->arc = NO_ARC;
  t->lm_state = 0;
  t->has_final = 0;
  t->final_cost = 0.0f;
  d->num_toks++;
  return t;
}

static void delete_token(OrcDecoder *d, Tok *t) {
  t->next = d->freed_toks;
  d->freed_toks = t;
  d->num_toks--;
}

static Lnk *new_link(OrcDecoder *d, Tok *next_tok, int32_t il, int32_t ol, float graph, float ac,
                     Lnk *next) {
  if (!d->freed_lnks) {
    size_t n = 4096;
    Lnk *b = (Lnk *)block_alloc(d, n * sizeof(Lnk));
    for (size_t i = 0; i + 1 < n; ++i) b[i].next = &b[i + 1];
    b[n - 1].next = NULL;
    d->freed_lnks = b;
  }
  Lnk *l = d->freed_lnks;
  d->freed_lnks = l->next;
  l->next_tok = next_tok;
  l->ilabel = il;
  l->olabel = ol;
  l->graph = graph;
  l->ac = ac;
  l->next = next;
  d->num_links++;
  return l;
}

static void delete_link(OrcDecoder *d, Lnk *l) {
  l->next = d->freed_lnks;
  d->freed_lnks = l;
  d->num_links--;
}

/* DeleteForwardLinks, inl.h:8-19 */
static void delete_forward_links(OrcDecoder *d, Tok *t) {
  Lnk *l = t->links;
  while (l) {
    Lnk *m = l->next;
    delete_link(d, l);
    l = m;
  }
  t->links = NULL;
}

/* ----------------------------------------------------------------- HashList */

/* HashList::SetSize, src/util/hash-list-inl.h:14-22 */
static void hl_set_size(OrcDecoder *d, size_t size) {
  d->hash_size = size;
  assert(d->list_head == NULL && d->bucket_list_tail == NO_BUCKET);
  if (size > d->n_buckets) {
    d->buckets = (Bucket *)realloc(d->buckets, size * sizeof(Bucket));
    for (size_t i = d->n_buckets; i < size; ++i) {
      d->buckets[i].prev_bucket = 0;
      d->buckets[i].last_elem = NULL;
    }
    d->n_buckets = size;
  }
}

/* HashList::Clear, hash-list-inl.h:24-38 */
static Elem *hl_clear(OrcDecoder *d) {
  for (size_t b = d->bucket_list_tail; b != NO_BUCKET; b = d->buckets[b].prev_bucket)
    d->buckets[b].last_elem = NULL;
  d->bucket_list_tail = NO_BUCKET;
  Elem *ans = d->list_head;
  d->list_head = NULL;
  return ans;
}

/* HashList::Delete, hash-list-inl.h:46-51 */
static void hl_delete(OrcDecoder *d, Elem *e) {
  e->tail = d->freed_elems;
  d->freed_elems = e;
}

/* HashList::DeleteElems, hash-list-inl.h:53-61 */
static void hl_delete_elems(OrcDecoder *d) {
  Elem *e = hl_clear(d), *t;
  for (; e; e = t) {
    t = e->tail;
    hl_delete(d, e);
  }
}

/* HashList::New, hash-list-inl.h:84-103 */
static Elem *hl_new(OrcDecoder *d) {
  if (!d->freed_elems) {
    size_t n = 1024;
    Elem *b = (Elem *)block_alloc(d, n * sizeof(Elem));
    for (size_t i = 0; i + 1 < n; ++i) b[i].tail = &b[i + 1];
    b[n - 1].tail = NULL;
    d->freed_elems = b;
  }
  Elem *e = d->freed_elems;
  d->freed_elems = e->tail;
  return e;
}

/* HashList::Insert, hash-list-inl.h:127-173: returns the existing element when the key
 * is present; otherwise appends to the key's bucket, buckets being chained in creation
 * order — this defines the reference's token visiting order. */
static Elem *hl_insert(OrcDecoder *d, uint64_t key, Tok *val) {
  size_t index = (size_t)key % d->hash_size;
  Bucket *bk = &d->buckets[index];
  if (bk->last_elem) {
    Elem *head = bk->prev_bucket == NO_BUCKET ? d->list_head
                                              : d->buckets[bk->prev_bucket].last_elem->tail;
    Elem *tail = bk->last_elem->tail;
    for (Elem *e = head; e != tail; e = e->tail)
      if (e->key == key) return e;
  }
  Elem *el = hl_new(d);
  el->key = key;
  el->val = val;
  if (!bk->last_elem) {
    if (d->bucket_list_tail == NO_BUCKET) {
      assert(d->list_head == NULL);
      d->list_head = el;
    } else {
      d->buckets[d->bucket_list_tail].last_elem->tail = el;
    }
    el->tail = NULL;
    bk->last_elem = el;
    bk->prev_bucket = d->bucket_list_tail;
    d->bucket_list_tail = index;
  } else {
    el->tail = bk->last_elem->tail;
    bk->last_elem->tail = el;
    bk->last_elem = el;
  }
  return el;
}

/* ------------------------------------------------------------------ DiffArpaLm */

#define KEY_STATE(k) ((int32_t)(uint32_t)(k))
#define KEY_LM(k) ((int32_t)(uint32_t)((k) >> 32))
#define MAKE_KEY(st, lm) ((uint64_t)(uint32_t)(st) + ((uint64_t)(uint32_t)(lm) << 32)) /* …-biglm.h:77-80 */

/* DiffArpaLm::Reset, src/newlm/diff-lm.h:39-46 */
static void difflm_reset(OrcDecoder *d) {
  if (!d->pair_map) {
    d->pair_map_size = 1 << 16;
    d->pair_map = (int32_t *)malloc(sizeof(int32_t) * (size_t)d->pair_map_size);
    d->cap_pairs = 1 << 15;
    d->pairs = (int32_t(*)[2])malloc(sizeof(int32_t[2]) * (size_t)d->cap_pairs);
  }
  for (int32_t i = 0; i < d->pair_map_size; ++i) d->pair_map[i] = -1;
  d->n_pairs = 0;
}

static int32_t difflm_intern(OrcDecoder *d, int32_t a, int32_t b) {
  uint32_t h = ((uint32_t)a * 7853u + (uint32_t)b) & (uint32_t)(d->pair_map_size - 1);
  for (;;) {
    int32_t id = d->pair_map[h];
    if (id < 0) break;
    if (d->pairs[id][0] == a && d->pairs[id][1] == b) return id;
    h = (h + 1) & (uint32_t)(d->pair_map_size - 1);
  }
  if (d->n_pairs + 1 >= d->cap_pairs) {
    fprintf(stderr, "[oracle] too many LM pair states\n");
    abort();
  }
  d->pairs[d->n_pairs][0] = a;
  d->pairs[d->n_pairs][1] = b;
  d->pair_map[h] = d->n_pairs;
  return d->n_pairs++;
}

/* NextLmState (…-biglm.h:54-70) over DiffArpaLm::GetArc (diff-lm.h:63-111).  Reference mode
 * reproduces the reference's argument quirk: both LMs are queried from FSA state number `s`
 * (the pair-state id), not from the members of the pair (diff-lm.h:75,80,86; SURVEY.md
 * Appendix B-6).  Canonical mode implements the intended semantics. */
static int32_t next_lm_state(OrcDecoder *d, int32_t lm_state, int32_t olabel, float *lm_score) {
  if (olabel == 0) {
    *lm_score = 0;
    return lm_state;
  }
  int32_t s1 = lm_state, s2 = lm_state;
  if (d->mode == ORC_MODE_CANONICAL) {
    s1 = d->pairs[lm_state][0];
    s2 = d->pairs[lm_state][1];
  }
  int32_t n1, n2;
  float v1, v2;
  clm_get_arc(d->lm1, s1, olabel, &n1, &v1);
  clm_get_arc(d->lm2, s2, olabel, &n2, &v2);
  *lm_score = v1 + v2; /* Times(w1, w2).Value1(), weigth.h:318-323 */
  return difflm_intern(d, n1, n2);
}

/* DiffArpaLm::Final, diff-lm.h:48-55 (this one does use the pair) */
static float difflm_final(OrcDecoder *d, int32_t lm_state) {
  return clm_final(d->lm1, d->pairs[lm_state][0]) + clm_final(d->lm2, d->pairs[lm_state][1]);
}

/* ------------------------------------------------------------------ decoder */

OrcDecoder *orc_decoder_create(const OrcGraph *g, const OrcConfig *cfg, int mode) {
  OrcDecoder *d = (OrcDecoder *)calloc(1, sizeof(OrcDecoder));
  d->g = g;
  d->cfg = *cfg;
  d->mode = mode;
  d->bucket_list_tail = NO_BUCKET;
  /* constructor, inl.h:22-30: _toks.SetSize(max_active * hash_ratio) */
  size_t sz = (size_t)((float)cfg->max_active * cfg->hash_ratio);
  if (sz > ((size_t)1 << 28)) sz = (size_t)1 << 28; /* guard: INT_MAX default would ask for 64 GB */
  if (sz < 1) sz = 1;
  hl_set_size(d, sz);
  return d;
}

static void clear_active_tokens(OrcDecoder *d);

/* OnlineLatticeDecoderMempoolBaseBiglm(fst, config, oldlm, newlm), …-biglm.h:21-30; the caller has
 * already scaled the old LM by -1 (kaldi-hclg-my-decoder-biglm.cc:55-60) */
OrcDecoder *orc_decoder_create_biglm(const OrcGraph *g, const OrcConfig *cfg, int mode, const OrcLm *lm1,
                                     const OrcLm *lm2) {
  OrcDecoder *d = orc_decoder_create(g, cfg, mode);
  d->lm1 = lm1;
  d->lm2 = lm2;
  return d;
}

void orc_decoder_destroy(OrcDecoder *d) {
  if (!d) return;
  for (size_t i = 0; i < d->n_blocks; ++i) free(d->blocks[i]);
  free(d->blocks);
  free(d->buckets);
  free(d->active);
  free(d->queue);
  free(d->tmp);
  free(d->stats);
  free(d->pairs);
  free(d->pair_map);
  free(d);
}

static void active_resize(OrcDecoder *d, size_t n) {
  if (n > d->cap_active) {
    size_t c = d->cap_active ? d->cap_active : 64;
    while (c < n) c *= 2;
    d->active = (TokList *)realloc(d->active, c * sizeof(TokList));
    d->cap_active = c;
  }
  for (size_t i = d->n_active; i < n; ++i) {
    d->active[i].toks = NULL;
    d->active[i].must_prune_links = 1;
    d->active[i].must_prune_toks = 1;
  }
  d->n_active = n;
}

/* ClearActiveTokens, inl.h:69-85 */
static void clear_active_tokens(OrcDecoder *d) {
  for (size_t i = 0; i < d->n_active; ++i) {
    for (Tok *t = d->active[i].toks; t;) {
      delete_forward_links(d, t);
      Tok *n = t->next;
      delete_token(d, t);
      t = n;
    }
  }
  d->n_active = 0;
  assert(d->num_toks == 0 && d->num_links == 0);
}

static void push_stat(OrcDecoder *d, const OrcFrameStat *s) {
  if (d->n_stats == d->cap_stats) {
    d->cap_stats = d->cap_stats ? d->cap_stats * 2 : 512;
    d->stats = (OrcFrameStat *)realloc(d->stats, d->cap_stats * sizeof(OrcFrameStat));
  }
  d->stats[d->n_stats++] = *s;
}

static inline int has_eps(const OrcGraph *g, int32_t s) { return g->n_ieps[s] != 0; }

/* canonical relaxation order: lower cost wins; equal cost -> lower global arc index */
static inline int better(const OrcDecoder *d, float tot, uint32_t arc, const Tok *old) {
  if (old->tot > tot) return 1;
  if (d->mode == ORC_MODE_CANONICAL && old->tot == tot && arc < old->arc) return 2;
  return 0;
}

/* FindOrAddToken, inl.h:88-136 */
static Elem *find_or_add_token(OrcDecoder *d, uint64_t key, int32_t frame_plus_one, float tot,
                               Tok *back, uint32_t arc, int *changed) {
  assert((size_t)frame_plus_one < d->n_active);
  Tok **toks = &d->active[frame_plus_one].toks;
  Elem *e = hl_insert(d, key, NULL);
  if (e->val == NULL) {
    Tok *t = new_token(d, tot, 0.0f, NULL, *toks, back);
    t->state = KEY_STATE(key);
    t->lm_state = KEY_LM(key);
    t->frame = frame_plus_one;
    t->arc = arc;
    *toks = t;
    e->val = t;
    if (changed) *changed = 1;
  } else {
    Tok *t = e->val;
    int b = better(d, tot, arc, t);
    if (b) {
      t->tot = tot;
      t->back = back;
      t->arc = arc;
    }
    if (changed) *changed = (b == 1); /* a tie-break swap leaves the cost unchanged */
  }
  return e;
}

/* exact k-th smallest (0-based) of a[0..n): what std::nth_element leaves at a[k] */
static float kth_smallest(float *a, size_t n, size_t k_) {
  long lo = 0, hi = (long)n - 1, k = (long)k_;
  while (lo < hi) { /* Hoare quickselect */
    float pivot = a[lo + (hi - lo) / 2];
    long i = lo, j = hi;
    while (i <= j) {
      while (a[i] < pivot) ++i;
      while (a[j] > pivot) --j;
      if (i <= j) {
        float t = a[i];
        a[i] = a[j];
        a[j] = t;
        ++i;
        --j;
      }
    }
    if (k <= j) hi = j;
    else if (k >= i) lo = i;
    else break; /* j < k < i: a[k] == pivot is in place */
  }
  return a[k];
}

/* GetCutoff, inl.h:138-234 */
static float get_cutoff(OrcDecoder *d, Elem *list_head, size_t *tok_count, float *adaptive_beam,
                        Elem **best_elem) {
  const OrcConfig *c = &d->cfg;
  float best_weight = ORC_INF;
  size_t count = 0;
  d->n_tmp = 0;
  for (Elem *e = list_head; e; e = e->tail, ++count) {
    float w = e->val->tot;
    if (d->n_tmp == d->cap_tmp) {
      d->cap_tmp = d->cap_tmp ? d->cap_tmp * 2 : 8192;
      d->tmp = (float *)realloc(d->tmp, d->cap_tmp * sizeof(float));
    }
    d->tmp[d->n_tmp++] = w;
    if (w < best_weight) { /* first strictly smaller in list order, inl.h:173-178 */
      best_weight = w;
      *best_elem = e;
    } else if (d->mode == ORC_MODE_CANONICAL && w == best_weight && *best_elem &&
               e->key < (*best_elem)->key) {
      *best_elem = e; /* canonical: lowest state id among equal best costs */
    }
  }
  *tok_count = count;
  float beam_cutoff = best_weight + c->beam;
  float min_active_cutoff = ORC_INF, max_active_cutoff = ORC_INF;
  if (d->n_tmp > (size_t)c->max_active) /* inl.h:188-195 */
    max_active_cutoff = kth_smallest(d->tmp, d->n_tmp, (size_t)c->max_active);
  if (max_active_cutoff < beam_cutoff) { /* inl.h:197-203 */
    *adaptive_beam = max_active_cutoff - best_weight + c->beam_delta;
    return max_active_cutoff;
  }
  if (d->n_tmp > (size_t)c->min_active) { /* inl.h:205-218 */
    if (c->min_active == 0) min_active_cutoff = best_weight;
    else min_active_cutoff = kth_smallest(d->tmp, d->n_tmp, (size_t)c->min_active);
  }
  if (min_active_cutoff > beam_cutoff) { /* inl.h:220-226 */
    *adaptive_beam = min_active_cutoff - best_weight + c->beam_delta;
    return min_active_cutoff;
  }
  *adaptive_beam = c->beam; /* inl.h:227-232 */
  return beam_cutoff;
}

/* PossiblyResizeHash, inl.h:236-244 */
static void possibly_resize_hash(OrcDecoder *d, size_t num_toks) {
  size_t new_sz = (size_t)((float)num_toks * d->cfg.hash_ratio);
  if (new_sz > d->hash_size) hl_set_size(d, new_sz);
}

/* ProcessEmitting, inl.h:246-351; biglm variant …-biglm.h:318-400 (graph cost += LM-difference
 * score of the arc's word, token key = (fst state, lm state)).  `ll` is the log-likelihood row
 * of this frame, column = ilabel - 1. */
static float process_emitting(OrcDecoder *d, const float *ll) {
  const OrcGraph *g = d->g;
  const int biglm = d->lm1 != NULL;
  int frame = (int)d->n_active - 1;
  active_resize(d, d->n_active + 1);
  Elem *final_toks = hl_clear(d);
  Elem *best_elem = NULL;
  float adaptive_beam = 0;
  size_t tok_cnt = 0;
  float cur_cutoff = get_cutoff(d, final_toks, &tok_cnt, &adaptive_beam, &best_elem);
  /* The biglm class shadows _toks (…-biglm.h:73) but calls the BASE PossiblyResizeHash
   * (…-biglm.h:335), which resizes the base class's unused hash: the biglm hash keeps its
   * constructor size (…-biglm.h:27). */
  if (!biglm) possibly_resize_hash(d, tok_cnt);
  memset(&d->pending, 0, sizeof(d->pending));
  d->pending.n_in = (uint32_t)tok_cnt;
  d->pending.cur_cutoff = cur_cutoff;
  d->pending.abeam = adaptive_beam;

  float next_cutoff = ORC_INF;
  /* best-token pre-pass, inl.h:282-300: (cost + w) - loglike;
   * biglm …-biglm.h:339-357: ((lm_score + cost) + w) - loglike */
  if (best_elem) {
    Tok *tok = best_elem->val;
    const int32_t bs = KEY_STATE(best_elem->key), blm = KEY_LM(best_elem->key);
    for (int64_t a = g->row_off[bs]; a < g->row_off[bs + 1]; ++a) {
      const OrcArc *arc = &g->arcs[a];
      if (arc->ilabel != 0) {
        d->pending.ll_calls++;
        float tot_score;
        if (biglm) {
          float lm_score;
          next_lm_state(d, blm, arc->olabel, &lm_score);
          tot_score = lm_score + tok->tot + arc->weight - ll[arc->ilabel - 1];
        } else if (g->clg && g->from_clg[a]) {
          /* online-clg-decoder-mempool-base.h:91: tok + clgarc.w + arc.w - loglike */
          tot_score = tok->tot + g->w_clg[a] + g->w_hmm[a] - ll[arc->ilabel - 1];
        } else {
          tot_score = tok->tot + arc->weight - ll[arc->ilabel - 1];
        }
        if (tot_score + adaptive_beam < next_cutoff) next_cutoff = tot_score + adaptive_beam;
      }
    }
  }
  if (d->mode == ORC_MODE_CANONICAL) {
    /* The final next_cutoff of the reference loop is order independent: an arc is skipped
     * only when tot >= next_cutoff, in which case tot + abeam could not lower it either
     * (inl.h:330-333).  Canonical mode computes it first and admits against it. */
    for (Elem *e = final_toks; e; e = e->tail) {
      Tok *tok = e->val;
      if (g->clg ? tok->tot < cur_cutoff : tok->tot <= cur_cutoff) { /* clg: …-clg-…-base.h:128 */
        const int32_t st = KEY_STATE(e->key), lms = KEY_LM(e->key);
        for (int64_t a = g->row_off[st]; a < g->row_off[st + 1]; ++a) {
          const OrcArc *arc = &g->arcs[a];
          if (arc->ilabel != 0) {
            float ac = -ll[arc->ilabel - 1];
            float graph_cost = arc->weight;
            if (biglm) {
              float lm_score;
              next_lm_state(d, lms, arc->olabel, &lm_score);
              graph_cost = arc->weight + lm_score;
            }
            float tot = tok->tot + ac + graph_cost;
            if (tot + adaptive_beam < next_cutoff) next_cutoff = tot + adaptive_beam;
          }
        }
      }
    }
  }
  /* main pass, inl.h:311-347 / …-biglm.h:363-396 */
  for (Elem *e = final_toks, *e_tail; e; e = e_tail) {
    const int32_t state = KEY_STATE(e->key), lm_state = KEY_LM(e->key);
    Tok *tok = e->val;
    if (g->clg ? tok->tot < cur_cutoff : tok->tot <= cur_cutoff) {
      for (int64_t a = g->row_off[state]; a < g->row_off[state + 1]; ++a) {
        const OrcArc *arc = &g->arcs[a];
        if (arc->ilabel != 0) {
          d->pending.ll_calls++;
          d->pending.arcs_expanded++;
          int32_t next_lm = 0;
          float graph_cost = arc->weight;
          if (biglm) {
            float lm_score;
            next_lm = next_lm_state(d, lm_state, arc->olabel, &lm_score);
            graph_cost = arc->weight + lm_score; /* …-biglm.h:379 */
          }
          float ac_cost = -ll[arc->ilabel - 1];
          float cur_cost = tok->tot;
          float tot_cost = cur_cost + ac_cost + graph_cost;
          if (g->clg ? tot_cost > next_cutoff : tot_cost >= next_cutoff) continue; /* clg: …-clg-…-base.h:156 */
          else if (tot_cost + adaptive_beam < next_cutoff)
            next_cutoff = tot_cost + adaptive_beam; /* never fires in canonical mode */
          Elem *nt = find_or_add_token(d, MAKE_KEY(arc->nextstate, next_lm), frame + 1, tot_cost, tok,
                                       (uint32_t)a, NULL);
          tok->links = new_link(d, nt->val, arc->ilabel, arc->olabel, graph_cost, ac_cost, tok->links);
          d->pending.arcs_admitted++;
        }
      }
    }
    e_tail = e->tail;
    hl_delete(d, e);
  }
  d->num_frames_decoded++;
  return next_cutoff;
}

/* ProcessNonemitting, inl.h:353-431 / …-biglm.h:402-466 */
static void process_nonemitting(OrcDecoder *d, float cutoff) {
  const OrcGraph *g = d->g;
  const int biglm = d->lm1 != NULL;
  int frame = (int)d->n_active - 1;
  assert(d->n_queue == 0);
  if (d->list_head == NULL && !d->warned) d->warned = 1; /* "no surviving tokens" */
  for (Elem *e = d->list_head; e; e = e->tail) {
    if (has_eps(g, KEY_STATE(e->key))) {
      if (d->n_queue == d->cap_queue) {
        d->cap_queue = d->cap_queue ? d->cap_queue * 2 : 4096;
        d->queue = (Elem **)realloc(d->queue, d->cap_queue * sizeof(Elem *));
      }
      d->queue[d->n_queue++] = e;
    }
  }
  while (d->n_queue) {
    Elem *elem = d->queue[--d->n_queue];
    const int32_t state = KEY_STATE(elem->key), lm_state = KEY_LM(elem->key);
    Tok *tok = elem->val;
    float cur_cost = tok->tot;
    if (cur_cost >= cutoff) continue; /* inl.h:391 */
    delete_forward_links(d, tok);     /* inl.h:399 */
    for (int64_t a = g->row_off[state]; a < g->row_off[state + 1]; ++a) {
      const OrcArc *arc = &g->arcs[a];
      if (arc->ilabel == 0) {
        d->pending.eps_arcs++;
        int32_t next_lm = 0;
        float graph_cost = arc->weight;
        if (biglm) {
          float lm_score;
          next_lm = next_lm_state(d, lm_state, arc->olabel, &lm_score);
          graph_cost = arc->weight + lm_score; /* …-biglm.h:446-448 */
        }
        float tot_cost = cur_cost + graph_cost;
        if (tot_cost < cutoff) { /* inl.h:415 */
          int changed = 0;
          Elem *nt = find_or_add_token(d, MAKE_KEY(arc->nextstate, next_lm), frame, tot_cost, tok,
                                       (uint32_t)a, &changed);
          tok->links = new_link(d, nt->val, 0, arc->olabel, graph_cost, 0.0f, tok->links);
          if (changed && has_eps(g, arc->nextstate)) {
            if (d->n_queue == d->cap_queue) {
              d->cap_queue = d->cap_queue ? d->cap_queue * 2 : 4096;
              d->queue = (Elem **)realloc(d->queue, d->cap_queue * sizeof(Elem *));
            }
            d->queue[d->n_queue++] = nt;
          }
        }
      }
    }
  }
  /* statistics (what oracle/ref_harness.cc records from the compiled reference) */
  OrcFrameStat s = d->pending;
  memset(&d->pending, 0, sizeof(d->pending));
  s.next_cutoff = cutoff;
  s.best = ORC_INF;
  for (Elem *e = d->list_head; e; e = e->tail) {
    s.n_raw++;
    if (e->val->tot < cutoff) s.n_within++;
    if (e->val->tot < s.best) s.best = e->val->tot;
  }
  push_stat(d, &s);
}

/* InitDecoding, inl.h:40-67 */
void orc_init_decoding(OrcDecoder *d) {
  clear_active_tokens(d);
  hl_delete_elems(d);
  d->n_queue = 0;
  d->n_tmp = 0;
  d->warned = 0;
  d->finalized = 0;
  d->final_costs_nonempty = 0;
  d->n_stats = 0;
  memset(&d->pending, 0, sizeof(d->pending));
  active_resize(d, 1);
  Tok *start_tok = new_token(d, 0.0f, 0.0f, NULL, NULL, NULL);
  start_tok->state = d->g->start;
  start_tok->frame = 0;
  d->active[0].toks = start_tok;
  int32_t start_lm = 0;
  if (d->lm1) { /* …-biglm.h:98-120: _diff_lm.Reset(); start pair = (graph start, difflm start) */
    difflm_reset(d);
    start_lm = difflm_intern(d, clm_start(d->lm1), clm_start(d->lm2));
    start_tok->lm_state = start_lm;
  }
  hl_insert(d, MAKE_KEY(d->g->start, start_lm), start_tok);
  process_nonemitting(d, d->cfg.beam);
  d->num_frames_decoded = 0;
}

/* PruneForwardLinks, inl.h:482-572 */
static void prune_forward_links(OrcDecoder *d, int frame_plus_one, int *extra_costs_changed,
                                int *links_pruned, float delta) {
  *extra_costs_changed = 0;
  *links_pruned = 0;
  if (d->active[frame_plus_one].toks == NULL && !d->warned) d->warned = 1;
  int changed = 1;
  while (changed) {
    changed = 0;
    for (Tok *tok = d->active[frame_plus_one].toks; tok; tok = tok->next) {
      Lnk *link, *prev_link = NULL;
      float tok_extra_cost = ORC_INF;
      for (link = tok->links; link;) {
        Tok *next_tok = link->next_tok;
        float link_extra_cost =
            next_tok->extra + ((tok->tot + link->ac + link->graph) - next_tok->tot);
        if (link_extra_cost > d->cfg.lattice_beam) { /* excise, inl.h:532-542 */
          Lnk *next_link = link->next;
          if (prev_link) prev_link->next = next_link;
          else tok->links = next_link;
          delete_link(d, link);
          link = next_link;
          *links_pruned = 1;
        } else {
          if (link_extra_cost < 0.0f) link_extra_cost = 0.0f;
          if (link_extra_cost < tok_extra_cost) tok_extra_cost = link_extra_cost;
          prev_link = link;
          link = link->next;
        }
      }
      if (fabsf(tok_extra_cost - tok->extra) > delta) changed = 1; /* inl.h:560 */
      tok->extra = tok_extra_cost;
    }
    if (changed) *extra_costs_changed = 1;
  }
}

/* PruneTokensForFrame, inl.h:578-607 */
static void prune_tokens_for_frame(OrcDecoder *d, int frame_plus_one) {
  Tok **toks = &d->active[frame_plus_one].toks;
  Tok *tok, *next_tok, *prev_tok = NULL;
  for (tok = *toks; tok; tok = next_tok) {
    next_tok = tok->next;
    if (tok->extra == ORC_INF) {
      if (prev_tok) prev_tok->next = tok->next;
      else *toks = tok->next;
      assert(tok->links == NULL);
      delete_token(d, tok);
    } else {
      prev_tok = tok;
    }
  }
}

/* PruneActiveTokens, inl.h:438-480 */
static void prune_active_tokens(OrcDecoder *d, float delta) {
  int cur_frame_plus_one = (int)d->n_active - 1;
  for (int f = cur_frame_plus_one - 1; f >= 0; f--) {
    if (d->active[f].must_prune_links) {
      int links_pruned = 0, extra_costs_changed = 0;
      prune_forward_links(d, f, &extra_costs_changed, &links_pruned, delta);
      if (extra_costs_changed && f > 0) d->active[f - 1].must_prune_links = 1;
      if (links_pruned) d->active[f].must_prune_toks = 1;
      d->active[f].must_prune_links = 0;
    }
    if (f + 1 < cur_frame_plus_one && d->active[f + 1].must_prune_toks) {
      prune_tokens_for_frame(d, f + 1);
      d->active[f + 1].must_prune_toks = 0;
    }
  }
}

/* AdvanceDecoding, inl.h:630-668 */
void orc_advance_decoding(OrcDecoder *d, const float *loglikes, int32_t stride,
                          int32_t frames_ready, int32_t max_num_frames) {
  assert(d->num_frames_decoded >= 0 && !d->finalized);
  assert(frames_ready >= d->num_frames_decoded);
  int target = frames_ready;
  if (max_num_frames >= 0 && d->num_frames_decoded + max_num_frames < target)
    target = d->num_frames_decoded + max_num_frames;
  while (d->num_frames_decoded < target) {
    int nfd = (int)d->n_active - 1; /* NumFramesDecoded(), online-decoder-base.h:133 */
    /* Canonical mode prunes once, at FinalizeDecoding, with exact extra costs: the periodic
     * prune works on lower bounds of the final extra costs (every path to the end passes the
     * current frame), so it only ever removes a subset of what the final prune removes. */
    if (d->mode == ORC_MODE_REFERENCE && nfd % d->cfg.prune_interval == 0)
      prune_active_tokens(d, d->cfg.lattice_beam * d->cfg.prune_scale);
    float cutoff = process_emitting(d, loglikes + (size_t)d->num_frames_decoded * stride);
    process_nonemitting(d, cutoff);
  }
}

int32_t orc_num_frames_decoded(const OrcDecoder *d) { return (int32_t)d->n_active - 1; }

/* ComputeFinalCosts, inl.h:670-720 (iterates the current-frame hash list); biglm variant
 * …-biglm.h:157-215: every final-state token gets final cost = DiffArpaLm::Final(lm state), and
 * best_cost_with_final is taken over ALL tokens.  Marks tokens instead
 * of filling the reference's unordered_map<Token*, BaseFloat>. */
static void compute_final_costs(OrcDecoder *d, int *nonempty, float *rel, float *best_out) {
  float best_cost = ORC_INF, best_with_final = ORC_INF;
  *nonempty = 0;
  for (Elem *e = d->list_head; e; e = e->tail) {
    Tok *tok = e->val;
    tok->has_final = 0;
    tok->final_cost = 0.0f;
    const int fst_final = KEY_STATE(e->key) == d->g->final_state; /* Fst::IsFinal, optimize-fst.h:189-192 */
    if (tok->tot < best_cost) best_cost = tok->tot;
    if (d->lm1) {
      float lm_final = difflm_final(d, KEY_LM(e->key));
      float cost_with_final = tok->tot + lm_final;
      if (cost_with_final < best_with_final) best_with_final = cost_with_final;
      if (fst_final) {
        tok->has_final = 1;
        tok->final_cost = lm_final;
        *nonempty = 1;
      }
    } else if (fst_final) {
      tok->has_final = 1; /* final cost 0: the weights live on the eps arcs into the super-final state */
      *nonempty = 1;
      if (tok->tot < best_with_final) best_with_final = tok->tot;
    }
  }
  if (rel) {
    if (best_cost == ORC_INF && best_with_final == ORC_INF) *rel = ORC_INF;
    else *rel = best_with_final - best_cost;
  }
  if (best_out) *best_out = best_with_final != ORC_INF ? best_with_final : best_cost;
}

/* PruneForwardLinksFinal, inl.h:725-824 (biglm: …-biglm.h:468-566, same arithmetic) */
static void prune_forward_links_final(OrcDecoder *d) {
  int frame_plus_one = (int)d->n_active - 1;
  compute_final_costs(d, &d->final_costs_nonempty, &d->final_relative_cost, &d->final_best_cost);
  d->finalized = 1;
  hl_delete_elems(d);
  int changed = 1;
  const float delta = d->mode == ORC_MODE_CANONICAL ? 0.0f : 1.0e-5f; /* canonical: exact fixed point */
  while (changed) {
    changed = 0;
    for (Tok *tok = d->active[frame_plus_one].toks; tok; tok = tok->next) {
      Lnk *link, *prev_link = NULL;
      float final_cost;
      if (!d->final_costs_nonempty) final_cost = 0.0f;
      else final_cost = tok->has_final ? tok->final_cost : ORC_INF;
      float tok_extra_cost = tok->tot + final_cost - d->final_best_cost;
      for (link = tok->links; link;) {
        Tok *next_tok = link->next_tok;
        float link_extra_cost =
            next_tok->extra + ((tok->tot + link->ac + link->graph) - next_tok->tot);
        if (link_extra_cost > d->cfg.lattice_beam) {
          Lnk *next_link = link->next;
          if (prev_link) prev_link->next = next_link;
          else tok->links = next_link;
          delete_link(d, link);
          link = next_link;
        } else {
          if (link_extra_cost < 0.0f) link_extra_cost = 0.0f;
          if (link_extra_cost < tok_extra_cost) tok_extra_cost = link_extra_cost;
          prev_link = link;
          link = link->next;
        }
      }
      if (tok_extra_cost > d->cfg.lattice_beam) tok_extra_cost = ORC_INF;
      if (fabsf(tok->extra - tok_extra_cost) > delta) changed = 1;
      tok->extra = tok_extra_cost;
    }
  }
}

/* FinalizeDecoding, inl.h:829-847 */
void orc_finalize_decoding(OrcDecoder *d) {
  int final_frame_plus_one = (int)d->n_active - 1;
  prune_forward_links_final(d);
  for (int f = final_frame_plus_one - 1; f >= 0; --f) {
    int b1, b2;
    prune_forward_links(d, f, &b1, &b2, 0.0f);
    prune_tokens_for_frame(d, f + 1);
  }
  prune_tokens_for_frame(d, 0);
}

/* GetBestPath = BestPathEnd (inl.h:1096-1158) + TraceBackBestPath (inl.h:1160-1200).
 * Output arcs are in path order start -> end, which is the order LatticeToVector walks the
 * linear lattice GetBestPath builds (inl.h:1080-1091). */
int32_t orc_get_best_path(OrcDecoder *d, int use_final_probs, int32_t *ilabel, int32_t *olabel,
                          float *graph, float *acoustic, int32_t cap) {
  if ((int)d->n_active - 1 <= 0) return -1;
  int nonempty = 0;
  if (d->finalized) {
    nonempty = d->final_costs_nonempty;
  } else if (use_final_probs) {
    compute_final_costs(d, &nonempty, NULL, NULL);
  }
  float best_cost = ORC_INF;
  Tok *best_tok = NULL;
  for (Tok *tok = d->active[d->n_active - 1].toks; tok; tok = tok->next) {
    float cost = tok->tot;
    if (use_final_probs && nonempty) { /* inl.h:1126-1139 */
      if (tok->has_final) cost += tok->final_cost;
      else cost = ORC_INF;
    }
    if (cost < best_cost) {
      best_cost = cost;
      best_tok = tok;
    } else if (d->mode == ORC_MODE_CANONICAL && best_tok && cost == best_cost && cost != ORC_INF &&
               (tok->state < best_tok->state ||
                (tok->state == best_tok->state && tok->lm_state < best_tok->lm_state))) {
      best_tok = tok;
    }
  }
  if (!best_tok) return -1;
  /* trace back, collecting arcs end -> start (including the zero arc of the start token) */
  int32_t n = 0;
  for (Tok *tok = best_tok; tok; tok = tok->back) {
    int32_t il = 0, ol = 0;
    float gc = 0.0f, ac = 0.0f;
    if (tok->back) {
      Lnk *link;
      for (link = tok->back->links; link; link = link->next) {
        if (link->next_tok == tok) { /* the FIRST (most recently added) link wins, inl.h:1169-1186 */
          il = link->ilabel;
          ol = link->olabel;
          gc = link->graph;
          ac = link->ac;
          break;
        }
      }
      if (!link) fprintf(stderr, "[oracle] error tracing best path back\n");
    }
    if (n < cap) {
      ilabel[n] = il;
      olabel[n] = ol;
      graph[n] = gc;
      acoustic[n] = ac;
    }
    ++n;
  }
  if (n > cap) return -2;
  for (int32_t i = 0, j = n - 1; i < j; ++i, --j) { /* reverse to path order */
    int32_t t;
    float f;
    t = ilabel[i]; ilabel[i] = ilabel[j]; ilabel[j] = t;
    t = olabel[i]; olabel[i] = olabel[j]; olabel[j] = t;
    f = graph[i]; graph[i] = graph[j]; graph[j] = f;
    f = acoustic[i]; acoustic[i] = acoustic[j]; acoustic[j] = f;
  }
  return n;
}

/* LatticeToVector, src/newfst/lattice-functions.cc:179-217 */
void orc_path_to_vector(const int32_t *ilabel, const int32_t *olabel, const float *graph,
                        const float *acoustic, int32_t n, int32_t *words, int32_t *n_words,
                        int32_t *ali, int32_t *n_ali, float *tot, float *lm) {
  float best_tot = 0, best_lm = 0;
  int32_t nw = 0, na = 0;
  for (int32_t i = 0; i < n; ++i) {
    if (ilabel[i] != 0) ali[na++] = ilabel[i];
    if (olabel[i] != 0) words[nw++] = olabel[i];
    best_lm += graph[i];
    best_tot += graph[i] + acoustic[i];
  }
  *n_words = nw;
  *n_ali = na;
  *tot = best_tot;
  *lm = best_lm;
}

/* debug: keys of the current frame in hash-list order */
int32_t orc_current_order(const OrcDecoder *d, int32_t *keys, int32_t cap) {
  int32_t n = 0;
  for (Elem *e = d->list_head; e; e = e->tail) {
    if (n < cap) keys[n] = e->key;
    ++n;
  }
  return n;
}

int32_t orc_frame_stats(const OrcDecoder *d, OrcFrameStat *out, int32_t cap) {
  int32_t n = (int32_t)d->n_stats;
  for (int32_t i = 0; i < n && i < cap; ++i) out[i] = d->stats[i];
  return n;
}

void orc_counts(const OrcDecoder *d, int64_t *num_toks, int64_t *num_links) {
  *num_toks = d->num_toks;
  *num_links = d->num_links;
}

int64_t orc_dump_lattice(const OrcDecoder *d, OrcLatTok *toks, int64_t tok_cap, OrcLatLink *links,
                         int64_t link_cap) {
  int64_t nt = 0, nl = 0;
  for (size_t f = 0; f < d->n_active; ++f) {
    for (Tok *t = d->active[f].toks; t; t = t->next) {
      if (nt < tok_cap) {
        toks[nt].frame = (int32_t)f;
        toks[nt].state = t->state;
        toks[nt].tot = t->tot;
        toks[nt].extra = t->extra;
      }
      ++nt;
      for (Lnk *l = t->links; l; l = l->next) {
        if (nl < link_cap) {
          links[nl].src_frame = (int32_t)f;
          links[nl].src_state = t->state;
          links[nl].dst_frame = l->next_tok->frame;
          links[nl].dst_state = l->next_tok->state;
          links[nl].ilabel = l->ilabel;
          links[nl].olabel = l->olabel;
          links[nl].graph = l->graph;
          links[nl].acoustic = l->ac;
          links[nl].src_tot = t->tot;
          links[nl].dst_tot = l->next_tok->tot;
        }
        ++nl;
      }
    }
  }
  return (nt << 32) | (nl & 0xFFFFFFFFLL);
}
