/* TEST INFRASTRUCTURE — CPU restatement of the reference decoder (the parity oracle).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (asr_decoder_b200/csrc) never links,
 * imports or calls it.
 *
 * Two token-order modes:
 *   ORC_MODE_REFERENCE  visits tokens in the reference's HashList list order
 *                       (src/util/hash-list-inl.h:127-173) and tightens next_cutoff while
 *                       expanding (src/my-decoder/online-decoder-base-inl.h:330-333).  It is
 *                       pinned bit-for-bit against the compiled reference (oracle/_ref).
 *   ORC_MODE_CANONICAL  the order-independent semantics the CUDA path implements: an arc is
 *                       admitted iff its cost is below the FINAL next_cutoff of the frame,
 *                       cost ties are broken by the lowest global arc index, best-token ties
 *                       by the lowest state id.
 */
#ifndef ASRD_WFST_ORACLE_H_
#define ASRD_WFST_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int32_t ilabel, olabel;
  float weight;
  int32_t nextstate;
} OrcArc; /* reference StdArc, src/newfst/arc.h:23-26 */

typedef struct {
  float beam;
  int32_t max_active;
  int32_t min_active;
  float lattice_beam;
  int32_t prune_interval;
  float beam_delta;
  float hash_ratio;
  float prune_scale;
} OrcConfig; /* src/my-decoder/lattice-faster-decoder-conf.h:21-44 */

typedef struct {
  uint32_t n_in;        /* tokens seen by GetCutoff */
  float cur_cutoff;     /* GetCutoff result */
  float abeam;          /* adaptive beam */
  float next_cutoff;    /* cutoff handed to ProcessNonemitting */
  uint32_t n_raw;       /* tokens alive after the eps closure */
  uint32_t n_within;    /* ... with cost < next_cutoff */
  float best;           /* best cost after closure */
  int64_t ll_calls;     /* LogLikelihood calls incl. the best-token pre-pass */
  int64_t arcs_expanded;/* emitting arcs of tokens with cost <= cur_cutoff */
  int64_t arcs_admitted;/* ... that produced a forward link */
  int64_t eps_arcs;     /* eps arcs scanned by the closure */
} OrcFrameStat;

typedef struct {
  int32_t frame;  /* 0 = before the first frame */
  int32_t state;
  float tot;
  float extra;
} OrcLatTok;

typedef struct {
  int32_t src_frame, src_state;
  int32_t dst_frame, dst_state;
  int32_t ilabel, olabel;
  float graph, acoustic;
  float src_tot, dst_tot; /* costs of the two tokens: tell apart the biglm tokens that share a state */
} OrcLatLink;

enum { ORC_MODE_REFERENCE = 0, ORC_MODE_CANONICAL = 1 };

typedef struct {
  int32_t wordid;
  float weight;
  int32_t tostateid;
} OrcLmArc; /* reference FsaArc, src/newlm/arpa2fsa.h:23-30 */

typedef struct OrcLm OrcLm;
typedef struct OrcGraph OrcGraph;
typedef struct OrcDecoder OrcDecoder;

OrcGraph *orc_graph_create(const OrcArc *arcs, const int64_t *row_off, const uint32_t *n_ieps,
                           int32_t n_states, int64_t n_arcs, int32_t start, int32_t final_state);
/* Marks the graph as a materialised CLG graph (asr_decoder_b200.fstio.materialize_clg): the decoders
 * built on it follow the reference's CLG decoder (src/my-decoder/online-clg-decoder-mempool-base.h):
 * tokens are expanded when cost < cur_cutoff (strict, :128), an arc is skipped only when its cost is
 * ABOVE the cutoff (:156), and the best-token pre-pass adds the CLG-arc and HMM-arc weights of a
 * two-level arc one after the other (:91). */
void orc_graph_set_clg(OrcGraph *g, const float *w_clg, const float *w_hmm, const unsigned char *from_clg);
void orc_graph_destroy(OrcGraph *g);

OrcDecoder *orc_decoder_create(const OrcGraph *g, const OrcConfig *cfg, int mode);
/* LM FSA (state 0 = unigram state, direct-indexed by word id); weights as stored by the reference
 * (natural-log probabilities, the old LM already scaled by -1 by the caller) */
OrcLm *orc_lm_create(int32_t bos, int32_t eos, int32_t n_states, const int32_t *arc_num,
                     const float *backoff_prob, const int32_t *backoff_id, const OrcLmArc *arcs,
                     int64_t n_arcs);
void orc_lm_destroy(OrcLm *lm);
OrcDecoder *orc_decoder_create_biglm(const OrcGraph *g, const OrcConfig *cfg, int mode, const OrcLm *lm1,
                                     const OrcLm *lm2);
void orc_decoder_destroy(OrcDecoder *d);

void orc_init_decoding(OrcDecoder *d);
/* loglikes: row-major [frames_ready x stride], column = ilabel - 1 */
void orc_advance_decoding(OrcDecoder *d, const float *loglikes, int32_t stride,
                          int32_t frames_ready, int32_t max_num_frames);
void orc_finalize_decoding(OrcDecoder *d);
int32_t orc_num_frames_decoded(const OrcDecoder *d);

/* GetBestPath: arcs in path order (start -> end). Returns arc count, or -1 if empty. */
int32_t orc_get_best_path(OrcDecoder *d, int use_final_probs, int32_t *ilabel, int32_t *olabel,
                          float *graph, float *acoustic, int32_t cap);
/* LatticeToVector over a best path (src/newfst/lattice-functions.cc:179-217). */
void orc_path_to_vector(const int32_t *ilabel, const int32_t *olabel, const float *graph,
                        const float *acoustic, int32_t n, int32_t *words, int32_t *n_words,
                        int32_t *ali, int32_t *n_ali, float *tot, float *lm);

int32_t orc_frame_stats(const OrcDecoder *d, OrcFrameStat *out, int32_t cap);
void orc_counts(const OrcDecoder *d, int64_t *num_toks, int64_t *num_links);
/* Raw lattice dump (one state per live token, one arc per live link). */
int64_t orc_dump_lattice(const OrcDecoder *d, OrcLatTok *toks, int64_t tok_cap,
                         OrcLatLink *links, int64_t link_cap);

#ifdef __cplusplus
}
#endif
#endif
