/* asrd.h — C ABI of the B200-native WFST token-passing beam-search decoder.
 *
 * This is the drop-in boundary for the hot path of datemoon/ASR-decoder's `my-decoder`
 * (SURVEY.md §8b).  Every entry point names the reference interface it replaces
 * (paths relative to the reference's src/).  Plain pointers and sizes only — no torch,
 * no C++ types.  The library is CUDA-only: there is no CPU fallback; every call fails
 * with ASRD_ERR_CUDA when no sm_100 device is usable.
 *
 * Threading: calls on DIFFERENT decoder handles may run concurrently; one host thread
 * can step thousands of streams through the batched entry points (the reference uses
 * one decoder object per pthread, src/v2-asrbin/v2-asr-service.cc:95-104).
 *
 * Semantics: the reference's per-frame token set depends on its hash-list visiting
 * order.  This library implements the order-independent
 * ("canonical") semantics: an arc is admitted iff its cost is below the frame's FINAL
 * next_cutoff; equal-cost recombination prefers the lowest arc index.  Measured against
 * the compiled reference under seven token orders (tests/golden/census.json, DESIGN.md
 * section 5.1): bit-identical one-best on the reference's CPU-sized configuration wherever
 * the reference agrees with itself; on the 1 M-state configuration, where the reference's
 * own answer depends on its token order for half of the utterances, the canonical answer
 * equals one of the reference's answers on 92 % of them and is otherwise within 1 % of the
 * path cost (cheaper or, on 4 % of the utterances, dearer).
 */
#ifndef ASRD_H_
#define ASRD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ASRD_ABI_VERSION 1

/* status codes (reference: bool returns / LOG_ERR throw / LOG_ASSERT abort,
 * src/util/log-message.cc:122-144) */
enum {
  ASRD_OK = 0,
  ASRD_ERR_BAD_ARG = -1,
  ASRD_ERR_CUDA = -2,         /* CUDA runtime error or no usable device */
  ASRD_ERR_NOMEM = -3,
  ASRD_ERR_HASH_OVERFLOW = -4,  /* per-frame state->token map full (raise hash_capacity) */
  ASRD_ERR_ARENA_OVERFLOW = -5, /* token / link arena full (raise token_capacity) */
  ASRD_ERR_FRAMES_OVERFLOW = -6,/* more frames than max_frames */
  ASRD_ERR_NO_TOKENS = -7,      /* reference: GetBestPath returns false (inl.h:1078-1079) */
  ASRD_ERR_PATH_OVERFLOW = -8,  /* best path longer than the caller's buffer */
  ASRD_ERR_STATE = -9,          /* call order violated (e.g. Advance after Finalize, inl.h:634) */
  ASRD_ERR_IO = -10,
  ASRD_ERR_LM_PAIRS_OVERFLOW = -11 /* biglm: LM state-pair table full (raise lm_pair_capacity) */
};

/* newfst StdArc, src/newfst/arc.h:23-26 (16 bytes, kept verbatim in HBM) */
typedef struct {
  int32_t ilabel;
  int32_t olabel;
  float weight;
  int32_t nextstate;
} asrd_arc;

/* LatticeFasterDecoderConfig, src/my-decoder/lattice-faster-decoder-conf.h:21-44.
 * hash_ratio only shapes the reference's HashList and has no effect here. */
typedef struct {
  float beam;
  int32_t max_active;
  int32_t min_active;
  float lattice_beam;
  int32_t prune_interval;
  float beam_delta;
  float hash_ratio;
  float prune_scale;
} asrd_config;

/* Device-side sizing; zero means "choose from the config". No reference counterpart
 * (the reference grows heap structures on demand). */
typedef struct {
  int32_t hash_capacity;   /* state->token slots per frame, rounded up to a power of two */
  int64_t token_capacity;  /* token records kept per utterance (all frames) */
  int32_t max_frames;      /* frames per utterance */
  int32_t collect_stats;   /* keep per-frame statistics for asrd_frame_stats */
  int32_t lm_pair_capacity;/* biglm: distinct (old LM state, new LM state) pairs one utterance may reach
                            * (DiffArpaLm's state table, src/newlm/diff-lm.h:92-103); rounded up to a
                            * power of two, default 65536 */
  int32_t prune_tokens;    /* 1: every config.prune_interval frames the tokens that can no longer reach the
                            * lattice (extra cost > lattice_beam against the current frontier) are dropped and
                            * the token arena is compacted — PruneActiveTokens, inl.h:438-480, called at
                            * inl.h:660-661.  Results do not change (one-best and raw lattice are bit-identical
                            * with and without); token_capacity then bounds the LIVE tokens, not the
                            * utterance.  0 (default): the arena keeps every token of the utterance.
                            * Plain (non-biglm) decoders only. */
  int32_t reserved[2];
} asrd_device_options;

/* per-frame statistics; index 0 = after InitDecoding (what the reference logs under
 * VLOG_COM(2), src/my-decoder/online-decoder-base-inl.h:306-308) */
typedef struct {
  uint32_t n_in;         /* tokens seen by GetCutoff for the frame that produced this one */
  float cur_cutoff;      /* GetCutoff result (inl.h:267) */
  float abeam;           /* adaptive beam */
  float next_cutoff;     /* final next_cutoff = cutoff of the eps closure */
  uint32_t n_tokens;     /* tokens alive after the closure (all have cost < next_cutoff) */
  float best;            /* best cost after the closure */
  uint32_t arcs_expanded;/* emitting arcs fetched for tokens with cost <= cur_cutoff */
  uint32_t arcs_admitted;/* ... whose cost was below the running cutoff when scored */
} asrd_frame_stat;

/* Raw lattice (DecoderItf::GetRawLattice, inl.h:868-975): one state per surviving token, one
 * arc per surviving forward link, after the lattice-beam pruning of FinalizeDecoding
 * (inl.h:725-847).  Links refer to tokens by index into the token array. */
typedef struct {
  int32_t frame;     /* 0 = before the first frame */
  int32_t state;     /* HCLG state */
  float cost;        /* forward cost (StdToken::_tot_cost) */
  float extra;       /* extra_cost after pruning (StdToken::_extra_cost) */
  int32_t is_final;  /* lattice state is final (inl.h:935-951) */
} asrd_lat_token;

typedef struct {
  int32_t src, dst;  /* token indices */
  int32_t ilabel, olabel;
  float graph, acoustic;
} asrd_lat_link;

/* LM FSA arc, reference FsaArc (src/newlm/arpa2fsa.h:23-30) */
typedef struct {
  int32_t wordid;
  float weight;    /* natural-log probability (cost = -weight) */
  int32_t tostateid;
} asrd_lm_arc;

typedef struct asrd_graph asrd_graph;
typedef struct asrd_lm asrd_lm;
typedef struct asrd_decoder asrd_decoder;

const char *asrd_strerror(int status);
int asrd_abi_version(void);
/* number of usable sm_100 devices; 0 or a negative status when none */
int asrd_device_count(void);
/* Optional, once, BEFORE the host process makes its first CUDA call: asks the driver for 32
 * hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS, unless the application already set it) so the
 * library's copy / scatter / frame-loop streams do not share a queue.  The library never changes
 * process-wide state on its own: not when it is loaded, not in any other entry point. */
int asrd_configure_process(void);

/* ---- graph: replaces Fst (src/newfst/optimize-fst.h:53-307) -------------------------- */

/* From the in-memory form Fst::ReadFst builds: per-state {num_arcs, niepsilons}
 * (StateInfo, optimize-fst.h:220-225) and the flat arc array.  Rows whose input-epsilon
 * arcs are not a prefix are stably partitioned on upload (relative order kept). */
int asrd_graph_create(const asrd_arc *arcs, const uint32_t *num_arcs, const uint32_t *niepsilons,
                      int32_t total_states, int64_t total_arcs, int32_t start, int32_t final_state,
                      int device, asrd_graph **out);
/* Fst::ReadFst(const char*), optimize-fst.h:208-219: same file format */
int asrd_graph_read(const char *path, int device, asrd_graph **out);
/* ClgFst::Init(clgfst, hmmfst), src/my-decoder/clg-fst.h:17-74: the CLG graph (newfst file whose
 * non-eps ilabels are HMM ids) and the HMM set (int32 count + that many newfst graphs), which the
 * reference expands on the fly, are written out as one static device graph over the reference's own
 * two-level state ids (clg-fst.h:82-165).  Decoders created on such a graph follow the reference's
 * CLG decoder, OnlineClgLatticeDecoderMempool (src/my-decoder/online-clg-decoder-mempool-base.h:
 * strict token cutoff :128, an arc is skipped only when above the cutoff :156, two-weight best-token
 * pre-pass :91); plain (non-biglm) decoders only. */
int asrd_graph_read_clg(const char *clg_path, const char *hmm_path, int device, asrd_graph **out);

/* ConstFst<StdArc,int>::Read + Fst(const ConstFst&) (src/newfst/const-fst.h:189-221,
 * src/newfst/optimize-fst.h:82-134): load an OpenFst "const" FST (e.g. a Kaldi HCLG.fst converted
 * with fstconvert --fst_type=const); final weights become leading 0:0 arcs to an appended
 * super-final state, exactly like the reference's conversion. */
int asrd_graph_read_const(const char *path, int device, asrd_graph **out);
int asrd_graph_destroy(asrd_graph *g);
int asrd_graph_info(const asrd_graph *g, int32_t *total_states, int64_t *total_arcs,
                    int32_t *start, int32_t *final_state, int64_t *device_bytes);

/* ---- LM: replaces ArpaLm / Fsa for the biglm path (src/newlm/arpa2fsa.h:217-480) -------- */

/* ARPA text LM -> the LM FSA file ArpaLm::Read takes (and CudaLm::Read / lm.read_lm here): what the
 * reference's arpa2fsa tool does with one thread (src/newlm/arpa2fsa.cc:311-739, arpa2fsa-bin.cc),
 * byte-identical output.  `wordlist`: "word id" per line with <s>, </s> and <unk>.  n-grams must
 * come grouped by history (as SRILM writes them): the reference asserts it, this returns
 * ASRD_ERR_IO.  Host code only — no device is touched. */
int asrd_lm_convert_arpa(const char *arpa_path, const char *wordlist_path, const char *out_path);

/* The arrays ArpaLm::Read loads (arpa2fsa.h:399-439, arpa2fsa.cc:70-176): per state
 * {arc_num, backoff_prob, backoff_id}, arcs grouped by state and sorted by word; state 0 is the
 * unigram state and MUST hold one arc per word id (direct index, arpa2fsa.h:211-214).  As in the
 * reference the caller rescales the OLD LM by -1 first (kaldi-hclg-my-decoder-biglm.cc:55-60). */
int asrd_lm_create(int32_t bos, int32_t eos, int32_t n_states, const int32_t *arc_num,
                   const float *backoff_prob, const int32_t *backoff_id, const asrd_lm_arc *arcs,
                   int64_t n_arcs, int device, asrd_lm **out);
int asrd_lm_destroy(asrd_lm *lm);

/* ---- decoder: replaces OnlineLatticeDecoderMempool behind DecoderItf ----------------- */

/* OnlineLatticeDecoderBase(FST*, const LatticeFasterDecoderConfig&), online-decoder-base.h:95.
 * The graph is shared and not owned (inl.h:24, _delete_fst(false)). */
int asrd_decoder_create(asrd_graph *g, const asrd_config *cfg, const asrd_device_options *opts,
                        asrd_decoder **out);
/* OnlineLatticeDecoderMempoolBaseBiglm(fst, config, oldlm, newlm)
 * (my-decoder/online-decoder-mempool-base-biglm.h:21-30): on-the-fly composition with the
 * LM-difference of two LMs; tokens are keyed by (HCLG state, LM state pair).  DiffArpaLm is
 * implemented with its intended semantics (both LMs advance from the members of the state
 * pair); the reference passes the pair-state id itself (newlm/diff-lm.h:75-86), which only
 * coincides on unigram-only LMs — SURVEY.md Appendix B-6.  GetBestPath and GetRawLattice both
 * work; lattice links carry arc weight + LM-difference score as their graph cost. */
int asrd_decoder_create_biglm(asrd_graph *g, const asrd_config *cfg, const asrd_device_options *opts,
                              asrd_lm *old_lm, asrd_lm *new_lm, asrd_decoder **out);
int asrd_decoder_destroy(asrd_decoder *d);

/* DecoderItf::InitDecoding (decoder-itf.h:15; inl.h:41-67), batched over n handles.
 * `stream` is a cudaStream_t (NULL = default stream). */
int asrd_init_decoding(asrd_decoder *const *decs, int32_t n, void *stream);

/* DecoderItf::AdvanceDecoding (decoder-itf.h:16; inl.h:630-668), batched.
 * Fails with ASRD_ERR_FRAMES_OVERFLOW — before decoding anything — when a stream would pass
 * max_frames, and with ASRD_ERR_BAD_ARG when num_indices is smaller than the graph's largest
 * ilabel: frames are never dropped and rows never read out of bounds.
 * loglikes[i] points at the row of frame NumFramesDecoded(i) of stream i — what
 * AmInterface::LogLikelihood(frame, index) would return for index = column + 1
 * (src/itf/decodable-itf.h:55-62; caller inl.h:295,326).  n_frames[i] rows with
 * row pitch stride[i] (floats) are available; all of them are decoded, at most
 * max_num_frames when that is >= 0.  num_indices = columns per row.
 * on_device != 0: the pointers are device pointers valid on `stream`.
 * on_device == 0: host pointers; rows are copied inside the call (asynchronously when the
 * memory is pinned) — the caller must not touch them before asrd_synchronize. */
int asrd_advance_decoding(asrd_decoder *const *decs, int32_t n, const float *const *loglikes,
                          const int32_t *n_frames, const int32_t *stride, int32_t num_indices,
                          int32_t max_num_frames, int32_t on_device, void *stream);

/* DecoderItf::FinalizeDecoding (decoder-itf.h:17; inl.h:829-847).  One-best mode keeps no
 * forward links on the device, so this only freezes the streams and picks the final token
 * (ComputeFinalCosts, inl.h:670-720). */
int asrd_finalize_decoding(asrd_decoder *const *decs, int32_t n, void *stream);

/* DecoderItf::NumFramesDecoded (decoder-itf.h:18) */
int32_t asrd_num_frames_decoded(const asrd_decoder *d);

/* DecoderItf::GetRawLattice (decoder-itf.h:23; inl.h:868-975) for one stream, including the
 * final lattice-beam pruning (PruneForwardLinksFinal / PruneForwardLinks / PruneTokensForFrame,
 * inl.h:482-607,725-847).  The device keeps tokens only; the forward links are regenerated from
 * the tokens, the graph, the log-likelihood history and the per-frame cutoffs in one backward
 * sweep, pruned there, and only the survivors are copied out.  Tokens come frame by frame with
 * eps links pointing forward inside a frame's block is NOT guaranteed: sort topologically if
 * needed.  n_toks / n_links receive the produced counts; ASRD_ERR_PATH_OVERFLOW when a cap was
 * too small (call again with larger buffers).  Requires FinalizeDecoding when use_final_probs. */
int asrd_get_raw_lattice(asrd_decoder *d, int32_t use_final_probs, asrd_lat_token *toks, int64_t tok_cap,
                         asrd_lat_link *links, int64_t link_cap, int64_t *n_toks, int64_t *n_links,
                         void *stream);

/* GetRawLattice for n streams in one call (the reference calls it once per decoder object,
 * kaldi-nnet3bin/kaldi-hclg-my-decoder.cc:134-139; on the device one stream's sweep occupies one
 * SM, so a service that finishes many utterances together hands them over together).  Stream i
 * owns toks[i*tok_cap ..) and links[i*link_cap ..); n_toks[i], n_links[i] and status[i] are what
 * asrd_get_raw_lattice would return for it (ASRD_ERR_NO_TOKENS = the reference's `false`,
 * ASRD_ERR_PATH_OVERFLOW = call again with larger windows).  Results are identical to n single
 * calls.  The return value covers the call itself (arguments, call order, CUDA). */
int asrd_get_raw_lattice_batch(asrd_decoder *const *decs, int32_t n, int32_t use_final_probs,
                               asrd_lat_token *toks, int64_t tok_cap, asrd_lat_link *links, int64_t link_cap,
                               int64_t *n_toks, int64_t *n_links, int32_t *status, void *stream);

/* DecoderItf::GetBestPath (decoder-itf.h:22; inl.h:1071-1200), batched.  For stream i the
 * arcs of the linear best-path lattice are written in path order (start -> end) to
 * ilabel/olabel/graph/acoustic[i*cap .. i*cap + n_arcs[i]), including the label-free arc
 * of the start token (inl.h:1193-1198).  status[i] receives the per-stream status
 * (ASRD_ERR_NO_TOKENS = the reference's `false`).  Synchronises `stream`. */
int asrd_get_best_path(asrd_decoder *const *decs, int32_t n, int32_t use_final_probs, int32_t cap,
                       int32_t *ilabel, int32_t *olabel, float *graph, float *acoustic,
                       int32_t *n_arcs, int32_t *status, void *stream);

/* LatticeToVector (src/newfst/lattice-functions.cc:179-217) over one best path. */
int asrd_path_to_vector(const int32_t *ilabel, const int32_t *olabel, const float *graph,
                        const float *acoustic, int32_t n_arcs, int32_t *words, int32_t *n_words,
                        int32_t *alignment, int32_t *n_alignment, float *tot_score, float *lm_score);

/* per-frame statistics (needs collect_stats); returns the number of entries available */
int32_t asrd_frame_stats(asrd_decoder *d, asrd_frame_stat *out, int32_t cap, void *stream);

/* sticky per-stream status (overflow flags raised by kernels); synchronises `stream` */
int asrd_decoder_status(asrd_decoder *d, void *stream);

int asrd_synchronize(void *stream);

/* pinned host memory for log-likelihood staging */
int asrd_host_alloc(void **ptr, int64_t bytes);
int asrd_host_free(void *ptr);

/* utterance totals summed over the n streams since their InitDecoding: emitting arcs fetched,
 * arcs admitted, token records kept.  Synchronises `stream`. */
int asrd_get_counters(asrd_decoder *const *decs, int32_t n, int64_t *arcs_expanded,
                      int64_t *arcs_admitted, int64_t *tokens, void *stream);

/* frames the on-chip frame loop had to redo through the HBM map (diagnostic; summed over the
 * streams of the last asrd_get_counters call) */
int64_t asrd_last_fallback_frames(void);
/* prune_tokens decoders: token records the arena prunes dropped (summed), and the largest arena
 * fill any one stream reached (token records), over the streams of the last asrd_get_counters call */
int64_t asrd_last_pruned_tokens(void);
/* token records the arena holds for every frame 0..NumFramesDecoded() right now (with prune_tokens:
 * what the prunes left).  Returns the number of frames, or a negative status; fills min(n, cap).
 * Synchronises `stream`. */
int32_t asrd_arena_frame_tokens(asrd_decoder *d, uint32_t *out, int32_t cap, void *stream);
int64_t asrd_last_peak_tokens(void);
/* SM cycles the arena prune spent per phase {map build, emitting links, eps rounds, survivors to the
 * front, closing up, -}, frames swept and eps rounds, summed over the streams of the last
 * asrd_get_counters call (diagnostic) */
void asrd_last_prune_cycles(int64_t *out8);
/* SM cycles the on-chip frame loop spent per phase {prologue (row + best-token pre-pass, or the
 * GetCutoff of a launch's first frame), expansion, eps closure, write-out, GetCutoff of the next
 * frame, HBM-map fallback frames}, summed over the streams of the last asrd_get_counters call */
void asrd_last_phase_cycles(int64_t *out6);

/* Per-kernel device timing with CUDA events on the launching stream (measurement aid for
 * bench.py's roofline; adds gaps between launches, so keep it off in timed regions). */
int asrd_profile_enable(int on);
int asrd_profile_reset(void);
/* kernel_ms[4], kernel_launches[4]: accumulated device time and launch counts of
 * {k_expand, k_post, k_stream, unused} since the last reset */
int asrd_profile_get(double *kernel_ms, int64_t *kernel_launches);

/* number of kernels launched by this library since load (bench.py "gpu_launches") */
int64_t asrd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ASRD_H_ */
