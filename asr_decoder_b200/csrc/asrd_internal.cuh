// Internal device/host structures of the B200 WFST decoder.  Not part of the ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "asrd.h"

namespace asrd {

constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;                 // lattice maps (32-bit keys)
constexpr unsigned long long kEmptyKey64 = 0xFFFFFFFFFFFFFFFFull;  // token map
constexpr uint32_t kNoArc = 0xFFFFFFFFu;
constexpr unsigned long long kInfVal = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kOrdInf = 0xFF800000u;  // f2ord(+inf)
constexpr uint32_t kDestEpsBit = 0x80000000u;  // device copy of StdArc::nextstate: top bit = dest has eps arcs
constexpr uint32_t kStateMask = 0x7FFFFFFFu;

constexpr int kExpandThreads = 256;   // CTA size of the expand kernel (8 warps, one token group each)
constexpr int kFinThreads = 256;      // CTA size of the finalize kernel
constexpr int kStreamThreads = 1024;  // CTA size of the per-stream kernels (closure, cutoff)
constexpr int kBestPathThreads = 512; // CTA size of the back-trace kernel
constexpr int kMaxBatch = 65535;      // streams per launch (gridDim.y of k_expand)
constexpr int kMinHashCapacity = 4096;

// One slot of the per-frame state->token map.  key = HCLG state (biglm: state | LM pair id << 32,
// the reference's PairId, my-decoder/online-decoder-mempool-base-biglm.h:77-80).
// val = (ordered cost << 32) | global arc id, recombined with a single 64-bit atomicMin: lowest
// cost wins, equal cost -> lowest arc id.
struct __align__(16) HashEntry {
  unsigned long long key;   // kEmptyKey64 when free
  unsigned long long val;
};

// Device view of one LM FSA (reference Fsa, newlm/arpa2fsa.h:217-480): state 0 is the unigram
// state and is direct-indexed by word id; other states hold word-sorted arcs and a back-off link.
struct LmView {
  const uint32_t *arc_off;     // [n_states + 1]
  const int32_t *arc_word;     // [n_arcs]
  const float *arc_weight;
  const int32_t *arc_to;
  const float *backoff_prob;   // [n_states]
  const int32_t *backoff_id;
  int32_t bos, eos, n_states;
};

// Device-resident graph: the reference's two flat arrays (src/newfst/optimize-fst.h:60-61)
// as CSR.  arcs[] are the 16-byte StdArc records verbatim (eps arcs first in every row).
struct GraphView {
  const int4 *arcs;        // {ilabel, olabel, weight bits, nextstate | kDestEpsBit if the destination has eps arcs}
  const uint2 *rows;       // [S+1] {row_off, emit_off}: eps span [x, y)
  const uint2 *erows;      // [S]   {emit_off, row_end}: emitting span [x, y) — one 8-byte load per token
  const uint4 *eps_rows;   // [S]   {row_off, emit_off, weight bits and nextstate word of the FIRST eps arc}: a
                           //       state with one eps arc (the common case) is relaxed with a single load
  const uint32_t *arc_src; // [A] source state of every arc
  const uint32_t *in_off;  // [S+1] incoming-arc index: arcs INTO state s are in_arc[in_off[s] .. in_off[s+1])
  const uint32_t *in_arc;  // [A] arc ids grouped by destination state: the eps arcs first, ascending inside a class
  const uint32_t *in_mid;  // [S] where the emitting arcs of the group start
  const uint32_t *par_bits;// [ceil(A/32)] arc has a same-class sibling with the same (src, dst)
  const uint32_t *eps_bits;// [ceil(S/32)] state has at least one input-epsilon arc
  int32_t n_states;
  uint32_t n_arcs;
  int32_t start;
  int32_t final_state;
  // CLG graphs (asrd_graph_read_clg: the graph ClgFst expands on the fly, written out over the
  // reference's own two-level state ids).  The reference's CLG decoder compares differently
  // (my-decoder/online-clg-decoder-mempool-base.h): tokens are expanded when cost < cur_cutoff
  // (:128, strict), an emitting arc is skipped only when its cost is ABOVE the cutoff (:156), and
  // the best-token pre-pass adds the CLG-arc and the HMM-arc weight of a two-level arc one after
  // the other (:91) — w_clg / w_hmm hold them for the arcs flagged in clg2_bits.
  int32_t clg;
  const float *w_clg;
  const float *w_hmm;
  const uint32_t *clg2_bits;
};

// Per-stream (per decoder object) state, resident in HBM.
struct StreamState {
  // ---- buffers
  HashEntry *hash;      // state->token map of the frame being built (recycled by the finalize phase)
  uint32_t *bm;         // claimed-slot bitmap of the map (capacity / 32 words)
  uint32_t *ebm;        // ... of which the state has eps arcs (seeds of the eps closure)
  uint32_t *queue[2];   // eps-closure frontier queues
  uint32_t *stamp;      // [capacity] eps-closure round stamp per slot (de-duplicates the frontier queue)
  uint32_t *tok_lm;     // biglm: token arena, LM pair id of every token
  unsigned long long *pair_map;  // biglm: interned (lm1 state, lm2 state) pairs; the slot index is the pair id
  uint32_t pair_mask;
  uint32_t pad2;
  uint2 *tok_sc;        // token arena: {state, cost bits}
  uint32_t *tok_arc;    // biglm: token arena, arc that set the token's cost (kNoArc for the start token);
                        // plain decoders keep {state, cost} only — the trace-back finds the arc through
                        // the graph's incoming-arc index (k_best_path_rev)
  uint32_t *tok_extra;  // prune_tokens: extra cost of every token as of the last arena prune (change detection)
  uint32_t *frame_off;  // [max_frames + 2] arena offset of every frame's token span
  float *frame_nc;      // [max_frames + 2] final next_cutoff of the step that produced each frame
  float *frame_cur;     // [max_frames + 2] GetCutoff result of each frame (which tokens were expanded)
  float *ll_hist;       // [max_frames x ll_stride] log-likelihood rows seen so far (trace-back needs them)
  asrd_frame_stat *stats;  // [max_frames + 1] or null
  uint32_t hash_mask;
  uint32_t hash_shift;  // 32 - log2(capacity)
  uint32_t token_capacity;
  int32_t max_frames;
  // ---- search state that survives between AdvanceDecoding calls
  int32_t frame;        // frames decoded so far == index of the current token span
  int32_t status;       // sticky ASRD_ERR_* raised by kernels
  uint32_t n_cur;       // tokens in the current frame
  int32_t finalized;
  unsigned long long best64;  // (ordered cost << 32 | state) of the best token of the current frame
  int32_t ll_stride;    // floats per history row
  int32_t target_frame; // this AdvanceDecoding call decodes while frame < target_frame
  // ---- utterance totals
  unsigned long long tot_arcs_expanded;
  unsigned long long tot_arcs_admitted;
  unsigned long long tot_fallback_frames;  // frames k_stream had to redo through the HBM map
  int32_t gc_frame;     // frontier frame of the last arena prune (0 after InitDecoding)
  uint32_t peak_tokens; // high-water mark of the arena (token records)
  unsigned long long tot_pruned_tokens;    // tokens dropped by the arena prunes of this utterance
  unsigned long long prune_cycles[8];      // k_prune: SM cycles per phase, frames swept, fixed-point rounds
  unsigned long long phase_cycles[6];      // k_stream: SM cycles per phase (prologue, expansion, closure, write-out, next cutoff, fallback)
};

// Per-stream descriptor of the frame step in flight: everything the grid-wide kernels need, in
// one contiguous 128-byte record per stream (one L2 round trip instead of a pointer chase
// through StreamState).  Rebuilt by k_post(PRO) / k_stream / k_init for every step; the running cutoff,
// survivor counter and best token are accumulated here with atomics.
struct __align__(16) FrameDesc {
  StreamState *st;
  const uint2 *toks;    // tokens of frame t (being expanded)
  const float *ll;      // log-likelihood row of frame t
  HashEntry *hn;        // map of frame t+1
  const uint32_t *toks_lm;  // biglm: LM pair ids of the tokens of frame t
  uint32_t *bm;         // claimed-slot bitmap of hn
  uint32_t *ebm;        // eps-seed bitmap of hn
  uint2 *out_sc;        // arena write window of frame t+1
  uint32_t *out_arc;
  unsigned long long best64;
  uint32_t n_cur;
  float cur_cut;
  float abeam;
  uint32_t next_cut_bits;  // running next_cutoff (ordered uint, atomicMin)
  uint32_t mask;
  uint32_t shift;
  uint32_t out_cap;
  uint32_t n_alive;
  uint32_t arcs_expanded;
  uint32_t arcs_admitted;
  int32_t stepping;
  int32_t t;            // frame being expanded; -1 while InitDecoding completes frame 0
};
static_assert(sizeof(FrameDesc) == 128, "FrameDesc is one 128-byte line");

// One slot of the per-frame lookup maps of the lattice back-sweep.
struct __align__(16) LatEntry {
  uint32_t key;        // state, kEmptyKey when free
  uint32_t cost_bits;  // token cost
  uint32_t extra_ord;  // ordered extra_cost (atomicMin), kOrdInf = token not (yet) reachable
  uint32_t idx;        // arena index of the token
};

struct LatticeOut {    // per-stream output window of k_lattice
  asrd_lat_token *toks;
  uint32_t *tok_arena_idx;   // arena index of every emitted token (links refer to arena indices)
  asrd_lat_link *links;
  LatEntry *map[2];
  uint32_t *map_pair[2];     // biglm: LM pair id of the token in every map slot
  uint32_t tok_cap, link_cap;
  uint32_t n_toks, n_links;   // produced (may exceed the caps: then nothing beyond the cap was written)
};

struct AdvanceParams {
  const float *ll;   // device pointer to the first new row of this chunk
  int32_t stride;    // floats between rows
  int32_t n_frames;  // rows in this chunk
  int32_t frame0;    // frame index of the first row (host-tracked, so chunks can be staged ahead)
  int32_t pad;
};

struct DecoderConfigDev {
  float beam;
  int32_t max_active;
  int32_t min_active;
  float lattice_beam;
  float beam_delta;
  int32_t collect_stats;
  int32_t debug_flags;   // measurement aids (ASRD_DEBUG_FLAGS)
};

// order-preserving float <-> uint32 map (handles negative costs): two instructions each way
__host__ __device__ inline uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  const uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);  // negative: flip all bits; else: set the sign bit
}
__host__ __device__ inline float ord2f(uint32_t o) {
  const uint32_t u = o ^ (~(uint32_t)((int32_t)o >> 31) | 0x80000000u);
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

}  // namespace asrd

struct asrd_lm {
  int refs;  // the handle + one per decoder built on it (guarded by g_ref_mu); freed when it reaches 0
  int device;
  asrd::LmView view;
  void *slab;
};

struct asrd_graph {
  int refs;  // the handle + one per decoder built on it (guarded by g_ref_mu); freed when it reaches 0
  int device;
  asrd::GraphView view;
  void *d_arcs, *d_rows, *d_erows, *d_eps_rows, *d_arc_src, *d_par, *d_eps, *d_in_off, *d_in_arc, *d_in_mid, *d_w_clg, *d_w_hmm, *d_clg2;
  int32_t max_ilabel;
  int64_t device_bytes;
  int64_t total_arcs;
};

struct asrd_decoder {
  int device;
  asrd_graph *graph;
  asrd_lm *lm1, *lm2;          // biglm: old LM (already scaled by -1 by the caller) and new LM
  asrd_config cfg;
  asrd_device_options opts;
  asrd::StreamState *d_state;  // device copy
  asrd::StreamState h_state;   // host mirror of the static fields
  void *slab;                  // one allocation carved into the buffers above
  int64_t slab_bytes;
  float *d_ll_hist;            // allocated at the first AdvanceDecoding (needs num_indices)
  int32_t ll_cols;
  int32_t frames_decoded;      // host-side mirror of StreamState::frame
  int32_t last_prune_frame;    // host-side mirror of StreamState::gc_frame (prune_tokens)
  int32_t finalized;
  int32_t initialized;
};
