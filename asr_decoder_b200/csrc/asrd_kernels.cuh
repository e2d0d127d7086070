// Hand-written sm_100a kernels of the WFST token-passing beam search.
//
//   k_stream    — the default path of plain decoders: ONE CTA per stream runs the whole frame loop
//                 of an AdvanceDecoding chunk with the per-frame state->token map in shared memory
//                 (GetCutoff, ProcessEmitting, ProcessNonemitting, survivor write-out;
//                 reference src/my-decoder/online-decoder-base-inl.h:138-431)
//   k_expand    — emitting-arc expansion over per-stream maps in HBM, all streams of a sub-batch
//                 per launch (ProcessEmitting's hot loop, inl.h:311-347); biglm decoders and, as
//                 the device function expand_frame, the overflow frames of k_stream
//   k_post      — one CTA per stream over the HBM map: eps closure (inl.h:353-431) as frontier
//                 rounds, survivors (cost < final cutoff) appended to the token arena with the
//                 arc the reference's trace-back would report (inl.h:1169-1186 + the lattice-beam
//                 link pruning of inl.h:524-542), map recycling, GetCutoff for the next frame
//                 (inl.h:138-234, exact radix select) and the best-token pre-pass (inl.h:282-300)
//   k_best_path — back-trace (BestPathEnd / TraceBackBestPath, inl.h:1096-1200)
//   k_lattice   — raw lattice with FinalizeDecoding's pruning (inl.h:725-975)
#pragma once

#include <math_constants.h>

#include "asrd_internal.cuh"

namespace asrd {

// ------------------------------------------------------------------ small helpers

__device__ __forceinline__ uint32_t hash_state(uint32_t s, uint32_t mask, uint32_t shift) {
  // Fibonacci hashing: the active states of a frame come in dense runs of neighbouring ids;
  // the multiplicative scramble spreads them uniformly over the map, which keeps linear-probe
  // chains short AND balances the per-1024-slot work groups of the closure and write-out walks of k_post.
  (void)mask;
  return (s * 0x9E3779B1u) >> shift;
}
__device__ __forceinline__ uint32_t hash_key(unsigned long long key, uint32_t mask, uint32_t shift) {
  // plain decoders: key == state, same slot as hash_state(state); biglm: the LM pair id is mixed in
  const uint32_t s = (uint32_t)key, p = (uint32_t)(key >> 32);
  return ((s + p * 0x85EBCA6Bu) * 0x9E3779B1u) >> shift;
}

// L2 residency control.  The graph (arcs + rows, ~88 MB at config 2) is re-read every frame by
// every stream at random and fits the 126 MB L2; the per-stream maps, token arena and
// log-likelihood rows stream through.  Graph loads carry an evict_last policy, streaming
// accesses evict_first, so the streaming traffic stops flushing the graph out of L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int4 ldg_arc(const int4 *ptr, uint64_t pol) {
  int4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(ptr), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 ldg_u2(const uint2 *ptr, uint64_t pol) {
  uint2 r;
  asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(ptr), "l"(pol));
  return r;
}

__device__ __forceinline__ unsigned long long pack_val(float cost, uint32_t arc) {
  return ((unsigned long long)f2ord(cost) << 32) | arc;
}

__device__ __forceinline__ bool par_bit(const uint32_t *bits, uint32_t arc) {
  return (__ldg(&bits[arc >> 5]) >> (arc & 31u)) & 1u;
}

// Find-or-claim the slot of `state` (FindOrAddToken, inl.h:88-136): CAS first, so a new
// state costs one L2 round trip and an existing one as well.
__device__ __forceinline__ bool hash_claim(HashEntry *tab, uint32_t mask, uint32_t h, unsigned long long state,
                                           uint32_t &slot, bool &is_new) {
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    const unsigned long long k = atomicCAS(&tab[h].key, kEmptyKey64, state);
    if (k == kEmptyKey64 || k == state) {
      is_new = (k == kEmptyKey64);
      slot = h;
      return true;
    }
    h = (h + 1) & mask;
  }
  return false;
}

__device__ __forceinline__ bool eps_bit(const uint32_t *bits, uint32_t state) {
  return (__ldg(&bits[state >> 5]) >> (state & 31u)) & 1u;
}

template <int NT>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t &total) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= d) incl += n;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < NW ? s_warp[lane] : 0;
    uint32_t wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, d);
      if (lane >= d) wi += n;
    }
    if (lane < NW) s_warp[lane] = wi - w;
    if (lane == NW - 1) s_warp[NW] = wi;
  }
  __syncthreads();
  uint32_t res = incl - v + s_warp[warp];
  total = s_warp[NW];
  __syncthreads();
  return res;
}

template <int NT>
__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v, unsigned long long *s_red) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
    v = o < v ? o : v;
  }
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < NW ? s_red[lane] : kInfVal;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
      v = o < v ? o : v;
    }
    if (lane == 0) s_red[0] = v;
  }
  __syncthreads();
  v = s_red[0];
  __syncthreads();
  return v;
}

template <int NT>
__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t *s_red) {
  constexpr int NW = NT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = __reduce_add_sync(0xFFFFFFFFu, v);
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < NW ? s_red[lane] : 0;
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if (lane == 0) s_red[0] = v;
  }
  __syncthreads();
  v = s_red[0];
  __syncthreads();
  return v;
}

// ------------------------------------------------------------------ biglm: LM-difference composition

struct LmPair {  // kernel argument: old LM (pre-scaled by -1 by the caller) and new LM
  LmView lm1, lm2;
};

// Fsa::GetArc (newlm/arpa2fsa.cc:244-262): word 0 = back-off link; the unigram state 0 is
// direct-indexed by word id (arpa2fsa.h:211-214), other states are binary-searched (:194-210).
__device__ __forceinline__ bool fsa_get_arc(const LmView &lm, int32_t id, int32_t word, float &weight, int32_t &to) {
  const uint32_t b = __ldg(&lm.arc_off[id]);
  if (id == 0) {
    weight = __ldg(&lm.arc_weight[b + word]);
    to = __ldg(&lm.arc_to[b + word]);
    return true;
  }
  int start = 0, end = (int)(__ldg(&lm.arc_off[id + 1]) - b) - 1;
  while (start <= end) {
    const int mid = (start + end) / 2;
    const int32_t w = __ldg(&lm.arc_word[b + mid]);
    if (w > word) end = mid - 1;
    else if (w < word) start = mid + 1;
    else {
      weight = __ldg(&lm.arc_weight[b + mid]);
      to = __ldg(&lm.arc_to[b + mid]);
      return true;
    }
  }
  return false;
}

// ComposeArpaLm::GetArc (newlm/compose-arpalm.cc:52-70): walk the back-off chain, summing the
// back-off weights in float, then the arc weight; Value1 = -(sum).
__device__ __forceinline__ void clm_get_arc(const LmView &lm, int32_t s, int32_t word, int32_t &next, float &value1) {
  float weight = 0.0f, w_arc = 0.0f;
  int32_t to = 0;
  while (!fsa_get_arc(lm, s, word, w_arc, to)) {
    w_arc = __ldg(&lm.backoff_prob[s]);
    to = __ldg(&lm.backoff_id[s]);
    s = to;
    weight += w_arc;
  }
  weight += w_arc;
  value1 = -1 * weight;
  next = to;
}

__device__ __forceinline__ float clm_final(const LmView &lm, int32_t s) {  // compose-arpalm.cc:15-29
  int32_t next;
  float v;
  clm_get_arc(lm, s, lm.eos, next, v);
  return v;
}

__device__ __forceinline__ int32_t clm_start(const LmView &lm) {  // compose-arpalm.cc:5-13
  float w;
  int32_t to;
  fsa_get_arc(lm, 0, lm.bos, w, to);
  return to;
}

// DiffArpaLm state table (newlm/diff-lm.h:92-103): (lm1 state, lm2 state) pairs interned in an
// open-addressing table; the SLOT INDEX is the pair id, so no separate numbering is needed.
__device__ __forceinline__ uint32_t pair_intern(unsigned long long *map, uint32_t pmask, int32_t n1, int32_t n2,
                                                int32_t *status) {
  const unsigned long long key = (((unsigned long long)(uint32_t)n1 << 32) | (uint32_t)n2) + 1ull;
  uint32_t h = ((uint32_t)n1 * 0x9E3779B1u + (uint32_t)n2 * 0x85EBCA6Bu) & pmask;
  // (bounded: a nearly full table must fail, not crawl — linear probing degrades to O(capacity))
  for (uint32_t probe = 0; probe <= pmask && probe < 2048u; ++probe) {
    unsigned long long k = __ldcg(&map[h]);
    if (k == 0ull) k = atomicCAS(&map[h], 0ull, key);
    if (k == 0ull || k == key) return h;
    h = (h + 1) & pmask;
  }
  atomicMin(status, ASRD_ERR_LM_PAIRS_OVERFLOW);
  return 0;
}

// NextLmState (…-biglm.h:54-70) with the INTENDED DiffArpaLm semantics: both LMs advance from the
// members of the pair (the reference passes the pair-state id itself, newlm/diff-lm.h:75-86 —
// SURVEY.md Appendix B-6; on unigram-only LMs the two coincide).  Returns the LM-difference score
// Times(w1, w2).Value1() and the successor LM states (not yet interned).
__device__ __forceinline__ float lm_step(const LmPair &lms, const unsigned long long *map, uint32_t pair_id,
                                         int32_t word, int32_t &n1, int32_t &n2) {
  const unsigned long long pk = __ldcg(&map[pair_id]) - 1ull;
  float v1, v2;
  clm_get_arc(lms.lm1, (int32_t)(uint32_t)(pk >> 32), word, n1, v1);
  clm_get_arc(lms.lm2, (int32_t)(uint32_t)pk, word, n2, v2);
  return v1 + v2;
}

__device__ __forceinline__ float lm_final(const LmPair &lms, const unsigned long long *map, uint32_t pair_id) {
  const unsigned long long pk = __ldcg(&map[pair_id]) - 1ull;  // DiffArpaLm::Final, diff-lm.h:48-55
  return clm_final(lms.lm1, (int32_t)(uint32_t)(pk >> 32)) + clm_final(lms.lm2, (int32_t)(uint32_t)pk);
}

// ------------------------------------------------------------------ warp work splitting

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t n = __shfl_up_sync(kFull, v, d);
    if (lane >= d) v += n;
  }
  return v;
}

// Lane that owns flattened item j, given each lane's exclusive prefix `off`: the last lane l
// with off_l <= j (lanes with zero items never win because their successor shares the offset).
// Must be called by the full warp.
__device__ __forceinline__ int warp_owner(uint32_t off, uint32_t j) {
  int l = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const uint32_t v = __shfl_sync(kFull, off, l + step);
    if (v <= j) l += step;
  }
  return l;
}

// ------------------------------------------------------------------ init / begin-advance

__device__ __forceinline__ void clear_map_by_bitmap(HashEntry *h, uint32_t *bm, uint32_t words, int warp,
                                                    int lane, int n_warps) {
  // Each warp takes 32 bitmap words at a time (one coalesced load), then lane b recycles slot b
  // of every non-empty word: 512 contiguous bytes per word, no dependent global loads.
  for (uint32_t w0 = warp * 32; w0 < words; w0 += n_warps * 32) {
    const uint32_t mine = (w0 + lane < words) ? bm[w0 + lane] : 0u;
    if (mine) bm[w0 + lane] = 0;
    unsigned nz = __ballot_sync(kFull, mine != 0);
    while (nz) {
      const int k = __ffs(nz) - 1;
      nz &= nz - 1;
      const uint32_t bits = __shfl_sync(kFull, mine, k);
      if ((bits >> lane) & 1u) {
        HashEntry e;
        e.key = kEmptyKey64; e.val = kInfVal;
        h[(w0 + k) * 32 + lane] = e;
      }
    }
  }
}

// InitDecoding (inl.h:41-67): forget the previous utterance and seed the start token; the
// closure / finalize / cutoff kernels that follow complete frame 0 with cutoff = beam.
template <bool BIGLM>
__global__ void __launch_bounds__(kStreamThreads)
k_init(StreamState *const *streams, FrameDesc *desc, GraphView g, DecoderConfigDev cfg, LmPair lms) {
  StreamState *st = streams[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t mask = st->hash_mask;
  const uint32_t words = (mask + 1) >> 5;
  // the map is normally left clean by the last finalize phase; an aborted utterance may not have
  clear_map_by_bitmap(st->hash, st->bm, words, warp, lane, kStreamThreads / 32);
  for (uint32_t i = tid; i < words; i += kStreamThreads) st->ebm[i] = 0;
  for (uint32_t i = tid; i <= mask; i += kStreamThreads) st->stamp[i] = 0;
  if (BIGLM)  // DiffArpaLm::Reset (newlm/diff-lm.h:39-46): the pair table is per utterance
    for (uint32_t i = tid; i <= st->pair_mask; i += kStreamThreads) st->pair_map[i] = 0ull;
  __syncthreads();
  if (tid == 0) {
    st->status = 0;
    st->finalized = 0;
    st->frame = -1;
    st->target_frame = 0;
    st->frame_off[0] = 0;
    st->n_cur = 0;
    st->best64 = kInfVal;
    st->tot_arcs_expanded = st->tot_arcs_admitted = st->tot_fallback_frames = 0;
    st->gc_frame = 0;
    st->peak_tokens = 0;
    st->tot_pruned_tokens = 0;
    for (int k = 0; k < 8; ++k) st->prune_cycles[k] = 0;
    for (int k = 0; k < 6; ++k) st->phase_cycles[k] = 0;
    uint32_t slot;
    bool is_new;
    HashEntry *hn = st->hash;
    unsigned long long start_key = (uint32_t)g.start;
    if (BIGLM)  // …-biglm.h:112: start pair = (graph start, DiffArpaLm::Start())
      start_key |= (unsigned long long)pair_intern(st->pair_map, st->pair_mask, clm_start(lms.lm1),
                                                   clm_start(lms.lm2), &st->status) << 32;
    hash_claim(hn, mask, hash_key(start_key, mask, st->hash_shift), start_key, slot, is_new);
    atomicMin(&hn[slot].val, pack_val(0.0f, kNoArc));
    st->bm[slot >> 5] = 1u << (slot & 31u);
    if (eps_bit(g.eps_bits, (uint32_t)g.start)) st->ebm[slot >> 5] = 1u << (slot & 31u);
    FrameDesc d;
    d.st = st;
    d.toks = nullptr;
    d.ll = nullptr;
    d.hn = hn;
    d.toks_lm = nullptr;
    d.bm = st->bm;
    d.ebm = st->ebm;
    d.out_sc = st->tok_sc;
    d.out_arc = st->tok_arc;
    d.best64 = kInfVal;
    d.n_cur = 0;
    d.cur_cut = 0.f;
    d.abeam = 0.f;
    d.next_cut_bits = f2ord(cfg.beam);
    d.mask = mask;
    d.shift = st->hash_shift;
    d.out_cap = st->token_capacity;
    d.n_alive = 0;
    d.arcs_expanded = d.arcs_admitted = 0;
    d.stepping = 1;
    d.t = -1;
    desc[blockIdx.x] = d;
  }
}

// Start of an AdvanceDecoding chunk: append the chunk's log-likelihood rows to every stream's
// history (the search reads them from there, and so does the trace-back) and set the frame
// target.  One CTA per (row block, stream).
__global__ void __launch_bounds__(256)
k_begin_advance(StreamState *const *streams, const AdvanceParams *params, int num_indices) {
  StreamState *st = streams[blockIdx.y];
  const AdvanceParams p = params[blockIdx.y];
  const int frame0 = p.frame0;
  int nf = p.n_frames;
  if (frame0 + nf > st->max_frames) {
    nf = st->max_frames - frame0;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicMin(&st->status, ASRD_ERR_FRAMES_OVERFLOW);
  }
  // chunks are staged ahead of the frame loops: the target only ever grows
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicMax(&st->target_frame, frame0 + (nf > 0 ? nf : 0));
  const int hs = st->ll_stride;
  float *dst = st->ll_hist + (size_t)frame0 * hs;
  const bool vec = ((num_indices & 3) == 0) && ((p.stride & 3) == 0) && ((hs & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(p.ll) & 15) == 0);
  for (int r = blockIdx.x; r < nf; r += gridDim.x) {
    const float *src = p.ll + (size_t)r * p.stride;
    float *drow = dst + (size_t)r * hs;
    if (vec) {
      const float4 *s4 = reinterpret_cast<const float4 *>(src);
      float4 *d4 = reinterpret_cast<float4 *>(drow);
      // streaming (evict-first) on both sides: the rows pass through L2 once and must not push
      // the graph out of it
      for (int c = threadIdx.x; c < (num_indices >> 2); c += blockDim.x) __stcs(&d4[c], __ldcs(&s4[c]));
    } else {
      for (int c = threadIdx.x; c < num_indices; c += blockDim.x) __stcs(&drow[c], __ldcs(&src[c]));
    }
  }
}

// ------------------------------------------------------------------ expand

// Grid (G, n_streams): blockIdx.y is the stream, its CTAs share the stream's token groups.
// One warp owns a group of 32 tokens: lane i loads token i and its emitting-arc span, a
// shuffle prefix sum flattens the spans, and the lanes then walk the flattened arc list so
// consecutive lanes fetch consecutive 16-byte arc records (LDG.128).  U x 32 arcs per warp are
// in flight per iteration.  The stream's log-likelihood row is staged in shared memory once
// per CTA (the only block barrier).  Token recombination: one 64-bit atomicMin per admitted
// arc on (ordered cost << 32 | arc id) in the per-frame state->token map.
// Body of the expansion: warp `warp_global` of `n_warps` walks the token groups of the frame
// described by *d (global or shared memory).  s_ll: the frame's log-likelihood row in shared
// memory (SMEM_LL) or unused.
template <int U, bool SMEM_LL, bool BIGLM>
__device__ __forceinline__ void expand_frame(FrameDesc *d, const GraphView &g, const float *s_ll,
                                             uint32_t warp_global, uint32_t n_warps, int flags, const LmPair &lms) {
  const uint32_t n_cur = d->n_cur;
  const uint32_t n_groups = (n_cur + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const float *__restrict__ ll = d->ll;
  const uint2 *__restrict__ toks = d->toks;
  const float cur_cut = d->cur_cut, abeam = d->abeam;
  HashEntry *hn = d->hn;
  uint32_t *bm = d->bm, *ebm = d->ebm;
  const uint32_t mask = d->mask, shift = d->shift;
  uint32_t *next_cut = &d->next_cut_bits;
  const uint32_t *__restrict__ toks_lm = d->toks_lm;
  unsigned long long *pair_map = BIGLM ? d->st->pair_map : nullptr;
  const uint32_t pair_mask = BIGLM ? d->st->pair_mask : 0;
  const bool hints = (flags & 1) != 0;
  const uint64_t pol_graph = hints ? l2_policy_evict_last() : 0, pol_stream = hints ? l2_policy_evict_first() : 0;

  uint32_t expanded = 0, admitted = 0;
  for (uint32_t grp = warp_global; grp < n_groups; grp += n_warps) {
    // running cutoff (inl.h:330), refreshed once per group; our own tightenings are applied
    // locally below, other warps' arrive with the next group
    float nc = ord2f(*reinterpret_cast<volatile uint32_t *>(next_cut));  // (*d may live in shared memory)
    // ---- lane i: token i of the group and its emitting span
    const uint32_t i = grp * 32 + lane;
    uint32_t deg = 0, base = 0, cost_bits = 0, my_pair = 0;
    if (i < n_cur) {
      // (L2 loads, not ld.global.nc: inside k_stream the tokens were written by this very kernel)
      const uint2 sc = hints ? ldg_u2(&toks[i], pol_stream) : __ldcg(&toks[i]);
      cost_bits = sc.y;
      if (BIGLM) my_pair = __ldcg(&toks_lm[i]);
      if (g.clg ? __uint_as_float(sc.y) < cur_cut : __uint_as_float(sc.y) <= cur_cut) {  // inclusive, inl.h:315 (CLG: strict)
        const uint2 er = hints ? ldg_u2(&g.erows[sc.x], pol_graph) : __ldg(&g.erows[sc.x]);
        base = er.x;
        deg = er.y - er.x;
      }
    }
    const uint32_t incl = warp_incl_scan(deg, lane);
    const uint32_t off = incl - deg;
    const uint32_t total = __shfl_sync(kFull, incl, 31);
    expanded += total;

    for (uint32_t jb = 0; jb < total; jb += 32 * U) {
      bool in[U];
      uint32_t a[U], tpair[U];
      float tcost[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = jb + u * 32 + lane;
        in[u] = j < total;
        const int l = warp_owner(off, j);
        const uint32_t off_l = __shfl_sync(kFull, off, l);
        const uint32_t base_l = __shfl_sync(kFull, base, l);
        tcost[u] = __uint_as_float(__shfl_sync(kFull, cost_bits, l));
        tpair[u] = BIGLM ? __shfl_sync(kFull, my_pair, l) : 0u;
        a[u] = base_l + (j - off_l);
      }
      int4 arc[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (in[u]) arc[u] = hints ? ldg_arc(&g.arcs[a[u]], pol_graph) : __ldg(&g.arcs[a[u]]);
      float tot[U];
      bool adm[U];
      uint32_t h0[U];
      uint4 e0[U];
      unsigned long long dkey[U];
      uint32_t cand_bits = 0xFFFFFFFFu;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        adm[u] = false;
        tot[u] = 0.f;
        h0[u] = 0;
        dkey[u] = 0;
        e0[u] = make_uint4(0, 0, 0, 0);
        if (in[u]) {
          const float ac = -(SMEM_LL ? s_ll[arc[u].x - 1] : __ldg(&ll[arc[u].x - 1]));
          float graph_cost = __int_as_float(arc[u].z);
          int32_t n1 = 0, n2 = 0;
          if (BIGLM && arc[u].y != 0)  // …-biglm.h:377-379: graph_cost = arc weight + LM-difference score
            graph_cost = __int_as_float(arc[u].z) + lm_step(lms, pair_map, tpair[u], arc[u].y, n1, n2);
          tot[u] = (tcost[u] + ac) + graph_cost;  // inl.h:326-329
          adm[u] = (g.clg ? tot[u] <= nc : tot[u] < nc) && !(flags & 2);  // inl.h:330 (CLG: skipped only when above); (flags & 2): measurement aid, no map traffic
          if (adm[u]) {
            const float cand = tot[u] + abeam;  // inl.h:332-333
            if (cand < nc) cand_bits = min(cand_bits, f2ord(cand));
            dkey[u] = (uint32_t)arc[u].w & kStateMask;
            if (BIGLM) {  // …-biglm.h:388: destination key = (arc target, next LM state)
              const uint32_t np = arc[u].y != 0 ? pair_intern(pair_map, pair_mask, n1, n2, &d->st->status) : tpair[u];
              dkey[u] |= (unsigned long long)np << 32;
            }
            h0[u] = hash_key(dkey[u], mask, shift);
            // first probe: the whole 16-byte entry {key, val} in one request
            e0[u] = __ldcg(reinterpret_cast<const uint4 *>(&hn[h0[u]]));
          }
        }
      }
      // warp-aggregated cutoff tightening (inl.h:332-333)
      if (__any_sync(kFull, cand_bits != 0xFFFFFFFFu)) {
        const uint32_t wmin = __reduce_min_sync(kFull, cand_bits);
        if (lane == 0 && !(flags & 4)) atomicMin(next_cut, wmin);
        nc = fminf(nc, ord2f(wmin));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (adm[u]) {
          const unsigned long long state = dkey[u];
          const unsigned long long pk = pack_val(tot[u], a[u]);
          uint32_t slot = h0[u];
          bool is_new = false, ok = (((unsigned long long)e0[u].y << 32) | e0[u].x) == state;
          // the map value only ever decreases: if what we saw already beats us, so does the
          // current value — no atomic needed (most admitted arcs lose the recombination)
          const bool lost = ok && (((unsigned long long)e0[u].w << 32) | e0[u].z) <= pk;
          if (!ok) ok = hash_claim(hn, mask, h0[u], state, slot, is_new);
          if (ok) {
            if (!lost) atomicMin(&hn[slot].val, pk);
            if (is_new) {
              atomicOr(&bm[slot >> 5], 1u << (slot & 31u));
              if ((uint32_t)arc[u].w & kDestEpsBit) atomicOr(&ebm[slot >> 5], 1u << (slot & 31u));
            }
          } else {
            atomicMin(&d->st->status, ASRD_ERR_HASH_OVERFLOW);
          }
          ++admitted;
        }
      }
    }
  }
  admitted = __reduce_add_sync(kFull, admitted);
  if (lane == 0 && expanded) {
    atomicAdd(&d->arcs_expanded, expanded);
    atomicAdd(&d->arcs_admitted, admitted);
  }
}

template <int U, bool SMEM_LL, bool BIGLM>
__global__ void __launch_bounds__(kExpandThreads, (U == 1 && !BIGLM) ? 8 : 5)
k_expand(FrameDesc *desc, GraphView g, int num_indices, int flags, LmPair lms) {
  extern __shared__ float s_ll[];
  FrameDesc *d = &desc[blockIdx.y];
  if (!d->stepping) return;
  const uint32_t n_groups = (d->n_cur + 31) >> 5;
  const int tid = threadIdx.x;
  if (blockIdx.x * (kExpandThreads / 32) >= n_groups) return;  // whole CTA idle
  if (SMEM_LL) {
    const float *__restrict__ ll = d->ll;
    for (int c = tid; c < num_indices; c += kExpandThreads) s_ll[c] = __ldg(&ll[c]);
    __syncthreads();
  }
  expand_frame<U, SMEM_LL, BIGLM>(d, g, s_ll, blockIdx.x * (kExpandThreads / 32) + (tid >> 5),
                                  gridDim.x * (kExpandThreads / 32), flags, lms);
}

// ------------------------------------------------------------------ cutoff

enum { kModeEpi = 2, kModePro = 4 };

// exact k-th smallest (0-based) cost among the tokens with ordered key < limit_ord: what
// std::nth_element yields in GetCutoff (inl.h:190-193, 211-216).  MSB-first radix select over
// (ordered key - min key), starting at the top byte of `range` (an upper bound of that span).
template <int NT, class OrdAt>
__device__ __forceinline__ float block_kth_smallest(OrdAt ord_at, uint32_t n, uint32_t k, uint32_t min_ord,
                                                    uint32_t limit_ord, uint32_t range, uint32_t *s_hist,
                                                    uint32_t *s_misc) {
  // ord_at(i), i < n: ordered cost key of item i, or 0xFFFFFFFF for "no token here"
  const int tid = threadIdx.x;
  int top = 24;
  while (top > 0 && (range >> top) == 0) top -= 8;
  uint32_t prefix = 0, pmask = 0, kk = k;
  for (int sh = top; sh >= 0; sh -= 8) {
    for (int b = tid; b < 256; b += NT) s_hist[b] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += NT) {
      const uint32_t key = ord_at(i);
      if (key < limit_ord && key != 0xFFFFFFFFu) {
        const uint32_t rk = key - min_ord;
        if ((rk & pmask) == prefix) atomicAdd(&s_hist[(rk >> sh) & 255u], 1u);
      }
    }
    __syncthreads();
    if (tid < 32) {  // warp 0: locate the bin holding rank kk
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        c[q] = s_hist[tid * 8 + q];
        sum += c[q];
      }
      const uint32_t incl = warp_incl_scan(sum, tid);
      uint32_t acc = incl - sum;
      if (kk >= acc && kk < incl) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (kk < acc + c[q]) {
            s_misc[0] = (uint32_t)(tid * 8 + q);
            s_misc[1] = kk - acc;
            break;
          }
          acc += c[q];
        }
      }
    }
    __syncthreads();
    prefix |= s_misc[0] << sh;
    pmask |= 255u << sh;
    kk = s_misc[1];
    // (no barrier here: s_misc is next written two barriers into the following pass)
  }
  __syncthreads();
  return ord2f(prefix + min_ord);
}

// GetCutoff (inl.h:138-234) over n tokens whose ordered cost keys are ord_at(i), i < n_items
// (0xFFFFFFFF = no token): the frame's tokens in the arena, or the slots of the on-chip map.
// best_ord = key of the best token.  Whole-CTA; every thread returns the same values.
template <int NT, class OrdAt>
__device__ __forceinline__ void get_cutoff(OrdAt ord_at, uint32_t n_items, uint32_t n, uint32_t best_ord,
                                           const DecoderConfigDev &cfg, uint32_t *s_red32, uint32_t *s_hist,
                                           uint32_t *s_misc, float &cur_cut, float &abeam,
                                           float all_below = 3.402823466e38f /* no bound known */) {
  const int tid = threadIdx.x;
  cur_cut = CUDART_INF_F;
  abeam = cfg.beam;
  if (n == 0) return;
  const float bc = ord2f(best_ord);
  const float beam_cut = bc + cfg.beam;  // inl.h:182
  const uint32_t beam_ord = f2ord(beam_cut);
  cur_cut = beam_cut;
  if (n <= (uint32_t)cfg.min_active && n <= (uint32_t)cfg.max_active) {
    // fewer tokens than min_active: min_active_cutoff stays +inf > beam_cutoff, nothing is
    // pruned and the adaptive beam is infinite (inl.h:183,205,220-226)
    cur_cut = CUDART_INF_F;
    abeam = CUDART_INF_F;
    return;
  }
  uint32_t lt = 0, le = 0;
  if (all_below <= beam_cut) {
    // every token is known to cost less than all_below (the frame's final next_cutoff, which the
    // search keeps at or below best + adaptive beam): no counting pass needed
    lt = le = n;
  } else {
  for (uint32_t i = tid; i < n_items; i += NT) {
    const uint32_t key = ord_at(i);
    if (key != 0xFFFFFFFFu) {
      const float c = ord2f(key);
      lt += c < beam_cut;
      le += c <= beam_cut;
    }
  }
  if (n_items < 65536u) {  // both counts in one block reduction
    const uint32_t both = block_sum_u32<NT>(lt | (le << 16), s_red32);
    lt = both & 0xFFFFu;
    le = both >> 16;
  } else {
    lt = block_sum_u32<NT>(lt, s_red32);
    le = block_sum_u32<NT>(le, s_red32);
  }
  }
  if (lt > (uint32_t)cfg.max_active) {
    // sorted[max_active] < beam_cutoff  <=>  more than max_active costs below it (inl.h:188-203)
    cur_cut = block_kth_smallest<NT>(ord_at, n_items, (uint32_t)cfg.max_active, best_ord, beam_ord,
                                     beam_ord - best_ord, s_hist, s_misc);
    abeam = cur_cut - bc + cfg.beam_delta;
  } else if (n <= (uint32_t)cfg.min_active) {
    cur_cut = CUDART_INF_F;
    abeam = CUDART_INF_F;
  } else if (cfg.min_active > 0 && le <= (uint32_t)cfg.min_active) {
    // sorted[min_active] > beam_cutoff  <=>  at most min_active costs <= it (inl.h:205-226)
    cur_cut = block_kth_smallest<NT>(ord_at, n_items, (uint32_t)cfg.min_active, best_ord, 0xFFFFFFFFu,
                                     0xFFFFFFFFu, s_hist, s_misc);
    abeam = cur_cut - bc + cfg.beam_delta;
  }
}

// Descriptor of the expansion of frame t (thread 0).
__device__ __forceinline__ void fill_desc(StreamState *st, FrameDesc *d, int t, uint32_t n, uint32_t tok_off,
                                          float cur_cut, float abeam, uint32_t next_bits, bool biglm) {
  const uint32_t out_base = st->frame_off[t + 1];
  FrameDesc nd;
  nd.st = st;
  nd.toks = st->tok_sc + tok_off;
  nd.ll = st->ll_hist + (size_t)t * st->ll_stride;
  nd.hn = st->hash;
  nd.toks_lm = biglm ? st->tok_lm + tok_off : nullptr;
  nd.bm = st->bm;
  nd.ebm = st->ebm;
  nd.out_sc = st->tok_sc + out_base;
  nd.out_arc = st->tok_arc + out_base;
  nd.best64 = kInfVal;
  st->frame_cur[t] = cur_cut;
  nd.n_cur = n;
  nd.cur_cut = cur_cut;
  nd.abeam = abeam;
  nd.next_cut_bits = next_bits;
  nd.mask = st->hash_mask;
  nd.shift = st->hash_shift;
  nd.out_cap = st->token_capacity > out_base ? st->token_capacity - out_base : 0;
  nd.n_alive = 0;
  nd.arcs_expanded = 0;
  nd.arcs_admitted = 0;
  nd.stepping = 1;
  nd.t = t;
  *d = nd;
}

// GetCutoff (inl.h:138-234) over the current frame's tokens, best-token pre-pass
// (inl.h:282-300), and the descriptor of the next expansion.  Whole-CTA device function.
template <int NT, bool BIGLM>
__device__ __forceinline__ void cutoff_prologue(StreamState *st, FrameDesc *d, const GraphView &g,
                                                const DecoderConfigDev &cfg, const LmPair &lms,
                                                unsigned long long *s_red64, uint32_t *s_red32,
                                                uint32_t *s_hist, uint32_t *s_misc) {
  const int tid = threadIdx.x;
  const int t = st->frame;
  if (t >= st->target_frame) {
    if (tid == 0) d->stepping = 0;
    return;
  }
  const uint32_t n = st->n_cur;
  const uint32_t tok_off = st->frame_off[t];
  const uint2 *toks = st->tok_sc + tok_off;
  const float *__restrict__ ll = st->ll_hist + (size_t)t * st->ll_stride;
  // best token: lowest cost, ties -> lowest state id (inl.h:169-179); accumulated by the write-out of the previous step
  const unsigned long long best64 = st->best64;
  float cur_cut, abeam;
  get_cutoff<NT>([&](uint32_t i) { return f2ord(__uint_as_float(toks[i].y)); }, n, n, (uint32_t)(best64 >> 32), cfg,
                 s_red32, s_hist, s_misc, cur_cut, abeam);
  uint32_t next_bits = kOrdInf;
  if (n > 0) {
    const float bc = ord2f((uint32_t)(best64 >> 32));
    // best-token pre-pass (inl.h:282-300): association (cost + w) - loglike;
    // biglm (…-biglm.h:339-357): ((lm_score + cost) + w) - loglike.  In biglm mode best64 carries
    // the token's index inside the frame instead of its state.
    uint32_t sb = (uint32_t)best64, bpair = 0;
    if (BIGLM) {
      bpair = st->tok_lm[tok_off + sb];
      sb = toks[sb].x;
    }
    const uint2 er = __ldg(&g.erows[sb]);
    uint32_t mn = kOrdInf;
    for (uint32_t a = er.x + tid; a < er.y; a += NT) {
      const int4 arc = __ldg(&g.arcs[a]);
      float tot;
      if (BIGLM && arc.y != 0) {
        int32_t n1, n2;
        const float lm_score = lm_step(lms, st->pair_map, bpair, arc.y, n1, n2);
        tot = lm_score + bc + __int_as_float(arc.z) - __ldg(&ll[arc.x - 1]);
      } else if (g.clg && ((__ldg(&g.clg2_bits[a >> 5]) >> (a & 31)) & 1u)) {
        // …-clg-decoder-mempool-base.h:91: tok + clgarc.w + arc.w - loglike
        tot = bc + __ldg(&g.w_clg[a]) + __ldg(&g.w_hmm[a]) - __ldg(&ll[arc.x - 1]);
      } else {
        tot = bc + __int_as_float(arc.z) - __ldg(&ll[arc.x - 1]);
      }
      mn = min(mn, f2ord(tot + abeam));
    }
    const unsigned long long m64 = block_min_u64<NT>((unsigned long long)mn, s_red64);
    next_bits = (uint32_t)m64;
  }
  if (tid == 0) fill_desc(st, d, t, n, tok_off, cur_cut, abeam, next_bits, BIGLM);
}

// ------------------------------------------------------------------ fused per-stream post phase

// Eps closure + survivor write-out + GetCutoff for one stream in ONE CTA: everything that follows the
// emitting expansion of a frame only touches the stream's own state, so a single resident CTA
// carries it from the eps closure to the descriptor of the next expansion without going back
// to the host-visible launch queue (three launches and their descriptor round trips saved).
// The survivor counter and the best token live in shared memory instead of global atomics.
struct PostSmem {  // block-level scratch of the per-stream phases
  unsigned long long red64[kStreamThreads / 32];
  unsigned long long best;
  uint32_t red32[kStreamThreads / 32 + 1];  // (+1: block_exclusive_scan keeps the total behind the warp sums)
  uint32_t hist[256];
  uint32_t misc[4];
  uint32_t qn[2];
  uint32_t alive;
};

// Frame bookkeeping once the survivors of frame t+1 are in the arena (thread 0 only).
__device__ __forceinline__ void frame_commit(StreamState *st, FrameDesc *d, const DecoderConfigDev &cfg, float nc,
                                             uint32_t n_alive, unsigned long long b64) {
  const int t = d->t;
  if (n_alive > d->out_cap) {
    n_alive = d->out_cap;
    atomicMin(&st->status, ASRD_ERR_ARENA_OVERFLOW);
  }
  const uint32_t out_base = st->frame_off[t + 1];
  if (cfg.collect_stats && st->stats) {
    asrd_frame_stat fs;
    fs.n_in = d->n_cur;
    fs.cur_cutoff = d->cur_cut;
    fs.abeam = d->abeam;
    fs.next_cutoff = nc;
    fs.n_tokens = n_alive;
    fs.best = b64 == kInfVal ? CUDART_INF_F : ord2f((uint32_t)(b64 >> 32));
    fs.arcs_expanded = d->arcs_expanded;
    fs.arcs_admitted = d->arcs_admitted;
    st->stats[t + 1] = fs;
  }
  st->tot_arcs_expanded += d->arcs_expanded;
  st->tot_arcs_admitted += d->arcs_admitted;
  st->frame_off[t + 2] = out_base + n_alive;
  st->frame_nc[t + 1] = nc;
  st->frame = t + 1;
  st->n_cur = n_alive;
  st->best64 = b64;
  d->stepping = 0;
}

// Eps closure + survivors -> arena + bookkeeping of the frame described by *d, over the
// stream's map in HBM.  Whole-CTA device function (kStreamThreads threads).
template <bool BIGLM>
__device__ __forceinline__ void post_epilogue(StreamState *st, FrameDesc *d, const GraphView &g,
                                              const DecoderConfigDev &cfg, const LmPair &lms, PostSmem &ps) {
  constexpr int NT = kStreamThreads;
  uint32_t *s_qn = ps.qn;
  uint32_t &s_alive = ps.alive;
  unsigned long long &s_best = ps.best;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const int t = d->t;
    const uint32_t mask = d->mask, shift = d->shift;
    const uint32_t groups = (mask + 1) >> 10;
    HashEntry *hn = d->hn;
    uint32_t *bm = d->bm;
    uint32_t *ebm = d->ebm;
    const float nc = ord2f(d->next_cut_bits);  // the FINAL next_cutoff of this frame
    uint32_t *q0 = st->queue[0], *q1 = st->queue[1];
    if (tid == 0) {
      s_qn[0] = s_qn[1] = 0;
      s_alive = 0;
      s_best = kInfVal;
    }
    __syncthreads();

    // ---- eps closure (ProcessNonemitting, inl.h:353-431)
    uint32_t *stamp = st->stamp;
    unsigned long long *pair_map = BIGLM ? st->pair_map : nullptr;
    const uint32_t pair_mask = BIGLM ? st->pair_mask : 0;
    const uint32_t stamp_base = (uint32_t)(t + 2) << 6;  // frame-unique round stamps: no clearing per frame
    auto relax_from = [&](uint32_t slot, uint32_t round) {
      const uint4 e = __ldcg(reinterpret_cast<const uint4 *>(&hn[slot]));
      const uint32_t state = e.x, pair = e.y;
      const float cost = ord2f(e.w);
      if (!(cost < nc)) return;  // inl.h:391
      const uint2 r = __ldg(&g.rows[state]);
      uint32_t *qout = ((round + 1) & 1) ? q1 : q0;
      for (uint32_t a = r.x; a < r.y; ++a) {
        const int4 arc = __ldg(&g.arcs[a]);
        float graph_cost = __int_as_float(arc.z);
        int32_t n1 = 0, n2 = 0;
        if (BIGLM && arc.y != 0)  // …-biglm.h:446-448
          graph_cost = __int_as_float(arc.z) + lm_step(lms, pair_map, pair, arc.y, n1, n2);
        const float tot = cost + graph_cost;  // inl.h:413-414
        if (tot < nc) {                        // inl.h:415
          uint32_t slot2;
          bool is_new;
          unsigned long long dst = (uint32_t)arc.w & kStateMask;
          if (BIGLM)
            dst |= (unsigned long long)(arc.y != 0 ? pair_intern(pair_map, pair_mask, n1, n2, &st->status) : pair) << 32;
          if (!hash_claim(hn, mask, hash_key(dst, mask, shift), dst, slot2, is_new)) {
            atomicMin(&st->status, ASRD_ERR_HASH_OVERFLOW);
            continue;
          }
          const unsigned long long pk = pack_val(tot, a);
          const unsigned long long old = atomicMin(&hn[slot2].val, pk);
          if (is_new) atomicOr(&bm[slot2 >> 5], 1u << (slot2 & 31u));
          const bool changed = (uint32_t)(pk >> 32) < (uint32_t)(old >> 32);  // inl.h:115-127
          const uint32_t sv = stamp_base + min(round + 1, 63u);
          if (changed && ((uint32_t)arc.w & kDestEpsBit) && (atomicMax(&stamp[slot2], sv) < sv || round >= 62u))
            qout[atomicAdd(&s_qn[(round + 1) & 1], 1u)] = slot2;  // inl.h:425-426
        }
      }
    };
    for (uint32_t w = tid; w < (groups << 5); w += NT) {  // round-1 seeds (inl.h:376-381)
      uint32_t bits = ebm[w];
      if (bits) {
        ebm[w] = 0;
        uint32_t pos = atomicAdd(&s_qn[1], (uint32_t)__popc(bits));
        while (bits) {
          const uint32_t b = __ffs(bits) - 1;
          bits &= bits - 1;
          q1[pos++] = (w << 5) + b;
        }
      }
    }
    for (uint32_t round = 1;; ++round) {
      __syncthreads();
      const uint32_t nq = s_qn[round & 1];
      if (nq == 0) break;
      __syncthreads();
      if (tid == 0) s_qn[(round + 1) & 1] = 0;
      __syncthreads();
      const uint32_t *qin = (round & 1) ? q1 : q0;
      for (uint32_t i = tid; i < nq; i += NT) relax_from(qin[i], round);
    }
    // (the loop exits right after a barrier: every relaxation is visible)

    // ---- survivors -> token arena
    {
      const uint32_t cap = d->out_cap;
      uint2 *out_sc = d->out_sc;
      uint32_t *out_arc = d->out_arc;
      uint32_t *out_lm = BIGLM ? st->tok_lm + st->frame_off[t + 1] : nullptr;
      unsigned long long best64 = kInfVal;
      for (uint32_t grp = warp; grp < groups; grp += NT / 32) {
        const uint32_t word = __ldcg(&bm[grp * 32 + lane]);
        if (!__any_sync(kFull, word != 0)) continue;
        if (word) bm[grp * 32 + lane] = 0;
        const uint32_t cnt = __popc(word);
        const uint32_t incl = warp_incl_scan(cnt, lane);
        const uint32_t off = incl - cnt;
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        for (uint32_t ib = 0; ib < total; ib += 32) {
          const uint32_t it = ib + lane;
          const int l = warp_owner(off, it);
          const uint32_t wl = __shfl_sync(kFull, word, l);
          const uint32_t offl = __shfl_sync(kFull, off, l);
          bool alive = false;
          uint32_t key = 0, pair = 0, rep = kNoArc;
          float cost = 0.f;
          if (it < total) {
            const uint32_t b = __fns(wl, 0, (int)(it - offl) + 1);
            const uint32_t slot = ((grp * 32 + l) << 5) + b;
            const uint4 e = __ldcg(reinterpret_cast<const uint4 *>(&hn[slot]));
            key = e.x;
            pair = e.y;
            rep = e.z;
            cost = ord2f(e.w);
            alive = g.clg ? cost <= nc : cost < nc;  // (CLG: an arc AT the cutoff was admitted)
          }
          const unsigned am = __ballot_sync(kFull, alive);
          if (am == 0) continue;
          uint32_t pos0 = 0;
          if (lane == 0) pos0 = atomicAdd(&s_alive, (uint32_t)__popc(am));
          pos0 = __shfl_sync(kFull, pos0, 0);
          if (alive) {
            const uint32_t idx = pos0 + __popc(am & ((1u << lane) - 1u));
            if (idx < cap) {
              out_sc[idx] = make_uint2(key, __float_as_uint(cost));
              if (BIGLM) {
                out_arc[idx] = rep;
                out_lm[idx] = pair;
              }
            }
            // best token: lowest cost, ties -> lowest state id; biglm keeps the token's index
            // inside the frame instead (the pre-pass needs its LM state as well)
            const unsigned long long b64 = ((unsigned long long)f2ord(cost) << 32) | (BIGLM ? idx : key);
            if (!BIGLM || idx < cap) best64 = b64 < best64 ? b64 : best64;
          }
        }
        // The map of this frame dies here (the next expansion reads the arena, and so does the
        // trace-back): recycle the group's slots while their sectors are still in L2.  Lane b
        // takes slot b of every non-empty word — 512 contiguous bytes per word, issued by other
        // lanes than the ones that loaded the entries (a same-thread load->store on one address
        // serialises the LSU; measured 5x slower).
        unsigned nz = __ballot_sync(kFull, word != 0);
        while (nz) {
          const int k = __ffs(nz) - 1;
          nz &= nz - 1;
          const uint32_t bits = __shfl_sync(kFull, word, k);
          if ((bits >> lane) & 1u)
            __stcg(reinterpret_cast<uint4 *>(&hn[((grp * 32 + k) << 5) + lane]),
                   make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu));
        }
      }
#pragma unroll
      for (int dlt = 16; dlt > 0; dlt >>= 1) {
        const unsigned long long o = __shfl_xor_sync(kFull, best64, dlt);
        best64 = o < best64 ? o : best64;
      }
      if (lane == 0 && best64 != kInfVal) atomicMin(&s_best, best64);
    }
    __syncthreads();
    if (tid == 0) frame_commit(st, d, cfg, nc, s_alive, s_best);
    __syncthreads();
  }
}

template <bool BIGLM>
__global__ void __launch_bounds__(kStreamThreads, BIGLM ? 1 : 2)
k_post(StreamState *const *streams, FrameDesc *desc, GraphView g, DecoderConfigDev cfg, int mode, LmPair lms) {
  __shared__ PostSmem ps;
  StreamState *st = streams[blockIdx.x];
  FrameDesc *d = &desc[blockIdx.x];
  if ((mode & kModeEpi) && d->stepping) post_epilogue<BIGLM>(st, d, g, cfg, lms, ps);
  if (mode & kModePro) cutoff_prologue<kStreamThreads, BIGLM>(st, d, g, cfg, lms, ps.red64, ps.red32, ps.hist, ps.misc);
}

}  // namespace asrd

#include "asrd_stream_kernel.cuh"  // k_stream: the on-chip frame loop

namespace asrd {

// ------------------------------------------------------------------ arena compaction (prune_tokens)

constexpr uint32_t kDeadBit = 0x80000000u;  // tok_sc.x of a token an arena prune dropped (until the compaction)

// Drops the tokens marked kDeadBit from frames lo..F of the arena and rewrites frame_off.  A tile
// is read into registers before anything of it is written: a token only ever moves towards the
// front.  Whole-CTA device function.
template <int NT>
__device__ __forceinline__ void compact_arena(StreamState *st, int lo, int F, uint32_t *s_scan) {
  const int tid = threadIdx.x;
  uint32_t run = st->frame_off[lo], old_b = run;
  __syncthreads();
  for (int f = lo; f <= F; ++f) {
    const uint32_t old_e = st->frame_off[f + 1];
    const uint32_t n = old_e - old_b, new_b = run;
    __syncthreads();  // everybody has read frame_off[f + 1] (written by the next round)
    for (uint32_t i0 = 0; i0 < n; i0 += NT) {
      const uint32_t i = i0 + tid;
      uint2 sc = make_uint2(kDeadBit, 0u);
      uint32_t ex = 0;
      if (i < n) {
        sc = st->tok_sc[old_b + i];
        ex = st->tok_extra[old_b + i];
      }
      const uint32_t alive = (sc.x & kDeadBit) ? 0u : 1u;
      uint32_t total;
      const uint32_t pos = block_exclusive_scan<NT>(alive, s_scan, total);
      if (alive) {
        st->tok_sc[run + pos] = sc;
        st->tok_extra[run + pos] = ex;
      }
      run += total;
    }
    if (tid == 0) st->frame_off[f] = new_b;
    old_b = old_e;
  }
  __syncthreads();
  if (tid == 0) {
    st->frame_off[F + 1] = run;
    st->tot_pruned_tokens += old_b - run;
    st->gc_frame = F;
  }
}

// ------------------------------------------------------------------ raw lattice

// GetRawLattice (inl.h:868-975) with the pruning of FinalizeDecoding (inl.h:725-847), one CTA
// per stream.  The device keeps tokens only.  Forward links are REGENERATED here: a link
// tok -> dst exists iff tok was expanded (cost <= cur_cutoff of its frame, inl.h:315; for eps
// arcs cost < the frame's closure cutoff, inl.h:391) and the arc's cost is below the final
// cutoff of the destination frame (inl.h:330, 415) — exactly the links the canonical search
// admits.  One backward sweep over the frames computes the extra costs (inl.h:524-562; eps links
// inside a frame iterate to the exact fixed point), drops links with
// link_extra_cost > lattice_beam and tokens without surviving links, and emits the survivors.
// Two small per-frame maps (state -> cost, extra, arena index) replace the reference's pointers.
// BIGLM: several tokens of a frame may share the HCLG state; the LM pair id of every slot sits in a
// parallel array and the probe sequence starts at the hash of (state, pair).
template <bool BIGLM>
__device__ __forceinline__ bool lat_find(const LatEntry *m, const uint32_t *mp, uint32_t mask, uint32_t shift,
                                         uint32_t state, uint32_t pair, uint32_t &slot) {
  uint32_t h = BIGLM ? hash_key(((unsigned long long)pair << 32) | state, mask, shift) : hash_state(state, mask, shift);
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    const uint32_t k = __ldcg(&m[h].key);
    if (k == state && (!BIGLM || __ldcg(&mp[h]) == pair)) {
      slot = h;
      return true;
    }
    if (k == kEmptyKey) return false;
    h = (h + 1) & mask;
  }
  return false;
}

//
// PRUNE (plain decoders, device option prune_tokens): the same sweep as PruneActiveTokens
// (inl.h:438-480, called every prune_interval frames at inl.h:660-661) — the frontier frame keeps
// every token with extra cost 0, nothing is emitted, tokens whose extra cost against the current
// frontier exceeds lattice_beam are marked dead and the arena is compacted afterwards.  An extra
// cost can only grow as the frontier moves on (later frontiers are reached THROUGH this one, float
// addition and min are monotone), so a token dropped here would be dropped by the final sweep too,
// and a link to it never survives: one-best and raw lattice are bit-identical with and without.
// The two lookup maps are the halves of the stream's (idle) HBM token map; the sweep stops at the
// first frame below the previous frontier whose extra costs and survivors did not change — the
// frames below it cannot change either.
template <bool BIGLM, bool PRUNE>
__global__ void __launch_bounds__(kStreamThreads, 2)
k_lattice(StreamState *const *streams, LatticeOut *outs, GraphView g, DecoderConfigDev cfg, int use_final,
          LmPair lms, int prune_interval) {
  static_assert(!(BIGLM && PRUNE), "arena pruning: plain decoders only");
  constexpr int NT = kStreamThreads;
  __shared__ unsigned long long s_red64[NT / 32];
  __shared__ uint32_t s_scan[NT / 32 + 1];
  __shared__ uint32_t s_changed, s_frame_changed;
  StreamState *st = streams[blockIdx.x];
  LatticeOut *out = PRUNE ? nullptr : &outs[blockIdx.x];
  const int tid = threadIdx.x;
  const int F = st->frame;
  if (!PRUNE && tid == 0) {
    out->n_toks = 0;
    out->n_links = 0;
  }
  if (F < 0) return;
  uint32_t mask = st->hash_mask, shift = st->hash_shift;
  LatEntry *pmap[2] = {nullptr, nullptr};
  int lo = 0;  // PRUNE: lowest frame the sweep changed
  if (PRUNE) {
    if (st->status < 0 || F - st->gc_frame < prune_interval) return;  // uniform
    // the lookup maps are the two halves of the stream's token map (idle between frame-loop
    // launches, left clean by them and by this kernel); a frame must fit one half at load <= 0.75
    mask >>= 1;
    shift += 1;
    pmap[0] = reinterpret_cast<LatEntry *>(st->hash);
    pmap[1] = pmap[0] + (mask + 1);
    uint32_t too_big = 0;
    for (int f = tid; f <= F; f += NT) too_big |= (st->frame_off[f + 1] - st->frame_off[f]) > (mask + 1) / 4 * 3;
    if (__syncthreads_or((int)too_big)) return;  // (retried at the next call; the arena just stays larger)
    if (tid == 0) {
      const uint32_t used = st->frame_off[F + 1];
      if (used > st->peak_tokens) st->peak_tokens = used;
    }
  }
  const float beam = cfg.lattice_beam;
  uint32_t *slots[2] = {st->queue[0], st->queue[1]};  // slot of token i of the frame in map[f & 1]

  unsigned long long *pair_map = BIGLM ? st->pair_map : nullptr;
  const uint32_t pair_mask = BIGLM ? st->pair_mask : 0;
  // ---- final costs (ComputeFinalCosts, inl.h:670-720): the token on the super-final state, if any.
  // biglm (…-biglm.h:157-215): a final token's final cost is DiffArpaLm::Final of its LM state, and
  // the best cost with final is taken over ALL tokens of the frame.
  float final_best = 0.f;
  bool any_final = false;
  if (!PRUNE) {
    const uint32_t b0 = st->frame_off[F], n0 = st->frame_off[F + 1] - b0;
    unsigned long long best_all = kInfVal, best_fin = kInfVal, best_wf = kInfVal;
    for (uint32_t i = tid; i < n0; i += NT) {
      const uint2 sc = st->tok_sc[b0 + i];
      const unsigned long long b = ((unsigned long long)f2ord(__uint_as_float(sc.y)) << 32) | sc.x;
      best_all = b < best_all ? b : best_all;
      if ((int32_t)sc.x == g.final_state) best_fin = b < best_fin ? b : best_fin;
      if (BIGLM) {
        const float wf = __uint_as_float(sc.y) + lm_final(lms, pair_map, st->tok_lm[b0 + i]);
        const unsigned long long bw = (unsigned long long)f2ord(wf) << 32;
        best_wf = bw < best_wf ? bw : best_wf;
      }
    }
    best_all = block_min_u64<NT>(best_all, s_red64);
    best_fin = block_min_u64<NT>(best_fin, s_red64);
    if (BIGLM) best_wf = block_min_u64<NT>(best_wf, s_red64);
    any_final = use_final && best_fin != kInfVal;
    if (BIGLM) {
      const float wf = ord2f((uint32_t)(best_wf >> 32));
      final_best = (use_final && best_wf != kInfVal && wf != CUDART_INF_F) ? wf : ord2f((uint32_t)(best_all >> 32));
    } else {
      final_best = ord2f((uint32_t)((any_final ? best_fin : best_all) >> 32));  // inl.h:709-719
    }
  }

  auto emit_link = [&](uint32_t src_idx, uint32_t dst_idx, const int4 &arc, float graph_cost, float ac) {
    if (PRUNE) return;
    const uint32_t p = atomicAdd(&out->n_links, 1u);
    if (p < out->link_cap) {
      asrd_lat_link l;
      l.src = (int32_t)src_idx;
      l.dst = (int32_t)dst_idx;
      l.ilabel = arc.x;
      l.olabel = arc.y;
      l.graph = graph_cost;
      l.acoustic = ac;
      out->links[p] = l;
    }
  };

  for (int f = F; f >= 0; --f) {
    LatEntry *mc = PRUNE ? pmap[f & 1] : out->map[f & 1];              // this frame
    LatEntry *mn = PRUNE ? pmap[(f + 1) & 1] : out->map[(f + 1) & 1];  // frame f + 1 (complete)
    uint32_t *pc = BIGLM ? out->map_pair[f & 1] : nullptr;
    const uint32_t *pn = BIGLM ? out->map_pair[(f + 1) & 1] : nullptr;
    uint32_t *sl = slots[f & 1];
    const uint32_t b0 = st->frame_off[f], n = st->frame_off[f + 1] - b0;
    const float nc_f = st->frame_nc[f];
    const uint32_t *tlm = BIGLM ? st->tok_lm + b0 : nullptr;  // LM pair ids of the frame's tokens
    // ---- recycle the map this frame reuses (it held frame f + 2)
    if (f + 2 <= F) {
      const uint32_t n2 = st->frame_off[f + 3] - st->frame_off[f + 2];
      for (uint32_t i = tid; i < n2; i += NT) {
        LatEntry e;  // (PRUNE: the empty pattern of the token map this memory belongs to)
        e.key = kEmptyKey; e.cost_bits = PRUNE ? 0xFFFFFFFFu : 0u; e.extra_ord = PRUNE ? 0xFFFFFFFFu : kOrdInf;
        e.idx = PRUNE ? 0xFFFFFFFFu : 0u;
        mc[sl[i]] = e;
      }
    }
    if (PRUNE && tid == 0) s_frame_changed = f >= st->gc_frame ? 1u : 0u;  // frames the last prune did not settle
    __syncthreads();
    // ---- insert the frame's tokens
    for (uint32_t i = tid; i < n; i += NT) {
      const uint2 sc = st->tok_sc[b0 + i];
      float init = CUDART_INF_F;
      const uint32_t pair = BIGLM ? tlm[i] : 0u;
      if (PRUNE) {
        if (f == F) init = 0.f;  // the frontier is not pruned (inl.h:455-470 stops above it)
      } else if (f == F) {  // PruneForwardLinksFinal, inl.h:758-775, 815-816
        float fc = 0.f;
        if (any_final) fc = (int32_t)sc.x == g.final_state ? (BIGLM ? lm_final(lms, pair_map, pair) : 0.f) : CUDART_INF_F;
        init = __uint_as_float(sc.y) + fc - final_best;
        if (init > beam) init = CUDART_INF_F;
      }
      uint32_t h = BIGLM ? hash_key(((unsigned long long)pair << 32) | sc.x, mask, shift) : hash_state(sc.x, mask, shift);
      for (;;) {
        if (atomicCAS(&mc[h].key, kEmptyKey, sc.x) == kEmptyKey) break;
        h = (h + 1) & mask;
      }
      if (BIGLM) pc[h] = pair;
      mc[h].cost_bits = sc.y;
      mc[h].extra_ord = f2ord(init);
      mc[h].idx = b0 + i;
      sl[i] = h;
    }
    __syncthreads();
    // ---- emitting links into frame f + 1 (its extra costs are final)
    if (f < F) {
      const float cur_cut = st->frame_cur[f];
      const float nc_next = st->frame_nc[f + 1];
      const float *__restrict__ ll = st->ll_hist + (size_t)f * st->ll_stride;
      for (uint32_t i = tid; i < n; i += NT) {
        const uint2 sc = st->tok_sc[b0 + i];
        const float cost = __uint_as_float(sc.y);
        if (g.clg ? !(cost < cur_cut) : !(cost <= cur_cut)) continue;  // inl.h:315 (CLG: strict)
        const uint2 er = __ldg(&g.erows[sc.x]);
        const uint32_t pair = BIGLM ? tlm[i] : 0u;
        float best = CUDART_INF_F;
        for (uint32_t a = er.x; a < er.y; ++a) {
          const int4 arc = __ldg(&g.arcs[a]);
          const float ac = -__ldg(&ll[arc.x - 1]);
          float graph_cost = __int_as_float(arc.z);
          int32_t n1 = 0, n2 = 0;
          if (BIGLM && arc.y != 0) graph_cost = __int_as_float(arc.z) + lm_step(lms, pair_map, pair, arc.y, n1, n2);
          const float tot = (cost + ac) + graph_cost;  // inl.h:326-329
          if (g.clg ? !(tot <= nc_next) : !(tot < nc_next)) continue;  // inl.h:330, final cutoff (CLG: inclusive)
          const uint32_t dpair = (BIGLM && arc.y != 0) ? pair_intern(pair_map, pair_mask, n1, n2, &st->status) : pair;
          uint32_t ds;
          if (!lat_find<BIGLM>(mn, pn, mask, shift, (uint32_t)arc.w & kStateMask, dpair, ds)) continue;
          const float dextra = ord2f(__ldcg(&mn[ds].extra_ord));
          float le = dextra + (tot - __uint_as_float(__ldcg(&mn[ds].cost_bits)));  // inl.h:524-526
          if (le > beam) continue;                                                  // inl.h:532
          if (le < 0.f) le = 0.f;                                                   // inl.h:545-551
          best = fminf(best, le);
          emit_link(b0 + i, __ldcg(&mn[ds].idx), arc, graph_cost, ac);
        }
        if (best < CUDART_INF_F) atomicMin(&mc[sl[i]].extra_ord, f2ord(best));
      }
    }
    // ---- eps links inside the frame: exact fixed point of the extra costs
    for (;;) {
      __syncthreads();
      if (tid == 0) s_changed = 0;
      __syncthreads();
      for (uint32_t i = tid; i < n; i += NT) {
        const uint2 sc = st->tok_sc[b0 + i];
        const float cost = __uint_as_float(sc.y);
        if (!(cost < nc_f)) continue;  // inl.h:391
        const uint2 r = __ldg(&g.rows[sc.x]);
        if (r.y == r.x) continue;
        float mine = ord2f(__ldcg(&mc[sl[i]].extra_ord));
        const uint32_t pair = BIGLM ? tlm[i] : 0u;
        for (uint32_t a = r.x; a < r.y; ++a) {
          const int4 arc = __ldg(&g.arcs[a]);
          float graph_cost = __int_as_float(arc.z);
          int32_t n1 = 0, n2 = 0;
          if (BIGLM && arc.y != 0) graph_cost = __int_as_float(arc.z) + lm_step(lms, pair_map, pair, arc.y, n1, n2);
          const float tot = cost + graph_cost;  // inl.h:413-414
          if (!(tot < nc_f)) continue;           // inl.h:415
          const uint32_t dpair = (BIGLM && arc.y != 0) ? pair_intern(pair_map, pair_mask, n1, n2, &st->status) : pair;
          uint32_t ds;
          if (!lat_find<BIGLM>(mc, pc, mask, shift, (uint32_t)arc.w & kStateMask, dpair, ds)) continue;
          float le = ord2f(__ldcg(&mc[ds].extra_ord)) + (tot - __uint_as_float(__ldcg(&mc[ds].cost_bits)));
          if (le > beam) continue;
          if (le < 0.f) le = 0.f;
          if (le < mine) {
            mine = le;
            atomicMin(&mc[sl[i]].extra_ord, f2ord(le));
            s_changed = 1;
          }
        }
      }
      __syncthreads();
      if (!s_changed) break;
    }
    if constexpr (PRUNE) {
      // ---- mark the tokens nothing in the lattice beam hangs on (PruneTokensForFrame, inl.h:591);
      // remember the extra costs: a frame where nothing changed ends the sweep
      for (uint32_t i = tid; i < n; i += NT) {
        const uint32_t ex = __ldcg(&mc[sl[i]].extra_ord);
        if (ex == kOrdInf) {
          st->tok_sc[b0 + i].x |= kDeadBit;
          s_frame_changed = 1;
        } else if (st->tok_extra[b0 + i] != ex) {
          st->tok_extra[b0 + i] = ex;
          s_frame_changed = 1;
        }
      }
      __syncthreads();
      const uint32_t changed = s_frame_changed;
      lo = f;
      __syncthreads();  // (thread 0 resets the flag at the top of the next round)
      if (!changed) break;  // uniform
    } else {
    // ---- emit the frame's surviving eps links and tokens
    for (uint32_t i = tid; i < n; i += NT) {
      const uint2 sc = st->tok_sc[b0 + i];
      const float cost = __uint_as_float(sc.y);
      const float extra = ord2f(__ldcg(&mc[sl[i]].extra_ord));
      if (!(extra < CUDART_INF_F)) continue;  // PruneTokensForFrame, inl.h:591
      const uint32_t p = atomicAdd(&out->n_toks, 1u);
      if (p < out->tok_cap) {
        asrd_lat_token t;
        t.frame = f;
        t.state = (int32_t)sc.x;
        t.cost = cost;
        t.extra = extra;
        t.is_final = (f == F && (!any_final || (int32_t)sc.x == g.final_state)) ? 1 : 0;  // inl.h:935-951
        out->toks[p] = t;
        out->tok_arena_idx[p] = b0 + i;  // links carry arena indices; the host maps them through this
      }
      if (!(cost < nc_f)) continue;
      const uint2 r = __ldg(&g.rows[sc.x]);
      const uint32_t pair = BIGLM ? tlm[i] : 0u;
      for (uint32_t a = r.x; a < r.y; ++a) {
        const int4 arc = __ldg(&g.arcs[a]);
        float graph_cost = __int_as_float(arc.z);
        int32_t n1 = 0, n2 = 0;
        if (BIGLM && arc.y != 0) graph_cost = __int_as_float(arc.z) + lm_step(lms, pair_map, pair, arc.y, n1, n2);
        const float tot = cost + graph_cost;
        if (!(tot < nc_f)) continue;
        const uint32_t dpair = (BIGLM && arc.y != 0) ? pair_intern(pair_map, pair_mask, n1, n2, &st->status) : pair;
        uint32_t ds;
        if (!lat_find<BIGLM>(mc, pc, mask, shift, (uint32_t)arc.w & kStateMask, dpair, ds)) continue;
        const float le = ord2f(__ldcg(&mc[ds].extra_ord)) + (tot - __uint_as_float(__ldcg(&mc[ds].cost_bits)));
        if (le > beam) continue;
        emit_link(b0 + i, __ldcg(&mc[ds].idx), arc, graph_cost, 0.f);
      }
    }
    __syncthreads();
    }
  }
  if (PRUNE) {
    // ---- hand the token map back clean: the two frames still in the lookup maps are lo and lo + 1
    __syncthreads();
    for (int f = lo; f <= min(lo + 1, F); ++f) {
      const uint32_t n = st->frame_off[f + 1] - st->frame_off[f];
      const uint32_t *sl = slots[f & 1];
      for (uint32_t i = tid; i < n; i += NT) {
        LatEntry e;
        e.key = e.cost_bits = e.extra_ord = e.idx = 0xFFFFFFFFu;
        pmap[f & 1][sl[i]] = e;
      }
    }
    compact_arena<NT>(st, lo, F, s_scan);
  }
}

// ------------------------------------------------------------------ counters

__global__ void k_counters(StreamState *const *streams, int n, unsigned long long *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const StreamState *st = streams[i];
  atomicAdd(&out[0], st->tot_arcs_expanded);
  atomicAdd(&out[1], st->tot_arcs_admitted);
  atomicAdd(&out[2], (unsigned long long)st->frame_off[st->frame + 1]);
  atomicAdd(&out[3], st->tot_fallback_frames);
  atomicAdd(&out[10], st->tot_pruned_tokens);
  for (int k = 0; k < 8; ++k) atomicAdd(&out[12 + k], st->prune_cycles[k]);
  atomicMax(&out[11], (unsigned long long)max(st->peak_tokens, st->frame_off[st->frame + 1]));
  for (int k = 0; k < 6; ++k) atomicAdd(&out[4 + k], st->phase_cycles[k]);
}

// ------------------------------------------------------------------ best path

// One CTA per stream.  Picks the end token (BestPathEnd, inl.h:1096-1158: the token on the
// super-final state when use_final_probs and one is alive, else the cheapest token; ties ->
// lowest state id) and follows the winning arcs back to the start token
// (TraceBackBestPath, inl.h:1160-1200).  Arcs are emitted end -> start.
//
// For every step pred -> tok the reference reports the most recently added surviving forward
// link (inl.h:1169-1186); links are prepended in arc order (inl.h:340-341, 421-422) and excised
// when link_extra_cost > lattice_beam (inl.h:524-542), which on the best path is
// (tot' - tok.cost) > lattice_beam.  With parallel arcs pred.state -> tok.state that is the
// highest-index admitted sibling within lattice_beam, not necessarily the cheapest one.
//
// biglm: tokens are keyed by (state, LM pair); the end token adds DiffArpaLm::Final of its LM
// state (…-biglm.h:185-191) and the predecessor is the token of the source state whose LM
// transition over the arc reproduces this token's LM state and cost.
template <bool BIGLM>
__global__ void __launch_bounds__(kBestPathThreads)
k_best_path(StreamState *const *streams, GraphView g, DecoderConfigDev cfg, int use_final, int cap,
            int32_t *o_il, int32_t *o_ol, float *o_gr, float *o_ac, int32_t *o_n, int32_t *o_status,
            LmPair lms, int finalized) {
  constexpr int NT = kBestPathThreads;
  __shared__ unsigned long long s_red64[NT / 32];
  __shared__ uint32_t s_found;
  StreamState *st = streams[blockIdx.x];
  const int tid = threadIdx.x;
  const size_t ob = (size_t)blockIdx.x * cap;
  int f = st->frame;
  if (st->status < 0 || f <= 0) {  // inl.h:1104-1108
    if (tid == 0) {
      o_n[blockIdx.x] = 0;
      o_status[blockIdx.x] = st->status < 0 ? st->status : ASRD_ERR_NO_TOKENS;
    }
    return;
  }
  const unsigned long long *pair_map = BIGLM ? st->pair_map : nullptr;
  // Token of `state` in `frame`.  biglm: several tokens may share the state; take the one whose
  // expansion over `arc` lands on (want_pair, want_cost) — `ll` is the row of that frame.
  auto find_token = [&](int frame, uint32_t state, const int4 &arc, uint32_t want_pair, float want_cost,
                        const float *ll) -> uint32_t {
    if (tid == 0) s_found = 0xFFFFFFFFu;
    __syncthreads();
    const uint32_t b = st->frame_off[frame], n = st->frame_off[frame + 1] - b;
    const uint2 *__restrict__ toks = st->tok_sc + b;
    for (uint32_t i0 = 0; i0 < n; i0 += NT * 4) {  // four independent loads in flight per thread
      uint2 k[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t i = i0 + u * NT + tid;
        k[u] = i < n ? __ldg(&toks[i]) : make_uint2(0xFFFFFFFFu, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k[u].x != state) continue;
        const uint32_t idx = b + i0 + u * NT + tid;
        if (BIGLM && arc.x != -1) {
          const uint32_t pp = st->tok_lm[idx];
          const float pc = __uint_as_float(k[u].y);
          float graph_cost = __int_as_float(arc.z);
          if (arc.y != 0) {
            int32_t n1, n2;
            graph_cost = __int_as_float(arc.z) + lm_step(lms, pair_map, pp, arc.y, n1, n2);
            // must land on the token's LM state pair
            if ((((unsigned long long)(uint32_t)n1 << 32) | (uint32_t)n2) + 1ull != __ldcg(&pair_map[want_pair])) continue;
          } else if (pp != want_pair) {
            continue;
          }
          const float tot = arc.x != 0 ? (pc + (-ll[arc.x - 1])) + graph_cost : pc + graph_cost;
          if (__float_as_uint(tot) != __float_as_uint(want_cost)) continue;
        }
        atomicMin(&s_found, idx);
      }
    }
    __syncthreads();
    const uint32_t idx = s_found;
    __syncthreads();
    return idx;
  };
  // end token
  const uint32_t b0 = st->frame_off[f], n0 = st->frame_off[f + 1] - b0;
  unsigned long long best_all = kInfVal, best_fin = kInfVal, best_wf = kInfVal;
  for (uint32_t i = tid; i < n0; i += NT) {
    const uint2 sc = st->tok_sc[b0 + i];
    const uint32_t low = BIGLM ? i : sc.x;  // biglm: index inside the frame (ties -> lowest index)
    const unsigned long long bb = ((unsigned long long)f2ord(__uint_as_float(sc.y)) << 32) | low;
    best_all = bb < best_all ? bb : best_all;
    float c = __uint_as_float(sc.y);
    if (BIGLM) {
      // …-biglm.h:185-188: best_cost_with_final runs over ALL tokens
      c += lm_final(lms, pair_map, st->tok_lm[b0 + i]);
      const unsigned long long bw = ((unsigned long long)f2ord(c) << 32) | low;
      best_wf = bw < best_wf ? bw : best_wf;
    }
    if ((int32_t)sc.x == g.final_state) {
      const unsigned long long bf = ((unsigned long long)f2ord(c) << 32) | low;  // inl.h:1133-1134
      best_fin = bf < best_fin ? bf : best_fin;
    }
  }
  best_all = block_min_u64<NT>(best_all, s_red64);
  best_fin = block_min_u64<NT>(best_fin, s_red64);
  if (BIGLM) best_wf = block_min_u64<NT>(best_wf, s_red64);
  const bool any_final = use_final && best_fin != kInfVal;
  const unsigned long long pick = any_final ? best_fin : best_all;
  // extra_cost of the end token after FinalizeDecoding (inl.h:775): 0 for the plain decoder; for
  // biglm (cost + final) - final_best with final_best taken over all tokens, so it can be positive,
  // and beyond lattice_beam the token is pruned away (inl.h:815-816) and nothing is returned.
  float delta = 0.f;
  if (BIGLM && finalized && pick != kInfVal) {
    const float with_final = any_final ? ord2f((uint32_t)(pick >> 32))
                                       : __uint_as_float(st->tok_sc[b0 + (uint32_t)pick].y);
    delta = with_final - ord2f((uint32_t)(best_wf >> 32));
  }
  if (pick == kInfVal || delta > cfg.lattice_beam) {
    if (tid == 0) {
      o_n[blockIdx.x] = 0;
      o_status[blockIdx.x] = ASRD_ERR_NO_TOKENS;
    }
    return;
  }
  const int4 no_arc = make_int4(-1, 0, 0, 0);
  uint32_t idx = BIGLM ? b0 + (uint32_t)pick : find_token(f, (uint32_t)pick, no_arc, 0, 0.f, nullptr);
  int n_out = 0;
  int status = idx == 0xFFFFFFFFu ? ASRD_ERR_STATE : ASRD_OK;
  while (status == ASRD_OK) {
    const uint2 sc = st->tok_sc[idx];
    const uint32_t state = sc.x;
    const float cost = __uint_as_float(sc.y);
    const uint32_t pair = BIGLM ? st->tok_lm[idx] : 0u;
    uint32_t rep = st->tok_arc[idx];
    if (n_out >= cap) {
      status = ASRD_ERR_PATH_OVERFLOW;
      break;
    }
    if (rep == kNoArc) {  // start token: label-free arc with unit weight (inl.h:1193-1198)
      if (tid == 0) {
        o_il[ob + n_out] = 0;
        o_ol[ob + n_out] = 0;
        o_gr[ob + n_out] = 0.f;
        o_ac[ob + n_out] = 0.f;
      }
      ++n_out;
      break;
    }
    int4 arc = __ldg(&g.arcs[rep]);
    const bool emitting = arc.x != 0;
    const uint32_t src = __ldg(&g.arc_src[rep]);
    const int fp = emitting ? f - 1 : f;
    if (fp < 0) {
      status = ASRD_ERR_STATE;
      break;
    }
    const float *__restrict__ ll = emitting ? st->ll_hist + (size_t)fp * st->ll_stride : nullptr;
    const uint32_t pidx = find_token(fp, src, arc, pair, cost, ll);
    if (pidx == 0xFFFFFFFFu) {
      status = ASRD_ERR_STATE;  // broken back-trace: cannot happen unless the arena overflowed
      break;
    }
    const float pc = __uint_as_float(st->tok_sc[pidx].y);
    const uint32_t ppair = BIGLM ? st->tok_lm[pidx] : 0u;
    float graph_cost = __int_as_float(arc.z);
    if (BIGLM && arc.y != 0) {
      int32_t n1, n2;
      graph_cost = __int_as_float(arc.z) + lm_step(lms, pair_map, ppair, arc.y, n1, n2);
    }
    if (par_bit(g.par_bits, rep)) {
      const float nc = st->frame_nc[f];
      const uint32_t hi = emitting ? __ldg(&g.erows[src]).y : __ldg(&g.rows[src]).y;
      for (uint32_t a2 = hi; a2-- > rep + 1;) {
        const int4 arc2 = __ldg(&g.arcs[a2]);
        if (((uint32_t)arc2.w & kStateMask) != state) continue;
        float gc2 = __int_as_float(arc2.z);
        if (BIGLM) {  // a sibling link reaches the SAME token only if it lands on the same LM state
          if ((arc2.y != 0) != (arc.y != 0)) continue;
          if (arc2.y != 0) {
            int32_t n1, n2;
            gc2 = __int_as_float(arc2.z) + lm_step(lms, pair_map, ppair, arc2.y, n1, n2);
            if ((((unsigned long long)(uint32_t)n1 << 32) | (uint32_t)n2) + 1ull != __ldcg(&pair_map[pair])) continue;
          }
        }
        float tot2;
        bool admitted;
        if (emitting) {
          tot2 = (pc + (-ll[arc2.x - 1])) + gc2;  // inl.h:326-329
          admitted = tot2 < nc;                    // inl.h:330
        } else {
          tot2 = pc + gc2;                         // inl.h:413-414
          admitted = pc < nc && tot2 < nc;         // inl.h:391,415
        }
        if (admitted && !((delta + (tot2 - cost)) > cfg.lattice_beam)) {  // inl.h:524-532
          rep = a2;
          arc = arc2;
          graph_cost = gc2;
          break;
        }
      }
    }
    if (tid == 0) {
      o_il[ob + n_out] = arc.x;
      o_ol[ob + n_out] = arc.y;
      o_gr[ob + n_out] = graph_cost;
      o_ac[ob + n_out] = emitting ? -ll[arc.x - 1] : 0.f;
    }
    ++n_out;
    f = fp;
    idx = pidx;
  }
  if (tid == 0) {
    o_n[blockIdx.x] = n_out;
    o_status[blockIdx.x] = status;
  }
}


// ------------------------------------------------------------------ best path, plain decoders

// Plain decoders keep {state, cost} per token and nothing else: the frame loop never records which
// arc set a token's cost.  The trace-back recovers it for the tokens of the best path only.  For
// token (frame f, state d, cost c) the winning arc is the LOWEST-INDEX arc a : s -> d for which a
// token p of state s exists with
//     emitting a:  p in frame f-1, p.cost <= cur_cutoff(f-1) (inl.h:315), (p.cost + ac) + w == c  (inl.h:326-329)
//     eps a:       p in frame f,   p.cost <  next_cutoff(f)  (inl.h:391),  p.cost + w == c        (inl.h:413-414)
// bit for bit — exactly the relaxation (lowest cost, then lowest arc index) the search performs with
// the arc packed next to the cost.  (An eps relaxation made while p's cost was not yet final gives
// tot' >= the one with the final cost, so it cannot be the only witness of c.)  Candidates come
// from the graph's incoming-arc index; their source states go into a small shared-memory hash and
// the tokens of frames f and f-1 are streamed past it once.
constexpr int kRevCand = 1024;  // incoming arcs examined per round
constexpr int kRevHash = 2048;

__global__ void __launch_bounds__(kBestPathThreads)
k_best_path_rev(StreamState *const *streams, GraphView g, DecoderConfigDev cfg, int use_final, int cap,
                int32_t *o_il, int32_t *o_ol, float *o_gr, float *o_ac, int32_t *o_n, int32_t *o_status) {
  constexpr int NT = kBestPathThreads;
  __shared__ unsigned long long s_red64[NT / 32];
  __shared__ unsigned long long s_win;  // (arc id << 32) | arena index of the predecessor
  __shared__ uint32_t s_found, s_classes;
  __shared__ uint32_t s_csrc[kRevCand], s_carc[kRevCand];
  __shared__ int32_t s_cnext[kRevCand], s_head[kRevHash];
  StreamState *st = streams[blockIdx.x];
  const int tid = threadIdx.x;
  const size_t ob = (size_t)blockIdx.x * cap;
  int f = st->frame;
  if (st->status < 0 || f <= 0) {  // inl.h:1104-1108
    if (tid == 0) {
      o_n[blockIdx.x] = 0;
      o_status[blockIdx.x] = st->status < 0 ? st->status : ASRD_ERR_NO_TOKENS;
    }
    return;
  }
  // ---- end token (BestPathEnd, inl.h:1096-1158): the token on the super-final state when
  // use_final_probs and one is alive, else the cheapest; ties -> lowest state id
  const uint32_t b0 = st->frame_off[f], n0 = st->frame_off[f + 1] - b0;
  unsigned long long best_all = kInfVal, best_fin = kInfVal;
  for (uint32_t i = tid; i < n0; i += NT) {
    const uint2 sc = st->tok_sc[b0 + i];
    const unsigned long long bb = ((unsigned long long)f2ord(__uint_as_float(sc.y)) << 32) | sc.x;
    best_all = bb < best_all ? bb : best_all;
    if ((int32_t)sc.x == g.final_state) best_fin = bb < best_fin ? bb : best_fin;
  }
  best_all = block_min_u64<NT>(best_all, s_red64);
  best_fin = block_min_u64<NT>(best_fin, s_red64);
  const unsigned long long pick = (use_final && best_fin != kInfVal) ? best_fin : best_all;
  if (pick == kInfVal) {
    if (tid == 0) {
      o_n[blockIdx.x] = 0;
      o_status[blockIdx.x] = ASRD_ERR_NO_TOKENS;
    }
    return;
  }
  if (tid == 0) s_found = 0xFFFFFFFFu;
  __syncthreads();
  for (uint32_t i = tid; i < n0; i += NT)
    if (st->tok_sc[b0 + i].x == (uint32_t)pick) atomicMin(&s_found, b0 + i);
  __syncthreads();
  uint32_t idx = s_found;
  int n_out = 0;
  int status = idx == 0xFFFFFFFFu ? ASRD_ERR_STATE : ASRD_OK;
  while (status == ASRD_OK) {
    const uint2 sc = st->tok_sc[idx];
    const uint32_t state = sc.x;
    const float cost = __uint_as_float(sc.y);
    if (n_out >= cap) {
      status = ASRD_ERR_PATH_OVERFLOW;
      break;
    }
    if (f == 0 && (int32_t)state == g.start && sc.y == 0u) {
      // start token: label-free arc with unit weight (inl.h:1193-1198)
      if (tid == 0) {
        o_il[ob + n_out] = 0;
        o_ol[ob + n_out] = 0;
        o_gr[ob + n_out] = 0.f;
        o_ac[ob + n_out] = 0.f;
      }
      ++n_out;
      break;
    }
    const float nc_f = st->frame_nc[f];
    const float cur_prev = f > 0 ? st->frame_cur[f - 1] : 0.f;
    const float *__restrict__ ll_prev = f > 0 ? st->ll_hist + (size_t)(f - 1) * st->ll_stride : nullptr;
    const uint32_t ib = __ldg(&g.in_off[state]), ie = __ldg(&g.in_off[state + 1]);
    if (tid == 0) s_win = kInfVal;
    for (uint32_t r0 = ib; r0 < ie; r0 += kRevCand) {
      const uint32_t ncand = min((uint32_t)kRevCand, ie - r0);
      for (int i = tid; i < kRevHash; i += NT) s_head[i] = -1;
      if (tid == 0) s_classes = 0;
      __syncthreads();
      for (uint32_t i = tid; i < ncand; i += NT) {
        const uint32_t a = __ldg(&g.in_arc[r0 + i]);
        const uint32_t src = __ldg(&g.arc_src[a]);
        const bool eps = __ldg(&g.arcs[a]).x == 0;
        if (!eps && f == 0) {  // no frame before the first
          s_csrc[i] = 0xFFFFFFFFu;
          continue;
        }
        s_csrc[i] = src;
        s_carc[i] = a;
        s_cnext[i] = atomicExch(&s_head[(src * 0x9E3779B1u) >> 21], (int32_t)i);
        atomicOr(&s_classes, eps ? 1u : 2u);
      }
      __syncthreads();
      const uint32_t classes = s_classes;
      for (int cls = 0; cls < 2; ++cls) {  // 0: eps candidates against frame f, 1: emitting against f - 1
        if (!(classes & (1u << cls))) continue;
        const int fr = cls == 0 ? f : f - 1;
        const uint32_t tb = st->frame_off[fr], tn = st->frame_off[fr + 1] - tb;
        const uint2 *__restrict__ toks = st->tok_sc + tb;
        for (uint32_t i0 = 0; i0 < tn; i0 += NT * 4) {  // four independent loads in flight per thread
          uint2 k[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * NT + tid;
            k[u] = i < tn ? __ldg(&toks[i]) : make_uint2(0xFFFFFFFFu, 0u);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (k[u].x == 0xFFFFFFFFu) continue;
            for (int32_t j = s_head[(k[u].x * 0x9E3779B1u) >> 21]; j >= 0; j = s_cnext[j]) {
              if (s_csrc[j] != k[u].x) continue;
              const uint32_t a = s_carc[j];
              const int4 arc = __ldg(&g.arcs[a]);
              if ((arc.x == 0) != (cls == 0)) continue;
              const float pc = __uint_as_float(k[u].y);
              float tot;
              if (cls == 0) {
                if (!(pc < nc_f)) continue;              // inl.h:391
                tot = pc + __int_as_float(arc.z);        // inl.h:413-414
              } else {
                if (g.clg ? !(pc < cur_prev) : !(pc <= cur_prev)) continue;  // inl.h:315 (CLG: strict)
                tot = (pc + (-ll_prev[arc.x - 1])) + __int_as_float(arc.z);  // inl.h:326-329
              }
              if (__float_as_uint(tot) != sc.y) continue;
              atomicMin(&s_win, ((unsigned long long)a << 32) | (tb + i0 + u * NT + tid));
            }
          }
        }
      }
      __syncthreads();
    }
    __syncthreads();
    const unsigned long long win = s_win;
    __syncthreads();
    if (win == kInfVal) {
      status = ASRD_ERR_STATE;  // broken back-trace: cannot happen unless the arena overflowed
      break;
    }
    uint32_t rep = (uint32_t)(win >> 32);
    const uint32_t pidx = (uint32_t)win;
    int4 arc = __ldg(&g.arcs[rep]);
    const bool emitting = arc.x != 0;
    const int fp = emitting ? f - 1 : f;
    const float pc = __uint_as_float(st->tok_sc[pidx].y);
    // For every step pred -> tok the reference reports the most recently added surviving forward
    // link (inl.h:1169-1186): with parallel arcs pred.state -> tok.state that is the highest-index
    // admitted sibling within lattice_beam, not necessarily the cheapest one (see k_best_path).
    if (par_bit(g.par_bits, rep)) {
      const uint32_t src = __ldg(&g.arc_src[rep]);
      const uint32_t hi = emitting ? __ldg(&g.erows[src]).y : __ldg(&g.rows[src]).y;
      for (uint32_t a2 = hi; a2-- > rep + 1;) {
        const int4 arc2 = __ldg(&g.arcs[a2]);
        if (((uint32_t)arc2.w & kStateMask) != state) continue;
        float tot2;
        bool admitted;
        if (emitting) {
          tot2 = (pc + (-ll_prev[arc2.x - 1])) + __int_as_float(arc2.z);  // inl.h:326-329
          admitted = g.clg ? tot2 <= nc_f : tot2 < nc_f;                   // inl.h:330 (CLG: inclusive)
        } else {
          tot2 = pc + __int_as_float(arc2.z);                              // inl.h:413-414
          admitted = pc < nc_f && tot2 < nc_f;                             // inl.h:391,415
        }
        if (admitted && !((tot2 - cost) > cfg.lattice_beam)) {  // inl.h:524-532
          rep = a2;
          arc = arc2;
          break;
        }
      }
    }
    if (tid == 0) {
      o_il[ob + n_out] = arc.x;
      o_ol[ob + n_out] = arc.y;
      o_gr[ob + n_out] = __int_as_float(arc.z);
      o_ac[ob + n_out] = emitting ? -ll_prev[arc.x - 1] : 0.f;
    }
    ++n_out;
    f = fp;
    idx = pidx;
  }
  if (tid == 0) {
    o_n[blockIdx.x] = n_out;
    o_status[blockIdx.x] = status;
  }
}

}  // namespace asrd

#include "asrd_prune_kernel.cuh"  // k_prune: the arena prune with the lookup map in shared memory
