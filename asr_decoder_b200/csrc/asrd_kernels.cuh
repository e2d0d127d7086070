// Hand-written sm_100a kernels of the WFST token-passing beam search.
//
// Per decoded frame two launches serve ALL streams of a batch:
//   k_expand    — load-balanced emitting-arc expansion (ProcessEmitting's hot loop,
//                 reference src/my-decoder/online-decoder-base-inl.h:311-347)
//   k_boundary  — one CTA per stream: eps closure (ProcessNonemitting, inl.h:353-431),
//                 survivor compaction into the token arena, sibling-link resolution for the
//                 trace-back (inl.h:1169-1186 + the lattice-beam link pruning of
//                 inl.h:524-542), GetCutoff for the next frame (inl.h:138-234) and the
//                 best-token pre-pass (inl.h:282-300).
// k_best_path walks the back-trace (BestPathEnd / TraceBackBestPath, inl.h:1096-1200).
#pragma once

#include <math_constants.h>

#include "asrd_internal.cuh"

namespace asrd {

// ------------------------------------------------------------------ small helpers

__device__ __forceinline__ uint32_t hash_state(uint32_t s, uint32_t mask, uint32_t shift) {
  // Locality preserving: neighbouring states (the common case in HCLG rows) land in
  // neighbouring slots, i.e. the same 32-byte sectors; the high bits decorrelate aliases.
  return (s + (s >> (32u - shift)) * 0x9E3779B1u) & mask;
}

__device__ __forceinline__ unsigned long long pack_val(float cost, uint32_t arc) {
  return ((unsigned long long)f2ord(cost) << 32) | arc;
}

__device__ __forceinline__ bool par_bit(const uint32_t *bits, uint32_t arc) {
  return (__ldg(&bits[arc >> 5]) >> (arc & 31u)) & 1u;
}

// Find-or-claim `state` and recombine with atomicMin (FindOrAddToken, inl.h:88-136).
__device__ __forceinline__ bool hash_insert(HashEntry *tab, uint32_t mask, uint32_t shift,
                                            uint32_t state, unsigned long long packed,
                                            uint32_t &slot, bool &is_new, unsigned long long &old) {
  uint32_t h = hash_state(state, mask, shift);
  is_new = false;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    uint32_t k = __ldcg(&tab[h].key);
    if (k == kEmptyKey) {
      k = atomicCAS(&tab[h].key, kEmptyKey, state);
      if (k == kEmptyKey) {
        is_new = true;
        k = state;
      }
    }
    if (k == state) {
      old = atomicMin(&tab[h].val, packed);
      slot = h;
      return true;
    }
    h = (h + 1) & mask;
  }
  return false;
}

__device__ __forceinline__ bool hash_find(const HashEntry *tab, uint32_t mask, uint32_t shift,
                                          uint32_t state, unsigned long long &val) {
  uint32_t h = hash_state(state, mask, shift);
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    uint32_t k = __ldcg(&tab[h].key);
    if (k == state) {
      val = __ldcg(&tab[h].val);
      return true;
    }
    if (k == kEmptyKey) return false;
    h = (h + 1) & mask;
  }
  return false;
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

// ------------------------------------------------------------------ begin-advance

__global__ void k_begin_advance(StreamState *const *streams, const AdvanceParams *params, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  StreamState *st = streams[i];
  AdvanceParams p = params[i];
  st->ll_base = p.ll;
  st->ll_stride = p.stride;
  st->ll_frame0 = st->frame;
  int target = st->frame + p.n_frames;
  if (target > st->max_frames) {
    target = st->max_frames;
    atomicMin(&st->status, ASRD_ERR_FRAMES_OVERFLOW);
  }
  st->target_frame = target;
}

// ------------------------------------------------------------------ expand

// Grid: persistent, a multiple of the SM count.  Every CTA scans the per-stream tile
// counts into shared memory, then takes tiles round-robin.  A tile = kTileTokens tokens of
// one stream; their emitting-arc spans are flattened through a shared-memory prefix so
// that consecutive threads fetch consecutive 16-byte arc records (LDG.128).
__global__ void __launch_bounds__(kExpandThreads)
k_expand(StreamState *const *streams, int n_streams, GraphView g) {
  extern __shared__ uint32_t s_dyn[];        // [n_streams + 1] tile prefix
  __shared__ uint32_t s_warp[kExpandThreads / 32 + 1];
  __shared__ uint32_t s_off[kTileTokens + 1];
  __shared__ uint32_t s_base[kTileTokens];
  __shared__ float s_cost[kTileTokens];
  __shared__ uint32_t s_cnt[2];

  const int tid = threadIdx.x;
  const int lane = tid & 31;

  // ---- tile prefix over streams
  {
    const int per = (n_streams + kExpandThreads - 1) / kExpandThreads;
    const int b = tid * per;
    uint32_t sum = 0;
    for (int s = b; s < b + per && s < n_streams; ++s) sum += streams[s]->tiles;
    uint32_t total;
    uint32_t excl = block_exclusive_scan<kExpandThreads>(sum, s_warp, total);
    for (int s = b; s < b + per && s < n_streams; ++s) {
      s_dyn[s] = excl;
      excl += streams[s]->tiles;
    }
    if (tid == 0) s_dyn[n_streams] = total;
    __syncthreads();
  }
  const uint32_t total_tiles = s_dyn[n_streams];

  for (uint32_t gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
    // stream of this tile: last s with prefix[s] <= gt
    int lo = 0, hi = n_streams;  // invariant: prefix[lo] <= gt < prefix[hi]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (s_dyn[mid] <= gt) lo = mid; else hi = mid;
    }
    StreamState *st = streams[lo];
    const uint32_t tile = gt - s_dyn[lo];
    const int frame = st->frame;
    const uint32_t n_cur = st->n_cur;
    const float cur_cut = st->cur_cut;
    const float abeam = st->abeam;
    const uint32_t tok_base = st->frame_off[frame];
    const float *__restrict__ ll = st->ll_base + (size_t)(frame - st->ll_frame0) * st->ll_stride;
    HashEntry *hn = st->hash[(frame + 1) & 1];
    uint32_t *slots = st->slots[(frame + 1) & 1];
    uint32_t *n_slots = &st->n_slots[(frame + 1) & 1];
    const uint32_t mask = st->hash_mask, shift = st->hash_shift;
    uint32_t *next_cut = &st->next_cut_bits;

    // ---- stage the tile: per-token emitting span
    const uint32_t i = tile * kTileTokens + tid;
    uint32_t deg = 0, base = 0;
    float cost = 0.f;
    if (i < n_cur) {
      uint2 sc = st->tok_sc[tok_base + i];
      cost = __uint_as_float(sc.y);
      if (cost <= cur_cut) {  // inclusive, inl.h:315
        uint2 r0 = __ldg(&g.rows[sc.x]);
        uint32_t end = __ldg(&g.rows[sc.x + 1]).x;
        base = r0.y;
        deg = end - r0.y;
      }
    }
    uint32_t total;
    uint32_t off = block_exclusive_scan<kExpandThreads>(deg, s_warp, total);
    s_off[tid] = off;
    s_base[tid] = base;
    s_cost[tid] = cost;
    if (tid == 0) {
      s_off[kTileTokens] = total;
      s_cnt[0] = 0;
    }
    __syncthreads();

    uint32_t admitted = 0;
    for (uint32_t jb = 0; jb < total; jb += kExpandThreads) {
      const uint32_t j = jb + tid;
      const bool in = j < total;
      bool need_cut = false, is_new = false;
      uint32_t cand_bits = 0xFFFFFFFFu, slot = 0;
      if (in) {
        // owner token: last t with s_off[t] <= j
        int l = 0, h = kTileTokens;
        while (h - l > 1) {
          int m = (l + h) >> 1;
          if (s_off[m] <= j) l = m; else h = m;
        }
        const uint32_t a = s_base[l] + (j - s_off[l]);
        const int4 arc = __ldg(&g.arcs[a]);
        const float ac = -__ldg(&ll[arc.x - 1]);
        const float tot = (s_cost[l] + ac) + __int_as_float(arc.z);  // inl.h:326-329
        const float nc = ord2f(*(volatile uint32_t *)next_cut);
        if (tot < nc) {  // inl.h:330 (running cutoff; the boundary kernel applies the final one)
          const float cand = tot + abeam;
          if (cand < nc) {
            need_cut = true;
            cand_bits = f2ord(cand);
          }
          unsigned long long old;
          if (!hash_insert(hn, mask, shift, (uint32_t)arc.w, pack_val(tot, a), slot, is_new, old))
            atomicMin(&st->status, ASRD_ERR_HASH_OVERFLOW);
          ++admitted;
        }
      }
      // warp-aggregated cutoff tightening (inl.h:332-333)
      if (__any_sync(0xFFFFFFFFu, need_cut)) {
        uint32_t wmin = __reduce_min_sync(0xFFFFFFFFu, cand_bits);
        if (lane == 0) atomicMin(next_cut, wmin);
      }
      // warp-aggregated append of newly claimed slots
      const unsigned newm = __ballot_sync(0xFFFFFFFFu, is_new);
      if (newm) {
        uint32_t pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(n_slots, __popc(newm));
        pos0 = __shfl_sync(0xFFFFFFFFu, pos0, 0);
        if (is_new) slots[pos0 + __popc(newm & ((1u << lane) - 1u))] = slot;
      }
    }
    admitted = __reduce_add_sync(0xFFFFFFFFu, admitted);
    if (lane == 0 && admitted) atomicAdd(&s_cnt[0], admitted);
    __syncthreads();
    if (tid == 0) {
      atomicAdd(&st->arcs_expanded, total);
      atomicAdd(&st->arcs_admitted, s_cnt[0]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ frame boundary

enum { kModeInit = 1, kModeEpi = 2, kModePro = 4 };

// exact k-th smallest (0-based) cost among the tokens [base, base+n) restricted to keys
// below `limit_ord` (exclusive; 0xFFFFFFFF = no limit): what std::nth_element yields in
// GetCutoff (inl.h:190-193, 211-216).  MSB-first radix select over (ordered key - min key).
template <int NT>
__device__ float block_kth_smallest(const uint2 *tok_sc, uint32_t n, uint32_t k, uint32_t min_ord,
                                    uint32_t limit_ord, uint32_t *s_hist, uint32_t *s_misc) {
  const int tid = threadIdx.x;
  // highest relative key
  uint32_t kmax = 0;
  for (uint32_t i = tid; i < n; i += NT) {
    uint32_t key = f2ord(__uint_as_float(tok_sc[i].y));
    if (key < limit_ord) kmax = max(kmax, key - min_ord);
  }
  kmax = __reduce_max_sync(0xFFFFFFFFu, kmax);
  if (tid == 0) s_misc[0] = 0;
  __syncthreads();
  if ((tid & 31) == 0) atomicMax(&s_misc[0], kmax);
  __syncthreads();
  kmax = s_misc[0];
  __syncthreads();
  int top = 24;
  while (top > 0 && (kmax >> top) == 0) top -= 8;
  uint32_t prefix = 0, pmask = 0, kk = k;
  for (int sh = top; sh >= 0; sh -= 8) {
    for (int b = tid; b < 256; b += NT) s_hist[b] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += NT) {
      uint32_t key = f2ord(__uint_as_float(tok_sc[i].y));
      if (key < limit_ord) {
        uint32_t rk = key - min_ord;
        if ((rk & pmask) == prefix) atomicAdd(&s_hist[(rk >> sh) & 255u], 1u);
      }
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      int b = 0;
      for (; b < 255; ++b) {
        if (acc + s_hist[b] > kk) break;
        acc += s_hist[b];
      }
      s_misc[0] = (uint32_t)b;
      s_misc[1] = kk - acc;
    }
    __syncthreads();
    prefix |= s_misc[0] << sh;
    pmask |= 255u << sh;
    kk = s_misc[1];
    __syncthreads();
  }
  return ord2f(prefix + min_ord);
}

__global__ void __launch_bounds__(kBoundaryThreads)
k_boundary(StreamState *const *streams, GraphView g, DecoderConfigDev cfg, int mode) {
  constexpr int NT = kBoundaryThreads;
  __shared__ uint32_t s_warp[NT / 32 + 1];
  __shared__ unsigned long long s_red64[NT / 32];
  __shared__ uint32_t s_red32[NT / 32];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_misc[4];
  __shared__ uint32_t s_nslots;
  __shared__ uint32_t s_qn[2];

  StreamState *st = streams[blockIdx.x];
  const int tid = threadIdx.x;
  const bool is_init = (mode & kModeInit) != 0;
  const uint32_t mask = st->hash_mask, shift = st->hash_shift;

  // =============================== epilogue of the frame just expanded (or InitDecoding)
  const bool stepping = is_init || ((mode & kModeEpi) && st->frame < st->target_frame);
  if (stepping) {
    const int t = is_init ? -1 : st->frame;  // tokens of frame t+1 are being completed
    const int next = (t + 1) & 1, cur = t & 1;
    HashEntry *hn = st->hash[next];
    HashEntry *hc = st->hash[cur];
    uint32_t *slots = st->slots[next];
    const float *__restrict__ ll =
        is_init ? nullptr : st->ll_base + (size_t)(t - st->ll_frame0) * st->ll_stride;
    float nc;
    if (is_init) {
      // InitDecoding (inl.h:41-67): forget the previous utterance, seed the start token,
      // closure with cutoff = beam.
      for (int w = 0; w < 2; ++w) {
        HashEntry *h = st->hash[w];
        const uint32_t *sl = st->slots[w];
        const uint32_t ns = st->n_slots[w];
        for (uint32_t i = tid; i < ns; i += NT) {
          HashEntry e;
          e.key = kEmptyKey; e.aux = 0; e.val = kInfVal;
          h[sl[i]] = e;
        }
      }
      __syncthreads();
      nc = cfg.beam;
      if (tid == 0) {
        st->n_slots[0] = st->n_slots[1] = 0;
        st->status = 0;
        st->finalized = 0;
        st->target_frame = 0;
        st->frame_off[0] = 0;
        uint32_t slot; bool is_new; unsigned long long old;
        hash_insert(hn, mask, shift, (uint32_t)g.start, pack_val(0.0f, kNoArc), slot, is_new, old);
        slots[0] = slot;
        s_nslots = 1;
      }
    } else {
      nc = ord2f(st->next_cut_bits);  // the FINAL next_cutoff of this frame
      if (tid == 0) s_nslots = st->n_slots[next];
    }
    if (tid == 0) s_qn[0] = s_qn[1] = 0;
    __syncthreads();

    // ---- eps closure (ProcessNonemitting, inl.h:353-431) as frontier rounds
    {
      const uint32_t n0 = s_nslots;
      for (uint32_t i = tid; i < n0; i += NT) {  // inl.h:376-381
        const uint32_t slot = slots[i];
        const uint32_t key = __ldcg(&hn[slot].key);
        const float cost = ord2f((uint32_t)(__ldcg(&hn[slot].val) >> 32));
        const uint2 r = __ldg(&g.rows[key]);
        if (cost < nc && r.y > r.x) {
          hn[slot].aux = 1;
          st->queue[1][atomicAdd(&s_qn[1], 1u)] = slot;
        }
      }
      for (uint32_t round = 1;; ++round) {
        __syncthreads();
        const uint32_t nq = s_qn[round & 1];
        if (nq == 0) break;
        __syncthreads();
        if (tid == 0) s_qn[(round + 1) & 1] = 0;
        __syncthreads();
        const uint32_t *qin = st->queue[round & 1];
        uint32_t *qout = st->queue[(round + 1) & 1];
        for (uint32_t i = tid; i < nq; i += NT) {
          const uint32_t slot = qin[i];
          const uint32_t state = __ldcg(&hn[slot].key);
          const float cost = ord2f((uint32_t)(__ldcg(&hn[slot].val) >> 32));
          if (!(cost < nc)) continue;  // inl.h:391
          const uint2 r = __ldg(&g.rows[state]);
          for (uint32_t a = r.x; a < r.y; ++a) {
            const int4 arc = __ldg(&g.arcs[a]);
            const float tot = cost + __int_as_float(arc.z);  // inl.h:413-414
            if (tot < nc) {                                   // inl.h:415
              uint32_t slot2; bool is_new; unsigned long long old;
              const unsigned long long pk = pack_val(tot, a);
              if (!hash_insert(hn, mask, shift, (uint32_t)arc.w, pk, slot2, is_new, old)) {
                atomicMin(&st->status, ASRD_ERR_HASH_OVERFLOW);
                continue;
              }
              if (is_new) slots[atomicAdd(&s_nslots, 1u)] = slot2;
              const bool changed = (uint32_t)(pk >> 32) < (uint32_t)(old >> 32);  // inl.h:115-127
              if (changed) {
                const uint2 r2 = __ldg(&g.rows[arc.w]);
                if (r2.y > r2.x && atomicExch(&hn[slot2].aux, round + 1) != round + 1)
                  qout[atomicAdd(&s_qn[(round + 1) & 1], 1u)] = slot2;  // inl.h:425-426
              }
            }
          }
        }
      }
    }
    __syncthreads();

    // ---- survivors (cost < final cutoff) -> token arena, in slot-list order
    const uint32_t ntot = s_nslots;
    const uint32_t out_base = st->frame_off[t + 1];
    const uint32_t cap = st->token_capacity;
    uint32_t running = 0;
    unsigned long long best64 = kInfVal;
    for (uint32_t c = 0; c < ntot; c += NT) {
      const uint32_t i = c + tid;
      bool alive = false;
      uint32_t key = 0;
      unsigned long long val = kInfVal;
      float cost = 0.f;
      if (i < ntot) {
        const uint32_t slot = slots[i];
        key = __ldcg(&hn[slot].key);
        val = __ldcg(&hn[slot].val);
        cost = ord2f((uint32_t)(val >> 32));
        alive = cost < nc;
      }
      uint32_t tot_c;
      const uint32_t pos = block_exclusive_scan<NT>(alive ? 1u : 0u, s_warp, tot_c);
      const uint32_t idx = out_base + running + pos;
      if (alive && idx < cap) {
        uint32_t rep = (uint32_t)val;
        float ac = 0.f;
        if (rep != kNoArc) {
          const int4 arc = __ldg(&g.arcs[rep]);
          const bool emitting = arc.x != 0;
          if (emitting) ac = -__ldg(&ll[arc.x - 1]);
          if (par_bit(g.par_bits, rep)) {
            // The reference reports, for the step pred -> tok, the most recently added
            // surviving forward link (inl.h:1169-1186); links are prepended in arc order
            // (inl.h:340-341) and excised when link_extra_cost > lattice_beam (inl.h:524-542),
            // which for a best-path token is (tot' - tok.cost) > lattice_beam.
            const uint32_t src = __ldg(&g.arc_src[rep]);
            const uint2 r = __ldg(&g.rows[src]);
            unsigned long long pv;
            if (emitting) {
              const uint32_t end = __ldg(&g.rows[src + 1]).x;
              if (hash_find(hc, mask, shift, src, pv)) {
                const float pc = ord2f((uint32_t)(pv >> 32));
                for (uint32_t a2 = end; a2-- > rep + 1;) {
                  const int4 arc2 = __ldg(&g.arcs[a2]);
                  if ((uint32_t)arc2.w != key) continue;
                  const float ac2 = -__ldg(&ll[arc2.x - 1]);
                  const float tot2 = (pc + ac2) + __int_as_float(arc2.z);
                  if (tot2 < nc && !((tot2 - cost) > cfg.lattice_beam)) {
                    rep = a2;
                    ac = ac2;
                    break;
                  }
                }
              }
            } else {
              if (hash_find(hn, mask, shift, src, pv)) {
                const float pc = ord2f((uint32_t)(pv >> 32));
                for (uint32_t a2 = r.y; a2-- > rep + 1;) {
                  const int4 arc2 = __ldg(&g.arcs[a2]);
                  if ((uint32_t)arc2.w != key) continue;
                  const float tot2 = pc + __int_as_float(arc2.z);
                  if (pc < nc && tot2 < nc && !((tot2 - cost) > cfg.lattice_beam)) {
                    rep = a2;
                    break;
                  }
                }
              }
            }
          }
        }
        st->tok_sc[idx] = make_uint2(key, __float_as_uint(cost));
        st->tok_aa[idx] = make_uint2(rep, __float_as_uint(ac));
        const unsigned long long b = ((unsigned long long)f2ord(cost) << 32) | key;
        best64 = b < best64 ? b : best64;
      }
      running += tot_c;
    }
    uint32_t n_alive = running;
    if (out_base + n_alive > cap) {
      n_alive = cap > out_base ? cap - out_base : 0;
      if (tid == 0) atomicMin(&st->status, ASRD_ERR_ARENA_OVERFLOW);
    }
    // ---- recycle the map of the previous frame (after every sibling look-up is done)
    __syncthreads();
    if (!is_init) {
      const uint32_t *sl = st->slots[cur];
      const uint32_t ns = st->n_slots[cur];
      for (uint32_t i = tid; i < ns; i += NT) {
        HashEntry e;
        e.key = kEmptyKey; e.aux = 0; e.val = kInfVal;
        hc[sl[i]] = e;
      }
    }
    best64 = block_min_u64<NT>(best64, s_red64);
    if (tid == 0) {
      if (cfg.collect_stats && st->stats) {
        asrd_frame_stat s;
        s.n_in = is_init ? 0 : st->n_cur;
        s.cur_cutoff = is_init ? 0.f : st->cur_cut;
        s.abeam = is_init ? 0.f : st->abeam;
        s.next_cutoff = nc;
        s.n_tokens = n_alive;
        s.best = ord2f((uint32_t)(best64 >> 32));
        s.arcs_expanded = is_init ? 0 : st->arcs_expanded;
        s.arcs_admitted = is_init ? 0 : st->arcs_admitted;
        st->stats[t + 1] = s;
      }
      if (is_init) {
        st->tot_arcs_expanded = 0;
        st->tot_arcs_admitted = 0;
      } else {
        st->tot_arcs_expanded += st->arcs_expanded;
        st->tot_arcs_admitted += st->arcs_admitted;
      }
      st->frame_off[t + 2] = out_base + n_alive;
      st->n_slots[next] = ntot;
      if (!is_init) st->n_slots[cur] = 0;
      st->frame = t + 1;
      st->n_cur = n_alive;
      st->tiles = 0;
    }
    __syncthreads();
  }

  // =============================== prologue of the next frame: GetCutoff + pre-pass
  if (mode & kModePro) {
    __syncthreads();
    const int t = st->frame;
    if (t >= st->target_frame) {
      if (tid == 0) st->tiles = 0;
      return;
    }
    const uint32_t n = st->n_cur;
    const uint2 *toks = st->tok_sc + st->frame_off[t];
    const float *__restrict__ ll = st->ll_base + (size_t)(t - st->ll_frame0) * st->ll_stride;
    // best token: lowest cost, ties -> lowest state id (inl.h:169-179)
    unsigned long long best64 = kInfVal;
    for (uint32_t i = tid; i < n; i += NT) {
      const uint2 sc = toks[i];
      const unsigned long long b = ((unsigned long long)f2ord(__uint_as_float(sc.y)) << 32) | sc.x;
      best64 = b < best64 ? b : best64;
    }
    best64 = block_min_u64<NT>(best64, s_red64);
    float cur_cut = CUDART_INF_F, abeam = cfg.beam;
    uint32_t next_bits = kOrdInf;
    if (n > 0) {
      const uint32_t best_ord = (uint32_t)(best64 >> 32);
      const float bc = ord2f(best_ord);
      const float beam_cut = bc + cfg.beam;  // inl.h:182
      uint32_t lt = 0, le = 0;
      for (uint32_t i = tid; i < n; i += NT) {
        const float c = __uint_as_float(toks[i].y);
        lt += c < beam_cut;
        le += c <= beam_cut;
      }
      lt = block_sum_u32<NT>(lt, s_red32);
      le = block_sum_u32<NT>(le, s_red32);
      cur_cut = beam_cut;
      if (lt > (uint32_t)cfg.max_active) {
        // sorted[max_active] < beam_cutoff  <=>  more than max_active costs below it (inl.h:188-203)
        cur_cut = block_kth_smallest<NT>(toks, n, (uint32_t)cfg.max_active, best_ord,
                                         f2ord(beam_cut), s_hist, s_misc);
        abeam = cur_cut - bc + cfg.beam_delta;
      } else if (n <= (uint32_t)cfg.min_active) {
        // fewer tokens than min_active: min_active_cutoff stays +inf > beam_cutoff, so nothing is
        // pruned and the adaptive beam is infinite (inl.h:183,205,220-226)
        cur_cut = CUDART_INF_F;
        abeam = CUDART_INF_F;
      } else if (cfg.min_active > 0 && le <= (uint32_t)cfg.min_active) {
        // sorted[min_active] > beam_cutoff  <=>  at most min_active costs <= it (inl.h:205-226)
        cur_cut = block_kth_smallest<NT>(toks, n, (uint32_t)cfg.min_active, best_ord, 0xFFFFFFFFu,
                                         s_hist, s_misc);
        abeam = cur_cut - bc + cfg.beam_delta;
      }
      // best-token pre-pass (inl.h:282-300): association (cost + w) - loglike
      const uint32_t sb = (uint32_t)best64;
      const uint2 r = __ldg(&g.rows[sb]);
      const uint32_t end = __ldg(&g.rows[sb + 1]).x;
      uint32_t mn = kOrdInf;
      for (uint32_t a = r.y + tid; a < end; a += NT) {
        const int4 arc = __ldg(&g.arcs[a]);
        const float tot = bc + __int_as_float(arc.z) - __ldg(&ll[arc.x - 1]);
        mn = min(mn, f2ord(tot + abeam));
      }
      unsigned long long m64 = block_min_u64<NT>((unsigned long long)mn, s_red64);
      next_bits = (uint32_t)m64;
    }
    if (tid == 0) {
      st->cur_cut = cur_cut;
      st->abeam = abeam;
      st->next_cut_bits = next_bits;
      st->arcs_expanded = 0;
      st->arcs_admitted = 0;
      st->tiles = (n + kTileTokens - 1) / kTileTokens;
    }
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
}

// ------------------------------------------------------------------ best path

// One CTA per stream.  Picks the end token (BestPathEnd, inl.h:1096-1158: the token on the
// super-final state when use_final_probs and one is alive, else the cheapest token; ties ->
// lowest state id) and follows the winning arcs back to the start token
// (TraceBackBestPath, inl.h:1160-1200).  Arcs are emitted end -> start.
__global__ void __launch_bounds__(256)
k_best_path(StreamState *const *streams, GraphView g, int use_final, int cap, int32_t *o_il,
            int32_t *o_ol, float *o_gr, float *o_ac, int32_t *o_n, int32_t *o_status) {
  constexpr int NT = 256;
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
  // end token
  const uint32_t b0 = st->frame_off[f], n0 = st->frame_off[f + 1] - b0;
  unsigned long long best_all = kInfVal, best_fin = kInfVal;
  for (uint32_t i = tid; i < n0; i += NT) {
    const uint2 sc = st->tok_sc[b0 + i];
    // key: (cost, state) for the argmin; the index is recovered by a second scan
    const unsigned long long b = ((unsigned long long)f2ord(__uint_as_float(sc.y)) << 32) | sc.x;
    best_all = b < best_all ? b : best_all;
    if ((int32_t)sc.x == g.final_state) best_fin = b < best_fin ? b : best_fin;
  }
  best_all = block_min_u64<NT>(best_all, s_red64);
  best_fin = block_min_u64<NT>(best_fin, s_red64);
  unsigned long long pick = (use_final && best_fin != kInfVal) ? best_fin : best_all;
  if (pick == kInfVal) {
    if (tid == 0) {
      o_n[blockIdx.x] = 0;
      o_status[blockIdx.x] = ASRD_ERR_NO_TOKENS;
    }
    return;
  }
  uint32_t want = (uint32_t)pick;  // state of the token to locate in frame f
  int n_out = 0;
  int status = ASRD_OK;
  while (true) {
    // locate the token of `want` in frame f
    if (tid == 0) s_found = 0xFFFFFFFFu;
    __syncthreads();
    const uint32_t b = st->frame_off[f], n = st->frame_off[f + 1] - b;
    for (uint32_t i = tid; i < n; i += NT)
      if (st->tok_sc[b + i].x == want) s_found = b + i;
    __syncthreads();
    const uint32_t idx = s_found;
    __syncthreads();
    if (idx == 0xFFFFFFFFu) {
      status = ASRD_ERR_STATE;  // broken back-trace: cannot happen unless the arena overflowed
      break;
    }
    const uint2 aa = st->tok_aa[idx];
    int32_t il = 0, ol = 0;
    float gr = 0.f;
    const float ac = __uint_as_float(aa.y);
    if (aa.x != kNoArc) {
      const int4 arc = __ldg(&g.arcs[aa.x]);
      il = arc.x;
      ol = arc.y;
      gr = __int_as_float(arc.z);
    }
    if (n_out >= cap) {
      status = ASRD_ERR_PATH_OVERFLOW;
      break;
    }
    if (tid == 0) {
      o_il[ob + n_out] = il;
      o_ol[ob + n_out] = ol;
      o_gr[ob + n_out] = gr;
      o_ac[ob + n_out] = ac;
    }
    ++n_out;
    if (aa.x == kNoArc) break;  // start token (inl.h:1193-1198)
    want = __ldg(&g.arc_src[aa.x]);
    if (il != 0) --f;
    if (f < 0) {
      status = ASRD_ERR_STATE;
      break;
    }
  }
  if (tid == 0) {
    o_n[blockIdx.x] = n_out;
    o_status[blockIdx.x] = status;
  }
}

}  // namespace asrd
