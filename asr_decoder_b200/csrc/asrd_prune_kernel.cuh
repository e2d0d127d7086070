// k_prune: PruneActiveTokens (online-decoder-base-inl.h:438-480, called every prune_interval
// frames at :660-661) for plain decoders with the device option prune_tokens.  Included by
// asrd_kernels.cuh after k_lattice, whose PRUNE mode is the same sweep over HBM maps (kept as the
// selectable second implementation, ASRD_PRUNE_KERNEL=0; the tests hold the two against each other).
#pragma once

namespace asrd {

// One CTA per stream sweeps the frames backwards from the frontier F.  The device keeps tokens
// only, so the forward links are regenerated like in k_lattice: tok -> dst exists iff tok was
// expanded (cost <= cur_cutoff of its frame, inl.h:315; eps arcs: cost < the frame's closure
// cutoff, inl.h:391) and the arc's cost is below the final cutoff of the destination frame
// (inl.h:330, 415).  link_extra_cost = dst.extra + ((tok.cost + arc) - dst.cost) (inl.h:524-526);
// a link above lattice_beam is dropped (inl.h:532), a token's extra cost is the minimum over its
// links (inl.h:545-562), a token without links dies (PruneTokensForFrame, inl.h:591).  The
// frontier keeps every token at extra cost 0.
//
// The sweep PULLS: only a few per cent of a frame's tokens stay alive a handful of frames behind
// the frontier, so instead of expanding every token of frame f forwards (26 k arcs) the survivors
// of frame f + 1 walk their INCOMING arcs (the graph's incoming-arc index, in_off / in_arc) and look
// the source states up in a state -> token map of frame f held in shared memory (key u32 + index
// u16 per slot, buckets of four keys like k_stream's).  Eps links inside the frame are pulled the
// same way from a worklist of the tokens whose extra cost was just lowered, round by round, to the
// exact fixed point (min-plus: the least fixed point does not depend on the order).  The extra
// costs of the frame being swept live in shared memory too.  Survivors are moved to the front of
// their frame's span as soon as the frame is done (so the next frame reads its seeds coalesced)
// and the spans are closed up at the end.
//
// An extra cost can only grow as the frontier moves on (later frontiers are reached THROUGH this
// one; float addition and min are monotone), so a token dropped here would be dropped by the
// final sweep too and a link to it never survives: one-best and raw lattice are bit-identical
// with and without pruning.  For the same reason the sweep may stop anywhere: `depth` >= 0 limits
// it to that many frames below the previous frontier (frames further back were swept at least
// twice and hold a few dozen tokens; re-sweeping them every time costs more than it frees).
constexpr uint32_t kPruneNone = 0xFFFFFFFFu;
constexpr int kPruneHubDeg = 96;    // destinations with more incoming arcs are walked by the whole CTA
constexpr int kPruneHubCap = 128;
constexpr int kPruneMaxFrames = 32768;

// shared-memory bytes besides the map: extra costs (4 B per token) and two queued-flag bitmaps
__host__ __device__ constexpr size_t prune_token_dyn_bytes(uint32_t ex_cap) { return (size_t)ex_cap * 4 + 2 * (size_t)(ex_cap / 8); }

struct PruneMap {
  uint32_t key_sa;  // shared-space address of key[4 * max buckets]; idx (u16) follows at idx_sa
  uint32_t idx_sa;
  uint32_t n_buckets;  // buckets in use for the frame the map holds (sized to the frame: clearing is per frame)
};

__device__ __forceinline__ void pm_insert(const PruneMap &m, uint32_t key, uint32_t idx) {
  const uint32_t hsh = key * 0x9E3779B1u;
  uint32_t b = __umulhi(hsh, m.n_buckets);
  for (;;) {  // (the caller keeps the load below 7/8: a free slot exists; the next bucket in line is tried,
              // which visits every bucket whatever the table size)
    const uint32_t ba = m.key_sa + b * 16u;
    const uint4 kk = lds_u4(ba);
    if (kk.w == kEmptyKey) {  // slots fill a bucket front to back
      const uint32_t emp = kk.x == kEmptyKey ? 0u : kk.y == kEmptyKey ? 1u : kk.z == kEmptyKey ? 2u : 3u;
      if (atoms_cas(ba + emp * 4u, kEmptyKey, key) == kEmptyKey) {
        sts_u16(m.idx_sa + (b * 4u + emp) * 2u, idx);
        return;
      }
      continue;  // somebody else's key took the slot: look at the bucket again
    }
    if (++b == m.n_buckets) b = 0;
  }
}

__device__ __forceinline__ uint32_t pm_find(const PruneMap &m, uint32_t key) {
  const uint32_t hsh = key * 0x9E3779B1u;
  uint32_t b = __umulhi(hsh, m.n_buckets);
  for (uint32_t probe = 0; probe < m.n_buckets; ++probe) {
    const uint32_t ba = m.key_sa + b * 16u;
    const uint4 kk = lds_u4(ba);
    uint32_t hit = kPruneNone;
    if (kk.x == key) hit = 0;
    else if (kk.y == key) hit = 1;
    else if (kk.z == key) hit = 2;
    else if (kk.w == key) hit = 3;
    if (hit != kPruneNone) {
      unsigned short v;
      asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=h"(v) : "r"(m.idx_sa + (b * 4u + hit) * 2u) : "memory");
      return v;
    }
    if (kk.w == kEmptyKey) return kPruneNone;  // a key lives in the first bucket of its sequence that had room
    if (++b == m.n_buckets) b = 0;
  }
  return kPruneNone;
}

// EMIT = the same sweep as GetRawLattice + the pruning of FinalizeDecoding (inl.h:725-847, 868-975),
// the fast twin of k_lattice for plain decoders: the frontier starts from the final costs
// (PruneForwardLinksFinal, inl.h:758-816) instead of 0, NOTHING in the arena is touched (no
// compaction: the seeds of a frame are an index list, the extra costs of frame f + 1 sit in a scratch
// buffer), and the surviving tokens and links are written to `outs` as they become final — emitting
// links while they are pulled (the destination's extra cost is final by then), eps links in one more
// pull over the frame's survivors after its fixed point.  A frame beyond the kernel's capacity
// reports n_toks = 0xFFFFFFFF and the host falls back to k_lattice.
template <bool EMIT>
__global__ void __launch_bounds__(kStreamThreads, 1)
k_prune(StreamState *const *streams, GraphView g, DecoderConfigDev cfg, int prune_interval, int depth,
        uint32_t max_buckets, uint32_t ex_cap, LatticeOut *outs, int use_final) {
  constexpr int NT = kStreamThreads;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ uint32_t s_nalive, s_nwl[2], s_nhub;
  __shared__ uint32_t s_hub[kPruneHubCap];
  __shared__ unsigned long long s_red64[NT / 32];
  StreamState *st = streams[blockIdx.x];
  LatticeOut *out = EMIT ? &outs[blockIdx.x] : nullptr;
  const int tid = threadIdx.x;
  const int F = st->frame;
  if (EMIT) {
    if (tid == 0) {
      out->n_toks = 0;
      out->n_links = 0;
    }
    if (F < 0) return;
  } else if (F < 0 || st->status < 0 || F - st->gc_frame < prune_interval) {
    return;  // uniform
  }
  const uint32_t max_slots = max_buckets * 4;
  uint32_t *s_key = reinterpret_cast<uint32_t *>(s_dyn);
  uint16_t *s_idx = reinterpret_cast<uint16_t *>(s_key + max_slots);
  uint32_t *s_ex = reinterpret_cast<uint32_t *>(s_idx + max_slots);  // (max_slots is a multiple of 8: 16-byte aligned)
  uint32_t *s_flag[2];
  s_flag[0] = s_ex + ex_cap;
  s_flag[1] = s_flag[0] + ex_cap / 32;
  PruneMap m;
  m.key_sa = smem_addr(s_key);
  m.idx_sa = smem_addr(s_idx);
  m.n_buckets = 32;
  const int lo_limit = (!EMIT && depth >= 0) ? max(0, st->gc_frame - depth) : 0;
  // global scratch (the stream's closure queues, idle between frame-loop launches): worklists and
  // the alive list (u16 token indices), survivors per frame, staging of the survivors of a frame
  const uint32_t H = st->hash_mask + 1;
  const uint32_t cap = min(min(ex_cap, max_slots / 8 * 7), min(H / 3, 65535u));  // tokens per frame this kernel takes
  uint16_t *wl[2];
  wl[0] = reinterpret_cast<uint16_t *>(st->queue[0]);
  wl[1] = wl[0] + cap;
  uint16_t *alive = wl[1] + cap;                      // 3 * cap * 2 <= 2 H bytes
  uint32_t *keep = st->queue[0] + H / 2;              // [frames swept]  (PRUNE)
  uint2 *stage_sc = reinterpret_cast<uint2 *>(st->queue[1]);
  uint32_t *stage_ex = st->queue[1] + 2 * (size_t)cap;  // 3 * cap * 4 <= 4 H bytes  (PRUNE)
  // EMIT: nothing is compacted — the survivors of frame f + 1 are an index list (two lists, taking
  // turns with `alive`), their extra costs a scratch array indexed like the frame
  uint16_t *alive_alt = alive + cap;                  // 4 * cap * 2 <= 8 H / 3 bytes of queue[0]
  uint32_t *exn = st->queue[1];                       // [cap] extra costs of frame f + 1
  {
    uint32_t too_big = (uint32_t)(F - lo_limit + 1) > min((uint32_t)kPruneMaxFrames, H / 2) ? 1u : 0u;
    for (int f = lo_limit + tid; f <= F; f += NT) too_big |= (st->frame_off[f + 1] - st->frame_off[f]) > cap;
    if (__syncthreads_or((int)too_big)) {  // (the HBM-map sweep takes the stream)
      if (EMIT && tid == 0) out->n_toks = 0xFFFFFFFFu;
      return;
    }
    if (!EMIT && tid == 0) {
      const uint32_t used = st->frame_off[F + 1];
      if (used > st->peak_tokens) st->peak_tokens = used;
    }
  }
  const float beam = cfg.lattice_beam;
  // SM cycles per phase (diagnostic; thread 0, right after a barrier): map build, emitting links,
  // eps rounds, survivors to the front, closing up; (unused); frames swept, eps rounds
  long long tph = clock64();
  unsigned long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto phase = [&](int k) {
    if (tid == 0) {
      const long long now = clock64();
      ph[k] += (unsigned long long)(now - tph);
      tph = now;
    }
  };
  // ---- the frontier.  PRUNE: it is not pruned, extra cost 0 everywhere (inl.h:455-470 stops above
  // it).  EMIT: it is swept like every frame, starting from the final costs (below).
  float final_best = 0.f;
  bool any_final = false;
  if (!EMIT) {
    const uint32_t b0 = st->frame_off[F], n = st->frame_off[F + 1] - b0;
    for (uint32_t i = tid; i < n; i += NT) st->tok_extra[b0 + i] = f2ord(0.f);
    if (tid == 0) keep[F - lo_limit] = n;
  } else {
    // ComputeFinalCosts (inl.h:670-720): the token on the super-final state, if any
    const uint32_t b0 = st->frame_off[F], n0 = st->frame_off[F + 1] - b0;
    unsigned long long best_all = kInfVal, best_fin = kInfVal;
    for (uint32_t i = tid; i < n0; i += NT) {
      const uint2 sc = __ldcg(&st->tok_sc[b0 + i]);
      const unsigned long long b = ((unsigned long long)f2ord(__uint_as_float(sc.y)) << 32) | sc.x;
      best_all = b < best_all ? b : best_all;
      if ((int32_t)sc.x == g.final_state) best_fin = b < best_fin ? b : best_fin;
    }
    best_all = block_min_u64<NT>(best_all, s_red64);
    best_fin = block_min_u64<NT>(best_fin, s_red64);
    any_final = use_final && best_fin != kInfVal;
    final_best = ord2f((uint32_t)((any_final ? best_fin : best_all) >> 32));  // inl.h:709-719
  }
  uint32_t k1 = st->frame_off[F + 1] - st->frame_off[F];  // survivors of frame f + 1 (PRUNE: at the front of its span)
  uint16_t *seeds = alive_alt;                             // EMIT: their indices
  __syncthreads();
  for (int f = EMIT ? F : F - 1; f >= lo_limit; --f) {
    const bool frontier = EMIT && f == F;
    const uint32_t b0 = st->frame_off[f], n = st->frame_off[f + 1] - b0;
    const uint32_t b1 = st->frame_off[f + 1];
    const float nc_f = st->frame_nc[f], nc_next = frontier ? 0.f : st->frame_nc[f + 1], cur_cut = frontier ? 0.f : st->frame_cur[f];
    const float *__restrict__ ll = st->ll_hist + (size_t)(frontier ? 0 : f) * st->ll_stride;
    uint16_t *alive_cur = EMIT ? (seeds == alive ? alive_alt : alive) : alive;
    if (tid == 0) {
      s_nalive = 0;
      s_nwl[0] = s_nwl[1] = 0;
      s_nhub = 0;
    }
    // ---- the map takes frame f (sized to it: load <= 1/2 where the room allows)
    m.n_buckets = min(max_buckets, max(32u, (n + 1) / 2));
    {
      uint4 *k4 = reinterpret_cast<uint4 *>(s_key);
      for (uint32_t i = tid; i < m.n_buckets; i += NT) k4[i] = make_uint4(kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey);
      for (uint32_t i = tid; i < n; i += NT) s_ex[i] = kOrdInf;
      for (uint32_t i = tid; i < (n + 31) / 32; i += NT) s_flag[0][i] = s_flag[1][i] = 0;
    }
    __syncthreads();
    for (uint32_t i0 = tid; i0 < n; i0 += 4 * NT) {  // (the loads of four tokens in flight together)
      uint32_t key[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) key[u] = __ldcg(&st->tok_sc[b0 + min(i0 + u * NT, n - 1)]).x;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * NT < n) pm_insert(m, key[u], i0 + u * NT);
    }
    __syncthreads();
    phase(0);
    ph[6] += 1;
    // The link src token i (frame f, cost `cost`) -> a destination token (cost dcost, extra cost
    // dextra; EMIT: in frame f + 1 over an emitting arc whose log-likelihood is llv, else in frame f
    // over an eps arc) with arc weight w: if the search admitted it, its extra cost goes to the source.
    auto emit_link = [&](uint32_t src_idx, uint32_t dst_idx, int32_t il, int32_t ol, float graph_cost, float ac) {
      const uint32_t p = atomicAdd(&out->n_links, 1u);
      if (p < out->link_cap) {
        asrd_lat_link l;
        l.src = (int32_t)src_idx;
        l.dst = (int32_t)dst_idx;
        l.ilabel = il;
        l.olabel = ol;
        l.graph = graph_cost;
        l.acoustic = ac;
        out->links[p] = l;
      }
    };
    // (LINKS, EMIT only: the extra costs are final and this is the pass that writes the surviving
    // eps links — same admission tests, nothing is relaxed)
    auto apply = [&](uint32_t i, float cost, float w, float llv, bool EMITTING, float dcost, float dextra, uint32_t par,
                     int32_t il, int32_t ol, uint32_t dst_arena, bool LINKS) {
      float tot;
      if (EMITTING) {
        if (g.clg ? !(cost < cur_cut) : !(cost <= cur_cut)) return;     // inl.h:315 (CLG decoder: strict)
        tot = (cost + (-llv)) + w;                                      // inl.h:326-329
        if (g.clg ? !(tot <= nc_next) : !(tot < nc_next)) return;       // inl.h:330, final cutoff (CLG: inclusive)
      } else {
        if (!(cost < nc_f)) return;       // inl.h:391
        tot = cost + w;                   // inl.h:413-414
        if (!(tot < nc_f)) return;        // inl.h:415
      }
      float le = dextra + (tot - dcost);  // inl.h:524-526
      if (le > beam) return;              // inl.h:532
      if (le < 0.f) le = 0.f;             // inl.h:545-551
      // (EMIT: an emitting link's destination extra cost is final: the link is (inl.h:532 passed))
      if (EMIT && EMITTING) emit_link(b0 + i, dst_arena, il, ol, w, -llv);
      if (EMIT && LINKS) {
        emit_link(b0 + i, dst_arena, il, ol, w, 0.f);
        return;
      }
      const uint32_t v = f2ord(le);
      const uint32_t old = atomicMin(&s_ex[i], v);
      if (v < old) {
        if (old == kOrdInf) alive_cur[atomicAdd(&s_nalive, 1u)] = (uint16_t)i;
        // its own incoming eps links have to be looked at (again): queue it once per round
        const uint32_t bit = 1u << (i & 31);
        if (!(atomicOr(&s_flag[par][i >> 5], bit) & bit)) wl[par][atomicAdd(&s_nwl[par], 1u)] = (uint16_t)i;
      }
    };
    auto link = [&](uint32_t i, const int4 &arc, bool EMITTING, float dcost, float dextra, uint32_t par, uint32_t dst_arena,
                    bool LINKS) {
      const float cost = __uint_as_float(__ldcg(&st->tok_sc[b0 + i]).y);
      apply(i, cost, __int_as_float(arc.z), EMITTING ? __ldg(&ll[arc.x - 1]) : 0.f, EMITTING, dcost, dextra, par, arc.x, arc.y,
            dst_arena, LINKS);
    };
    auto relax_in = [&](uint32_t a, bool EMITTING, float dcost, float dextra, uint32_t par, uint32_t dst_arena,
                        bool LINKS) {  // incoming arc a
      const uint32_t i = pm_find(m, __ldg(&g.arc_src[a]));
      if (i != kPruneNone) link(i, __ldg(&g.arcs[a]), EMITTING, dcost, dextra, par, dst_arena, LINKS);
    };
    // The incoming arcs (of the class) of one destination token, this lane's share of them: lane `sub`
    // of the `nsub` lanes that work on the destination takes every nsub-th arc.  A prune is a chain of
    // dependent loads (index -> arc -> source token -> log-likelihood), so everything that can be
    // requested together is: the ranges, then up to four arc ids, then their source states AND arc
    // records, then (after the shared-memory lookups) the costs and log-likelihoods of the hits.
    // Hubs are left to the whole CTA.
    auto pull = [&](uint32_t state, uint32_t hub_tag, bool EMITTING, float dcost, float dextra, uint32_t par,
                    uint32_t sub, uint32_t nsub, uint32_t dst_arena, bool LINKS) {
      const uint32_t mid = __ldg(&g.in_mid[state]);
      const uint32_t ib = EMITTING ? mid : __ldg(&g.in_off[state]), ie = EMITTING ? __ldg(&g.in_off[state + 1]) : mid;
      if (ie - ib > (uint32_t)kPruneHubDeg) {
        if (sub != 0) return;
        const uint32_t h = atomicAdd(&s_nhub, 1u);
        if (h < (uint32_t)kPruneHubCap) {
          s_hub[h] = hub_tag;
          return;
        }
        nsub = 1;  // (no room in the hub list: this lane walks all of it)
      }
      for (uint32_t r0 = ib + sub; r0 < ie; r0 += 4 * nsub) {
        uint32_t a[4], src[4], idx[4];
        int4 arc[4];
        float cost[4], llv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = __ldg(&g.in_arc[min(r0 + u * nsub, ie - 1)]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          src[u] = __ldg(&g.arc_src[a[u]]);
          arc[u] = __ldg(&g.arcs[a[u]]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) idx[u] = r0 + u * nsub < ie ? pm_find(m, src[u]) : kPruneNone;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool hit = idx[u] != kPruneNone;
          cost[u] = __uint_as_float(__ldcg(&st->tok_sc[b0 + (hit ? idx[u] : 0u)]).y);
          llv[u] = EMITTING ? __ldg(&ll[hit ? arc[u].x - 1 : 0]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (idx[u] != kPruneNone)
            apply(idx[u], cost[u], __int_as_float(arc[u].z), llv[u], EMITTING, dcost, dextra, par, arc[u].x, arc[u].y, dst_arena,
                  LINKS);
      }
    };
    // Hubs (whole CTA; base: arena offset of the destinations' frame).  A hub with fewer incoming
    // arcs than the frame has tokens is walked from its side; else the frame's tokens are walked
    // forwards and only their arcs INTO the hub looked at (the super-final state has an incoming
    // arc from every final state of the graph).
    auto next_extra = [&](uint32_t j) { return EMIT ? __ldcg(&exn[j]) : __ldcg(&st->tok_extra[b1 + j]); };
    auto pull_hubs = [&](bool EMITTING, uint32_t base, uint32_t par, bool LINKS) {
      const uint32_t nh = min(s_nhub, (uint32_t)kPruneHubCap);
      for (uint32_t h = 0; h < nh; ++h) {
        const uint32_t j = s_hub[h];
        const uint2 sc = __ldcg(&st->tok_sc[base + j]);
        const float dcost = __uint_as_float(sc.y);
        const float dextra = ord2f(EMITTING ? next_extra(j) : *reinterpret_cast<volatile uint32_t *>(&s_ex[j]));
        const uint32_t mid = __ldg(&g.in_mid[sc.x]);
        const uint32_t ib = EMITTING ? mid : __ldg(&g.in_off[sc.x]), ie = EMITTING ? __ldg(&g.in_off[sc.x + 1]) : mid;
        if (ie - ib <= 4u * n) {
          for (uint32_t r = ib + tid; r < ie; r += NT) relax_in(__ldg(&g.in_arc[r]), EMITTING, dcost, dextra, par, base + j, LINKS);
        } else {
          for (uint32_t i = tid; i < n; i += NT) {
            const uint2 tk = __ldcg(&st->tok_sc[b0 + i]);
            const float cost = __uint_as_float(tk.y);
            if (EMITTING ? !(cost <= cur_cut) : !(cost < nc_f)) continue;
            if (!EMITTING && !((__ldg(&g.eps_bits[tk.x >> 5]) >> (tk.x & 31)) & 1u)) continue;
            const uint2 span = EMITTING ? __ldg(&g.erows[tk.x]) : __ldg(&g.rows[tk.x]);
            for (uint32_t a = span.x; a < span.y; ++a) {
              const int4 arc = __ldg(&g.arcs[a]);
              if (((uint32_t)arc.w & kStateMask) == sc.x) link(i, arc, EMITTING, dcost, dextra, par, base + j, LINKS);
            }
          }
        }
      }
    };
    if (!frontier) {
      // ---- emitting links: the survivors of frame f + 1 pull from their sources in frame f
      // (few survivors: several lanes share one destination, so that the chain of a destination with
      // many incoming arcs is not what every other warp waits for at the barrier)
      const uint32_t nsub = k1 > NT / 2 ? 1u : k1 > NT / 4 ? 2u : k1 > NT / 8 ? 4u : 8u;
      for (uint32_t w = tid; w < k1 * nsub; w += NT) {
        const uint32_t q = w / nsub;
        const uint32_t j = EMIT ? (uint32_t)__ldcg(&seeds[q]) : q;
        const uint2 sc = __ldcg(&st->tok_sc[b1 + j]);
        pull(sc.x, j, true, __uint_as_float(sc.y), ord2f(next_extra(j)), 0u, w % nsub, nsub, b1 + j, false);
      }
      __syncthreads();
      if (s_nhub) {
        pull_hubs(true, b1, 0u, false);
        __syncthreads();
      }
    } else {
      // ---- EMIT, frame F: PruneForwardLinksFinal (inl.h:758-775, 815-816): a token starts from its cost
      // with the final cost against the best one, or is dropped when that is beyond the lattice beam
      for (uint32_t i = tid; i < n; i += NT) {
        const uint2 sc = __ldcg(&st->tok_sc[b0 + i]);
        float fc = 0.f;
        if (any_final) fc = (int32_t)sc.x == g.final_state ? 0.f : CUDART_INF_F;
        float init = __uint_as_float(sc.y) + fc - final_best;
        if (init > beam) init = CUDART_INF_F;
        if (init < CUDART_INF_F) {
          s_ex[i] = f2ord(init);
          alive_cur[atomicAdd(&s_nalive, 1u)] = (uint16_t)i;
          atomicOr(&s_flag[0][i >> 5], 1u << (i & 31));
          wl[0][atomicAdd(&s_nwl[0], 1u)] = (uint16_t)i;
        }
      }
      __syncthreads();
    }
    phase(1);
    // ---- eps links inside the frame: tokens whose extra cost was lowered pull from their eps
    // sources, round by round, until nothing moves (the exact fixed point of inl.h:524-562)
    for (uint32_t round = 0;; ++round) {
      const uint32_t par = round & 1u;
      const uint32_t nq = s_nwl[par];
      if (nq == 0) break;  // uniform
      __syncthreads();
      if (tid == 0) {
        s_nwl[par ^ 1u] = 0;
        s_nhub = 0;
      }
      __syncthreads();
      const uint32_t nsub = nq > NT / 2 ? 1u : nq > NT / 4 ? 2u : nq > NT / 8 ? 4u : 8u;
      for (uint32_t w = tid; w < nq * nsub; w += NT) {
        const uint32_t i = __ldcg(&wl[par][w / nsub]);
        if (w % nsub == 0) atomicAnd(&s_flag[par][i >> 5], ~(1u << (i & 31)));
        const uint2 sc = __ldcg(&st->tok_sc[b0 + i]);
        pull(sc.x, i, false, __uint_as_float(sc.y), ord2f(*reinterpret_cast<volatile uint32_t *>(&s_ex[i])), par ^ 1u,
             w % nsub, nsub, b0 + i, false);
      }
      __syncthreads();
      if (s_nhub) {
        pull_hubs(false, b0, par ^ 1u, false);
        __syncthreads();
      }
      ph[7] += 1;
    }
    phase(2);
    const uint32_t k = s_nalive;
    if constexpr (EMIT) {
      // ---- the frame's extra costs are final: its surviving tokens, and — one more pull over them —
      // its surviving eps links (inl.h:532 with the final extra costs on both ends)
      for (uint32_t q = tid; q < k; q += NT) {
        const uint32_t i = __ldcg(&alive_cur[q]);
        const uint2 sc = __ldcg(&st->tok_sc[b0 + i]);
        const uint32_t p = atomicAdd(&out->n_toks, 1u);
        if (p < out->tok_cap) {
          asrd_lat_token t;
          t.frame = f;
          t.state = (int32_t)sc.x;
          t.cost = __uint_as_float(sc.y);
          t.extra = ord2f(s_ex[i]);
          t.is_final = (f == F && (!any_final || (int32_t)sc.x == g.final_state)) ? 1 : 0;  // inl.h:935-951
          out->toks[p] = t;
          out->tok_arena_idx[p] = b0 + i;
        }
      }
      __syncthreads();  // (everybody is past the last eps round's look at the hub counter)
      if (tid == 0) s_nhub = 0;
      __syncthreads();
      {
        const uint32_t nsub = k > NT / 2 ? 1u : k > NT / 4 ? 2u : k > NT / 8 ? 4u : 8u;
        for (uint32_t w = tid; w < k * nsub; w += NT) {
          const uint32_t i = __ldcg(&alive_cur[w / nsub]);
          const uint2 sc = __ldcg(&st->tok_sc[b0 + i]);
          pull(sc.x, i, false, __uint_as_float(sc.y), ord2f(s_ex[i]), 0u, w % nsub, nsub, b0 + i, true);
        }
      }
      __syncthreads();
      if (s_nhub) {
        pull_hubs(false, b0, 0u, true);
        __syncthreads();
      }
      // the next frame's seeds: this frame's survivors, their extra costs in the scratch array
      for (uint32_t q = tid; q < k; q += NT) {
        const uint32_t i = __ldcg(&alive_cur[q]);
        exn[i] = s_ex[i];
      }
      seeds = alive_cur;
      k1 = k;
      __syncthreads();
      phase(3);
    } else {
      // ---- survivors (PruneTokensForFrame, inl.h:591) to the front of the frame's span, through a
      // staging buffer: the next frame reads them as its seeds, coalesced
      for (uint32_t q = tid; q < k; q += NT) {
        const uint32_t i = __ldcg(&alive[q]);
        stage_sc[q] = __ldcg(&st->tok_sc[b0 + i]);
        stage_ex[q] = s_ex[i];
      }
      __syncthreads();
      for (uint32_t q = tid; q < k; q += NT) {
        st->tok_sc[b0 + q] = __ldcg(&stage_sc[q]);
        st->tok_extra[b0 + q] = __ldcg(&stage_ex[q]);
      }
      if (tid == 0) keep[f - lo_limit] = k;
      k1 = k;
      __syncthreads();
      phase(3);
    }
  }
  if constexpr (EMIT) return;
  // ---- close the spans up from frame lo_limit upwards (a tile is read into registers before
  // anything of it is written: tokens only ever move towards the front)
  {
    uint32_t run = st->frame_off[lo_limit], old_b = run;
    __syncthreads();
    for (int f = lo_limit; f <= F; ++f) {
      const uint32_t old_e = st->frame_off[f + 1];
      const uint32_t kf = __ldcg(&keep[f - lo_limit]), new_b = run;
      __syncthreads();  // everybody has read frame_off[f + 1] (written by the next round)
      if (new_b != old_b) {
        for (uint32_t i0 = 0; i0 < kf; i0 += 4 * NT) {
          uint2 sc[4];
          uint32_t ex[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = min(i0 + u * NT + tid, kf - 1);
            sc[u] = __ldcg(&st->tok_sc[old_b + i]);
            ex[u] = __ldcg(&st->tok_extra[old_b + i]);
          }
          __syncthreads();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * NT + tid;
            if (i < kf) {
              st->tok_sc[new_b + i] = sc[u];
              st->tok_extra[new_b + i] = ex[u];
            }
          }
          __syncthreads();
        }
      }
      if (tid == 0) st->frame_off[f] = new_b;
      run += kf;
      old_b = old_e;
    }
    __syncthreads();
    if (tid == 0) {
      st->frame_off[F + 1] = run;
      st->tot_pruned_tokens += old_b - run;
      st->gc_frame = F;
    }
  }
  __syncthreads();
  phase(4);
  if (tid == 0)
    for (int k = 0; k < 8; ++k) st->prune_cycles[k] += ph[k];
}

}  // namespace asrd
