// k_stream: the on-chip frame loop of the plain decoders (one CTA per stream, the per-frame
// state->token map in shared memory).  Included by asrd_kernels.cuh after the HBM-map device
// functions it falls back to (expand_frame, post_epilogue, cutoff_prologue).
#pragma once

namespace asrd {

// ------------------------------------------------------------------ on-chip frame loop

// k_stream: ONE CTA per stream runs the whole frame loop of an AdvanceDecoding chunk with the
// per-frame state->token map in SHARED memory (the stream's private recombination state never
// leaves the SM): GetCutoff + pre-pass, emitting expansion, eps closure and the survivor
// write-out of every frame without going back to the launch queue.  HBM traffic per frame is
// what the search really needs — arc records, row offsets, the previous frame's tokens, the
// log-likelihood row — plus the 8-byte {state, cost} append of the survivors to the token arena.
//
// The map holds COSTS ONLY: key[slot] = destination state, cost[slot] = order-preserving uint of
// the best cost so far, recombined with the native 32-bit shared-memory atomicMin (ATOMS.MIN).
// Which arc won is not recorded anywhere in the hot loop: the trace-back recovers it from the
// graph's incoming-arc index (k_best_path_rev), for the few hundred tokens of the best path
// instead of for every token of every frame.
//
// A frame whose distinct destination states exceed the on-chip capacity is redone through the
// HBM map of the stream by the same CTA (expand_frame + post_epilogue), so results never depend
// on which path ran.  Plain (non-biglm) decoders only.
constexpr int kQuad = 4;                  // arcs per work item: four lanes fetch 64 contiguous bytes
constexpr int kItemBuf = 64;              // work items per warp buffer
constexpr int kStageCap = 64;             // admitted-arc staging ring per warp
constexpr int kWarpScratch = (kItemBuf + kStageCap) * 8;  // bytes per warp
constexpr int kEpsQueueCap = 4096;        // eps-closure worklist entries (u16 slot ids), two buffers
constexpr int kHistBins = 2048;           // cost histogram of GetCutoff (aliases the warp scratch)
constexpr int kCandCap = 2048;            // candidates of the exact k-th selection (aliases the warp scratch)
constexpr uint32_t kFreeCost = 0xFFFFFFFFu;
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr int kMaxProbe = 48;             // buckets probed before a frame is declared too big for the map
constexpr uint32_t kMaxStreamArcs = 1u << 30;  // work items pack (arc index << 2 | count - 1)
static_assert(kHistBins * 4 + kCandCap * 4 <= (kStreamThreads / 32) * kWarpScratch, "histogram + candidates alias the warp scratch");

// shared-memory bytes of k_stream besides the map: log-likelihood row, warp scratch, closure queues
__host__ __device__ constexpr size_t stream_fixed_dyn_bytes(int ll_floats) {
  return (size_t)((ll_floats + 3) & ~3) * 4 + (size_t)(kStreamThreads / 32) * kWarpScratch + 2 * (size_t)kEpsQueueCap * 2;
}

__device__ __forceinline__ uint4 lds_volatile_u4(const uint32_t *p) {
  uint4 r;
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return r;
}
__device__ __forceinline__ uint32_t lds_volatile_u32(const uint32_t *p) {
  return *reinterpret_cast<const volatile uint32_t *>(p);
}

struct SmemMap {
  uint32_t *key;    // [4 * n_buckets] state | kDestEpsBit, kEmptyKey when free
  uint32_t *cost;   // [4 * n_buckets] ordered cost, kFreeCost when free
  uint32_t n_buckets;
};

// FindOrAddToken's lookup half (inl.h:88-136) on the on-chip map: slot of key dstw, claiming a
// free one if the key is new (n_new counts the claims of this thread).  Buckets of four keys —
// one 16-byte shared load compares four slots —, double hashing between buckets.  A key lives
// in the first bucket of its probe sequence that had a free slot when it was inserted; slots are
// never freed within a frame, so a lookup may stop at the first bucket that still has one.
// kNoSlot: the probe budget ran out (the frame is then redone through the HBM map).
__device__ __forceinline__ uint32_t smem_find_or_claim(const SmemMap &m, uint32_t dstw, uint32_t &n_new) {
  const uint32_t hsh = (dstw & kStateMask) * 0x9E3779B1u;
  uint32_t b = __umulhi(hsh, m.n_buckets);
  const uint32_t step = ((hsh >> 4) & 31u) + 1u;
#pragma unroll 1
  for (int probe = 0; probe < kMaxProbe; ++probe) {
    const uint4 kk = lds_volatile_u4(&m.key[b * 4]);
    if (kk.x == dstw) return b * 4;
    if (kk.y == dstw) return b * 4 + 1;
    if (kk.z == dstw) return b * 4 + 2;
    if (kk.w == dstw) return b * 4 + 3;
    const int emp = kk.x == kEmptyKey ? 0 : kk.y == kEmptyKey ? 1 : kk.z == kEmptyKey ? 2 : kk.w == kEmptyKey ? 3 : -1;
    if (emp >= 0) {
      const uint32_t old = atomicCAS(&m.key[b * 4 + emp], kEmptyKey, dstw);
      if (old == kEmptyKey) {
        ++n_new;
        return b * 4 + emp;
      }
      if (old == dstw) return b * 4 + emp;
      continue;  // somebody else's key took the slot: look at the bucket again
    }
    b += step;
    if (b >= m.n_buckets) b -= m.n_buckets;
  }
  return kNoSlot;
}

template <bool SMEM_LL>
__global__ void __launch_bounds__(kStreamThreads, 1)
k_stream(StreamState *const *streams, const AdvanceParams *params, GraphView g, DecoderConfigDev cfg,
         int num_indices, uint32_t n_buckets) {
  constexpr int NT = kStreamThreads;
  constexpr int NW = NT / 32;
  constexpr int U = 2;  // work-item batches (8 items = 32 arc slots each) in flight per warp
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ PostSmem ps;
  __shared__ FrameDesc s_d;
  __shared__ uint32_t s_claims, s_overflow, s_best_ord, s_best_state, s_alive, s_ncand;
  __shared__ uint32_t s_qn[3];  // closure worklist lengths, indexed by round % 3
  __shared__ struct {  // GetCutoff result of the next frame, computed on chip
    unsigned long long best;
    float cur, abeam;
    uint32_t n, off;
  } s_h;
  const uint32_t n_slots = n_buckets * 4;
  SmemMap m;
  m.key = reinterpret_cast<uint32_t *>(s_dyn);
  m.cost = m.key + n_slots;
  m.n_buckets = n_buckets;
  float *s_ll = reinterpret_cast<float *>(m.cost + n_slots);
  unsigned char *s_scratch = reinterpret_cast<unsigned char *>(s_ll + (SMEM_LL ? ((num_indices + 3) & ~3) : 0));
  uint16_t *s_eq = reinterpret_cast<uint16_t *>(s_scratch + NW * kWarpScratch);  // [2][kEpsQueueCap]
  uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_scratch);                    // [kHistBins]   (write-out phase only)
  uint32_t *s_cand = s_hist + kHistBins;                                         // [kCandCap]
  // (test hook: a smaller on-chip budget forces overflow frames)
  const uint32_t claim_limit = (cfg.debug_flags >> 8) ? min((uint32_t)(cfg.debug_flags >> 8), n_slots - n_slots / 16)
                                                       : n_slots - n_slots / 16;
  StreamState *st = streams[blockIdx.x];
  FrameDesc *d = &s_d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const LmPair lms = {};
  // The overflow flag carries WHEN it was raised (1 = expansion, r + 2 = closure round r; first
  // writer wins): a warp that is already one phase ahead must not change the decision the slower
  // warps are still taking about the phase behind it (the decisions have to be uniform).
  auto raise_overflow = [&](uint32_t tag) { atomicCAS(&s_overflow, 0u, tag); };
  // this launch decodes the rows of ONE staged chunk: later chunks may already be raising
  // target_frame while their rows are still being copied
  const int limit = min(params[blockIdx.x].frame0 + params[blockIdx.x].n_frames, st->max_frames);

  for (uint32_t i = tid; i < n_slots; i += NT) {
    m.key[i] = kEmptyKey;
    m.cost[i] = kFreeCost;
  }
  if (tid == 0) s_d.stepping = 0;
  __syncthreads();

  // per-phase SM cycles (diagnostic): accumulated in shared memory by thread 0 right after a
  // barrier, written back once when the launch ends
  __shared__ unsigned long long s_phase[6];
  long long tph = clock64();
  if (tid < 6) s_phase[tid] = 0;
  auto phase = [&](int k) {
    if (tid == 0) {
      const long long now = clock64();
      s_phase[k] += (unsigned long long)(now - tph);
      tph = now;
    }
  };
  // Frame index and, after an on-chip frame, the GetCutoff result of the NEXT frame (computed from
  // the survivors while they are written out): no HBM round trip per frame.
  int t = st->frame;
  bool have_cut = false;
  const float *const ll_hist = st->ll_hist;
  const int ll_stride = st->ll_stride;
  for (;;) {
    if (t >= limit) break;
    if (tid == 0) {
      s_claims = 0;
      s_overflow = (cfg.debug_flags & 8) ? 1u : 0u;  // test hook: every frame through the HBM map
      s_best_ord = 0xFFFFFFFFu;
      s_best_state = 0xFFFFFFFFu;
      s_alive = 0;
      s_ncand = 0;
      s_qn[0] = s_qn[1] = s_qn[2] = 0;
    }
    if (t + 1 < limit) {  // the next frame's row: into L2 while this frame is searched
      const char *nxt = reinterpret_cast<const char *>(ll_hist + (size_t)(t + 1) * ll_stride);
      if (tid * 128 < num_indices * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + tid * 128));
    }
    if (!have_cut) {
      // ---- first frame of the launch / after an HBM-map frame: GetCutoff + best-token pre-pass
      // over the arena tokens; descriptor of the step into shared memory
      cutoff_prologue<NT, false>(st, d, g, cfg, lms, ps.red64, ps.red32, ps.hist, ps.misc);
      __syncthreads();
      if (!s_d.stepping) break;  // uniform: frame == target_frame
      if (SMEM_LL) {
        const float *__restrict__ llr = s_d.ll;
        for (int c = tid; c < num_indices; c += NT) s_ll[c] = __ldcs(&llr[c]);
        __syncthreads();
      }
    } else {
      // ---- cutoff known: stage the row, then the best-token pre-pass (inl.h:282-300) reads it
      // from shared memory
      const float *__restrict__ llr = ll_hist + (size_t)t * ll_stride;
      const uint32_t h_n = s_h.n;
      const float h_abeam = s_h.abeam;
      const unsigned long long h_best = s_h.best;
      if (tid == 0) fill_desc(st, d, t, h_n, s_h.off, s_h.cur, h_abeam, kOrdInf, false);
      if (SMEM_LL)
        for (int c = tid; c < num_indices; c += NT) s_ll[c] = __ldcs(&llr[c]);
      uint32_t mn = kOrdInf;
      if (h_n > 0) {
        const float bc = ord2f((uint32_t)(h_best >> 32));
        const uint2 er = __ldg(&g.erows[(uint32_t)h_best]);
        if (SMEM_LL) __syncthreads();
        for (uint32_t a = er.x + tid; a < er.y; a += NT) {
          const int4 arc = __ldg(&g.arcs[a]);
          const float tot = bc + __int_as_float(arc.z) - (SMEM_LL ? s_ll[arc.x - 1] : __ldg(&llr[arc.x - 1]));
          mn = min(mn, f2ord(tot + h_abeam));
        }
      } else if (SMEM_LL) {
        __syncthreads();
      }
      const unsigned long long m64 = block_min_u64<NT>((unsigned long long)mn, ps.red64);
      if (tid == 0) s_d.next_cut_bits = (uint32_t)m64;
      __syncthreads();
    }
    phase(0);
    const float *__restrict__ ll = s_d.ll;
    const uint32_t n_cur = s_d.n_cur;
    const uint32_t n_groups = (n_cur + 31) >> 5;
    const uint2 *__restrict__ toks = s_d.toks;
    const float cur_cut = s_d.cur_cut, abeam = s_d.abeam;
    uint32_t *next_cut = &s_d.next_cut_bits;
    uint32_t my_best = 0xFFFFFFFFu;  // lowest cost this warp has reported to s_best_ord

    // ---- emitting expansion (ProcessEmitting, inl.h:311-347) into the on-chip map.
    // A warp takes a group of 32 tokens: lane i loads token i and its emitting span and cuts the
    // span into work items of up to four consecutive arcs, written to a warp-private buffer at
    // the lane's prefix-sum offset.  Item k is then fetched by the four lanes 4k..4k+3 — 64
    // contiguous bytes, one 16-byte LDG.128 per lane —, so mapping a lane to its arc costs one
    // 8-byte shared load instead of a binary search over the prefix sums.  About a third of the
    // arcs pass the running cutoff; they are staged in a warp-private ring and go to the map 32
    // at a time with every lane busy.
    {
      uint2 *ibuf = reinterpret_cast<uint2 *>(s_scratch + warp * kWarpScratch);
      uint2 *ring = ibuf + kItemBuf;
      uint32_t head = 0, n_staged = 0;
      uint32_t expanded = 0, admitted = 0;
      auto flush = [&](uint32_t cnt) {  // the first cnt (<= 32) staged arcs go to the map
        const bool act = (uint32_t)lane < cnt;
        uint2 e = make_uint2(0u, 0xFFFFFFFFu);
        if (act) e = ring[(head + lane) & (kStageCap - 1)];
        head += cnt;
        n_staged -= cnt;
        uint32_t n_new = 0, slot = kNoSlot;
        if (act) slot = smem_find_or_claim(m, e.x, n_new);
        const uint32_t wmin = __reduce_min_sync(kFull, e.y);
        if (wmin < my_best) {  // best cost of the frame (inl.h:169-179), tracked where costs enter the map
          my_best = wmin;
          if (lane == 0) atomicMin(&s_best_ord, wmin);
        }
        bool seed = false;
        if (slot != kNoSlot) {
          atomicMin(&m.cost[slot], e.y);
          seed = n_new != 0 && (e.x & kDestEpsBit) != 0;  // a new state with eps arcs: closure seed (inl.h:376-381)
        }
        const unsigned sm = __ballot_sync(kFull, seed);
        const unsigned fm = __ballot_sync(kFull, act && slot == kNoSlot);
        const uint32_t nn = __reduce_add_sync(kFull, n_new);
        uint32_t qb = 0;
        if (lane == 0) {
          if ((nn && atomicAdd(&s_claims, nn) + nn > claim_limit) || fm) raise_overflow(1u);
          if (sm) qb = atomicAdd(&s_qn[0], (uint32_t)__popc(sm));
        }
        if (sm) {
          qb = __shfl_sync(kFull, qb, 0);
          if (seed) {
            const uint32_t qi = qb + (uint32_t)__popc(sm & lt_mask);
            if (qi < (uint32_t)kEpsQueueCap) s_eq[qi] = (uint16_t)slot;
            else raise_overflow(1u);
          }
        }
        __syncwarp();
      };
      float nc = ord2f(lds_volatile_u32(next_cut));
      // Two more groups are in flight behind the one being walked: the tokens of group +2 and the
      // emitting-arc spans of group +1 (the span load needs the token's state), so a new group
      // starts without waiting on HBM.
      uint32_t t1_cost = 0, t1_base = 0, t1_deg = 0;  // next group: cost, span
      uint2 t2 = make_uint2(0, 0);                    // group after that: {state, cost}
      bool t2_ok = false;
      auto load_tokens = [&](uint32_t grp) {
        const uint32_t i = grp * 32 + lane;
        t2_ok = grp < n_groups && i < n_cur;
        if (t2_ok) t2 = __ldcg(&toks[i]);  // written by this kernel one frame ago: no ld.global.nc
      };
      auto load_spans = [&]() {
        t1_cost = t2.y;
        t1_base = 0;
        t1_deg = 0;
        if (t2_ok && __uint_as_float(t2.y) <= cur_cut) {  // inclusive, inl.h:315
          const uint2 er = __ldg(&g.erows[t2.x]);
          t1_base = er.x;
          t1_deg = er.y - er.x;
        }
      };
      load_tokens(warp);
      load_spans();
      load_tokens(warp + NW);
      for (uint32_t grp = warp; grp < n_groups; grp += NW) {
        // warp-uniform decision (the lanes may not have reconverged after the map updates)
        if (__any_sync(kFull, lds_volatile_u32(&s_overflow) != 0u)) break;
        // running cutoff (inl.h:330): other warps' tightenings arrive once per group
        nc = fminf(nc, ord2f(lds_volatile_u32(next_cut)));
        const uint32_t cost_bits = t1_cost, base = t1_base, deg = t1_deg;
        load_spans();
        load_tokens(grp + 2 * NW);
        const uint32_t nit = (deg + kQuad - 1) / kQuad;
        const uint32_t incl = warp_incl_scan(nit, lane);
        const uint32_t off = incl - nit;
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        expanded += deg;
        for (uint32_t ib = 0; ib < total; ib += kItemBuf) {
          // ---- this lane's items with index in [ib, ib + kItemBuf) -> buffer
          for (uint32_t k = off < ib ? ib - off : 0u; k < nit && off + k < ib + kItemBuf; ++k) {
            const uint32_t left = deg - k * kQuad;
            ibuf[off + k - ib] = make_uint2(((base + k * kQuad) << 2) | ((left < (uint32_t)kQuad ? left : (uint32_t)kQuad) - 1u),
                                            cost_bits);
          }
          __syncwarp();
          const uint32_t nbuf = total - ib < (uint32_t)kItemBuf ? total - ib : (uint32_t)kItemBuf;
          for (uint32_t i0 = 0; i0 < nbuf; i0 += 8 * U) {
            bool in[U];
            float tcost[U];
            int4 arc[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const uint32_t it = i0 + u * 8 + (lane >> 2);
              uint2 item = make_uint2(0u, 0u);
              const bool have = it < nbuf;
              if (have) item = ibuf[it];
              const uint32_t r = lane & 3u;
              in[u] = have && r <= (item.x & 3u);
              tcost[u] = __uint_as_float(item.y);
              arc[u] = make_int4(1, 0, 0, 0);
              if (in[u]) arc[u] = __ldg(&g.arcs[(item.x >> 2) + r]);
            }
            float tot[U];
            bool adm[U];
            uint32_t cand_bits = 0xFFFFFFFFu;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const float ac = -(SMEM_LL ? s_ll[arc[u].x - 1] : __ldg(&ll[arc[u].x - 1]));
              tot[u] = (tcost[u] + ac) + __int_as_float(arc[u].z);  // inl.h:326-329
              adm[u] = in[u] && tot[u] < nc;                        // inl.h:330
              const float cand = tot[u] + abeam;                    // inl.h:332-333
              if (adm[u] && cand < nc) cand_bits = min(cand_bits, f2ord(cand));
            }
            if (__any_sync(kFull, cand_bits != 0xFFFFFFFFu)) {
              const uint32_t wmin = __reduce_min_sync(kFull, cand_bits);
              if (lane == 0) atomicMin(next_cut, wmin);
              nc = fminf(nc, ord2f(wmin));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const unsigned am = __ballot_sync(kFull, adm[u]);
              if (am == 0) continue;
              if (adm[u])
                ring[(head + n_staged + (uint32_t)__popc(am & lt_mask)) & (kStageCap - 1)] =
                    make_uint2((uint32_t)arc[u].w, f2ord(tot[u]));
              const uint32_t nnew = (uint32_t)__popc(am);
              n_staged += nnew;
              admitted += nnew;
              __syncwarp();
              if (n_staged >= 32u) flush(32u);
            }
          }
          __syncwarp();  // every lane is done with the item buffer
        }
      }
      if (n_staged) flush(n_staged);  // the last, partial set (n_staged < 32)
      expanded = __reduce_add_sync(kFull, expanded);
      if (lane == 0 && expanded) {
        atomicAdd(&s_d.arcs_expanded, expanded);
        atomicAdd(&s_d.arcs_admitted, admitted);
      }
    }
    __syncthreads();
    phase(1);
    const float nc = ord2f(s_d.next_cut_bits);  // the FINAL next_cutoff of this frame
    const uint32_t nc_ord = s_d.next_cut_bits;

    // ---- eps closure (ProcessNonemitting, inl.h:353-431) over worklists of slots: round r relaxes
    // the eps arcs of the slots queued for it (round 0: every new state with eps arcs, queued when
    // it was claimed) and queues the destinations whose cost it lowered for round r + 1.  A slot
    // whose cost is lowered twice is queued twice; relaxing is idempotent.  The min-plus fixed
    // point is unique, so the costs equal the reference's LIFO order.
    if (const uint32_t ovf0 = lds_volatile_u32(&s_overflow); !(ovf0 != 0u && ovf0 <= 1u)) {
      // (the warp scratch is idle from here on: clear the cost histogram of the write-out phase)
      for (int i = tid; i < kHistBins; i += NT) s_hist[i] = 0;
      for (uint32_t round = 0;; ++round) {
        const uint32_t nq = min(s_qn[round % 3u], (uint32_t)kEpsQueueCap);
        const uint16_t *qin = s_eq + (round & 1u) * kEpsQueueCap;
        uint16_t *qout = s_eq + ((round + 1u) & 1u) * kEpsQueueCap;
        uint32_t *qn_out = &s_qn[(round + 1u) % 3u];
        for (uint32_t i = tid; i < nq; i += NT) {
          const uint32_t sl = qin[i];
          const uint32_t co = lds_volatile_u32(&m.cost[sl]);
          if (!(co < nc_ord)) continue;  // inl.h:391
          const float cost = ord2f(co);
          const uint2 r = __ldg(&g.rows[m.key[sl] & kStateMask]);
          for (uint32_t a = r.x; a < r.y; ++a) {
            const int4 arc = __ldg(&g.arcs[a]);
            const float tot = cost + __int_as_float(arc.z);  // inl.h:413-414
            if (!(tot < nc)) continue;                       // inl.h:415
            const uint32_t to = f2ord(tot);
            uint32_t n_new = 0;
            const uint32_t s2 = smem_find_or_claim(m, (uint32_t)arc.w, n_new);
            if (s2 == kNoSlot) {
              raise_overflow(round + 2u);
              continue;
            }
            if (n_new && atomicAdd(&s_claims, 1u) + 1u > claim_limit) raise_overflow(round + 2u);
            const uint32_t old = atomicMin(&m.cost[s2], to);
            if (to < old) {  // cost changed (inl.h:115-127)
              atomicMin(&s_best_ord, to);
              if ((uint32_t)arc.w & kDestEpsBit) {  // inl.h:425-426
                const uint32_t qi = atomicAdd(qn_out, 1u);
                if (qi < (uint32_t)kEpsQueueCap) qout[qi] = (uint16_t)s2;
                else raise_overflow(round + 2u);
              }
            }
          }
        }
        // one barrier per round: round r reads s_qn[r % 3] and raises s_qn[(r + 1) % 3]; the
        // counter round r + 1 raises is lowered here — its last readers passed the previous barrier
        if (tid == 0) s_qn[(round + 2u) % 3u] = 0;
        __syncthreads();
        const uint32_t ovf = lds_volatile_u32(&s_overflow);
        if (s_qn[(round + 1u) % 3u] == 0 || (ovf != 0u && ovf <= round + 2u)) break;
      }
    }
    phase(2);

    if (s_overflow) {
      // ---- too many distinct destinations for the on-chip map: wipe it and redo the frame
      // through the stream's HBM map (identical results; the running cutoff stays valid)
      __syncthreads();
      for (uint32_t i = tid; i < n_slots; i += NT) {
        m.key[i] = kEmptyKey;
        m.cost[i] = kFreeCost;
      }
      if (tid == 0) {
        s_d.arcs_expanded = 0;
        s_d.arcs_admitted = 0;
        st->tot_fallback_frames += 1;
      }
      __syncthreads();
      // (two arc steps in flight per lane: a single CTA is latency-bound on the HBM map)
      expand_frame<2, SMEM_LL, false>(d, g, s_ll, (uint32_t)warp, NW, 0, lms);
      __syncthreads();
      post_epilogue<false>(st, d, g, cfg, lms, ps);
      phase(5);
      ++t;
      have_cut = false;
      continue;
    }

    // ---- the frame's tokens are final.  ONE pass over the map: survivors (cost < next_cutoff) are
    // appended to the token arena, their costs binned for GetCutoff, the best token's state
    // resolved (lowest cost — known from the tracking above —, ties -> lowest state id,
    // inl.h:169-179), the slots recycled.
    {
      const uint32_t best_ord = s_best_ord;
      const float bc = ord2f(best_ord);
      // GetCutoff (inl.h:138-234) needs the exact max_active-th smallest cost when more tokens
      // than that survive.  All survivors lie in [best, next_cutoff): a 2048-bin histogram over
      // that range (monotone in the cost) locates the bin holding that rank; its few members are
      // then selected exactly.  Only when at least max_active states were claimed at all.
      const bool want_hist = s_claims > (uint32_t)cfg.max_active && best_ord != 0xFFFFFFFFu && nc < CUDART_INF_F;
      const float h_scale = want_hist ? (float)kHistBins / fmaxf(nc - bc, 1e-6f) : 0.f;
      auto bin_of = [&](float c) -> uint32_t {
        const float x = (c - bc) * h_scale;
        return x >= (float)(kHistBins - 1) ? (uint32_t)(kHistBins - 1) : (uint32_t)(int)fmaxf(x, 0.f);
      };
      const uint32_t cap = s_d.out_cap;
      uint2 *out_sc = s_d.out_sc;
      const uint32_t rows = n_slots >> 5;  // n_slots is a multiple of 32
      for (uint32_t row = warp; row < rows; row += NW) {
        const uint32_t slot = row * 32 + lane;
        const uint32_t kw = m.key[slot];
        if (!__any_sync(kFull, kw != kEmptyKey)) continue;
        bool alive = false;
        uint32_t co = kFreeCost;
        if (kw != kEmptyKey) {
          co = m.cost[slot];
          alive = co < nc_ord;
          m.key[slot] = kEmptyKey;
          m.cost[slot] = kFreeCost;
        }
        const unsigned am = __ballot_sync(kFull, alive);
        if (am == 0) continue;
        uint32_t pos = 0;
        if (lane == 0) pos = atomicAdd(&s_alive, (uint32_t)__popc(am));
        pos = __shfl_sync(kFull, pos, 0);
        if (alive) {
          const float c = ord2f(co);
          const uint32_t idx = pos + (uint32_t)__popc(am & lt_mask);
          if (idx < cap) out_sc[idx] = make_uint2(kw & kStateMask, __float_as_uint(c));
          if (want_hist) atomicAdd(&s_hist[bin_of(c)], 1u);
          if (co == best_ord) atomicMin(&s_best_state, kw & kStateMask);
        }
      }
      __syncthreads();
      phase(3);
      const uint32_t n_alive = s_alive;
      const unsigned long long best64 =
          n_alive ? (((unsigned long long)best_ord << 32) | s_best_state) : kInfVal;
      const uint32_t n_kept = n_alive < cap ? n_alive : cap;
      float n_cur_cut = CUDART_INF_F, n_abeam = cfg.beam;
      const float beam_cut = bc + cfg.beam;  // inl.h:182
      if (n_alive == 0) {
        // nothing survived: the next frame has nothing to expand
      } else if (n_alive <= (uint32_t)cfg.min_active && n_alive <= (uint32_t)cfg.max_active) {
        n_cur_cut = CUDART_INF_F;  // inl.h:183,205,220-226
        n_abeam = CUDART_INF_F;
      } else if (nc <= beam_cut && n_alive <= (uint32_t)cfg.max_active) {
        n_cur_cut = beam_cut;  // every token is below best + beam and max_active does not bind (inl.h:227-232)
        n_abeam = cfg.beam;
      } else if (nc <= beam_cut && want_hist && n_kept == n_alive) {
        // ---- sorted[max_active] over the survivors (inl.h:188-203): bin of that rank, then exact
        const uint32_t k = (uint32_t)cfg.max_active;
        uint32_t total = 0;
        const uint32_t h0 = s_hist[2 * tid], h1 = s_hist[2 * tid + 1];
        const uint32_t ex = block_exclusive_scan<NT>(h0 + h1, ps.red32, total);
        if (k >= ex && k < ex + h0 + h1) {
          const bool second = k >= ex + h0;
          ps.misc[0] = 2 * tid + (second ? 1 : 0);
          ps.misc[1] = k - ex - (second ? h0 : 0u);
          ps.misc[2] = second ? h1 : h0;
        }
        __syncthreads();
        const uint32_t kbin = ps.misc[0], kk = ps.misc[1], kcount = ps.misc[2];
        if (kcount <= (uint32_t)kCandCap) {
          for (uint32_t i = tid; i < n_alive; i += NT) {
            const float c = __uint_as_float(__ldcg(&out_sc[i]).y);
            if (bin_of(c) == kbin) s_cand[atomicAdd(&s_ncand, 1u)] = f2ord(c);
          }
          __syncthreads();
          n_cur_cut = block_kth_smallest<NT>([&](uint32_t i) { return s_cand[i]; }, kcount, kk, best_ord, 0xFFFFFFFFu,
                                             nc_ord - best_ord, ps.hist, ps.misc);
        } else {  // (a degenerate cost distribution: select over all survivors)
          n_cur_cut = block_kth_smallest<NT>([&](uint32_t i) { return f2ord(__uint_as_float(__ldcg(&out_sc[i]).y)); },
                                             n_alive, k, best_ord, 0xFFFFFFFFu, nc_ord - best_ord, ps.hist, ps.misc);
        }
        n_abeam = n_cur_cut - bc + cfg.beam_delta;
      } else {
        // ---- general case (the adaptive beam of this frame was wider than the beam, or the arena
        // is full): GetCutoff over the survivors in the arena
        get_cutoff<NT>([&](uint32_t i) { return f2ord(__uint_as_float(__ldcg(&out_sc[i]).y)); }, n_kept, n_kept,
                       best_ord, cfg, ps.red32, ps.hist, ps.misc, n_cur_cut, n_abeam);
      }
      if (tid == 0) {
        s_h.off = st->frame_off[t + 1];
        s_h.n = n_kept;
        s_h.best = best64;
        s_h.cur = n_cur_cut;
        s_h.abeam = n_abeam;
        frame_commit(st, d, cfg, nc, n_alive, best64);
      }
      have_cut = true;
      ++t;
    }
    __syncthreads();
    phase(4);
  }
  __syncthreads();
  if (tid < 6) st->phase_cycles[tid] += s_phase[tid];
}

}  // namespace asrd
