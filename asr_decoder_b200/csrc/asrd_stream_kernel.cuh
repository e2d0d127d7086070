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
#ifndef ASRD_ITEM_ARCS
#define ASRD_ITEM_ARCS 4
#endif

constexpr int kItemArcs = ASRD_ITEM_ARCS;  // arcs per work item (2 or 4): that many lanes fetch one item's consecutive records
constexpr int kBatch = 32 / kItemArcs;    // items per step (32 arc slots)
constexpr int kItemRing = 8 * kBatch;     // work-item ring per warp (refilled when fewer than two steps are left)
constexpr int kStageCap = 64;             // admitted-arc staging ring per warp
constexpr int kWarpScratch = (kItemRing + kStageCap) * 8;  // bytes per warp
constexpr int kEpsQueueCap = 4096;        // eps-closure worklist entries (u16 slot ids), two buffers
constexpr int kHistBins = 2048;           // cost histogram of GetCutoff (aliases the warp scratch)
constexpr int kCandCap = 2048;            // candidates of the exact k-th selection (aliases the warp scratch)
constexpr uint32_t kFreeCost = 0xFFFFFFFFu;
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr int kMaxProbe = 48;             // buckets probed before a frame is declared too big for the map
constexpr int kItemShift = kItemArcs == 4 ? 2 : 1;
constexpr uint32_t kMaxStreamArcs = 1u << (32 - kItemShift);  // work items pack (arc index << shift | count - 1)
static_assert(kHistBins * 4 + kCandCap * 4 <= (kStreamThreads / 32) * kWarpScratch, "histogram + candidates alias the warp scratch");
static_assert(kItemRing >= 4 * kBatch && kStageCap >= 64, "ring sizes");

// Counters of the frame in flight, at the start of the dynamic shared memory (one base register
// addresses them all; static __shared__ scalars cost a window-address computation per access).
struct StreamHot {
  uint32_t claims;      // distinct states claimed in the map
  uint32_t overflow;    // 0, or WHEN the frame outgrew the map (1 = expansion, r + 2 = closure round r)
  uint32_t best_ord;    // lowest ordered cost that entered the map
  uint32_t best_state;  // lowest state id among the tokens with that cost
  uint32_t alive;       // survivors appended to the arena
  uint32_t ncand;
  uint32_t qn[3];       // closure worklist lengths, indexed by round % 3
  uint32_t next_group;  // token groups handed out to the warps so far (expansion)
  uint32_t pad[6];
};
static_assert(sizeof(StreamHot) == 64, "StreamHot is 64 bytes");

// shared-memory bytes of k_stream besides the map: counters, log-likelihood row, warp scratch, closure queues
__host__ __device__ constexpr size_t stream_fixed_dyn_bytes(int ll_floats) {
  return sizeof(StreamHot) + (size_t)((ll_floats + 3) & ~3) * 4 + (size_t)(kStreamThreads / 32) * kWarpScratch +
         2 * (size_t)kEpsQueueCap * 2;
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
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
// asynchronous 16-byte global -> shared copy (LDGSTS): no register, no scoreboard wait at the request
__device__ __forceinline__ void cp_async16(uint32_t smem_a, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_a), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Shared memory by 32-bit shared-state-space address: the hot loop keeps a handful of base
// addresses in registers and every access is one LDS / STS / ATOMS with a register base (through
// generic pointers each access site recomputes the shared window base first).
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f32_ro(uint32_t a) {  // read-only within the phase (the log-likelihood row)
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
  uint2 v;
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u2(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t x) {
  asm volatile("st.volatile.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)x) : "memory");
}
__device__ __forceinline__ uint32_t atoms_cas(uint32_t a, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t atoms_min(uint32_t a, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.min.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ void reds_min(uint32_t a, uint32_t val) {
  asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(val) : "memory");
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t a, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(val) : "memory");
  return old;
}

struct SmemMap {
  uint32_t *key;    // [4 * n_buckets] state | kDestEpsBit, kEmptyKey when free
  uint32_t *cost;   // [4 * n_buckets] ordered cost, kFreeCost when free
  uint32_t n_buckets;
  uint32_t key_sa;  // shared-space address of key[]; cost[] follows at key_sa + 16 * n_buckets
};

// FindOrAddToken's lookup half (inl.h:88-136) on the on-chip map: slot of key dstw, claiming a
// free one if the key is new (is_new).  Buckets of four keys — one 16-byte shared load compares
// four slots —, double hashing between buckets.  A key lives in the first bucket of its probe
// sequence that had a free slot when it was inserted; slots are never freed within a frame and
// fill a bucket front to back, so a lookup may stop at the first bucket that still has a free
// slot.  kNoSlot: the probe budget ran out (the frame is then redone through the HBM map).
__device__ __forceinline__ uint32_t smem_find_or_claim(const SmemMap &m, uint32_t dstw, bool &is_new) {
  const uint32_t hsh = (dstw & kStateMask) * 0x9E3779B1u;
  uint32_t b = __umulhi(hsh, m.n_buckets);
  const uint32_t step = ((hsh >> 4) & 31u) + 1u;
#pragma unroll 1
  for (int probe = 0; probe < kMaxProbe; ++probe) {
    const uint32_t ba = m.key_sa + b * 16u;
    const uint4 kk = lds_u4(ba);
    if (kk.x == dstw) return b * 4;
    if (kk.y == dstw) return b * 4 + 1;
    if (kk.z == dstw) return b * 4 + 2;
    if (kk.w == dstw) return b * 4 + 3;
    if (kk.w == kEmptyKey) {  // the bucket has a free slot: the first one
      const uint32_t emp = kk.x == kEmptyKey ? 0u : kk.y == kEmptyKey ? 1u : kk.z == kEmptyKey ? 2u : 3u;
      const uint32_t old = atoms_cas(ba + emp * 4u, kEmptyKey, dstw);
      if (old == kEmptyKey) {
        is_new = true;
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

// CLG: the graph is a materialised CLG graph (asrd_graph_read_clg) and the loop follows the
// reference's CLG decoder (my-decoder/online-clg-decoder-mempool-base.h): tokens are expanded when
// cost < cur_cutoff (strict, :128), an emitting arc is skipped only when ABOVE the cutoff (:156) —
// so tokens AT the final cutoff survive —, and the best-token pre-pass adds the two weights of an
// arc that leaves a CLG state through an HMM one after the other (:91).
template <bool SMEM_LL, bool CLG>
__global__ void __launch_bounds__(kStreamThreads, 1)
k_stream(StreamState *const *streams, const AdvanceParams *params, GraphView g, DecoderConfigDev cfg,
         int num_indices, uint32_t n_buckets) {
  constexpr int NT = kStreamThreads;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ PostSmem ps;
  __shared__ FrameDesc s_d;
  __shared__ struct {  // GetCutoff result of the next frame, computed on chip
    unsigned long long best;
    float cur, abeam;
    uint32_t n, off;
  } s_h;
  const uint32_t n_slots = n_buckets * 4;
  StreamHot *hot = reinterpret_cast<StreamHot *>(s_dyn);
  SmemMap m;
  // fixed-size parts first (compile-time offsets from the base), the map — whose size depends on
  // the launch — last
  constexpr uint32_t kOffScratch = sizeof(StreamHot);
  constexpr uint32_t kOffEq = kOffScratch + NW * kWarpScratch;
  constexpr uint32_t kOffLl = kOffEq + 2 * kEpsQueueCap * 2;
  unsigned char *s_scratch = s_dyn + kOffScratch;
  uint16_t *s_eq = reinterpret_cast<uint16_t *>(s_dyn + kOffEq);  // [2][kEpsQueueCap]
  float *s_ll = reinterpret_cast<float *>(s_dyn + kOffLl);
  m.key = reinterpret_cast<uint32_t *>(s_ll + (SMEM_LL ? ((num_indices + 3) & ~3) : 0));
  m.cost = m.key + n_slots;
  m.n_buckets = n_buckets;
  m.key_sa = smem_addr(m.key);
  uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_scratch);                    // [kHistBins]   (write-out phase only)
  uint32_t *s_cand = s_hist + kHistBins;                                         // [kCandCap]
  // every warp may run ahead of the shared claim counter by what it has not reported yet
  const uint32_t slack = n_slots / 16 > 1024u ? n_slots / 16 : 1024u;
  // (test hook: a smaller on-chip budget forces overflow frames)
  const uint32_t claim_limit = (cfg.debug_flags >> 8) ? min((uint32_t)(cfg.debug_flags >> 8), n_slots - slack)
                                                       : n_slots - slack;
  StreamState *st = streams[blockIdx.x];
  FrameDesc *d = &s_d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const LmPair lms = {};
  // The overflow flag carries WHEN it was raised (first writer wins): a warp that is already one
  // phase ahead must not change the decision the slower warps are still taking about the phase
  // behind it (the decisions have to be uniform).
  auto raise_overflow = [&](uint32_t tag) { atomicCAS(&hot->overflow, 0u, tag); };
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
  bool row_staged = false;  // the row of frame t is already in s_ll (copied behind the previous frame's closure)
  const float *const ll_hist = st->ll_hist;
  const int ll_stride = st->ll_stride;
  for (;;) {
    if (t >= limit) break;
    if (tid == 0) {
      hot->claims = 0;
      hot->overflow = (cfg.debug_flags & 8) ? 1u : 0u;  // test hook: every frame through the HBM map
      hot->best_ord = 0xFFFFFFFFu;
      hot->best_state = 0xFFFFFFFFu;
      hot->alive = 0;
      hot->ncand = 0;
      hot->qn[0] = hot->qn[1] = hot->qn[2] = 0;
      hot->next_group = 0;
    }
    if (t + 1 < limit) {  // the next frame's row: into L2 while this frame is searched
      const char *nxt = reinterpret_cast<const char *>(ll_hist + (size_t)(t + 1) * ll_stride);
      if (tid * 128 < num_indices * 4) prefetch_l2(nxt + tid * 128);
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
      if (SMEM_LL) {
        if (row_staged) cp_async_wait_all();  // (made visible to the other threads by the barrier below)
        else
          for (int c = tid; c < num_indices; c += NT) s_ll[c] = __ldcs(&llr[c]);
      }
      row_staged = false;
      uint32_t mn = kOrdInf;
      if (h_n > 0) {
        const float bc = ord2f((uint32_t)(h_best >> 32));
        const uint2 er = __ldg(&g.erows[(uint32_t)h_best]);
        if (SMEM_LL) __syncthreads();
        for (uint32_t a = er.x + tid; a < er.y; a += NT) {
          const int4 arc = __ldg(&g.arcs[a]);
          const float llv = SMEM_LL ? s_ll[arc.x - 1] : __ldg(&llr[arc.x - 1]);
          float tot = bc + __int_as_float(arc.z) - llv;
          if (CLG && ((__ldg(&g.clg2_bits[a >> 5]) >> (a & 31)) & 1u))
            tot = bc + __ldg(&g.w_clg[a]) + __ldg(&g.w_hmm[a]) - llv;  // …-clg-…-base.h:91
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

    // ---- eps relaxation (ProcessNonemitting's inner step, inl.h:413-426), used by the closure rounds.
    // (Relaxing eagerly inside the expansion — whenever a state with eps arcs gets cheaper — was
    // measured: the closure phase all but vanishes, but the divergent chain walks inside the map
    // update cost three times what they save, 58 vs 48 ms per step.)  relax_one: dstw gets cost + w if that is below `cut`;
    // true when the destination's cost was lowered and it has eps arcs itself.  relax_chain: every
    // eps arc of a state (er: its eps row, first arc inline) from `cost`, then the lowered
    // destinations one after the other (depth first, like the reference's LIFO); when a state fans
    // out to several such destinations the extra ones go to worklist `qsel` for a later round.
    auto relax_one = [&](float cost, float w, uint32_t dstw, float cut, uint32_t tag, float &new_cost, uint32_t &s2) -> bool {
      const float tot = cost + w;  // inl.h:413-414
      if (!(tot < cut)) return false;  // inl.h:415
      const uint32_t to = f2ord(tot);
      bool is_new = false;
      s2 = smem_find_or_claim(m, dstw, is_new);
      if (s2 == kNoSlot) {
        raise_overflow(tag);
        return false;
      }
      if (is_new && atomicAdd(&hot->claims, 1u) + 1u > claim_limit) raise_overflow(tag);
      const uint32_t old = atomicMin(&m.cost[s2], to);
      if (!(to < old)) return false;  // cost unchanged (inl.h:115-127)
      atomicMin(&hot->best_ord, to);
      new_cost = tot;
      return (dstw & kDestEpsBit) != 0;  // inl.h:425-426
    };
    auto relax_chain = [&](float cost, uint4 er, float cut, uint32_t tag, uint32_t qsel) {
      for (int hop = 0; hop < 64; ++hop) {
        uint32_t next_w = 0;
        float next_cost = 0.f;
        bool have_next = false;
        for (uint32_t a = er.x; a < er.y; ++a) {
          float w;
          uint32_t dstw;
          if (a == er.x) {
            w = __uint_as_float(er.z);
            dstw = er.w;
          } else {
            const int4 arc = __ldg(&g.arcs[a]);
            w = __int_as_float(arc.z);
            dstw = (uint32_t)arc.w;
          }
          float c2;
          uint32_t s2;
          if (relax_one(cost, w, dstw, cut, tag, c2, s2)) {
            if (!have_next && hop < 63) {
              have_next = true;
              next_w = dstw;
              next_cost = c2;
            } else {
              const uint32_t qi = atomicAdd(&hot->qn[qsel % 3u], 1u);
              if (qi < (uint32_t)kEpsQueueCap) s_eq[(qsel & 1u) * kEpsQueueCap + qi] = (uint16_t)s2;
              else raise_overflow(tag);
            }
          }
        }
        if (!have_next) return;
        cost = next_cost;
        er = __ldg(&g.eps_rows[next_w & kStateMask]);
      }
    };

    // ---- emitting expansion (ProcessEmitting, inl.h:311-347) into the on-chip map.
    // A warp takes groups of 32 tokens: lane i loads token i and its emitting span and cuts the
    // span into work items of up to kItemArcs consecutive arcs, appended to a warp-private ring at
    // the lane's prefix-sum offset.  The ring is drained one STEP (32 arc slots = kBatch items) at
    // a time — item k by the kItemArcs lanes k * kItemArcs .., each one 16-byte LDG.128 —, so
    // mapping a lane to its arc costs one 8-byte shared load instead of a binary search over the
    // prefix sums, and what a group leaves over is carried into the next one (every step but the
    // last is full).  The loads of step i+1 are issued BEFORE step i is scored and merged, so
    // the HBM / L2 latency of the arc fetch hides behind the shared-memory work.  About a third
    // of the arcs pass the running cutoff; they are staged in a second warp-private ring and go to
    // the map 32 at a time with every lane busy.
    {
      // (Register diet: at 64 registers per thread every spill costs an L2 round trip here — the
      // L1 is all but gone to the shared-memory carve-out.  Only two shared-memory base addresses
      // are kept; flags are folded into values (an arc slot without an arc carries cost +inf) or
      // recomputed from the group ids.)
      const uint32_t base_sa = smem_addr(hot);
      const uint32_t iring_sa = base_sa + kOffScratch + warp * kWarpScratch;
      const uint32_t cost_sa = m.key_sa + n_buckets * 16u;
      constexpr uint32_t kHotClaims = 0, kHotOverflow = 4, kHotBest = 8, kHotQn0 = 24, kHotNextGroup = 36;
      static_assert(offsetof(StreamHot, overflow) == kHotOverflow && offsetof(StreamHot, best_ord) == kHotBest &&
                        offsetof(StreamHot, qn) == kHotQn0 && offsetof(StreamHot, claims) == kHotClaims &&
                        offsetof(StreamHot, next_group) == kHotNextGroup,
                    "StreamHot layout");
      auto sring_at = [&](uint32_t i) { return iring_sa + kItemRing * 8 + ((i & (kStageCap - 1)) << 3); };
      auto iring_at = [&](uint32_t i) { return iring_sa + ((i & (kItemRing - 1)) << 3); };
      auto cost_at = [&](uint32_t slot) { return cost_sa + slot * 4u; };
      auto ll_at = [&](uint32_t c) { return base_sa + kOffLl + c * 4u; };
      uint32_t ihead = 0, itail = 0, shead = 0, stail = 0;
      uint32_t expanded = 0, unreported = 0;
      uint32_t my_best = 0xFFFFFFFFu;  // lowest cost this warp has reported to hot->best_ord
      float nc = ord2f(lds_u32(smem_addr(next_cut)));
      auto flush = [&](uint32_t cnt) {  // the first cnt (<= 32) staged arcs go to the map
        const bool act = (uint32_t)lane < cnt;
        uint2 e = make_uint2(0u, 0xFFFFFFFFu);
        if (act) e = lds_u2(sring_at(shead + lane));
        shead += cnt;
        bool is_new = false;
        if (act) {
          const uint32_t slot = smem_find_or_claim(m, e.x, is_new);
          if (slot != kNoSlot) {
            reds_min(cost_at(slot), e.y);
            if (is_new && (e.x & kDestEpsBit)) {
              // a new state with eps arcs: closure seed (inl.h:376-381); its eps row will be wanted soon
              const uint32_t qi = atomicAdd(&hot->qn[0], 1u);
              if (qi < (uint32_t)kEpsQueueCap) s_eq[qi] = (uint16_t)slot;
              else raise_overflow(1u);
              prefetch_l2(&g.eps_rows[e.x & kStateMask]);
            }
          } else {
            raise_overflow(1u);
          }
        }
        const uint32_t wmin = __reduce_min_sync(kFull, e.y);
        if (wmin < my_best) {  // best cost of the frame (inl.h:169-179), tracked where costs enter the map
          my_best = wmin;
          if (lane == 0) reds_min(base_sa + kHotBest, wmin);
        }
        // claims are reported to the shared counter 32 at a time (the limit leaves that slack)
        unreported += (uint32_t)__popc(__ballot_sync(kFull, is_new));
        if (unreported >= 32u) {
          if (lane == 0 && atoms_add(base_sa + kHotClaims, unreported) + unreported > claim_limit) raise_overflow(1u);
          unreported = 0;
        }
      };
      // a step: its arc records and the cost of the token each one leaves (+inf: no arc in this slot)
      struct Step {
        int4 arc;
        float tcost;
      };
      // (Every lane loads — lanes without an arc re-read their item's last one, or record 0 —: a
      // predicated load with a default value makes ptxas write the default into the load's
      // destination registers AFTER the request, and that write waits for the data: measured,
      // the warp then sits out the whole fetch latency at the request.)
      auto issue = [&](uint32_t cnt) -> Step {  // cnt (<= kBatch) items from the head of the item ring
        Step sp;
        const uint32_t it = (uint32_t)lane >> kItemShift;
        const bool have = it < cnt;
        const uint2 item = lds_u2(iring_at(ihead + (have ? it : 0u)));
        const uint32_t r = (uint32_t)lane & (kItemArcs - 1), last = item.x & (kItemArcs - 1);
        sp.tcost = (have && r <= last) ? __uint_as_float(item.y) : CUDART_INF_F;
        sp.arc = __ldg(&g.arcs[(item.x >> kItemShift) + (r < last ? r : last)]);
        ihead += cnt;
        return sp;
      };
      auto process = [&](const Step &sp) {
        // (a slot without an arc may hold an eps record, ilabel 0: the row index is clamped)
        const int li = sp.arc.x > 0 ? sp.arc.x - 1 : 0;
        const float ac = -(SMEM_LL ? lds_f32_ro(ll_at((uint32_t)li)) : __ldg(&ll[li]));
        const float tot = (sp.tcost + ac) + __int_as_float(sp.arc.z);  // inl.h:326-329
        // (the olabel is not needed, but its register must stay reserved until the record has arrived:
        // ptxas reuses the register of an unused component as a temporary right after the request, and
        // that write waits for the whole fetch.  No olabel is 0x80000001; the test costs one predicate input.)
        const bool adm = (CLG ? tot <= nc : tot < nc) && sp.arc.y != (int)0x80000001;  // inl.h:330 (tot is +inf where there is no arc)
        const float cand = adm ? tot + abeam : CUDART_INF_F;       // inl.h:332-333
        if (__any_sync(kFull, cand < nc)) {
          const uint32_t wmin = __reduce_min_sync(kFull, f2ord(cand));
          if (lane == 0) reds_min(smem_addr(next_cut), wmin);
          nc = fminf(nc, ord2f(wmin));
        }
        const unsigned am = __ballot_sync(kFull, adm);
        if (adm) sts_u2(sring_at(stail + (uint32_t)__popc(am & lanemask_lt())), (uint32_t)sp.arc.w, f2ord(tot));
        stail += (uint32_t)__popc(am);
        __syncwarp();
        if (stail - shead >= 32u) flush(32u);
      };
      // Two steps scored together: both log-likelihood reads and both float chains are in flight at
      // once, one vote / reduction tightens the cutoff for both (the minimum is the same number), and
      // only the staging — which may have to flush in between — stays sequential.
      auto process2 = [&](const Step &a, const Step &b) {
        const int liA = a.arc.x > 0 ? a.arc.x - 1 : 0, liB = b.arc.x > 0 ? b.arc.x - 1 : 0;
        const float acA = -(SMEM_LL ? lds_f32_ro(ll_at((uint32_t)liA)) : __ldg(&ll[liA]));
        const float acB = -(SMEM_LL ? lds_f32_ro(ll_at((uint32_t)liB)) : __ldg(&ll[liB]));
        const float totA = (a.tcost + acA) + __int_as_float(a.arc.z);  // inl.h:326-329
        const float totB = (b.tcost + acB) + __int_as_float(b.arc.z);
        const bool admA = (CLG ? totA <= nc : totA < nc) && a.arc.y != (int)0x80000001;  // inl.h:330
        const bool admB = (CLG ? totB <= nc : totB < nc) && b.arc.y != (int)0x80000001;
        const float cand = fminf(admA ? totA + abeam : CUDART_INF_F, admB ? totB + abeam : CUDART_INF_F);  // inl.h:332-333
        if (__any_sync(kFull, cand < nc)) {
          const uint32_t wmin = __reduce_min_sync(kFull, f2ord(cand));
          if (lane == 0) reds_min(smem_addr(next_cut), wmin);
          nc = fminf(nc, ord2f(wmin));
        }
        const unsigned amA = __ballot_sync(kFull, admA), amB = __ballot_sync(kFull, admB);
        if (admA) sts_u2(sring_at(stail + (uint32_t)__popc(amA & lanemask_lt())), (uint32_t)a.arc.w, f2ord(totA));
        stail += (uint32_t)__popc(amA);
        __syncwarp();
        if (stail - shead >= 32u) flush(32u);
        if (admB) sts_u2(sring_at(stail + (uint32_t)__popc(amB & lanemask_lt())), (uint32_t)b.arc.w, f2ord(totB));
        stail += (uint32_t)__popc(amB);
        __syncwarp();
        if (stail - shead >= 32u) flush(32u);
      };
      // Two more groups are in flight behind the one being cut into items: the tokens of group id2
      // and the emitting-arc spans of group id1 (the span load needs the token's state), so a new
      // group starts without waiting on HBM.  Unconditional loads with clamped indices: a
      // predicated load that keeps the old value needs a move out of the load's destination, which
      // waits for the data.
      uint32_t t1_cost = 0;
      uint2 t1_er = make_uint2(0, 0), t2 = make_uint2(0, 0);
      uint32_t id1, id2;
      auto load_tokens = [&](uint32_t id) {
        const uint32_t i = id * 32 + lane;
        t2 = __ldcg(&toks[i < n_cur ? i : 0u]);  // written by this kernel one frame ago: no ld.global.nc
      };
      auto load_spans = [&](uint32_t id) {  // id: the group whose tokens are in t2
        t1_cost = t2.y;
        t1_er = __ldg(&g.erows[id * 32 + lane < n_cur ? t2.x : 0u]);
      };
      // Token groups are handed out dynamically (one shared counter): the warp schedulers favour
      // some warps over others, and with a fixed share per warp the favoured ones would wait at the
      // barrier while the rest finish alone, with nothing left to hide their latencies behind.
      auto acquire = [&]() -> uint32_t {
        uint32_t id = 0;
        if (lane == 0) id = atoms_add(base_sa + kHotNextGroup, 1u);
        return __shfl_sync(kFull, id, 0);
      };
      id1 = acquire();
      load_tokens(id1);
      load_spans(id1);
      id2 = acquire();
      load_tokens(id2);
      // The group being cut into items keeps its place in registers (a group with more items than
      // the ring has room for is appended slice by slice).
      uint32_t g_cost = 0, g_base = 0, g_deg = 0, g_off = 0, g_total = 0, g_r0 = 0;
      auto top_up = [&]() {  // keep two steps' worth of items in the ring while there are tokens left
        for (;;) {
          const uint32_t avail = itail - ihead;
          if (avail >= 2u * kBatch) return;
          if (g_r0 >= g_total) {
            if (id1 >= n_groups) return;
            // next group of tokens; warp-uniform decision (the lanes may not have reconverged after the map updates)
            if (__any_sync(kFull, lds_u32(base_sa + kHotOverflow) != 0u)) {
              id1 = id2 = 0xFFFFFFFFu;  // the frame is being redone anyway: drop what is queued
              ihead = itail;
              return;
            }
            // running cutoff (inl.h:330): other warps' tightenings arrive once per group
            nc = fminf(nc, ord2f(lds_u32(smem_addr(next_cut))));
            g_cost = t1_cost;
            g_base = t1_er.x;
            // inclusive token test, inl.h:315
            g_deg = (id1 * 32 + lane < n_cur &&
                     (CLG ? __uint_as_float(t1_cost) < cur_cut : __uint_as_float(t1_cost) <= cur_cut))
                        ? t1_er.y - t1_er.x
                        : 0u;
            // the span's arc records are requested a few steps from now: have them in L2 by then
            // (four in ten come from HBM otherwise, and two steps in flight do not cover that;
            // measured 47.0 vs 47.9 ms per step.  Prefetching for the NEXT group instead, one step
            // later when its spans have arrived, was measured too: +0.7 ms)
            if (g_deg) {
              prefetch_l2(&g.arcs[g_base]);
              if (((g_base + g_deg - 1) >> 3) != (g_base >> 3)) prefetch_l2(&g.arcs[g_base + g_deg - 1]);  // (eight per 128-byte line)
            }
            load_spans(id2);
            id1 = id2;
            id2 = acquire();
            load_tokens(id2);
            const uint32_t nit = (g_deg + kItemArcs - 1) / kItemArcs;
            const uint32_t incl = warp_incl_scan(nit, lane);
            g_off = incl - nit;
            g_total = __shfl_sync(kFull, incl, 31);
            g_r0 = 0;
            expanded += g_deg;
            if (g_total == 0) continue;
          }
          // items [g_r0, g_r0 + take) of the group -> ring
          const uint32_t room = (uint32_t)kItemRing - avail;
          const uint32_t take = room < g_total - g_r0 ? room : g_total - g_r0;
          const uint32_t nit = (g_deg + kItemArcs - 1) / kItemArcs;
          for (uint32_t k = g_off < g_r0 ? g_r0 - g_off : 0u; k < nit && g_off + k < g_r0 + take; ++k) {
            const uint32_t left = g_deg - k * kItemArcs;
            sts_u2(iring_at(itail + g_off + k - g_r0),
                   ((g_base + k * kItemArcs) << kItemShift) | ((left < (uint32_t)kItemArcs ? left : (uint32_t)kItemArcs) - 1u), g_cost);
          }
          itail += take;
          g_r0 += take;
          __syncwarp();
        }
      };
      Step sA, sB;
      for (;;) {  // two steps requested back to back, then scored
        top_up();
        const uint32_t avail = itail - ihead;
        if (avail == 0) break;
        sA = issue(avail < (uint32_t)kBatch ? avail : (uint32_t)kBatch);
        if (avail > (uint32_t)kBatch) {
          sB = issue(avail - kBatch < (uint32_t)kBatch ? avail - kBatch : (uint32_t)kBatch);
          process2(sA, sB);  // (measured against process(sA); process(sB): -1.6 % per step)
        } else {
          process(sA);
        }
      }
      while (stail != shead) flush(stail - shead < 32u ? stail - shead : 32u);  // what is still staged
      if (unreported && lane == 0 && atoms_add(base_sa + kHotClaims, unreported) + unreported > claim_limit) raise_overflow(1u);
      expanded = __reduce_add_sync(kFull, expanded);
      if (lane == 0 && expanded) {
        atomicAdd(&s_d.arcs_expanded, expanded);
        atomicAdd(&s_d.arcs_admitted, stail);  // every staged arc was admitted against the running cutoff
      }
    }
    __syncthreads();
    phase(1);
    const float nc = ord2f(s_d.next_cut_bits);  // the FINAL next_cutoff of this frame
    const uint32_t nc_ord = s_d.next_cut_bits;

    // ---- eps closure (ProcessNonemitting, inl.h:353-431) over worklists of slots: round 0 relaxes
    // the eps arcs of every new state with eps arcs (queued when it was claimed).  A thread that
    // lowers the cost of a destination which has eps arcs itself FOLLOWS it at once (depth first, like
    // the reference's LIFO) instead of waiting for the next round: a chain of eps arcs costs one load
    // per hop and no barrier; only when a state fans out to several such destinations do the extra
    // ones go to the worklist of the next round.  A slot whose cost is lowered twice is relaxed
    // twice; relaxing is idempotent.  The min-plus fixed point is unique, so the costs equal the
    // reference's.  The eps row of a state carries its first eps arc inline (most states have
    // one): one load per relaxed state.
    if (const uint32_t ovf0 = lds_volatile_u32(&hot->overflow); !(ovf0 != 0u && ovf0 <= 1u)) {
      // The expansion was the last reader of this frame's log-likelihood row: the next frame's row
      // is copied into its place now, asynchronously, behind the closure and the write-out.
      if (SMEM_LL && t + 1 < limit && (ll_stride & 3) == 0) {
        const float *nrow = ll_hist + (size_t)(t + 1) * ll_stride;
        const uint32_t ll_sa = smem_addr(s_ll);
        for (int c = tid * 4; c < num_indices; c += NT * 4) cp_async16(ll_sa + c * 4, nrow + c);
        row_staged = true;
      }
      for (uint32_t round = 0;; ++round) {
        const uint32_t nq = min(hot->qn[round % 3u], (uint32_t)kEpsQueueCap);
        const uint16_t *qin = s_eq + (round & 1u) * kEpsQueueCap;
        for (uint32_t i0 = tid; i0 < nq; i0 += 4 * NT) {  // up to four worklist entries per thread, loads batched
          float cost[4];
          uint4 er[4];
          bool go[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // (unconditional loads, clamped: see load_tokens)
            const uint32_t i = i0 + j * NT;
            const uint32_t sl = qin[i < nq ? i : 0u];
            const uint32_t co = lds_volatile_u32(&m.cost[sl]);
            go[j] = i < nq && co < nc_ord;  // inl.h:391
            cost[j] = ord2f(co);
            er[j] = __ldg(&g.eps_rows[m.key[sl] & kStateMask]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (go[j]) relax_chain(cost[j], er[j], nc, round + 2u, round + 1u);
        }
        // one barrier per round: round r reads qn[r % 3] and raises qn[(r + 1) % 3]; the
        // counter round r + 1 raises is lowered here — its last readers passed the previous barrier
        if (tid == 0) hot->qn[(round + 2u) % 3u] = 0;
        __syncthreads();
        const uint32_t ovf = lds_volatile_u32(&hot->overflow);
        if (hot->qn[(round + 1u) % 3u] == 0 || (ovf != 0u && ovf <= round + 2u)) break;
      }
    }
    phase(2);

    if (hot->overflow) {
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
      if (SMEM_LL && row_staged) {  // the next frame's row has taken this one's place: fetch it again
        cp_async_wait_all();
        __syncthreads();
        const float *__restrict__ llr = s_d.ll;
        for (int c = tid; c < num_indices; c += NT) s_ll[c] = __ldcs(&llr[c]);
      }
      row_staged = false;
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

    // ---- the frame's tokens are final.  ONE pass over the map, a bucket (four slots, two 16-byte
    // shared loads) per lane: survivors (cost < next_cutoff) are appended to the token arena, the
    // best token's state resolved (lowest cost — known from the tracking above —, ties -> lowest
    // state id, inl.h:169-179), the slots recycled.
    {
      const uint32_t best_ord = hot->best_ord;
      const float bc = ord2f(best_ord);
      const uint32_t cap = s_d.out_cap;
      uint2 *out_sc = s_d.out_sc;
      // GetCutoff (inl.h:138-234) needs the exact max_active-th smallest cost when more tokens than
      // that survive: only possible when more states than that were claimed at all
      const bool want_hist = hot->claims > (uint32_t)cfg.max_active && best_ord != 0xFFFFFFFFu && nc < CUDART_INF_F;
      const uint32_t blocks = n_buckets >> 5;  // n_buckets is a multiple of 32
      // Fewer claimed states than buckets (the peaked / narrow-beam regimes): most buckets are empty.
      // Each warp first gathers the occupied buckets of its share into its (idle) scratch and then
      // walks those, 32 at a time, instead of every bucket (-3 % per step at beam 10, -5 % with peaked
      // scores; nothing changes for the busy frames, which take the other branch).
      const bool sparse = !want_hist && hot->claims * 2u < n_buckets;
      uint32_t n_occ = 0;
      const uint32_t occ_sa = smem_addr(s_scratch + warp * kWarpScratch);
      if (sparse) {
        for (uint32_t blk = warp; blk < blocks; blk += NW) {
          const uint32_t bkt = blk * 32 + lane;
          const bool occ = m.key[bkt * 4] != kEmptyKey;
          const unsigned om = __ballot_sync(kFull, occ);
          if (occ) sts_u16(occ_sa + (n_occ + (uint32_t)__popc(om & lanemask_lt())) * 2u, bkt);
          n_occ += (uint32_t)__popc(om);
        }
        __syncwarp();
      } else {
        for (int i = tid; i < kHistBins; i += NT) s_hist[i] = 0;  // (the warp scratch is idle from here on)
      }
      const uint32_t n_iter = sparse ? (n_occ + 31) >> 5 : (blocks > (uint32_t)warp ? (blocks - warp + NW - 1) / NW : 0u);
      for (uint32_t it = 0; it < n_iter; ++it) {
        uint32_t s0;
        bool have = true;
        if (sparse) {
          have = it * 32 + lane < n_occ;
          unsigned short bk = 0;
          if (have) asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=h"(bk) : "r"(occ_sa + (it * 32 + lane) * 2u) : "memory");
          s0 = (uint32_t)bk * 4;
        } else {
          s0 = ((warp + it * NW) * 32 + lane) * 4;
        }
        const uint4 kw = have ? *reinterpret_cast<const uint4 *>(&m.key[s0]) : make_uint4(kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey);
        uint32_t cnt = 0;
        uint4 co = make_uint4(kFreeCost, kFreeCost, kFreeCost, kFreeCost);
        if (kw.x != kEmptyKey) {  // (slots fill a bucket front to back)
          co = *reinterpret_cast<const uint4 *>(&m.cost[s0]);
          *reinterpret_cast<uint4 *>(&m.key[s0]) = make_uint4(kEmptyKey, kEmptyKey, kEmptyKey, kEmptyKey);
          *reinterpret_cast<uint4 *>(&m.cost[s0]) = make_uint4(kFreeCost, kFreeCost, kFreeCost, kFreeCost);
          // (CLG: an arc AT the cutoff was admitted; a free slot's cost is 0xFFFFFFFF, above +inf's key)
          const uint32_t lim = CLG ? nc_ord + (nc_ord != 0xFFFFFFFFu ? 1u : 0u) : nc_ord;
          cnt = (co.x < lim) + (co.y < lim) + (co.z < lim) + (co.w < lim);
        }
        const uint32_t incl = warp_incl_scan(cnt, lane);
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        if (total == 0) continue;
        uint32_t pos = 0;
        if (lane == 0) pos = atomicAdd(&hot->alive, total);
        pos = __shfl_sync(kFull, pos, 0) + incl - cnt;
        auto put = [&](uint32_t k, uint32_t c) {
          if (CLG ? (c <= nc_ord && c != kFreeCost) : c < nc_ord) {
            if (pos < cap) out_sc[pos] = make_uint2(k & kStateMask, __float_as_uint(ord2f(c)));
            ++pos;
            if (c == best_ord) atomicMin(&hot->best_state, k & kStateMask);
          }
        };
        if (cnt) {
          put(kw.x, co.x);
          put(kw.y, co.y);
          put(kw.z, co.z);
          put(kw.w, co.w);
        }
      }
      __syncthreads();
      phase(3);
      const uint32_t n_alive = hot->alive;
      const unsigned long long best64 =
          n_alive ? (((unsigned long long)best_ord << 32) | hot->best_state) : kInfVal;
      const uint32_t n_kept = n_alive < cap ? n_alive : cap;
      float n_cur_cut = CUDART_INF_F, n_abeam = cfg.beam;
      const float beam_cut = bc + cfg.beam;  // inl.h:182
      auto arena_ord = [&](uint32_t i) { return f2ord(__uint_as_float(__ldcg(&out_sc[i]).y)); };
      if (n_alive == 0) {
        // nothing survived: the next frame has nothing to expand
      } else if (n_alive <= (uint32_t)cfg.min_active && n_alive <= (uint32_t)cfg.max_active) {
        n_cur_cut = CUDART_INF_F;  // inl.h:183,205,220-226
        n_abeam = CUDART_INF_F;
      } else if (nc <= beam_cut && n_alive <= (uint32_t)cfg.max_active) {
        n_cur_cut = beam_cut;  // every token is below best + beam and max_active does not bind (inl.h:227-232)
        n_abeam = cfg.beam;
      } else if (nc <= beam_cut && want_hist && n_kept == n_alive) {
        // ---- sorted[max_active] over the survivors (inl.h:188-203).  All of them lie in
        // [best, next_cutoff): a 2048-bin histogram over that range (monotone in the cost) locates
        // the bin holding that rank; its few members are then selected exactly.  Both passes read
        // the survivors back from the arena (coalesced, L2-resident: they were written just now).
        const float h_scale = (float)kHistBins / fmaxf(nc - bc, 1e-6f);
        auto bin_of = [&](float c) -> uint32_t {
          const float x = (c - bc) * h_scale;
          return x >= (float)(kHistBins - 1) ? (uint32_t)(kHistBins - 1) : (uint32_t)(int)fmaxf(x, 0.f);
        };
        // (the survivors' costs are read from the arena once and kept in registers for both passes
        // when there are at most eight per thread)
        constexpr int kKeep = 8;
        const bool keep = n_alive <= (uint32_t)(kKeep * NT);
        float cv[kKeep];
#pragma unroll
        for (int j = 0; j < kKeep; ++j) {
          const uint32_t i = tid + j * NT;
          cv[j] = __uint_as_float(__ldcg(&out_sc[i < n_alive ? i : 0u]).y);
        }
#pragma unroll
        for (int j = 0; j < kKeep; ++j)
          if (tid + j * NT < n_alive) atomicAdd(&s_hist[bin_of(cv[j])], 1u);
        for (uint32_t i = tid + kKeep * NT; i < n_alive; i += NT)
          atomicAdd(&s_hist[bin_of(__uint_as_float(__ldcg(&out_sc[i]).y))], 1u);
        __syncthreads();
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
#pragma unroll
          for (int j = 0; j < kKeep; ++j)
            if (tid + j * NT < n_alive && bin_of(cv[j]) == kbin) s_cand[atomicAdd(&hot->ncand, 1u)] = f2ord(cv[j]);
          for (uint32_t i = tid + kKeep * NT; i < n_alive; i += NT) {
            const float c = __uint_as_float(__ldcg(&out_sc[i]).y);
            if (bin_of(c) == kbin) s_cand[atomicAdd(&hot->ncand, 1u)] = f2ord(c);
          }
          __syncthreads();
          if (kcount <= 64u) {
            // a handful of candidates (the usual case): rank by counting, two warps
            if (tid < (int)kcount) {
              const uint32_t mine = s_cand[tid];
              uint32_t rank = 0;
              for (uint32_t j = 0; j < kcount; ++j) {
                const uint32_t o = s_cand[j];
                rank += (o < mine) || (o == mine && j < (uint32_t)tid);
              }
              if (rank == kk) ps.misc[3] = mine;
            }
            __syncthreads();
            n_cur_cut = ord2f(ps.misc[3]);
          } else {
            n_cur_cut = block_kth_smallest<NT>([&](uint32_t i) { return s_cand[i]; }, kcount, kk, best_ord, 0xFFFFFFFFu,
                                               nc_ord - best_ord, ps.hist, ps.misc);
          }
        } else {  // (a degenerate cost distribution: select over all survivors)
          n_cur_cut = block_kth_smallest<NT>(arena_ord, n_alive, k, best_ord, 0xFFFFFFFFu, nc_ord - best_ord, ps.hist,
                                             ps.misc);
        }
        (void)keep;
        n_abeam = n_cur_cut - bc + cfg.beam_delta;
        if (CLG && !(n_cur_cut < beam_cut)) {  // (a survivor AT next_cutoff == best + beam: inl.h:196 does not fire)
          n_cur_cut = beam_cut;
          n_abeam = cfg.beam;
        }
      } else {
        // ---- general case (the adaptive beam of this frame was wider than the beam, or the arena
        // is full): GetCutoff over the survivors in the arena
        get_cutoff<NT>(arena_ord, n_kept, n_kept, best_ord, cfg, ps.red32, ps.hist, ps.misc, n_cur_cut, n_abeam);
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
