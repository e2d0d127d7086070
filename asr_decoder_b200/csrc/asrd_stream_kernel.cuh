// k_stream: the on-chip frame loop of the plain decoders (one CTA per stream, the per-frame
// state->token map in shared memory).  Included by asrd_kernels.cuh after the HBM-map device
// functions it falls back to (expand_frame, post_epilogue, cutoff_prologue).
#pragma once

namespace asrd {

// ------------------------------------------------------------------ on-chip frame loop

// k_stream: ONE CTA per stream runs the whole frame loop of an AdvanceDecoding call with the
// per-frame state->token map in SHARED memory (the stream's private recombination state never
// leaves the SM): GetCutoff + pre-pass, emitting expansion, eps closure and the survivor
// write-out of every frame without going back to the launch queue.  HBM traffic per frame is
// what the search really needs — arc records, row offsets, the previous frame's tokens, the
// log-likelihood row — plus the append of the survivors to the token arena; the map's random
// 32-byte sectors (the DRAM row-activation bound of the k_expand/k_post path) are gone.
// A frame whose distinct destination states exceed the on-chip capacity is redone through the
// HBM map of the stream by the same CTA (expand_frame + post_epilogue), so results never depend
// on which path ran.  Plain (non-biglm) decoders only.
constexpr int kSmemLog2 = 14;
constexpr uint32_t kSmemSlots = 1u << kSmemLog2;
constexpr size_t kSmemMapBytes = (size_t)kSmemSlots * 12 + 2 * (kSmemSlots / 32) * 4;  // vals, keys, round bitmaps
constexpr uint32_t kSmemClaimLimit = kSmemSlots - kStreamThreads - 64;  // every warp may overshoot by 32 claims per map update

struct SmemMap {
  unsigned long long *val;  // (ordered cost << 32) | arc id, kInfVal when free
  uint32_t *key;            // state | kDestEpsBit, kEmptyKey when free
  uint32_t *qbits;          // [2][kSmemSlots / 32] slots to relax in the eps-closure round of that parity
  uint32_t *claims;
  uint32_t *overflow;
  uint32_t claim_limit;
};

// position of the r-th (0-based) set bit of mask (r < popc(mask))
__device__ __forceinline__ int select_nth(uint32_t mask, int r) {
  int pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const uint32_t low = mask & ((1u << w) - 1u);
    const int c = __popc(low);
    const bool up = r >= c;
    r -= up ? c : 0;
    mask = up ? (mask >> w) : low;
    pos += up ? w : 0;
  }
  return pos;
}

__device__ __forceinline__ uint4 lds_volatile_u4(const uint32_t *p) {
  uint4 r;
  asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return r;
}

// FindOrAddToken (inl.h:88-136) on the on-chip map, called by the FULL warp: lanes with `act`
// relax destination dstw (= state | kDestEpsBit) with value pk.
// The key array is probed in BUCKETS of four slots (one 16-byte shared load compares four keys;
// double hashing between buckets): heavy frames fill the map to 85 %, where slot-wise linear
// probing needs ~20 probes per insert and the slowest lane of the warp many more.  The probe
// loop is warp-uniform (a per-lane loop ran at 3 active threads per instruction); new claims are
// counted once per call.  A key lives in the first bucket of its probe sequence that had a free
// slot when it was inserted; slots are never freed within a frame, so a lookup may stop at the
// first bucket that still has one.  Lanes give up when the claim budget is exhausted (the frame
// is then redone through HBM).
__device__ __forceinline__ void smem_relax(const SmemMap &m, bool act, uint32_t dstw, unsigned long long pk,
                                           uint32_t next_round, uint32_t *s_any, int lane) {
  constexpr uint32_t kBuckets = kSmemSlots / 4;
  const uint32_t hsh = (dstw & kStateMask) * 0x9E3779B1u;
  uint32_t b = hsh >> (32 - kSmemLog2 + 2);
  const uint32_t step = (hsh >> 3) | 1u;  // odd: the sequence visits every bucket
  bool pend = act;
  uint32_t slot = 0xFFFFFFFFu, nclaim = 0;
  while (__any_sync(kFull, pend)) {
    if (pend) {
      const uint4 kk = lds_volatile_u4(&m.key[b * 4]);
      const int hit = kk.x == dstw ? 0 : kk.y == dstw ? 1 : kk.z == dstw ? 2 : kk.w == dstw ? 3 : -1;
      const int emp = kk.x == kEmptyKey ? 0 : kk.y == kEmptyKey ? 1 : kk.z == kEmptyKey ? 2 : kk.w == kEmptyKey ? 3 : -1;
      if (hit >= 0) {
        slot = b * 4 + hit;
        pend = false;
      } else if (emp >= 0) {
        if (*reinterpret_cast<volatile uint32_t *>(m.overflow)) {
          pend = false;
        } else {
          const uint32_t old = atomicCAS(&m.key[b * 4 + emp], kEmptyKey, dstw);
          if (old == kEmptyKey || old == dstw) {
            slot = b * 4 + emp;
            pend = false;
            nclaim += old == kEmptyKey;
          }  // else: somebody else's key took the slot — look at the bucket again
        }
      } else {
        b = (b + step) & (kBuckets - 1);
      }
    }
  }
  if (__any_sync(kFull, nclaim != 0)) {
    const uint32_t c = __reduce_add_sync(kFull, nclaim);
    // (the flag carries the closure round that raised it, see the round loop of k_stream)
    if (lane == 0 && atomicAdd(m.claims, c) + c > m.claim_limit) atomicCAS(m.overflow, 0u, next_round);
  }
  if (slot != 0xFFFFFFFFu) {
    // the value only ever decreases: an arc that cannot win needs no atomic
    const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&m.val[slot]);
    if (cur > pk) {
      const unsigned long long old = atomicMin(&m.val[slot], pk);
      if ((dstw & kDestEpsBit) && (uint32_t)(pk >> 32) < (uint32_t)(old >> 32)) {  // cost changed: (re)queue, inl.h:115-127,425
        atomicOr(&m.qbits[(next_round & 1u) * (kSmemSlots / 32) + (slot >> 5)], 1u << (slot & 31u));
        if (s_any) *s_any = 1u;
      }
    }
  }
  __syncwarp();
}

template <int U, bool SMEM_LL>
__global__ void __launch_bounds__(kStreamThreads, 1)
k_stream(StreamState *const *streams, const AdvanceParams *params, GraphView g, DecoderConfigDev cfg,
         int num_indices) {
  constexpr int NT = kStreamThreads;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ PostSmem ps;
  __shared__ FrameDesc s_d;
  __shared__ uint32_t s_claims, s_overflow, s_any[3];
  __shared__ struct {  // GetCutoff result of the next frame, computed on chip
    unsigned long long best;
    float cur, abeam;
    uint32_t n, off;
  } s_h;
  // per-warp scratch: staging of admitted arcs during the expansion (kStage x {u64 value, u32
  // destination}); the eps closure reuses it as its compaction buffer of stamped slots (64 x u16)
  constexpr int kStage = 44;
  __shared__ __align__(8) unsigned char s_warp_scratch[kStreamThreads / 32][kStage * 12];
  SmemMap m;
  m.val = reinterpret_cast<unsigned long long *>(s_dyn);
  m.key = reinterpret_cast<uint32_t *>(s_dyn + (size_t)kSmemSlots * 8);
  m.qbits = reinterpret_cast<uint32_t *>(s_dyn + (size_t)kSmemSlots * 12);
  m.claims = &s_claims;
  m.overflow = &s_overflow;
  m.claim_limit = (cfg.debug_flags >> 8) ? min((uint32_t)(cfg.debug_flags >> 8), kSmemClaimLimit) : kSmemClaimLimit;  // (test hook: smaller on-chip budget)
  float *s_ll = reinterpret_cast<float *>(s_dyn + kSmemMapBytes);
  StreamState *st = streams[blockIdx.x];
  FrameDesc *d = &s_d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const LmPair lms = {};
  // this launch decodes the rows of ONE staged chunk: later chunks may already be raising
  // target_frame while their rows are still being copied
  const int limit = min(params[blockIdx.x].frame0 + params[blockIdx.x].n_frames, st->max_frames);

  for (uint32_t i = tid; i < kSmemSlots; i += NT) {
    m.val[i] = kInfVal;
    m.key[i] = kEmptyKey;
  }
  for (uint32_t i = tid; i < 2 * (kSmemSlots / 32); i += NT) m.qbits[i] = 0;
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
  // the map while it is still in shared memory): uniform registers, no HBM round trip per frame.
  int t = st->frame;
  bool have_cut = false;
  const float *const ll_hist = st->ll_hist;
  const int ll_stride = st->ll_stride;
  for (;;) {
    if (t >= limit) break;
    if (tid == 0) {
      s_claims = 0;
      s_overflow = (cfg.debug_flags & 8) ? 1u : 0u;  // test hook: every frame through the HBM map
      s_any[0] = s_any[1] = s_any[2] = 0;
    }
    if (!have_cut) {
      // ---- first frame of the launch / after an HBM-map frame: GetCutoff + best-token pre-pass
      // over the arena tokens; descriptor of the step into shared memory
      cutoff_prologue<NT, false>(st, d, g, cfg, lms, ps.red64, ps.red32, ps.hist, ps.misc);
      __syncthreads();
      if (!s_d.stepping) break;  // uniform: frame == target_frame
      phase(0);
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
      phase(0);
    }
    const float *__restrict__ ll = s_d.ll;
    phase(1);
    const uint32_t n_cur = s_d.n_cur;
    const uint32_t n_groups = (n_cur + 31) >> 5;
    const uint2 *__restrict__ toks = s_d.toks;
    const float cur_cut = s_d.cur_cut, abeam = s_d.abeam;
    uint32_t *next_cut = &s_d.next_cut_bits;

    // ---- emitting expansion (ProcessEmitting, inl.h:311-347) into the on-chip map
    // Software pipeline per warp: the arc records of step i+1 (the next U x 32 flattened arcs,
    // possibly of the next token group) are requested before step i is scored and merged into
    // the map, so the HBM/L2 latency of the arc fetch overlaps the shared-memory work.
    {
      uint32_t expanded = 0, admitted = 0;
      // Only about a third of the arcs are admitted: they are staged in a warp-private buffer and
      // the map is updated 32 arcs at a time with every lane busy.
      unsigned long long *st_pk = reinterpret_cast<unsigned long long *>(s_warp_scratch[warp]);
      uint32_t *st_w = reinterpret_cast<uint32_t *>(s_warp_scratch[warp] + kStage * 8);
      uint32_t n_staged = 0;
      const uint32_t lt_mask = (1u << lane) - 1u;
      auto flush = [&](uint32_t k) {  // the first k (<= 32) staged arcs go to the map, the rest moves up
        const bool act = (uint32_t)lane < k;
        const uint32_t w = act ? st_w[lane] : 0u;
        const unsigned long long pk = act ? st_pk[lane] : 0ull;
        const uint32_t rem = n_staged - k;
        __syncwarp();
        if (rem) {
          uint32_t tw = 0;
          unsigned long long tpk = 0;
          if ((uint32_t)lane < rem) {
            tw = st_w[k + lane];
            tpk = st_pk[k + lane];
          }
          __syncwarp();
          if ((uint32_t)lane < rem) {
            st_w[lane] = tw;
            st_pk[lane] = tpk;
          }
          __syncwarp();
        }
        n_staged = rem;
        smem_relax(m, act, w, pk, 1u, nullptr, lane);
      };
      float nc = ord2f(*reinterpret_cast<volatile uint32_t *>(next_cut));
      // fetch cursor: the token group whose arcs are being requested.  Two more groups are in
      // flight behind it: the tokens of group +2 and the emitting-arc spans of group +1 (the
      // span load needs the token's state), so a new group starts without waiting on HBM.
      uint32_t f_grp = warp, f_off = 0, f_base = 0, f_cost = 0, f_total = 0, f_jb = 0;
      bool f_open = false;
      uint32_t t1_cost = 0, t1_base = 0, t1_deg = 0;  // group f_grp + 32: cost, span
      uint2 t2 = make_uint2(0, 0);                    // group f_grp + 64: {state, cost}
      bool t2_ok = false;
      auto load_tokens = [&](uint32_t grp) {  // stage A
        const uint32_t i = grp * 32 + lane;
        t2_ok = grp < n_groups && i < n_cur;
        if (t2_ok) t2 = __ldcg(&toks[i]);  // written by this kernel one frame ago: no ld.global.nc
      };
      auto load_spans = [&]() {  // stage B: consumes stage A
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
      load_tokens(warp + NT / 32);
      // the step in flight
      bool n_have = false, n_in[U];
      uint32_t n_a[U];
      float n_tc[U];
      int4 n_arc[U];
      auto issue = [&]() {
        n_have = false;
        while (!f_open || f_jb >= f_total) {
          if (f_open) f_grp += NT / 32;
          f_open = false;
          if (f_grp >= n_groups) return;
          // warp-uniform decision (the lanes may not have reconverged after the map updates)
          if (__any_sync(kFull, *reinterpret_cast<volatile uint32_t *>(&s_overflow) != 0u)) {
            f_grp = n_groups;
            return;
          }
          // running cutoff (inl.h:330): other warps' tightenings arrive once per group
          nc = fminf(nc, ord2f(*reinterpret_cast<volatile uint32_t *>(next_cut)));
          f_cost = t1_cost;
          f_base = t1_base;
          const uint32_t deg = t1_deg;
          load_spans();
          load_tokens(f_grp + 2 * (NT / 32));
          const uint32_t incl = warp_incl_scan(deg, lane);
          f_off = incl - deg;
          f_total = __shfl_sync(kFull, incl, 31);
          f_jb = 0;
          f_open = true;
          expanded += f_total;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t j = f_jb + u * 32 + lane;
          n_in[u] = j < f_total;
          const int l = warp_owner(f_off, j);
          const uint32_t off_l = __shfl_sync(kFull, f_off, l);
          const uint32_t base_l = __shfl_sync(kFull, f_base, l);
          n_tc[u] = __uint_as_float(__shfl_sync(kFull, f_cost, l));
          n_a[u] = base_l + (j - off_l);
          if (n_in[u]) n_arc[u] = __ldg(&g.arcs[n_a[u]]);
        }
        f_jb += 32 * U;
        n_have = true;
      };
      issue();
      while (n_have) {
        bool in[U];
        uint32_t a[U];
        float tcost[U];
        int4 arc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          in[u] = n_in[u];
          a[u] = n_a[u];
          tcost[u] = n_tc[u];
          arc[u] = n_arc[u];
        }
        issue();
        float tot[U];
        bool adm[U];
        uint32_t cand_bits = 0xFFFFFFFFu;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          adm[u] = false;
          tot[u] = 0.f;
          if (in[u]) {
            const float ac = -(SMEM_LL ? s_ll[arc[u].x - 1] : __ldg(&ll[arc[u].x - 1]));
            tot[u] = (tcost[u] + ac) + __int_as_float(arc[u].z);  // inl.h:326-329
            adm[u] = tot[u] < nc;
            if (adm[u]) {
              const float cand = tot[u] + abeam;  // inl.h:332-333
              if (cand < nc) cand_bits = min(cand_bits, f2ord(cand));
            }
          }
        }
        if (__any_sync(kFull, cand_bits != 0xFFFFFFFFu)) {
          const uint32_t wmin = __reduce_min_sync(kFull, cand_bits);
          if (lane == 0) atomicMin(next_cut, wmin);
          nc = fminf(nc, ord2f(wmin));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned nmask = __ballot_sync(kFull, adm[u]);
          if (nmask == 0) continue;
          admitted += adm[u];
          const uint32_t nnew = (uint32_t)__popc(nmask);
          if (n_staged + nnew > (uint32_t)kStage) flush(n_staged);  // (n_staged < 32 here)
          if (adm[u]) {
            const uint32_t r = n_staged + (uint32_t)__popc(nmask & lt_mask);
            st_w[r] = (uint32_t)arc[u].w;
            st_pk[r] = pack_val(tot[u], a[u]);
          }
          n_staged += nnew;
          __syncwarp();
          if (n_staged >= 32u) flush(32u);
        }
      }
      flush(n_staged);  // the last, partial set
      admitted = __reduce_add_sync(kFull, admitted);
      if (lane == 0 && expanded) {
        atomicAdd(&s_d.arcs_expanded, expanded);
        atomicAdd(&s_d.arcs_admitted, admitted);
      }
    }
    __syncthreads();
    phase(2);
    const float nc = ord2f(s_d.next_cut_bits);  // the FINAL next_cutoff of this frame

    // ---- eps closure (ProcessNonemitting, inl.h:353-431): round r relaxes the slots stamped r
    if (!s_overflow) {
      uint16_t *wq = reinterpret_cast<uint16_t *>(s_warp_scratch[warp]);  // this warp's compaction buffer of stamped slots
      for (uint32_t round = 1;; ++round) {
        const uint32_t nr = round + 1;
        uint32_t nq = 0;
        // the warp owns 512 consecutive slots = 16 words of this round's bitmap; stamped slots are
        // compacted into wq and relaxed 32 at a time, so the row loads of a batch are issued
        // together and the eps arcs of the batch are flattened over the lanes like the emitting
        // arcs above
        constexpr int kWordsPerWarp = (int)(kSmemSlots / 32 / (kStreamThreads / 32));
        uint32_t *qw = &m.qbits[(round & 1u) * (kSmemSlots / 32) + (uint32_t)warp * kWordsPerWarp];
        uint32_t myword = 0;
        if (lane < kWordsPerWarp) {
          myword = qw[lane];
          if (myword) qw[lane] = 0;
        }
        const unsigned nzw = __ballot_sync(kFull, myword != 0);
        for (int k = 0; k < kWordsPerWarp; ++k) {
          const bool last = k == kWordsPerWarp - 1;
          if (!((nzw >> k) & 1u) && !(last && nq > 0u)) continue;
          const uint32_t sm = __shfl_sync(kFull, myword, k);
          const uint32_t slot = ((uint32_t)warp * kWordsPerWarp + (uint32_t)k) * 32u + lane;
          const bool stamped = (sm >> lane) & 1u;
          if (sm) {
            if (stamped) wq[nq + __popc(sm & ((1u << lane) - 1u))] = (uint16_t)slot;
            nq += __popc(sm);
            __syncwarp();
          }
          while (nq >= 32u || (last && nq > 0u)) {
            const uint32_t cnt = nq < 32u ? nq : 32u;
            uint32_t deg = 0, base = 0, cost_bits = 0;
            if ((uint32_t)lane < cnt) {
              const uint32_t sl = wq[lane];
              const uint32_t state = m.key[sl] & kStateMask;
              const uint32_t co = (uint32_t)(*reinterpret_cast<volatile unsigned long long *>(&m.val[sl]) >> 32);
              cost_bits = __float_as_uint(ord2f(co));
              if (ord2f(co) < nc) {  // inl.h:391
                const uint2 r = __ldg(&g.rows[state]);
                base = r.x;
                deg = r.y - r.x;
              }
            }
            __syncwarp();
            if (nq > 32u) {  // keep the remainder for the next batch
              const uint16_t keep = (uint32_t)lane + 32u < nq ? wq[lane + 32] : (uint16_t)0;
              __syncwarp();
              wq[lane] = keep;
              __syncwarp();
            }
            nq -= cnt;
            const uint32_t incl = warp_incl_scan(deg, lane);
            const uint32_t off = incl - deg;
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            for (uint32_t jb = 0; jb < total; jb += 32) {
              const uint32_t j = jb + lane;
              const bool in = j < total;
              const int l = warp_owner(off, j);
              const uint32_t off_l = __shfl_sync(kFull, off, l);
              const uint32_t base_l = __shfl_sync(kFull, base, l);
              const float cost = __uint_as_float(__shfl_sync(kFull, cost_bits, l));
              const uint32_t a = base_l + (j - off_l);
              int4 arc = make_int4(0, 0, 0, 0);
              if (in) arc = __ldg(&g.arcs[a]);
              const float tot = cost + __int_as_float(arc.z);  // inl.h:413-414
              smem_relax(m, in && tot < nc, (uint32_t)arc.w, pack_val(tot, a), nr, &s_any[nr % 3u], lane);  // inl.h:415
            }
          }
        }
        // one barrier per round: round r raises s_any[(r + 1) % 3] and reads it after the barrier;
        // the flag the NEXT round raises is lowered here — its last readers passed the previous
        // barrier, its next writers wait behind this one
        if (tid == 0) s_any[(round + 2u) % 3u] = 0;
        __syncthreads();
        // an overflow raised by a warp that is already in round r + 1 carries the tag r + 2 and must
        // not stop the slower warps one round early (the decision has to be uniform)
        const uint32_t ovf = *reinterpret_cast<volatile uint32_t *>(&s_overflow);
        if (s_any[nr % 3u] == 0 || (ovf != 0 && ovf <= nr)) break;
      }
    }

    phase(3);
    if (s_overflow) {
      // ---- too many distinct destinations for the on-chip map: wipe it and redo the frame
      // through the stream's HBM map (identical results; the running cutoff stays valid)
      for (uint32_t i = tid; i < kSmemSlots; i += NT) {
        m.val[i] = kInfVal;
        m.key[i] = kEmptyKey;
      }
      for (uint32_t i = tid; i < 2 * (kSmemSlots / 32); i += NT) m.qbits[i] = 0;
      if (tid == 0) {
        s_d.arcs_expanded = 0;
        s_d.arcs_admitted = 0;
        st->tot_fallback_frames += 1;
      }
      __syncthreads();
      // (two arc steps in flight per lane: a single CTA is latency-bound on the HBM map)
      expand_frame<2, SMEM_LL, false>(d, g, s_ll, (uint32_t)warp, NT / 32, 0, lms);
      __syncthreads();
      post_epilogue<false>(st, d, g, cfg, lms, ps);
      phase(5);
      ++t;
      have_cut = false;
      continue;
    }

    // ---- the frame's tokens are final.  While the map is still on chip: count the survivors,
    // find the best token (lowest cost, ties -> lowest state id, inl.h:169-179) and run GetCutoff
    // for the next frame over the map slots; then append the survivors to the token arena in slot
    // order (deterministic) and recycle the slots.
    {
      auto ord_at = [&](uint32_t i) -> uint32_t {
        if (m.key[i] == kEmptyKey) return 0xFFFFFFFFu;
        const uint32_t o = (uint32_t)(m.val[i] >> 32);
        return ord2f(o) < nc ? o : 0xFFFFFFFFu;
      };
      constexpr uint32_t kPerWarp = kSmemSlots / (NT / 32);  // each warp owns a contiguous run of slots
      const uint32_t slot0 = (uint32_t)warp * kPerWarp;
      unsigned long long best64 = kInfVal;
      uint32_t cnt = 0;
      for (uint32_t k = lane; k < kPerWarp; k += 32) {
        const uint32_t slot = slot0 + k;
        const uint32_t o = ord_at(slot);
        if (o != 0xFFFFFFFFu) {
          ++cnt;
          const unsigned long long b64 = ((unsigned long long)o << 32) | (m.key[slot] & kStateMask);
          best64 = b64 < best64 ? b64 : best64;
        }
      }
      cnt = __reduce_add_sync(kFull, cnt);
      if (lane == 0) ps.red32[warp] = cnt;
      best64 = block_min_u64<NT>(best64, ps.red64);  // (barriers inside: ps.red32 is complete)
      const uint32_t wc = ps.red32[lane];            // NT / 32 == 32 warps
      const uint32_t n_alive = __reduce_add_sync(kFull, wc);
      uint32_t pos = __reduce_add_sync(kFull, lane < warp ? wc : 0u);  // arena offset of this warp's run
      __syncthreads();
      float n_cur_cut, n_abeam;
      get_cutoff<NT>(ord_at, kSmemSlots, n_alive, (uint32_t)(best64 >> 32), cfg, ps.red32, ps.hist, ps.misc,
                     n_cur_cut, n_abeam, nc);
      const uint32_t cap = s_d.out_cap;
      uint2 *out_sc = s_d.out_sc;
      uint32_t *out_arc = s_d.out_arc;
      for (uint32_t k = lane; k < kPerWarp; k += 32) {
        const uint32_t slot = slot0 + k;
        const uint32_t kw = m.key[slot];
        bool alive = false;
        unsigned long long v = kInfVal;
        if (kw != kEmptyKey) {
          v = m.val[slot];
          alive = ord2f((uint32_t)(v >> 32)) < nc;
          m.key[slot] = kEmptyKey;
          m.val[slot] = kInfVal;
        }
        const unsigned am = __ballot_sync(kFull, alive);
        if (alive) {
          const uint32_t idx = pos + __popc(am & ((1u << lane) - 1u));
          if (idx < cap) {
            out_sc[idx] = make_uint2(kw & kStateMask, __float_as_uint(ord2f((uint32_t)(v >> 32))));
            out_arc[idx] = (uint32_t)v;
          }
        }
        pos += __popc(am);
      }
      if (tid == 0) {
        s_h.off = st->frame_off[t + 1];
        s_h.n = n_alive < cap ? n_alive : cap;
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
