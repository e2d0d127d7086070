// Host side of the C ABI (include/asrd.h).  Thin: validates arguments, owns device
// memory, uploads the graph, sequences the per-frame kernels.  No CPU fallback: every
// search step runs in the kernels of asrd_kernels.cuh.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "asrd_kernels.cuh"

using namespace asrd;

namespace {

std::atomic<long long> g_launches{0};
std::mutex g_ref_mu;  // reference counts of graphs and LMs (shared by decoders, src/my-decoder/online-decoder-base-inl.h:24)
thread_local std::string g_last_error;

#define CU_CHECK(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      g_last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                \
      fprintf(stderr, "[asrd] CUDA error %s at %s:%d\n", g_last_error.c_str(), __FILE__, \
              __LINE__);                                                                 \
      return ASRD_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)


constexpr int kHostChunkFrames = 16;

// per-device side stream + events for overlapping host->device staging with the search
constexpr int kMaxWorkers = 32;
struct DeviceCtx {
  cudaStream_t copy_stream = nullptr;
  cudaStream_t sync = nullptr;  // event-only stream: joins the workers' per-chunk completion events
  cudaStream_t prep = nullptr;  // high priority: row scatter (k_begin_advance) must not queue behind the frame loops
  cudaStream_t worker[kMaxWorkers] = {nullptr};
  cudaEvent_t ev_ready = nullptr, ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxWorkers] = {nullptr};
  // Host rows are staged through two device buffers OWNED BY THE CONTEXT (not stream-ordered scratch
  // of the call): the copy stream may then run ahead of the caller's stream, so the host->device
  // copies of the NEXT AdvanceDecoding call overlap the search of the current one (a service steps
  // thousands of streams chunk by chunk: with per-call scratch every call exposed its first copy).
  float *stage[2] = {nullptr, nullptr};
  size_t stage_floats = 0;
  bool stage_busy[2] = {false, false};  // ev_done[b] guards the last scatter that read stage[b]
  std::mutex issue_mu;  // one AdvanceDecoding call at a time ISSUES work on a device (the events above are shared)
};
std::mutex g_ctx_mu;
DeviceCtx g_ctx[16];

int GetCtx(int device, DeviceCtx **out) {
  if (device < 0 || device >= 16) return ASRD_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DeviceCtx &c = g_ctx[device];
  if (!c.copy_stream) {
    CU_CHECK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    int prio_least = 0, prio_greatest = 0;
    CU_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    CU_CHECK(cudaStreamCreateWithPriority(&c.prep, cudaStreamNonBlocking, prio_greatest));
    CU_CHECK(cudaStreamCreateWithFlags(&c.sync, cudaStreamNonBlocking));
    CU_CHECK(cudaEventCreateWithFlags(&c.ev_ready, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      CU_CHECK(cudaEventCreateWithFlags(&c.ev_copied[i], cudaEventDisableTiming));
      CU_CHECK(cudaEventCreateWithFlags(&c.ev_done[i], cudaEventDisableTiming));
    }
    CU_CHECK(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
    for (int i = 0; i < kMaxWorkers; ++i) {
      CU_CHECK(cudaStreamCreateWithPriority(&c.worker[i], cudaStreamNonBlocking, prio_least));
      CU_CHECK(cudaEventCreateWithFlags(&c.ev_join[i], cudaEventDisableTiming));
    }
  }
  *out = &c;
  return ASRD_OK;
}

// Optional per-kernel timing with CUDA events on the launching stream (asrd_profile_*):
// class 0 = k_expand, 1 = k_post, 2 = k_stream, 3 = k_lattice in arena-prune mode.  Off by default: events
// between launches add gaps.
std::atomic<int> g_profile{0};
std::mutex g_prof_mu;
double g_prof_ms[4] = {0, 0, 0, 0};
long long g_prof_n[4] = {0, 0, 0, 0};

struct Profiler {
  bool on;
  std::vector<cudaEvent_t> ev;  // begin/end pairs
  std::vector<int> cls;
  explicit Profiler(int enabled) : on(enabled != 0) {}
  ~Profiler() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
  void Begin(int c, cudaStream_t s) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    ev.push_back(e);
    cls.push_back(c);
  }
  void End(cudaStream_t s) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    ev.push_back(e);
  }
  int Finish(cudaStream_t s) {
    if (!on) return ASRD_OK;
    CU_CHECK(cudaStreamSynchronize(s));
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (size_t i = 0; i < cls.size(); ++i) {
      float ms = 0.f;
      CU_CHECK(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
      g_prof_ms[cls[i]] += ms;
      g_prof_n[cls[i]] += 1;
    }
    return ASRD_OK;
  }
};

// The frame loops of the sub-batches, the host->device staging copies and the row scatter run on
// a dozen CUDA streams.  With the driver's default of 8 hardware work queues several of them share
// a queue, and a sub-batch whose stream lands on the queue of the copy stream runs behind every
// staged copy (measured: +12 ms per 256-stream step, one sub-batch finishing 15 ms after the
// others).  CUDA_DEVICE_MAX_CONNECTIONS is read when the driver creates the context, so a host
// that wants the extra queues calls asrd_configure_process() (or sets the variable itself) before
// its first CUDA call; the library no longer touches the environment when it is loaded.

std::atomic<int64_t> g_last_fallback_frames{0};
std::atomic<int64_t> g_last_pruned_tokens{0}, g_last_peak_tokens{0};
std::atomic<int64_t> g_last_prune_cycles[8];
std::atomic<int64_t> g_last_phase_cycles[6];

int EnvInt(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

int g_num_sms = 0;

// The library's stream-ordered scratch (staging buffers, descriptors, result windows) comes from a
// PRIVATE memory pool per device, kept cached across the synchronising calls; the device's default
// pool — which other code in the host process may use — is left alone.
cudaMemPool_t g_pool[64] = {nullptr};
std::mutex g_pool_mu;

int EnsureDevice(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_last_error = "no CUDA device";
    return ASRD_ERR_CUDA;
  }
  if (device < 0 || device >= n || device >= 64) return ASRD_ERR_BAD_ARG;
  CU_CHECK(cudaSetDevice(device));
  if (!g_num_sms) {
    cudaDeviceProp p;
    CU_CHECK(cudaGetDeviceProperties(&p, device));
    g_num_sms = p.multiProcessorCount;
  }
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pool[device]) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CU_CHECK(cudaMemPoolCreate(&g_pool[device], &props));
    unsigned long long keep = ~0ull;
    CU_CHECK(cudaMemPoolSetAttribute(g_pool[device], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  return ASRD_OK;
}

cudaMemPool_t CurrentPool() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64) ? g_pool[dev] : nullptr;
}

uint32_t NextPow2(uint64_t v) {
  uint32_t p = 1;
  while (p < v && p < (1u << 30)) p <<= 1;
  return p;
}

struct Scratch {  // stream-ordered device scratch released on scope exit
  cudaStream_t s;
  std::vector<void *> ptrs;
  explicit Scratch(cudaStream_t st) : s(st) {}
  ~Scratch() {
    for (void *p : ptrs) cudaFreeAsync(p, s);
  }
  template <typename T>
  cudaError_t Alloc(T **out, size_t count) {
    void *p = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&p, std::max<size_t>(count * sizeof(T), 16), CurrentPool(), s);
    if (e == cudaSuccess) ptrs.push_back(p);
    *out = (T *)p;
    return e;
  }
};

int UploadStreams(asrd_decoder *const *decs, int n, cudaStream_t s, Scratch &sc, StreamState ***out) {
  std::vector<StreamState *> h(n);
  for (int i = 0; i < n; ++i) h[i] = decs[i]->d_state;
  StreamState **d = nullptr;
  CU_CHECK(sc.Alloc(&d, (size_t)n));
  CU_CHECK(cudaMemcpyAsync(d, h.data(), sizeof(StreamState *) * n, cudaMemcpyHostToDevice, s));
  *out = d;
  return ASRD_OK;
}

DecoderConfigDev DevCfg(const asrd_decoder *d) {
  DecoderConfigDev c;
  c.beam = d->cfg.beam;
  c.max_active = d->cfg.max_active;
  c.min_active = d->cfg.min_active;
  c.lattice_beam = d->cfg.lattice_beam;
  c.beam_delta = d->cfg.beam_delta;
  c.collect_stats = d->opts.collect_stats;
  c.debug_flags = EnvInt("ASRD_DEBUG_FLAGS", 0);
  return c;
}

LmPair Lms(const asrd_decoder *d) {
  LmPair p;
  memset(&p, 0, sizeof(p));
  if (d->lm1) {
    p.lm1 = d->lm1->view;
    p.lm2 = d->lm2->view;
  }
  return p;
}

int CheckBatch(asrd_decoder *const *decs, int n) {
  if (!decs || n <= 0 || n > kMaxBatch) return ASRD_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i) {
    if (!decs[i]) return ASRD_ERR_BAD_ARG;
    if (decs[i]->graph != decs[0]->graph) return ASRD_ERR_BAD_ARG;  // one graph per launch
    if (memcmp(&decs[i]->cfg, &decs[0]->cfg, sizeof(asrd_config)) != 0) return ASRD_ERR_BAD_ARG;
    if (decs[i]->opts.hash_capacity != decs[0]->opts.hash_capacity) return ASRD_ERR_BAD_ARG;
    if (decs[i]->opts.collect_stats != decs[0]->opts.collect_stats) return ASRD_ERR_BAD_ARG;
    if (decs[i]->opts.prune_tokens != decs[0]->opts.prune_tokens) return ASRD_ERR_BAD_ARG;
    if (decs[i]->lm1 != decs[0]->lm1 || decs[i]->lm2 != decs[0]->lm2) return ASRD_ERR_BAD_ARG;
  }
  return ASRD_OK;
}

// ---- kernel launch plumbing -------------------------------------------------------------

typedef void (*ExpandFn)(FrameDesc *, GraphView, int, int, LmPair);

struct ExpandPlan {
  ExpandFn fn;
  dim3 grid;
  size_t dyn;
  int flags;
};

// Picks the k_expand instantiation (arcs in flight per lane; log-likelihood row staged in
// shared memory when it fits) and sizes the grid to about one resident wave: blockIdx.y is
// the stream, gridDim.x CTAs share one stream's token groups.
int PlanExpand(int n_streams, int num_indices, bool biglm, ExpandPlan *plan) {
  const int u = EnvInt("ASRD_EXPAND_U", 1);
  const bool smem_ll = EnvInt("ASRD_SMEM_LL", 1) != 0 && (size_t)num_indices * 4 <= 96 * 1024;
  ExpandFn fn;
  if (biglm) fn = smem_ll ? k_expand<1, true, true> : k_expand<1, false, true>;
  else if (smem_ll) fn = u >= 4 ? k_expand<4, true, false> : (u == 2 ? k_expand<2, true, false> : k_expand<1, true, false>);
  else fn = u >= 4 ? k_expand<4, false, false> : (u == 2 ? k_expand<2, false, false> : k_expand<1, false, false>);
  const size_t dyn = smem_ll ? (size_t)num_indices * 4 : 0;
  if (dyn > 48 * 1024)
    CU_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int per_sm = 0;
  CU_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kExpandThreads, dyn));
  const int resident = g_num_sms * std::max(per_sm, 1);
  int gx = EnvInt("ASRD_EXPAND_G", 0);
  if (gx <= 0) gx = (4 * resident + n_streams - 1) / n_streams;  // ~4 waves of small CTAs balance best
  plan->fn = fn;
  plan->grid = dim3((unsigned)std::max(gx, 1), (unsigned)n_streams, 1);
  plan->dyn = dyn;
  plan->flags = EnvInt("ASRD_EXPAND_FLAGS", 0);
  return ASRD_OK;
}


typedef void (*StreamFn)(StreamState *const *, const AdvanceParams *, GraphView, DecoderConfigDev, int, uint32_t);

struct StreamPlan {
  StreamFn fn = nullptr;  // null: use the k_expand / k_post pair
  size_t dyn = 0;
  uint32_t n_buckets = 0;
};

// The on-chip frame loop (k_stream) serves plain decoders; everything else runs the HBM-map
// kernels.  The map takes whatever shared memory the log-likelihood row, the warp scratch and the
// closure queues leave (8 bytes per slot, buckets of four): ~20 k slots at 3000 pdfs.  Rows too
// wide to leave a useful map are read from global memory instead.
int PlanStream(const asrd_graph *graph, int num_indices, bool biglm, StreamPlan *plan) {
  plan->fn = nullptr;
  if (biglm || !EnvInt("ASRD_STREAM_KERNEL", 1)) return ASRD_OK;
  if (graph->total_arcs >= (int64_t)kMaxStreamArcs) return ASRD_OK;  // work items pack (arc index << 2 | count)
  cudaFuncAttributes fa;
  CU_CHECK(cudaFuncGetAttributes(&fa, k_stream<true, false>));
  int dev = 0, max_optin = 0;
  CU_CHECK(cudaGetDevice(&dev));
  CU_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t room = (size_t)max_optin > fa.sharedSizeBytes + 64 ? (size_t)max_optin - fa.sharedSizeBytes - 64 : 0;
  constexpr size_t kMinMapBytes = (size_t)8192 * 8;
  bool smem_ll = stream_fixed_dyn_bytes(num_indices) + 2 * kMinMapBytes <= room;
  const size_t fixed = stream_fixed_dyn_bytes(smem_ll ? num_indices : 0);
  if (fixed + kMinMapBytes > room) return ASRD_OK;
  // n_buckets must be a multiple of 32 (the write-out walks 32 buckets per warp step) and fit a u16 slot id
  uint32_t n_buckets = (uint32_t)std::min<size_t>((room - fixed) / 32, 16352) & ~31u;
  const int force = EnvInt("ASRD_STREAM_BUCKETS", 0);  // (measurement aid)
  if (force >= 64 && (uint32_t)force < n_buckets) n_buckets = (uint32_t)force & ~31u;
  const bool clg = graph->view.clg != 0;
  plan->fn = smem_ll ? (clg ? k_stream<true, true> : k_stream<true, false>) : (clg ? k_stream<false, true> : k_stream<false, false>);
  plan->n_buckets = n_buckets;
  plan->dyn = fixed + (size_t)n_buckets * 32;
  CU_CHECK(cudaFuncSetAttribute(plan->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->dyn));
  return ASRD_OK;
}

// The arena prune (device option prune_tokens): k_prune keeps its lookup map (6 bytes per slot),
// the extra costs of the frame being swept (4 bytes per token) and two flag bitmaps in shared
// memory: frames of up to ~20 k tokens.  Streams with a larger frame are taken by the HBM-map
// sweep (k_lattice in PRUNE mode), launched right behind it (a no-op for the streams k_prune
// served); ASRD_PRUNE_KERNEL=0 sends every stream there.  ASRD_PRUNE_DEPTH: frames below the
// previous frontier a sweep goes back (default: prune_interval; -1: to frame 0 every time).
typedef void (*PruneFn)(StreamState *const *, GraphView, DecoderConfigDev, int, int, uint32_t, uint32_t, LatticeOut *, int);

struct PrunePlan {
  PruneFn fn = nullptr;  // null: k_lattice<false, true> only
  PruneFn emit_fn = nullptr;  // the same sweep as raw-lattice extraction (k_prune<true>)
  size_t dyn = 0;
  uint32_t n_buckets = 0, ex_cap = 0;
};

int PlanPrune(PrunePlan *plan) {
  plan->fn = nullptr;
  if (!EnvInt("ASRD_PRUNE_KERNEL", 1)) return ASRD_OK;
  cudaFuncAttributes fa;
  CU_CHECK(cudaFuncGetAttributes(&fa, k_prune<true>));  // (the larger static footprint of the two)
  int dev = 0, max_optin = 0;
  CU_CHECK(cudaGetDevice(&dev));
  CU_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t room = (size_t)max_optin > fa.sharedSizeBytes + 64 ? (size_t)max_optin - fa.sharedSizeBytes - 64 : 0;
  // per token: 4 + 1/4 bytes of extra cost and flags, 6 bytes per map slot at a load of at most 7/8
  uint32_t ex_cap = (uint32_t)std::min<size_t>(20480, room / 12) & ~255u;
  const int force = EnvInt("ASRD_PRUNE_CAP", 0);  // (test hook: smaller frames already go to the HBM-map sweep)
  if (force >= 256 && (uint32_t)force < ex_cap) ex_cap = (uint32_t)force & ~255u;
  if (ex_cap < 4096 && !force) return ASRD_OK;
  const uint32_t n_buckets = (uint32_t)std::min<size_t>((room - prune_token_dyn_bytes(ex_cap)) / 24, 16384) & ~1u;  // 4 slots x (4 + 2) bytes
  plan->fn = k_prune<false>;
  plan->emit_fn = k_prune<true>;
  plan->n_buckets = n_buckets;
  plan->ex_cap = ex_cap;
  plan->dyn = (size_t)n_buckets * 24 + prune_token_dyn_bytes(ex_cap);
  CU_CHECK(cudaFuncSetAttribute(plan->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->dyn));
  CU_CHECK(cudaFuncSetAttribute(plan->emit_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan->dyn));
  return ASRD_OK;
}

}  // namespace

extern "C" {

const char *asrd_strerror(int status) {
  switch (status) {
    case ASRD_OK: return "ok";
    case ASRD_ERR_BAD_ARG: return "bad argument";
    case ASRD_ERR_CUDA: return g_last_error.empty() ? "CUDA error" : g_last_error.c_str();
    case ASRD_ERR_NOMEM: return "out of memory";
    case ASRD_ERR_HASH_OVERFLOW: return "state->token map overflow (raise hash_capacity)";
    case ASRD_ERR_LM_PAIRS_OVERFLOW: return "biglm LM state-pair table full (raise lm_pair_capacity)";
    case ASRD_ERR_ARENA_OVERFLOW: return "token arena overflow (raise token_capacity)";
    case ASRD_ERR_FRAMES_OVERFLOW: return "more frames than max_frames";
    case ASRD_ERR_NO_TOKENS: return "no surviving tokens / nothing decoded";
    case ASRD_ERR_PATH_OVERFLOW: return "best path longer than the output buffer";
    case ASRD_ERR_STATE: return "call order violated";
    case ASRD_ERR_IO: return "I/O error";
  }
  return "unknown status";
}

int asrd_abi_version(void) { return ASRD_ABI_VERSION; }

int asrd_configure_process(void) {
  // more hardware work queues than the driver's default of 8 (see the note above); a value chosen
  // by the application wins.  Only effective before the process creates its CUDA context.
  return setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0) == 0 ? ASRD_OK : ASRD_ERR_BAD_ARG;
}

int asrd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return ASRD_ERR_CUDA;
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major >= 10) ++ok;
  }
  return ok;
}

int64_t asrd_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------- graph

// CLG graphs: per arc (in the order of `arcs`, which must then be eps-first already) the two weights
// of an arc that leaves a CLG state through an HMM, and which arcs those are
struct ClgExtras {
  const float *w_clg, *w_hmm;
  const unsigned char *two_level;
};

static int GraphCreate(const asrd_arc *arcs, const uint32_t *num_arcs, const uint32_t *niepsilons,
                       int32_t total_states, int64_t total_arcs, int32_t start, int32_t final_state,
                       int device, const ClgExtras *clg, asrd_graph **out);

int asrd_graph_create(const asrd_arc *arcs, const uint32_t *num_arcs, const uint32_t *niepsilons,
                      int32_t total_states, int64_t total_arcs, int32_t start, int32_t final_state,
                      int device, asrd_graph **out) {
  return GraphCreate(arcs, num_arcs, niepsilons, total_states, total_arcs, start, final_state, device, nullptr, out);
}

static int GraphCreate(const asrd_arc *arcs, const uint32_t *num_arcs, const uint32_t *niepsilons,
                       int32_t total_states, int64_t total_arcs, int32_t start, int32_t final_state,
                       int device, const ClgExtras *clg, asrd_graph **out) {
  if (!arcs || !num_arcs || !niepsilons || !out || total_states <= 0 || total_arcs < 0 ||
      total_arcs >= 0xFFFFFFFFll || start < 0 || start >= total_states)
    return ASRD_ERR_BAD_ARG;
  int rc = EnsureDevice(device);
  if (rc) return rc;
  const int32_t S = total_states;
  const int64_t A = total_arcs;
  std::vector<uint2> rows((size_t)S + 1);
  std::vector<asrd_arc> parc((size_t)std::max<int64_t>(A, 1));
  std::vector<uint32_t> src((size_t)std::max<int64_t>(A, 1));
  std::vector<uint32_t> par((size_t)(A + 31) / 32 + 1, 0u);
  std::vector<uint32_t> epsb((size_t)(S + 31) / 32 + 1, 0u);
  std::vector<std::pair<int32_t, uint32_t>> tmp;
  int64_t off = 0;
  for (int32_t s = 0; s < S; ++s) {
    const uint32_t n = num_arcs[s];
    if (off + n > A) return ASRD_ERR_BAD_ARG;
    // stable partition: input-epsilon arcs first (no-op for convert_fst / Fst(ConstFst) graphs,
    // reference src/newfst/optimize-fst.h:107-119)
    uint32_t ne = 0;
    for (uint32_t k = 0; k < n; ++k)
      if (arcs[off + k].ilabel == 0) parc[off + ne++] = arcs[off + k];
    uint32_t w = ne;
    for (uint32_t k = 0; k < n; ++k)
      if (arcs[off + k].ilabel != 0) parc[off + w++] = arcs[off + k];
    if (clg)  // (the per-arc extras follow the caller's order: the rows must be eps-first already)
      for (uint32_t k = 0; k < n; ++k)
        if ((arcs[off + k].ilabel == 0) != (k < ne)) return ASRD_ERR_BAD_ARG;
    rows[s] = make_uint2((uint32_t)off, (uint32_t)(off + ne));
    if (ne) epsb[(size_t)s >> 5] |= 1u << (s & 31);
    for (uint32_t k = 0; k < n; ++k) {
      src[off + k] = (uint32_t)s;
      const asrd_arc &a = parc[off + k];
      if (a.nextstate < 0 || a.nextstate >= S || a.ilabel < 0) return ASRD_ERR_BAD_ARG;
    }
    // sibling flags: same class (eps / emitting), same destination
    for (int cls = 0; cls < 2; ++cls) {
      const uint32_t b = cls == 0 ? 0 : ne, e = cls == 0 ? ne : n;
      if (e - b < 2) continue;
      tmp.clear();
      for (uint32_t k = b; k < e; ++k) tmp.emplace_back(parc[off + k].nextstate, (uint32_t)(off + k));
      std::sort(tmp.begin(), tmp.end());
      for (size_t k = 0; k + 1 < tmp.size(); ++k)
        if (tmp[k].first == tmp[k + 1].first) {
          par[tmp[k].second >> 5] |= 1u << (tmp[k].second & 31);
          par[tmp[k + 1].second >> 5] |= 1u << (tmp[k + 1].second & 31);
        }
    }
    off += n;
  }
  if (off != A) return ASRD_ERR_BAD_ARG;
  rows[S] = make_uint2((uint32_t)A, (uint32_t)A);
  std::vector<uint2> erows((size_t)S);
  for (int32_t s = 0; s < S; ++s) erows[s] = make_uint2(rows[s].y, rows[s + 1].x);
  // device copy only: flag arcs whose destination has eps arcs (saves a bitmap probe per admitted arc)
  for (int64_t a = 0; a < A; ++a) {
    const uint32_t ns = (uint32_t)parc[a].nextstate;
    if ((epsb[ns >> 5] >> (ns & 31)) & 1u) parc[a].nextstate = (int32_t)(ns | kDestEpsBit);
  }

  // incoming-arc index (counting sort by destination; arc ids ascend inside a group): the
  // trace-back of plain decoders finds the arc that set a token's cost through it (k_best_path_rev)
  std::vector<uint32_t> in_off((size_t)S + 1, 0u), in_arc((size_t)std::max<int64_t>(A, 1));
  int32_t max_ilabel = 0;
  for (int64_t a = 0; a < A; ++a) {
    ++in_off[((uint32_t)parc[a].nextstate & kStateMask) + 1];
    max_ilabel = std::max(max_ilabel, parc[a].ilabel);
  }
  for (int32_t s2 = 0; s2 < S; ++s2) in_off[s2 + 1] += in_off[s2];
  // inside a group the eps arcs come first (in_mid: where the emitting ones start): the arena prune
  // walks the two classes separately (k_prune)
  std::vector<uint32_t> in_mid((size_t)S, 0u);
  {
    for (int64_t a = 0; a < A; ++a)
      if (parc[a].ilabel == 0) ++in_mid[(uint32_t)parc[a].nextstate & kStateMask];
    for (int32_t s2 = 0; s2 < S; ++s2) in_mid[s2] += in_off[s2];
    std::vector<uint32_t> fill_eps(in_off.begin(), in_off.end() - 1), fill_emit(in_mid);
    for (int64_t a = 0; a < A; ++a) {
      const uint32_t d = (uint32_t)parc[a].nextstate & kStateMask;
      in_arc[parc[a].ilabel == 0 ? fill_eps[d]++ : fill_emit[d]++] = (uint32_t)a;
    }
  }

  // eps rows with the first eps arc inline (the closure of k_stream)
  std::vector<uint4> eps_rows((size_t)S);
  for (int32_t s2 = 0; s2 < S; ++s2) {
    uint4 r = make_uint4(rows[s2].x, rows[s2].y, 0u, 0u);
    if (r.y > r.x) {
      memcpy(&r.z, &parc[r.x].weight, 4);
      r.w = (uint32_t)parc[r.x].nextstate;
    }
    eps_rows[s2] = r;
  }

  asrd_graph *g = new asrd_graph();
  memset(g, 0, sizeof(*g));
  g->refs = 1;
  g->device = device;
  g->total_arcs = A;
  g->max_ilabel = max_ilabel;
  const size_t b_arcs = sizeof(asrd_arc) * parc.size(), b_rows = sizeof(uint2) * rows.size(),
               b_src = sizeof(uint32_t) * src.size(), b_par = sizeof(uint32_t) * par.size(),
               b_eps = sizeof(uint32_t) * epsb.size(), b_erows = sizeof(uint2) * std::max<size_t>(erows.size(), 1),
               b_ioff = sizeof(uint32_t) * in_off.size(), b_iarc = sizeof(uint32_t) * in_arc.size(),
               b_imid = sizeof(uint32_t) * in_mid.size(),
               b_epsr = sizeof(uint4) * eps_rows.size();
  struct Up { void **dst; const void *src; size_t bytes; };
  const Up ups[] = {{&g->d_arcs, parc.data(), b_arcs}, {&g->d_rows, rows.data(), b_rows},
                    {&g->d_erows, erows.data(), sizeof(uint2) * erows.size()}, {&g->d_arc_src, src.data(), b_src},
                    {&g->d_par, par.data(), b_par}, {&g->d_eps, epsb.data(), b_eps},
                    {&g->d_in_off, in_off.data(), b_ioff}, {&g->d_in_arc, in_arc.data(), b_iarc},
                    {&g->d_in_mid, in_mid.data(), b_imid},
                    {&g->d_eps_rows, eps_rows.data(), b_epsr}};
  for (const Up &u : ups) {
    if (cudaMalloc(u.dst, std::max<size_t>(u.bytes, 16)) != cudaSuccess) {
      cudaGetLastError();
      asrd_graph_destroy(g);
      return ASRD_ERR_NOMEM;
    }
    const cudaError_t e = cudaMemcpy(*u.dst, u.src, u.bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      g_last_error = std::string("graph upload: ") + cudaGetErrorString(e);
      asrd_graph_destroy(g);
      return ASRD_ERR_CUDA;
    }
  }
  g->device_bytes = (int64_t)(b_arcs + b_rows + b_erows + b_src + b_par + b_eps + b_ioff + b_iarc + b_imid + b_epsr);
  if (clg) {
    std::vector<uint32_t> bits((size_t)(A + 31) / 32 + 1, 0u);
    for (int64_t a = 0; a < A; ++a)
      if (clg->two_level[a]) bits[(size_t)a >> 5] |= 1u << (a & 31);
    const Up cups[] = {{&g->d_w_clg, clg->w_clg, sizeof(float) * (size_t)A}, {&g->d_w_hmm, clg->w_hmm, sizeof(float) * (size_t)A},
                       {&g->d_clg2, bits.data(), sizeof(uint32_t) * bits.size()}};
    for (const Up &u : cups) {
      if (cudaMalloc(u.dst, std::max<size_t>(u.bytes, 16)) != cudaSuccess ||
          cudaMemcpy(*u.dst, u.src, u.bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        asrd_graph_destroy(g);
        return ASRD_ERR_NOMEM;
      }
      g->device_bytes += (int64_t)u.bytes;
    }
    g->view.clg = 1;
    g->view.w_clg = (const float *)g->d_w_clg;
    g->view.w_hmm = (const float *)g->d_w_hmm;
    g->view.clg2_bits = (const uint32_t *)g->d_clg2;
  }
  g->view.arcs = (const int4 *)g->d_arcs;
  g->view.rows = (const uint2 *)g->d_rows;
  g->view.erows = (const uint2 *)g->d_erows;
  g->view.eps_rows = (const uint4 *)g->d_eps_rows;
  g->view.arc_src = (const uint32_t *)g->d_arc_src;
  g->view.in_off = (const uint32_t *)g->d_in_off;
  g->view.in_arc = (const uint32_t *)g->d_in_arc;
  g->view.in_mid = (const uint32_t *)g->d_in_mid;
  g->view.par_bits = (const uint32_t *)g->d_par;
  g->view.eps_bits = (const uint32_t *)g->d_eps;
  g->view.n_states = S;
  g->view.n_arcs = (uint32_t)A;
  g->view.start = start;
  g->view.final_state = final_state;
  *out = g;
  return ASRD_OK;
}

int asrd_graph_read(const char *path, int device, asrd_graph **out) {
  // Fst::ReadFst, src/newfst/optimize-fst.h:226-280
  if (!path || !out) return ASRD_ERR_BAD_ARG;
  FILE *fp = fopen(path, "rb");
  if (!fp) return ASRD_ERR_IO;
  int32_t hdr[6];
  if (fread(hdr, 4, 6, fp) != 6 || hdr[2] <= 0 || hdr[3] < 0) {
    fclose(fp);
    return ASRD_ERR_IO;
  }
  const int32_t S = hdr[2];
  const int64_t A = hdr[3];
  std::vector<uint32_t> info((size_t)S * 3);
  std::vector<asrd_arc> arcs((size_t)std::max<int64_t>(A, 1));
  bool ok = fread(info.data(), 12, (size_t)S, fp) == (size_t)S &&
            fread(arcs.data(), 16, (size_t)A, fp) == (size_t)A;
  fclose(fp);
  if (!ok) return ASRD_ERR_IO;
  std::vector<uint32_t> na(S), ne(S);
  for (int32_t s = 0; s < S; ++s) {
    na[s] = info[(size_t)s * 3];
    ne[s] = info[(size_t)s * 3 + 1];
  }
  return asrd_graph_create(arcs.data(), na.data(), ne.data(), S, A, hdr[0], hdr[1], device, out);
}

namespace {
struct FlatFst {  // one newfst graph as Fst::ReadFst(FILE*) reads it (optimize-fst.h:226-280)
  int32_t start = 0, final_state = 0, n_states = 0;
  std::vector<uint32_t> num_arcs, off;
  std::vector<asrd_arc> arcs;
};

bool ReadFlat(FILE *fp, FlatFst *f) {
  int32_t hdr[6];
  if (fread(hdr, 4, 6, fp) != 6 || hdr[2] <= 0 || hdr[3] < 0) return false;
  f->start = hdr[0];
  f->final_state = hdr[1];
  f->n_states = hdr[2];
  std::vector<uint32_t> info((size_t)hdr[2] * 3);
  f->arcs.resize((size_t)hdr[3]);
  if (fread(info.data(), 12, (size_t)hdr[2], fp) != (size_t)hdr[2]) return false;
  if (hdr[3] && fread(f->arcs.data(), 16, (size_t)hdr[3], fp) != (size_t)hdr[3]) return false;
  f->num_arcs.resize((size_t)hdr[2]);
  f->off.assign((size_t)hdr[2] + 1, 0u);
  for (int32_t s = 0; s < hdr[2]; ++s) {
    f->num_arcs[s] = info[(size_t)s * 3];
    f->off[s + 1] = f->off[s] + f->num_arcs[s];
  }
  return f->off[hdr[2]] == (uint32_t)hdr[3];
}
}  // namespace

// ClgFst::Init(clgfst, hmmfst) (my-decoder/clg-fst.h:17-74): the CLG graph and the HMM set, written
// out as ONE static graph over the reference's own two-level state ids — CLG state s keeps its id,
// the copy of state k of the HMM inside CLG arc a is a + offset * (k + 1), offset = total arcs + 1
// (GetState / MapClgTokenStateId, clg-fst.h:82-165) — with the arcs in the order the reference's
// CLG decoder enumerates them (online-clg-decoder-mempool-base.h:122-205):
//   CLG state, eps arc: unchanged;
//   CLG state, arc a with HMM id h: one arc per EMITTING arc e of state 0 of HMM h — ilabel of e (a
//     pdf), olabel of the CLG arc, weight e.w + clg.w, destination a + offset if e is the self-loop
//     of state 0, else a + 2 offset;
//   HMM copy (a, k): the arcs of state k of the HMM without olabels (RmOlalel): an emitting arc stays
//     (self-loop) or goes to (a, k + 1) whatever its target, the eps arc (HMM end) goes to the CLG
//     arc's destination.
// Decoders on such a graph follow the reference's CLG decoder (see GraphView::clg).
int asrd_graph_read_clg(const char *clg_path, const char *hmm_path, int device, asrd_graph **out) {
  if (!clg_path || !hmm_path || !out) return ASRD_ERR_BAD_ARG;
  FlatFst clg;
  std::vector<FlatFst> hmms;
  {
    FILE *fp = fopen(clg_path, "rb");
    if (!fp) return ASRD_ERR_IO;
    const bool ok = ReadFlat(fp, &clg);
    fclose(fp);
    if (!ok) return ASRD_ERR_IO;
    fp = fopen(hmm_path, "rb");
    if (!fp) return ASRD_ERR_IO;
    int32_t n = 0;
    bool okh = fread(&n, 4, 1, fp) == 1 && n >= 0;
    hmms.resize(okh ? (size_t)n : 0);
    for (int32_t i = 0; okh && i < n; ++i) okh = ReadFlat(fp, &hmms[i]);
    fclose(fp);
    if (!okh) return ASRD_ERR_IO;
  }
  const int64_t A = (int64_t)clg.arcs.size();
  const int64_t offset = A + 1;
  int32_t kmax = 1;
  for (const FlatFst &h : hmms) kmax = std::max(kmax, h.n_states);
  const int64_t n_ids = offset * ((int64_t)kmax + 1);
  if (clg.n_states > offset || n_ids >= 0x7FFFFFFFll) return ASRD_ERR_BAD_ARG;  // clg-fst.h:25 asserts the same
  struct Out { asrd_arc arc; float w_clg, w_hmm; unsigned char two; };
  std::vector<std::vector<Out>> rows((size_t)n_ids);
  auto push = [&](int64_t id, int32_t il, int32_t ol, float w, int64_t to, float wc, float wh, bool two) {
    Out o;
    o.arc.ilabel = il; o.arc.olabel = ol; o.arc.weight = w; o.arc.nextstate = (int32_t)to;
    o.w_clg = wc; o.w_hmm = wh; o.two = two ? 1 : 0;
    rows[(size_t)id].push_back(o);
  };
  for (int32_t s = 0; s < clg.n_states; ++s) {
    for (int pass = 0; pass < 2; ++pass)  // eps arcs first (stable), as the device layout wants them
      for (uint32_t a = clg.off[s]; a < clg.off[s + 1]; ++a) {
        const asrd_arc &ca = clg.arcs[a];
        if ((ca.ilabel == 0) != (pass == 0)) continue;
        if (ca.ilabel == 0) {
          push(s, 0, ca.olabel, ca.weight, ca.nextstate, 0.f, ca.weight, false);
          continue;
        }
        if (ca.ilabel < 1 || (size_t)ca.ilabel > hmms.size()) return ASRD_ERR_BAD_ARG;
        const FlatFst &h = hmms[(size_t)ca.ilabel - 1];
        for (uint32_t e = h.off[0]; e < h.off[1]; ++e) {
          const asrd_arc &ea = h.arcs[e];
          if (ea.ilabel == 0) continue;
          push(s, ea.ilabel, ca.olabel, ea.weight + ca.weight, ea.nextstate == 0 ? a + offset : a + 2 * offset,
               ca.weight, ea.weight, true);
        }
      }
    for (uint32_t a = clg.off[s]; a < clg.off[s + 1]; ++a) {
      const asrd_arc &ca = clg.arcs[a];
      if (ca.ilabel == 0) continue;
      const FlatFst &h = hmms[(size_t)ca.ilabel - 1];
      for (int32_t k = 0; k < h.n_states; ++k) {
        const int64_t sid = a + offset * ((int64_t)k + 1);
        for (int pass = 0; pass < 2; ++pass)
          for (uint32_t e = h.off[k]; e < h.off[k + 1]; ++e) {
            const asrd_arc &ea = h.arcs[e];
            if ((ea.ilabel == 0) != (pass == 0)) continue;
            if (ea.ilabel == 0) push(sid, 0, 0, ea.weight, ca.nextstate, 0.f, ea.weight, false);
            else push(sid, ea.ilabel, 0, ea.weight, ea.nextstate == k ? sid : sid + offset, 0.f, ea.weight, false);
          }
      }
    }
  }
  std::vector<uint32_t> na((size_t)n_ids), ne((size_t)n_ids);
  int64_t total = 0;
  for (int64_t i = 0; i < n_ids; ++i) {
    na[i] = (uint32_t)rows[i].size();
    uint32_t e = 0;
    for (const Out &o : rows[i]) e += o.arc.ilabel == 0;
    ne[i] = e;
    total += na[i];
  }
  std::vector<asrd_arc> arcs((size_t)std::max<int64_t>(total, 1));
  std::vector<float> wc((size_t)std::max<int64_t>(total, 1)), wh((size_t)std::max<int64_t>(total, 1));
  std::vector<unsigned char> two((size_t)std::max<int64_t>(total, 1));
  int64_t p = 0;
  for (int64_t i = 0; i < n_ids; ++i)
    for (const Out &o : rows[i]) {
      arcs[p] = o.arc; wc[p] = o.w_clg; wh[p] = o.w_hmm; two[p] = o.two;
      ++p;
    }
  const ClgExtras ex = {wc.data(), wh.data(), two.data()};
  return GraphCreate(arcs.data(), na.data(), ne.data(), (int32_t)n_ids, total, clg.start, clg.final_state, device, &ex, out);
}

int asrd_graph_read_const(const char *path, int device, asrd_graph **out) {
  // ConstFst<StdArc, int>::Read (src/newfst/const-fst.h:46-80,189-221) + Fst(const ConstFst&)
  // (src/newfst/optimize-fst.h:82-134): an OpenFst "const" FST over the standard (tropical) arc
  // becomes the flat graph — one super-final state appended, every final state gets
  // 0:0/final_weight -> super-final as its FIRST arc.  Like the reference's reader: no symbol
  // tables, no alignment padding.
  if (!path || !out) return ASRD_ERR_BAD_ARG;
  FILE *fp = fopen(path, "rb");
  if (!fp) return ASRD_ERR_IO;
  auto fail = [&]() {
    fclose(fp);
    return ASRD_ERR_IO;
  };
  auto read_string = [&](std::string *str) {
    int32_t n = 0;
    if (fread(&n, 4, 1, fp) != 1 || n < 0 || n > 64) return false;
    str->resize((size_t)n);
    return n == 0 || fread(&(*str)[0], 1, (size_t)n, fp) == (size_t)n;
  };
  int32_t magic = 0, version = 0, flags = 0;
  uint64_t properties = 0;
  int64_t start = 0, n_states = 0, n_arcs = 0;
  std::string fsttype, arctype;
  if (fread(&magic, 4, 1, fp) != 1 || magic != 2125659606 || !read_string(&fsttype) || !read_string(&arctype) ||
      fsttype != "const" || arctype != "standard" || fread(&version, 4, 1, fp) != 1 || fread(&flags, 4, 1, fp) != 1 ||
      fread(&properties, 8, 1, fp) != 1 || fread(&start, 8, 1, fp) != 1 || fread(&n_states, 8, 1, fp) != 1 ||
      fread(&n_arcs, 8, 1, fp) != 1 || n_states <= 0 || n_states >= 0x7FFFFFF0ll || n_arcs < 0 || n_arcs >= 0xFFFFFFF0ll)
    return fail();
  struct ConstState {  // const-fst.h:206-215
    float weight;
    int32_t pos, narcs, niepsilons, noepsilons;
  };
  std::vector<ConstState> cs((size_t)n_states);
  std::vector<asrd_arc> in((size_t)std::max<int64_t>(n_arcs, 1));
  if (fread(cs.data(), sizeof(ConstState), (size_t)n_states, fp) != (size_t)n_states ||
      fread(in.data(), sizeof(asrd_arc), (size_t)n_arcs, fp) != (size_t)n_arcs)
    return fail();
  fclose(fp);
  const int32_t S = (int32_t)n_states + 1;
  int64_t n_final = 0;
  for (int64_t s = 0; s < n_states; ++s) n_final += cs[s].weight != std::numeric_limits<float>::infinity();
  std::vector<asrd_arc> arcs((size_t)std::max<int64_t>(n_arcs + n_final, 1));
  std::vector<uint32_t> na((size_t)S, 0u), ne((size_t)S, 0u);
  int64_t w = 0;
  for (int64_t s = 0; s < n_states; ++s) {
    const ConstState &c = cs[s];
    if (c.pos < 0 || c.narcs < 0 || (int64_t)c.pos + c.narcs > n_arcs) return ASRD_ERR_IO;
    const bool fin = c.weight != std::numeric_limits<float>::infinity();  // Weight::Zero() marks non-final states
    if (fin) {
      asrd_arc a;
      a.ilabel = 0;
      a.olabel = 0;
      a.weight = c.weight;
      a.nextstate = S - 1;
      arcs[w++] = a;
    }
    std::copy(in.begin() + c.pos, in.begin() + c.pos + c.narcs, arcs.begin() + w);
    w += c.narcs;
    na[s] = (uint32_t)c.narcs + (fin ? 1u : 0u);
    ne[s] = (uint32_t)c.niepsilons + (fin ? 1u : 0u);
  }
  return asrd_graph_create(arcs.data(), na.data(), ne.data(), S, w, (int32_t)start, S - 1, device, out);
}

// The handle is reference counted: decoders built on a graph keep it alive, so the caller may
// release graph and decoders in any order (the reference shares one read-only Fst among its
// decoder threads and never frees it under them).
int asrd_graph_destroy(asrd_graph *g) {
  if (!g) return ASRD_OK;
  {
    std::lock_guard<std::mutex> lk(g_ref_mu);
    if (--g->refs > 0) return ASRD_OK;
  }
  cudaSetDevice(g->device);
  cudaFree(g->d_arcs);
  cudaFree(g->d_rows);
  cudaFree(g->d_erows);
  cudaFree(g->d_eps_rows);
  cudaFree(g->d_arc_src);
  cudaFree(g->d_par);
  cudaFree(g->d_eps);
  cudaFree(g->d_in_off);
  cudaFree(g->d_in_arc);
  cudaFree(g->d_in_mid);
  cudaFree(g->d_w_clg);
  cudaFree(g->d_w_hmm);
  cudaFree(g->d_clg2);
  delete g;
  return ASRD_OK;
}

int asrd_graph_info(const asrd_graph *g, int32_t *total_states, int64_t *total_arcs, int32_t *start,
                    int32_t *final_state, int64_t *device_bytes) {
  if (!g) return ASRD_ERR_BAD_ARG;
  if (total_states) *total_states = g->view.n_states;
  if (total_arcs) *total_arcs = g->total_arcs;
  if (start) *start = g->view.start;
  if (final_state) *final_state = g->view.final_state;
  if (device_bytes) *device_bytes = g->device_bytes;
  return ASRD_OK;
}

// ------------------------------------------------------------------------- decoder

static int DecoderCreate(asrd_graph *g, const asrd_config *cfg, const asrd_device_options *opts, asrd_lm *lm1,
                         asrd_lm *lm2, asrd_decoder **out) {
  if (!g || !cfg || !out) return ASRD_ERR_BAD_ARG;
  if ((lm1 == nullptr) != (lm2 == nullptr)) return ASRD_ERR_BAD_ARG;
  if (lm1 && (lm1->device != g->device || lm2->device != g->device)) return ASRD_ERR_BAD_ARG;
  // LatticeFasterDecoderConfig::Check, src/my-decoder/lattice-faster-decoder-conf.h:62-67
  if (!(cfg->beam > 0.0f && cfg->max_active > 1 && cfg->lattice_beam > 0.0f && cfg->beam_delta > 0.0f) ||
      cfg->min_active < 0)
    return ASRD_ERR_BAD_ARG;
  int rc = EnsureDevice(g->device);
  if (rc) return rc;
  asrd_decoder *d = new asrd_decoder();
  memset((void *)d, 0, sizeof(*d));
  d->device = g->device;
  d->graph = g;
  d->lm1 = lm1;
  d->lm2 = lm2;
  d->cfg = *cfg;
  if (opts) d->opts = *opts;
  asrd_device_options &o = d->opts;
  if (o.hash_capacity <= 0) {
    // every surviving token can fan out; 8x max_active keeps the load factor low
    uint64_t want = (uint64_t)std::min<int64_t>((int64_t)cfg->max_active, 1 << 20) * 8;
    o.hash_capacity = (int32_t)std::min<uint64_t>(std::max<uint64_t>(NextPow2(want), 4096), 1u << 24);
  } else {
    o.hash_capacity = (int32_t)std::max<uint32_t>(NextPow2((uint64_t)o.hash_capacity), kMinHashCapacity);
  }
  if (o.max_frames <= 0) o.max_frames = 2048;
  if (o.token_capacity <= 0)  // 8 bytes per token record {state, cost}; ~1.5 x max_active survivors per frame
    o.token_capacity = std::min<int64_t>((int64_t)o.max_frames * std::min<int64_t>(cfg->max_active, 1 << 16) * 3 / 2,
                                         (int64_t)1 << 23);
  if (o.token_capacity >= 0xFFFFFFF0ll) {
    delete d;
    return ASRD_ERR_BAD_ARG;
  }
  const size_t H = (size_t)o.hash_capacity;
  auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
  // biglm: interned (lm1 state, lm2 state) pairs of one utterance (the reference's unordered_map
  // is unbounded, newlm/diff-lm.h:92-103); sized by lm_pair_capacity, default 2^16
  size_t pair_cap = 0;
  if (lm1) {
    if (o.lm_pair_capacity <= 0) o.lm_pair_capacity = 1 << 16;
    pair_cap = NextPow2((uint64_t)std::max(o.lm_pair_capacity, 1024));
    o.lm_pair_capacity = (int32_t)pair_cap;
  }
  const size_t b_hash = align(H * sizeof(HashEntry)), b_list = align(H * 4), b_bm = align(H / 8),
               b_tok = align((size_t)o.token_capacity * 8),
               b_arc = lm1 ? align((size_t)o.token_capacity * 4) : 0,  // plain decoders keep no arc per token
               b_off = align(((size_t)o.max_frames + 4) * 4),
               b_stats = o.collect_stats ? align(((size_t)o.max_frames + 1) * sizeof(asrd_frame_stat)) : 0,
               b_state = align(sizeof(StreamState));
  const size_t b_lm = b_arc, b_pair = align(pair_cap * 8);
  if (g->view.clg && lm1) {  // CLG graphs: plain decoders
    delete d;
    return ASRD_ERR_BAD_ARG;
  }
  if (o.prune_tokens && (lm1 || cfg->prune_interval <= 0)) {  // arena pruning: plain decoders, a positive interval
    delete d;
    return ASRD_ERR_BAD_ARG;
  }
  const size_t b_extra = o.prune_tokens ? align((size_t)o.token_capacity * 4) : 0;
  const size_t total = b_state + b_hash + 2 * b_bm + 3 * b_list + b_tok + b_arc + 3 * b_off + b_stats + b_lm + b_pair + b_extra;
  if (cudaMalloc(&d->slab, total) != cudaSuccess) {
    cudaGetLastError();
    delete d;
    return ASRD_ERR_NOMEM;
  }
  d->slab_bytes = (int64_t)total;
  char *p = (char *)d->slab;
  StreamState &h = d->h_state;
  memset(&h, 0, sizeof(h));
  d->d_state = (StreamState *)p; p += b_state;
  h.hash = (HashEntry *)p; p += b_hash;
  h.bm = (uint32_t *)p; p += b_bm;
  h.ebm = (uint32_t *)p; p += b_bm;
  for (int i = 0; i < 2; ++i) { h.queue[i] = (uint32_t *)p; p += b_list; }
  h.stamp = (uint32_t *)p; p += b_list;
  if (lm1) {
    h.tok_lm = (uint32_t *)p; p += b_lm;
    h.pair_map = (unsigned long long *)p; p += b_pair;
    h.pair_mask = (uint32_t)pair_cap - 1;
  }
  h.tok_sc = (uint2 *)p; p += b_tok;
  h.tok_arc = lm1 ? (uint32_t *)p : nullptr; p += b_arc;
  h.tok_extra = o.prune_tokens ? (uint32_t *)p : nullptr; p += b_extra;
  h.frame_off = (uint32_t *)p; p += b_off;
  h.frame_nc = (float *)p; p += b_off;
  h.frame_cur = (float *)p; p += b_off;
  h.stats = o.collect_stats ? (asrd_frame_stat *)p : nullptr;
  h.hash_mask = (uint32_t)H - 1;
  uint32_t lg = 0;
  while ((1u << lg) < H) ++lg;
  h.hash_shift = 32 - lg;
  h.token_capacity = (uint32_t)o.token_capacity;
  h.max_frames = o.max_frames;
  // empty maps: key = 0xFFFFFFFF, val = +inf; empty bitmaps
  if (cudaMemset(h.hash, 0xFF, b_hash) != cudaSuccess ||
      cudaMemset(h.bm, 0, 2 * b_bm) != cudaSuccess ||
      cudaMemcpy(d->d_state, &h, sizeof(h), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d->slab);
    delete d;
    return ASRD_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_ref_mu);
    ++g->refs;
    if (lm1) {
      ++lm1->refs;
      ++lm2->refs;
    }
  }
  *out = d;
  return ASRD_OK;
}

int asrd_decoder_create(asrd_graph *g, const asrd_config *cfg, const asrd_device_options *opts,
                        asrd_decoder **out) {
  return DecoderCreate(g, cfg, opts, nullptr, nullptr, out);
}

int asrd_decoder_create_biglm(asrd_graph *g, const asrd_config *cfg, const asrd_device_options *opts,
                              asrd_lm *old_lm, asrd_lm *new_lm, asrd_decoder **out) {
  if (!old_lm || !new_lm) return ASRD_ERR_BAD_ARG;
  return DecoderCreate(g, cfg, opts, old_lm, new_lm, out);
}

int asrd_lm_create(int32_t bos, int32_t eos, int32_t n_states, const int32_t *arc_num,
                   const float *backoff_prob, const int32_t *backoff_id, const asrd_lm_arc *arcs,
                   int64_t n_arcs, int device, asrd_lm **out) {
  if (!arc_num || !backoff_prob || !backoff_id || !arcs || !out || n_states <= 0 || n_arcs <= 0 ||
      n_arcs >= 0x7FFFFFFFll || bos <= 0 || eos <= 0)
    return ASRD_ERR_BAD_ARG;
  int rc = EnsureDevice(device);
  if (rc) return rc;
  std::vector<uint32_t> off((size_t)n_states + 1, 0);
  for (int32_t i = 0; i < n_states; ++i) {
    if (arc_num[i] < 0 || backoff_id[i] < 0 || backoff_id[i] >= n_states) return ASRD_ERR_BAD_ARG;
    off[i + 1] = off[i] + (uint32_t)arc_num[i];
  }
  if ((int64_t)off[n_states] != n_arcs || bos >= arc_num[0] || eos >= arc_num[0]) return ASRD_ERR_BAD_ARG;
  std::vector<int32_t> word((size_t)n_arcs), to((size_t)n_arcs);
  std::vector<float> weight((size_t)n_arcs);
  for (int64_t a = 0; a < n_arcs; ++a) {
    word[a] = arcs[a].wordid;
    weight[a] = arcs[a].weight;
    to[a] = arcs[a].tostateid;
    if (to[a] < 0 || to[a] >= n_states) return ASRD_ERR_BAD_ARG;
  }
  for (int32_t w = 0; w < arc_num[0]; ++w)
    if (word[w] != w) return ASRD_ERR_BAD_ARG;  // the unigram state is direct-indexed by word id
  auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t b_off = align(off.size() * 4), b_a = align((size_t)n_arcs * 4), b_s = align((size_t)n_states * 4);
  asrd_lm *lm = new asrd_lm();
  memset(lm, 0, sizeof(*lm));
  lm->refs = 1;
  lm->device = device;
  if (cudaMalloc(&lm->slab, b_off + 3 * b_a + 2 * b_s) != cudaSuccess) {
    cudaGetLastError();
    delete lm;
    return ASRD_ERR_NOMEM;
  }
  char *p = (char *)lm->slab;
  struct Up { const void **view; const void *src; size_t bytes, stride; };
  const Up ups[] = {{(const void **)&lm->view.arc_off, off.data(), off.size() * 4, b_off},
                    {(const void **)&lm->view.arc_word, word.data(), (size_t)n_arcs * 4, b_a},
                    {(const void **)&lm->view.arc_weight, weight.data(), (size_t)n_arcs * 4, b_a},
                    {(const void **)&lm->view.arc_to, to.data(), (size_t)n_arcs * 4, b_a},
                    {(const void **)&lm->view.backoff_prob, backoff_prob, (size_t)n_states * 4, b_s},
                    {(const void **)&lm->view.backoff_id, backoff_id, (size_t)n_states * 4, b_s}};
  for (const Up &u : ups) {
    *u.view = p;
    const cudaError_t e = cudaMemcpy(p, u.src, u.bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      g_last_error = std::string("LM upload: ") + cudaGetErrorString(e);
      asrd_lm_destroy(lm);
      return ASRD_ERR_CUDA;
    }
    p += u.stride;
  }
  lm->view.bos = bos;
  lm->view.eos = eos;
  lm->view.n_states = n_states;
  *out = lm;
  return ASRD_OK;
}

int asrd_lm_destroy(asrd_lm *lm) {
  if (!lm) return ASRD_OK;
  {
    std::lock_guard<std::mutex> lk(g_ref_mu);
    if (--lm->refs > 0) return ASRD_OK;
  }
  cudaSetDevice(lm->device);
  cudaFree(lm->slab);
  delete lm;
  return ASRD_OK;
}

int asrd_decoder_destroy(asrd_decoder *d) {
  if (!d) return ASRD_OK;
  cudaSetDevice(d->device);
  cudaFree(d->slab);
  cudaFree(d->d_ll_hist);
  asrd_graph_destroy(d->graph);  // drops this decoder's references
  asrd_lm_destroy(d->lm1);
  asrd_lm_destroy(d->lm2);
  delete d;
  return ASRD_OK;
}

int asrd_init_decoding(asrd_decoder *const *decs, int32_t n, void *stream) {
  int rc = CheckBatch(decs, n);
  if (rc) return rc;
  if ((rc = EnsureDevice(decs[0]->graph->device))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(s);
  StreamState **d_streams;
  if ((rc = UploadStreams(decs, n, s, sc, &d_streams))) return rc;
  const GraphView gv = decs[0]->graph->view;
  const DecoderConfigDev cfg = DevCfg(decs[0]);
  FrameDesc *d_desc;
  CU_CHECK(sc.Alloc(&d_desc, (size_t)n));
  const LmPair lms = Lms(decs[0]);
  if (decs[0]->lm1) {
    k_init<true><<<n, kStreamThreads, 0, s>>>(d_streams, d_desc, gv, cfg, lms);
    k_post<true><<<n, kStreamThreads, 0, s>>>(d_streams, d_desc, gv, cfg, kModeEpi, lms);
  } else {
    k_init<false><<<n, kStreamThreads, 0, s>>>(d_streams, d_desc, gv, cfg, lms);
    k_post<false><<<n, kStreamThreads, 0, s>>>(d_streams, d_desc, gv, cfg, kModeEpi, lms);
  }
  g_launches += 2;
  CU_CHECK(cudaGetLastError());
  for (int i = 0; i < n; ++i) {
    decs[i]->frames_decoded = 0;
    decs[i]->last_prune_frame = 0;
    decs[i]->finalized = 0;
    decs[i]->initialized = 1;
  }
  return ASRD_OK;
}

int asrd_advance_decoding(asrd_decoder *const *decs, int32_t n, const float *const *loglikes,
                          const int32_t *n_frames, const int32_t *stride, int32_t num_indices,
                          int32_t max_num_frames, int32_t on_device, void *stream) {
  int rc = CheckBatch(decs, n);
  if (rc) return rc;
  if (!loglikes || !n_frames || !stride || num_indices <= 0) return ASRD_ERR_BAD_ARG;
  if ((rc = EnsureDevice(decs[0]->graph->device))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<int32_t> nf(n);
  int32_t max_nf = 0;
  for (int i = 0; i < n; ++i) {
    if (!decs[i]->initialized || decs[i]->finalized) return ASRD_ERR_STATE;  // inl.h:634-635
    if (n_frames[i] < 0 || stride[i] < num_indices || (n_frames[i] > 0 && !loglikes[i])) return ASRD_ERR_BAD_ARG;
    nf[i] = n_frames[i];
    if (max_num_frames >= 0) nf[i] = std::min(nf[i], max_num_frames);  // inl.h:643-648
    // frames beyond max_frames are never dropped silently: the call fails before anything is decoded
    if (nf[i] > decs[i]->opts.max_frames - decs[i]->frames_decoded) return ASRD_ERR_FRAMES_OVERFLOW;
    max_nf = std::max(max_nf, nf[i]);
  }
  // the graph's ilabels index the rows (column = ilabel - 1): a narrower matrix would be read out of bounds
  if (num_indices < decs[0]->graph->max_ilabel) return ASRD_ERR_BAD_ARG;
  if (max_nf == 0) return ASRD_OK;
  // log-likelihood history (read by the search and by the trace-back); sized on first use
  for (int i = 0; i < n; ++i) {
    asrd_decoder *d = decs[i];
    if (d->d_ll_hist && d->ll_cols != num_indices) return ASRD_ERR_BAD_ARG;  // columns must not change
    if (!d->d_ll_hist) {
      const int32_t hs = (num_indices + 3) & ~3;
      if (cudaMalloc((void **)&d->d_ll_hist, (size_t)d->opts.max_frames * hs * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return ASRD_ERR_NOMEM;
      }
      d->ll_cols = num_indices;
      d->h_state.ll_hist = d->d_ll_hist;
      d->h_state.ll_stride = hs;
      CU_CHECK(cudaMemcpyAsync(&d->d_state->ll_hist, &d->h_state.ll_hist, sizeof(float *), cudaMemcpyHostToDevice, s));
      CU_CHECK(cudaMemcpyAsync(&d->d_state->ll_stride, &d->h_state.ll_stride, sizeof(int32_t), cudaMemcpyHostToDevice, s));
    }
  }
  const GraphView gv = decs[0]->graph->view;
  const DecoderConfigDev cfg = DevCfg(decs[0]);
  const bool biglm = decs[0]->lm1 != nullptr;
  const LmPair lms = Lms(decs[0]);
  DeviceCtx *ctx = nullptr;
  if ((rc = GetCtx(decs[0]->graph->device, &ctx))) return rc;
  // Host threads may call the ABI concurrently on different decoder handles (the reference runs one
  // decoder per worker thread): the side streams and events of the device are shared, so the issue
  // phases of two calls must not interleave.  The device work itself still overlaps.
  std::lock_guard<std::mutex> issue_lock(ctx->issue_mu);
  Scratch sc(s);
  StreamState **d_streams;
  if ((rc = UploadStreams(decs, n, s, sc, &d_streams))) return rc;
  AdvanceParams *d_params;
  FrameDesc *d_desc;
  const size_t row = (size_t)num_indices;
  // resident rows are cut into chunks too: the chunk launches of the sub-batches interleave on the
  // SMs, so a batch that is not a multiple of the SM count does not leave a ragged last wave
  const int32_t chunk = std::min<int32_t>(max_nf, std::max(1, on_device ? EnvInt("ASRD_DEVICE_CHUNK", 32) : EnvInt("ASRD_HOST_CHUNK", kHostChunkFrames)));
  const int n_chunks = (max_nf + chunk - 1) / chunk;
  CU_CHECK(sc.Alloc(&d_params, (size_t)n * n_chunks));
  CU_CHECK(sc.Alloc(&d_desc, (size_t)n));

  // Sub-batch pipelining: the batch is cut into sub-batches that run their frame loops on
  // separate worker streams.  The per-stream post phase (k_post) is a latency-bound chain;
  // the expansion (k_expand) is throughput-bound and fills the machine, so expansions of
  // different sub-batches serialise by themselves while each k_post overlaps the expansion of
  // the other sub-batches.
  int sub = EnvInt("ASRD_SUBBATCH", 32);
  if (sub <= 0 || sub > n) sub = n;
  const int n_sub = (n + sub - 1) / sub;
  const int n_workers = std::max(1, std::min({n_sub, EnvInt("ASRD_WORKERS", 8), kMaxWorkers}));
  const bool single = n_sub == 1;
  std::vector<ExpandPlan> plans(n_sub);
  for (int b = 0; b < n_sub; ++b)
    if ((rc = PlanExpand(std::min(sub, n - b * sub), num_indices, biglm, &plans[b]))) return rc;
  StreamPlan splan;
  if ((rc = PlanStream(decs[0]->graph, num_indices, biglm, &splan))) return rc;

  // Host log-likelihoods: staged chunk by chunk through two device buffers, copied on a
  // side stream so the H2D of chunk k+1 overlaps the search of chunk k.
  float *d_stage[2] = {nullptr, nullptr};
  // Host rows of neighbouring streams that sit at one fixed distance from each other (one
  // [k, T, P] block, or several such blocks) are copied with ONE pitched copy per run and chunk
  // instead of one per stream.
  struct CopyRun { int first, count; };
  std::vector<CopyRun> runs;
  if (!on_device) {
    const size_t need = (size_t)n * chunk * row;
    if (ctx->stage_floats < need) {  // (grow-only; rare: every user of the old buffers has to be done)
      CU_CHECK(cudaDeviceSynchronize());
      for (int b = 0; b < 2; ++b) {
        cudaFree(ctx->stage[b]);
        ctx->stage[b] = nullptr;
        ctx->stage_busy[b] = false;
      }
      ctx->stage_floats = 0;
      for (int b = 0; b < 2; ++b)
        if (cudaMalloc((void **)&ctx->stage[b], need * sizeof(float)) != cudaSuccess) {
          cudaGetLastError();
          return ASRD_ERR_NOMEM;
        }
      ctx->stage_floats = need;
    }
    d_stage[0] = ctx->stage[0];
    d_stage[1] = ctx->stage[1];
    for (int i = 0; i < n;) {
      int j = i + 1;
      if (j < n && stride[i] == num_indices && nf[i] > 0) {
        const ptrdiff_t pitch = loglikes[j] - loglikes[i];
        while (j < n && nf[j] == nf[i] && stride[j] == num_indices && loglikes[j] - loglikes[j - 1] == pitch &&
               pitch >= (ptrdiff_t)((size_t)nf[i] * row))
          ++j;
      }
      runs.push_back(CopyRun{i, j - i});
      i = j;
    }
  }
  std::vector<AdvanceParams> hp((size_t)n * n_chunks);
  std::vector<int32_t> steps_of((size_t)n_chunks * n_sub, 0);
  for (int k = 0; k < n_chunks; ++k) {
    const int32_t f0 = k * chunk;
    for (int i = 0; i < n; ++i) {
      AdvanceParams &p = hp[(size_t)k * n + i];
      const int32_t c = std::max(0, std::min(chunk, nf[i] - f0));
      p.n_frames = c;
      p.frame0 = decs[i]->frames_decoded + std::min(f0, nf[i]);
      if (on_device) {
        p.ll = loglikes[i] + (size_t)f0 * stride[i];
        p.stride = stride[i];
      } else {
        p.ll = d_stage[k & 1] + (size_t)i * chunk * row;
        p.stride = num_indices;
      }
      int32_t &st = steps_of[(size_t)k * n_sub + i / sub];
      st = std::max(st, c);
    }
  }
  CU_CHECK(cudaMemcpyAsync(d_params, hp.data(), sizeof(AdvanceParams) * hp.size(), cudaMemcpyHostToDevice, s));
  // The row scatter of every chunk runs on a high-priority stream: the frame loops keep every SM
  // busy (k_stream takes a whole SM), and a scatter queued at normal priority would wait for a
  // free SM, stalling the staging double buffer and with it the host->device copies.
  cudaStream_t ps = single ? s : ctx->prep;
  if (!single) {
    CU_CHECK(cudaEventRecord(ctx->ev_fork, s));  // stream table, parameters and staging buffers exist
    CU_CHECK(cudaStreamWaitEvent(ps, ctx->ev_fork, 0));
  }

  Profiler prof(g_profile.load());
  const bool trace = EnvInt("ASRD_TRACE", 0) != 0;
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  std::vector<cudaEvent_t> tr_copy, tr_rows, tr_w0, tr_join;  // ASRD_TRACE: device timeline of the pipeline
  cudaEvent_t tr_start = nullptr;
  if (trace) {
    cudaEventCreate(&tr_start);
    cudaEventRecord(tr_start, s);
  }
  auto tr_mark = [&](std::vector<cudaEvent_t> &v, cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    v.push_back(e);
  };
  struct EventList {  // rows of chunk k are in the histories; destroyed on every exit path
    std::vector<cudaEvent_t> v;
    ~EventList() {
      for (cudaEvent_t e : v) cudaEventDestroy(e);
    }
  } ev_rows_holder;
  std::vector<cudaEvent_t> &ev_rows = ev_rows_holder.v;
  for (int k = 0; k < n_chunks; ++k) {
    const int32_t f0 = k * chunk;
    float *stage = d_stage[k & 1];
    // ---- rows of this chunk -> device (copy stream) -> per-stream histories (stream s)
    if (!on_device) {
      // the buffer's previous rows (of this call or of an earlier one) must be in the histories
      if (ctx->stage_busy[k & 1]) CU_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[k & 1], 0));
      for (const CopyRun &r : runs) {
        const int i = r.first;
        const int32_t c = hp[(size_t)k * n + i].n_frames;
        if (c <= 0) continue;
        if (r.count == 1)
          CU_CHECK(cudaMemcpy2DAsync(stage + (size_t)i * chunk * row, row * 4, loglikes[i] + (size_t)f0 * stride[i],
                                     (size_t)stride[i] * 4, row * 4, (size_t)c, cudaMemcpyHostToDevice,
                                     ctx->copy_stream));
        else  // c rows of every stream of the run are one contiguous piece: `count` pieces at a fixed pitch
          CU_CHECK(cudaMemcpy2DAsync(stage + (size_t)i * chunk * row, (size_t)chunk * row * 4,
                                     loglikes[i] + (size_t)f0 * row, (size_t)(loglikes[i + 1] - loglikes[i]) * 4,
                                     (size_t)c * row * 4, (size_t)r.count, cudaMemcpyHostToDevice, ctx->copy_stream));
      }
      CU_CHECK(cudaEventRecord(ctx->ev_copied[k & 1], ctx->copy_stream));
      tr_mark(tr_copy, ctx->copy_stream);
      CU_CHECK(cudaStreamWaitEvent(ps, ctx->ev_copied[k & 1], 0));
    }
    if (trace) fprintf(stderr, "[asrd] chunk %d copies issued at %.2f ms\n", k, since());
    if (EnvInt("ASRD_SKIP_SCATTER", 0))  // measurement aid: rows assumed to be in the histories already
      k_begin_advance<<<dim3(1, (unsigned)n), 32, 0, ps>>>(d_streams, d_params + (size_t)k * n, 0);
    else
      k_begin_advance<<<dim3(8, (unsigned)n), 256, 0, ps>>>(d_streams, d_params + (size_t)k * n, num_indices);
    ++g_launches;
    // the rows now live in the per-stream histories: the staging buffer may be refilled
    if (!on_device) {
      CU_CHECK(cudaEventRecord(ctx->ev_done[k & 1], ps));
      ctx->stage_busy[k & 1] = true;
    }
    cudaEvent_t ev = nullptr;
    CU_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ev_rows.push_back(ev);
    CU_CHECK(cudaEventRecord(ev, ps));
    tr_mark(tr_rows, ps);
  }
  // ---- frame loops: one per sub-batch, on the worker streams (or on s when there is one)
  for (int k = 0; k < n_chunks; ++k) {
    for (int w = 0; w < (single ? 0 : n_workers); ++w) CU_CHECK(cudaStreamWaitEvent(ctx->worker[w], ev_rows[k], 0));
    int32_t max_steps = 0;
    for (int b = 0; b < n_sub; ++b) max_steps = std::max(max_steps, steps_of[(size_t)k * n_sub + b]);
    if (splan.fn) {  // on-chip frame loop: one launch per sub-batch carries the whole chunk
      for (int b = 0; b < n_sub; ++b) {
        if (steps_of[(size_t)k * n_sub + b] == 0) continue;
        cudaStream_t ws = single ? s : ctx->worker[b % n_workers];
        const int nb = std::min(sub, n - b * sub);
        prof.Begin(2, ws);
        splan.fn<<<nb, kStreamThreads, splan.dyn, ws>>>(d_streams + (size_t)b * sub, d_params + (size_t)k * n + (size_t)b * sub, gv, cfg,
                                                         num_indices, splan.n_buckets);
        prof.End(ws);
        ++g_launches;
      }
      max_steps = 0;
      tr_mark(tr_w0, single ? s : ctx->worker[0]);
    }
    for (int32_t f = -1; f < max_steps; ++f) {
      for (int b = 0; b < n_sub; ++b) {
        const int32_t steps = steps_of[(size_t)k * n_sub + b];
        if (steps == 0 || f >= steps) continue;
        cudaStream_t ws = single ? s : ctx->worker[b % n_workers];
        const int nb = std::min(sub, n - b * sub);
        StreamState **bs = d_streams + (size_t)b * sub;
        FrameDesc *bd = d_desc + (size_t)b * sub;
        if (f < 0) {  // GetCutoff + pre-pass of the chunk's first frame
          prof.Begin(1, ws);
          if (biglm) k_post<true><<<nb, kStreamThreads, 0, ws>>>(bs, bd, gv, cfg, kModePro, lms);
          else k_post<false><<<nb, kStreamThreads, 0, ws>>>(bs, bd, gv, cfg, kModePro, lms);
          prof.End(ws);
          ++g_launches;
          continue;
        }
        prof.Begin(0, ws);
        plans[b].fn<<<plans[b].grid, kExpandThreads, plans[b].dyn, ws>>>(bd, gv, num_indices, plans[b].flags, lms);
        prof.End(ws);
        prof.Begin(1, ws);
        const int pm = kModeEpi | (f + 1 < steps ? kModePro : 0);
        if (biglm) k_post<true><<<nb, kStreamThreads, 0, ws>>>(bs, bd, gv, cfg, pm, lms);
        else k_post<false><<<nb, kStreamThreads, 0, ws>>>(bs, bd, gv, cfg, pm, lms);
        prof.End(ws);
        g_launches += 2;
      }
    }
    CU_CHECK(cudaGetLastError());
  }
  if (!single) {  // join the workers back into the caller's stream
    for (int w = 0; w < n_workers; ++w) {
      tr_mark(tr_join, ctx->worker[w]);
      CU_CHECK(cudaEventRecord(ctx->ev_join[w], ctx->worker[w]));
      CU_CHECK(cudaStreamWaitEvent(s, ctx->ev_join[w], 0));
    }
  }
  // ---- PruneActiveTokens (inl.h:438-480, every prune_interval frames at inl.h:660-661): ONE launch
  // over every stream of the call, behind the frame loops.  (A prune takes between a fraction of a
  // millisecond and several, depending on how wide the stream's recent frames are: launched per
  // sub-batch, the slowest CTA of each launch held its worker stream while the SMs of the finished
  // ones sat idle.)  The kernel decides per stream (frames since its last prune); the host only
  // skips the launch when no stream of the batch can be due.
  if (decs[0]->opts.prune_tokens && !biglm) {
    bool due = false;
    for (int i = 0; i < n && !due; ++i)
      due = decs[i]->frames_decoded + nf[i] - decs[i]->last_prune_frame >= decs[0]->cfg.prune_interval;
    if (due) {
      PrunePlan pplan;
      if ((rc = PlanPrune(&pplan))) return rc;
      const int prune_depth = EnvInt("ASRD_PRUNE_DEPTH", decs[0]->cfg.prune_interval);
      prof.Begin(3, s);
      if (pplan.fn) {
        pplan.fn<<<n, kStreamThreads, pplan.dyn, s>>>(d_streams, gv, cfg, decs[0]->cfg.prune_interval, prune_depth,
                                                      pplan.n_buckets, pplan.ex_cap, nullptr, 0);
        ++g_launches;
      }
      // (streams k_prune served are no longer due: their CTAs return at once)
      k_lattice<false, true><<<n, kStreamThreads, 0, s>>>(d_streams, nullptr, gv, cfg, 0, lms, decs[0]->cfg.prune_interval);
      prof.End(s);
      ++g_launches;
      CU_CHECK(cudaGetLastError());
      for (int i = 0; i < n; ++i)
        if (decs[i]->frames_decoded + nf[i] - decs[i]->last_prune_frame >= decs[0]->cfg.prune_interval)
          decs[i]->last_prune_frame = decs[i]->frames_decoded + nf[i];
    }
  }
  for (int i = 0; i < n; ++i) decs[i]->frames_decoded += nf[i];
  if (trace) {
    fprintf(stderr, "[asrd] all issued at %.2f ms\n", since());
    cudaStreamSynchronize(s);
    fprintf(stderr, "[asrd] stream done at %.2f ms\n", since());
    for (size_t k = 0; k < tr_rows.size(); ++k) {
      float a = -1.f, b = -1.f, c = -1.f;
      if (k < tr_copy.size()) cudaEventElapsedTime(&a, tr_start, tr_copy[k]);
      cudaEventElapsedTime(&b, tr_start, tr_rows[k]);
      if (k < tr_w0.size()) cudaEventElapsedTime(&c, tr_start, tr_w0[k]);
      fprintf(stderr, "[asrd] chunk %zu: copied %.2f  rows in history %.2f  worker-0 chunk done %.2f ms\n", k, a, b, c);
    }
    for (size_t w = 0; w < tr_join.size(); ++w) {
      float a = -1.f;
      cudaEventElapsedTime(&a, tr_start, tr_join[w]);
      fprintf(stderr, "[asrd] worker %zu done %.2f ms\n", w, a);
    }
    for (auto *v : {&tr_copy, &tr_rows, &tr_w0, &tr_join})
      for (cudaEvent_t e : *v) cudaEventDestroy(e);
    cudaEventDestroy(tr_start);
  }
  if ((rc = prof.Finish(s))) return rc;
  return ASRD_OK;
}

int asrd_finalize_decoding(asrd_decoder *const *decs, int32_t n, void *stream) {
  (void)stream;
  int rc = CheckBatch(decs, n);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    if (!decs[i]->initialized) return ASRD_ERR_STATE;
    decs[i]->finalized = 1;
  }
  return ASRD_OK;
}

int32_t asrd_num_frames_decoded(const asrd_decoder *d) { return d ? d->frames_decoded : ASRD_ERR_BAD_ARG; }

int asrd_get_best_path(asrd_decoder *const *decs, int32_t n, int32_t use_final_probs, int32_t cap,
                       int32_t *ilabel, int32_t *olabel, float *graph, float *acoustic, int32_t *n_arcs,
                       int32_t *status, void *stream) {
  int rc = CheckBatch(decs, n);
  if (rc) return rc;
  if (cap <= 0 || !ilabel || !olabel || !graph || !acoustic || !n_arcs || !status) return ASRD_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i) {
    if (!decs[i]->initialized) return ASRD_ERR_STATE;
    if (decs[i]->finalized && !use_final_probs) return ASRD_ERR_STATE;  // inl.h:1100-1102
    // (the biglm end-token pruning depends on FinalizeDecoding: one launch serves one state)
    if (decs[i]->finalized != decs[0]->finalized) return ASRD_ERR_BAD_ARG;
  }
  if ((rc = EnsureDevice(decs[0]->graph->device))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(s);
  StreamState **d_streams;
  if ((rc = UploadStreams(decs, n, s, sc, &d_streams))) return rc;
  const size_t tot = (size_t)n * cap;
  int32_t *d_il, *d_ol, *d_n, *d_st;
  float *d_gr, *d_ac;
  CU_CHECK(sc.Alloc(&d_il, tot));
  CU_CHECK(sc.Alloc(&d_ol, tot));
  CU_CHECK(sc.Alloc(&d_gr, tot));
  CU_CHECK(sc.Alloc(&d_ac, tot));
  CU_CHECK(sc.Alloc(&d_n, (size_t)n));
  CU_CHECK(sc.Alloc(&d_st, (size_t)n));
  if (decs[0]->lm1)
    k_best_path<true><<<n, kBestPathThreads, 0, s>>>(d_streams, decs[0]->graph->view, DevCfg(decs[0]),
                                                     use_final_probs, cap, d_il, d_ol, d_gr, d_ac, d_n, d_st,
                                                     Lms(decs[0]), decs[0]->finalized);
  else
    k_best_path_rev<<<n, kBestPathThreads, 0, s>>>(d_streams, decs[0]->graph->view, DevCfg(decs[0]), use_final_probs,
                                                   cap, d_il, d_ol, d_gr, d_ac, d_n, d_st);
  ++g_launches;
  CU_CHECK(cudaGetLastError());
  CU_CHECK(cudaMemcpyAsync(n_arcs, d_n, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaMemcpyAsync(status, d_st, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaStreamSynchronize(s));
  // arcs come back end -> start; one copy per array, then flip each stream to path order
  CU_CHECK(cudaMemcpyAsync(ilabel, d_il, 4 * tot, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaMemcpyAsync(olabel, d_ol, 4 * tot, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaMemcpyAsync(graph, d_gr, 4 * tot, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaMemcpyAsync(acoustic, d_ac, 4 * tot, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaStreamSynchronize(s));
  for (int i = 0; i < n; ++i) {
    const int32_t m = std::max(0, std::min(n_arcs[i], cap));
    const size_t b = (size_t)i * cap;
    std::reverse(ilabel + b, ilabel + b + m);
    std::reverse(olabel + b, olabel + b + m);
    std::reverse(graph + b, graph + b + m);
    std::reverse(acoustic + b, acoustic + b + m);
  }
  return ASRD_OK;
}

namespace {

// Canonical order of one stream's raw lattice: tokens by (frame, state), links by (src, dst, arc
// labels) — the kernels emit them in scheduling order.  Links carry arena indices; map them to
// token indices.
int CanonicalLattice(asrd_lat_token *toks, const uint32_t *arena, uint32_t n_toks, asrd_lat_link *links, uint32_t n_links) {
  std::vector<uint32_t> order(n_toks);
  for (uint32_t i = 0; i < n_toks; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    if (toks[a].frame != toks[b].frame) return toks[a].frame < toks[b].frame;
    if (toks[a].state != toks[b].state) return toks[a].state < toks[b].state;
    if (toks[a].cost != toks[b].cost) return toks[a].cost < toks[b].cost;  // biglm: several LM states per HCLG state
    return arena[a] < arena[b];
  });
  std::vector<asrd_lat_token> sorted(n_toks);
  std::vector<std::pair<uint32_t, uint32_t>> a2i(n_toks);
  for (uint32_t i = 0; i < n_toks; ++i) {
    sorted[i] = toks[order[i]];
    a2i[i] = std::make_pair(arena[order[i]], i);
  }
  std::copy(sorted.begin(), sorted.end(), toks);
  std::sort(a2i.begin(), a2i.end());
  auto lookup = [&](uint32_t arena_idx) -> int32_t {
    auto it = std::lower_bound(a2i.begin(), a2i.end(), std::make_pair(arena_idx, 0u));
    return (it != a2i.end() && it->first == arena_idx) ? (int32_t)it->second : -1;
  };
  for (uint32_t i = 0; i < n_links; ++i) {
    links[i].src = lookup((uint32_t)links[i].src);
    links[i].dst = lookup((uint32_t)links[i].dst);
    if (links[i].src < 0 || links[i].dst < 0) return ASRD_ERR_STATE;
  }
  std::sort(links, links + n_links, [](const asrd_lat_link &a, const asrd_lat_link &b) {
    if (a.src != b.src) return a.src < b.src;
    if (a.dst != b.dst) return a.dst < b.dst;
    if (a.ilabel != b.ilabel) return a.ilabel < b.ilabel;
    if (a.olabel != b.olabel) return a.olabel < b.olabel;
    return a.graph < b.graph;
  });
  return ASRD_OK;
}

}  // namespace

int asrd_get_raw_lattice_batch(asrd_decoder *const *decs, int32_t n, int32_t use_final_probs, asrd_lat_token *toks,
                               int64_t tok_cap, asrd_lat_link *links, int64_t link_cap, int64_t *n_toks,
                               int64_t *n_links, int32_t *status, void *stream) {
  if (!n_toks || !n_links || !status || tok_cap < 0 || link_cap < 0 || tok_cap > 0x7FFFFFFF || link_cap > 0x7FFFFFFF)
    return ASRD_ERR_BAD_ARG;
  int rc = CheckBatch(decs, n);
  if (rc) return rc;
  std::vector<asrd_decoder *> act;   // streams with something decoded
  std::vector<int> act_idx;
  for (int i = 0; i < n; ++i) {
    if (!decs[i]->initialized) return ASRD_ERR_STATE;
    if (decs[i]->finalized && !use_final_probs) return ASRD_ERR_STATE;  // inl.h:879-884
    n_toks[i] = n_links[i] = 0;
    status[i] = ASRD_ERR_NO_TOKENS;
    if (decs[i]->frames_decoded > 0 && decs[i]->d_ll_hist) {
      act.push_back(decs[i]);
      act_idx.push_back(i);
    }
  }
  const int m = (int)act.size();
  if (!m) return ASRD_OK;
  asrd_decoder *d0 = act[0];
  if ((rc = EnsureDevice(d0->graph->device))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(s);
  StreamState **d_streams;
  if ((rc = UploadStreams(act.data(), m, s, sc, &d_streams))) return rc;
  const size_t H = (size_t)d0->opts.hash_capacity;
  const size_t tc = (size_t)std::max<int64_t>(tok_cap, 1), lc = (size_t)std::max<int64_t>(link_cap, 1);
  asrd_lat_token *d_toks;
  uint32_t *d_arena;
  asrd_lat_link *d_links;
  LatticeOut *d_out;
  CU_CHECK(sc.Alloc(&d_out, (size_t)m));
  CU_CHECK(sc.Alloc(&d_toks, tc * m));
  CU_CHECK(sc.Alloc(&d_arena, tc * m));
  CU_CHECK(sc.Alloc(&d_links, lc * m));
  std::vector<LatticeOut> h(m), r(m);
  for (int j = 0; j < m; ++j) {
    memset(&h[j], 0, sizeof(LatticeOut));
    h[j].toks = d_toks + tc * j;
    h[j].tok_arena_idx = d_arena + tc * j;
    h[j].links = d_links + lc * j;
    h[j].tok_cap = (uint32_t)tok_cap;
    h[j].link_cap = (uint32_t)link_cap;
  }
  CU_CHECK(cudaMemcpyAsync(d_out, h.data(), sizeof(LatticeOut) * m, cudaMemcpyHostToDevice, s));
  // Plain decoders: the pull sweep with its lookup map in shared memory (k_prune<true>: work in
  // proportion to the tokens that survive, the arena untouched), one CTA per stream; a stream with
  // a frame beyond that kernel's capacity, and every biglm decoder, goes through the HBM-map sweep
  // (k_lattice).
  std::vector<int> redo;
  bool swept = false;
  if (!d0->lm1 && EnvInt("ASRD_LATTICE_KERNEL", 1)) {
    PrunePlan pplan;
    if ((rc = PlanPrune(&pplan))) return rc;
    if (pplan.emit_fn) {
      pplan.emit_fn<<<m, kStreamThreads, pplan.dyn, s>>>(d_streams, d0->graph->view, DevCfg(d0), 0, -1, pplan.n_buckets,
                                                          pplan.ex_cap, d_out, use_final_probs ? 1 : 0);
      ++g_launches;
      CU_CHECK(cudaGetLastError());
      CU_CHECK(cudaMemcpyAsync(r.data(), d_out, sizeof(LatticeOut) * m, cudaMemcpyDeviceToHost, s));
      CU_CHECK(cudaStreamSynchronize(s));
      for (int j = 0; j < m; ++j)
        if (r[j].n_toks == 0xFFFFFFFFu) redo.push_back(j);
      swept = true;
    }
  }
  if (!swept)
    for (int j = 0; j < m; ++j) redo.push_back(j);
  if (!redo.empty()) {
    const int k = (int)redo.size();
    LatEntry *maps;  // the two per-frame lookup maps of the HBM-map sweep, per stream
    uint32_t *pairs = nullptr;
    CU_CHECK(sc.Alloc(&maps, 2 * H * k));
    if (d0->lm1) CU_CHECK(sc.Alloc(&pairs, 2 * H * k));
    CU_CHECK(cudaMemsetAsync(maps, 0xFF, 2 * H * k * sizeof(LatEntry), s));
    std::vector<LatticeOut> hs(k), rs(k);
    std::vector<asrd_decoder *> sub(k);
    for (int q = 0; q < k; ++q) {
      hs[q] = h[redo[q]];
      hs[q].map[0] = maps + 2 * H * q;
      hs[q].map[1] = maps + 2 * H * q + H;
      if (pairs) {
        hs[q].map_pair[0] = pairs + 2 * H * q;
        hs[q].map_pair[1] = pairs + 2 * H * q + H;
      }
      sub[q] = act[redo[q]];
    }
    StreamState **d_sub;
    LatticeOut *d_out_sub;
    if ((rc = UploadStreams(sub.data(), k, s, sc, &d_sub))) return rc;
    CU_CHECK(sc.Alloc(&d_out_sub, (size_t)k));
    CU_CHECK(cudaMemcpyAsync(d_out_sub, hs.data(), sizeof(LatticeOut) * k, cudaMemcpyHostToDevice, s));
    if (d0->lm1)
      k_lattice<true, false><<<k, kStreamThreads, 0, s>>>(d_sub, d_out_sub, d0->graph->view, DevCfg(d0), use_final_probs ? 1 : 0, Lms(d0), 0);
    else
      k_lattice<false, false><<<k, kStreamThreads, 0, s>>>(d_sub, d_out_sub, d0->graph->view, DevCfg(d0), use_final_probs ? 1 : 0, Lms(d0), 0);
    ++g_launches;
    CU_CHECK(cudaGetLastError());
    CU_CHECK(cudaMemcpyAsync(rs.data(), d_out_sub, sizeof(LatticeOut) * k, cudaMemcpyDeviceToHost, s));
    CU_CHECK(cudaStreamSynchronize(s));
    for (int q = 0; q < k; ++q) r[redo[q]] = rs[q];
  }
  std::vector<std::vector<uint32_t>> arena(m);
  for (int j = 0; j < m; ++j) {
    const int i = act_idx[j];
    n_toks[i] = r[j].n_toks;
    n_links[i] = r[j].n_links;
    if (r[j].n_toks > h[j].tok_cap || r[j].n_links > h[j].link_cap) { status[i] = ASRD_ERR_PATH_OVERFLOW; continue; }
    if (r[j].n_toks == 0) { status[i] = ASRD_ERR_NO_TOKENS; continue; }
    if (!toks || !links) { status[i] = ASRD_ERR_BAD_ARG; continue; }
    status[i] = ASRD_OK;
    arena[j].resize(r[j].n_toks);
    CU_CHECK(cudaMemcpyAsync(toks + (size_t)tok_cap * i, h[j].toks, sizeof(asrd_lat_token) * r[j].n_toks, cudaMemcpyDeviceToHost, s));
    CU_CHECK(cudaMemcpyAsync(arena[j].data(), h[j].tok_arena_idx, 4 * (size_t)r[j].n_toks, cudaMemcpyDeviceToHost, s));
    if (r[j].n_links)
      CU_CHECK(cudaMemcpyAsync(links + (size_t)link_cap * i, h[j].links, sizeof(asrd_lat_link) * r[j].n_links, cudaMemcpyDeviceToHost, s));
  }
  CU_CHECK(cudaStreamSynchronize(s));
  // host side: one stream's sort is independent of the next one's
  auto canon = [&](int j) {
    const int i = act_idx[j];
    if (status[i] == ASRD_OK)
      status[i] = CanonicalLattice(toks + (size_t)tok_cap * i, arena[j].data(), r[j].n_toks, links + (size_t)link_cap * i, r[j].n_links);
  };
  const int n_thr = std::min<int>({m, (int)std::max(1u, std::thread::hardware_concurrency()), 16});
  if (n_thr <= 1) {
    for (int j = 0; j < m; ++j) canon(j);
  } else {
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thr; ++t)
      pool.emplace_back([&] {
        for (int j; (j = next.fetch_add(1)) < m;) canon(j);
      });
    for (auto &t : pool) t.join();
  }
  return ASRD_OK;
}

int asrd_get_raw_lattice(asrd_decoder *d, int32_t use_final_probs, asrd_lat_token *toks, int64_t tok_cap,
                         asrd_lat_link *links, int64_t link_cap, int64_t *n_toks, int64_t *n_links,
                         void *stream) {
  if (!d) return ASRD_ERR_BAD_ARG;
  asrd_decoder *one[1] = {d};
  int32_t st = ASRD_OK;
  const int rc = asrd_get_raw_lattice_batch(one, 1, use_final_probs, toks, tok_cap, links, link_cap, n_toks, n_links, &st, stream);
  return rc ? rc : st;
}

int asrd_path_to_vector(const int32_t *ilabel, const int32_t *olabel, const float *graph,
                        const float *acoustic, int32_t n_arcs, int32_t *words, int32_t *n_words,
                        int32_t *alignment, int32_t *n_alignment, float *tot_score, float *lm_score) {
  // LatticeToVector, src/newfst/lattice-functions.cc:179-217
  if (n_arcs < 0 || !n_words || !n_alignment || !tot_score || !lm_score) return ASRD_ERR_BAD_ARG;
  float tot = 0.f, lm = 0.f;
  int32_t nw = 0, na = 0;
  for (int32_t i = 0; i < n_arcs; ++i) {
    if (ilabel[i] != 0 && alignment) alignment[na] = ilabel[i];
    if (ilabel[i] != 0) ++na;
    if (olabel[i] != 0 && words) words[nw] = olabel[i];
    if (olabel[i] != 0) ++nw;
    lm += graph[i];
    tot += graph[i] + acoustic[i];
  }
  *n_words = nw;
  *n_alignment = na;
  *tot_score = tot;
  *lm_score = lm;
  return ASRD_OK;
}

int32_t asrd_frame_stats(asrd_decoder *d, asrd_frame_stat *out, int32_t cap, void *stream) {
  if (!d) return ASRD_ERR_BAD_ARG;
  if (!d->opts.collect_stats) return 0;
  const int32_t n = d->initialized ? d->frames_decoded + 1 : 0;
  if (out && cap > 0 && n > 0) {
    if (EnsureDevice(d->graph->device)) return ASRD_ERR_CUDA;
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemcpyAsync(out, d->h_state.stats, sizeof(asrd_frame_stat) * (size_t)std::min(n, cap),
                        cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
      return ASRD_ERR_CUDA;
  }
  return n;
}

int32_t asrd_arena_frame_tokens(asrd_decoder *d, uint32_t *out, int32_t cap, void *stream) {
  if (!d) return ASRD_ERR_BAD_ARG;
  const int32_t n = d->initialized ? d->frames_decoded + 1 : 0;
  if (out && cap > 0 && n > 0) {
    if (EnsureDevice(d->graph->device)) return ASRD_ERR_CUDA;
    cudaStream_t s = (cudaStream_t)stream;
    const int32_t m = std::min(n, cap);
    std::vector<uint32_t> off((size_t)m + 1);
    if (cudaMemcpyAsync(off.data(), d->h_state.frame_off, 4 * ((size_t)m + 1), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
      return ASRD_ERR_CUDA;
    for (int32_t f = 0; f < m; ++f) out[f] = off[f + 1] - off[f];
  }
  return n;
}

int asrd_decoder_status(asrd_decoder *d, void *stream) {
  if (!d) return ASRD_ERR_BAD_ARG;
  int rc = EnsureDevice(d->graph->device);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  int32_t st = 0;
  CU_CHECK(cudaMemcpyAsync(&st, &d->d_state->status, 4, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaStreamSynchronize(s));
  return st;
}

int asrd_profile_enable(int on) {
  g_profile.store(on ? 1 : 0);
  return ASRD_OK;
}

int asrd_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < 4; ++i) {
    g_prof_ms[i] = 0;
    g_prof_n[i] = 0;
  }
  return ASRD_OK;
}

int asrd_profile_get(double *kernel_ms, int64_t *kernel_launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < 4; ++i) {
    if (kernel_ms) kernel_ms[i] = g_prof_ms[i];
    if (kernel_launches) kernel_launches[i] = g_prof_n[i];
  }
  return ASRD_OK;
}

int asrd_get_counters(asrd_decoder *const *decs, int32_t n, int64_t *arcs_expanded, int64_t *arcs_admitted,
                      int64_t *tokens, void *stream) {
  int rc = CheckBatch(decs, n);
  if (rc) return rc;
  if ((rc = EnsureDevice(decs[0]->graph->device))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  Scratch sc(s);
  StreamState **d_streams;
  if ((rc = UploadStreams(decs, n, s, sc, &d_streams))) return rc;
  unsigned long long *d_out;
  CU_CHECK(sc.Alloc(&d_out, 20));
  CU_CHECK(cudaMemsetAsync(d_out, 0, 160, s));
  k_counters<<<(n + 127) / 128, 128, 0, s>>>(d_streams, n, d_out);
  ++g_launches;
  unsigned long long h[20];
  CU_CHECK(cudaMemcpyAsync(h, d_out, 160, cudaMemcpyDeviceToHost, s));
  CU_CHECK(cudaStreamSynchronize(s));
  if (arcs_expanded) *arcs_expanded = (int64_t)h[0];
  if (arcs_admitted) *arcs_admitted = (int64_t)h[1];
  if (tokens) *tokens = (int64_t)h[2];
  g_last_fallback_frames.store((int64_t)h[3]);
  g_last_pruned_tokens.store((int64_t)h[10]);
  g_last_peak_tokens.store((int64_t)h[11]);
  for (int k = 0; k < 8; ++k) g_last_prune_cycles[k].store((int64_t)h[12 + k]);
  for (int k = 0; k < 6; ++k) g_last_phase_cycles[k].store((int64_t)h[4 + k]);
  return ASRD_OK;
}

int64_t asrd_last_fallback_frames(void) { return g_last_fallback_frames.load(); }
int64_t asrd_last_pruned_tokens(void) { return g_last_pruned_tokens.load(); }
int64_t asrd_last_peak_tokens(void) { return g_last_peak_tokens.load(); }
void asrd_last_prune_cycles(int64_t *out8) {
  for (int k = 0; k < 8; ++k) out8[k] = g_last_prune_cycles[k].load();
}

void asrd_last_phase_cycles(int64_t *out6) {
  for (int k = 0; k < 6; ++k) out6[k] = g_last_phase_cycles[k].load();
}

int asrd_synchronize(void *stream) {
  CU_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return ASRD_OK;
}

int asrd_host_alloc(void **ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return ASRD_ERR_BAD_ARG;
  CU_CHECK(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
  return ASRD_OK;
}

int asrd_host_free(void *ptr) {
  if (ptr) CU_CHECK(cudaFreeHost(ptr));
  return ASRD_OK;
}

}  // extern "C"
