// TEST INFRASTRUCTURE — not product code.
//
// Driver for the UNMODIFIED reference decoder, compiled in place from
// /root/reference (see oracle/Makefile; outputs only under oracle/_ref/).
// It instantiates the reference's own OnlineLatticeDecoderMempool
// (src/my-decoder/online-decoder-mempool-base.h:77) exactly as the reference's
// offline bin does (src/kaldi-nnet3bin/kaldi-hclg-my-decoder.cc:95-122):
//   InitDecoding -> AdvanceDecoding -> FinalizeDecoding -> GetBestPath -> LatticeToVector
// and prints one JSON object per utterance.  Per-frame statistics are obtained
// by sub-classing (protected virtuals GetCutoff / ProcessNonemitting); no
// reference source is modified or copied.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may execute the binary built from this file.
//
// File formats (ours, shared with asr_decoder_b200/fstio.py):
//   graph    : newfst flat format (src/newfst/optimize-fst.h:226-280)
//   loglikes : int32 magic 0x4c4c5341, int32 n_utt, then per utterance
//              int32 T, int32 P, float32[T*P] (row-major, column = ilabel-1)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <chrono>
#include <atomic>

#include "src/my-decoder/online-decoder-mempool-base.h"
#ifdef ASRD_REF_CLG
// -DASRD_REF_CLG (oracle/_ref/ref_decode_clg): the same driver around the reference's CLG decoder —
// ClgFst (src/my-decoder/clg-fst.h:9-189: CLG graph + HMM set, expanded on the fly) under
// OnlineClgLatticeDecoderMempool (src/my-decoder/online-clg-decoder-mempool-base.h:10-281), as
// OnlineClgLatticeFastDecoder selects it (src/kaldi-nnet3/kaldi-online-nnet3-my-decoder.h:250-283).
#include "src/my-decoder/clg-fst.h"
#include "src/my-decoder/online-clg-decoder-mempool-base.h"
#endif
#include "src/newfst/lattice-functions.h"
#include "src/newfst/lattice-determinize-api.h"

using namespace datemoon;

namespace {

struct Utt {
  int T = 0, P = 0;
  std::vector<float> ll;
};

// 10-line matrix decodable: LogLikelihood(f, i) = M[f][i-1].
class MatrixDecodable : public AmInterface {
 public:
  MatrixDecodable(const Utt *u, int frames_ready)
      : u_(u), ready_(frames_ready), calls(0) {}
  virtual BaseFloat LogLikelihood(int32 frame, int32 index) {
    ++calls;
    return u_->ll[(size_t)frame * u_->P + (index - 1)];
  }
  virtual bool IsLastFrame(int32 frame) const { return frame == u_->T - 1; }
  virtual int32 NumFramesReady() const { return ready_; }
  virtual int32 NumIndices() const { return u_->P; }
  void SetReady(int r) { ready_ = r; }
  long long calls;

 private:
  const Utt *u_;
  int ready_;
};

struct FrameStat {
  unsigned n_in = 0;       // tokens seen by GetCutoff for this frame
  float cur_cutoff = 0;    // GetCutoff result
  float abeam = 0;         // adaptive beam
  float next_cutoff = 0;   // cutoff handed to ProcessNonemitting
  unsigned n_raw = 0;      // tokens alive after the eps closure
  unsigned n_within = 0;   // ... of which cost < next_cutoff
  float best = 0;          // best cost after closure
  long long ll_calls = 0;  // LogLikelihood calls during this frame
};

#ifdef ASRD_REF_CLG
typedef OnlineClgLatticeDecoderMempool RefDecoder;
typedef ClgFst RefGraph;
#else
typedef OnlineLatticeDecoderMempool RefDecoder;
typedef Fst RefGraph;
#endif

class Probe : public RefDecoder {
 public:
  typedef RefDecoder Base;
  Probe(RefGraph *fst, const LatticeFasterDecoderConfig &c) : Base(fst, c), collect(false) {}
  bool collect;
  int dump_frame = -1;           // debug: print the hash-list key order of this frame to stderr
  std::vector<FrameStat> stats;  // index 0 = after InitDecoding
  FrameStat pending;

  virtual BaseFloat GetCutoff(Elem *list_head, size_t *tok_count,
                              BaseFloat *adaptive_beam, Elem **best_elem) {
    size_t cnt = 0;
    BaseFloat ab = 0;
    BaseFloat r = Base::GetCutoff(list_head, &cnt, &ab, best_elem);
    if (tok_count) *tok_count = cnt;
    if (adaptive_beam) *adaptive_beam = ab;
    if (collect) {
      pending = FrameStat();
      pending.n_in = (unsigned)cnt;
      pending.cur_cutoff = r;
      pending.abeam = ab;
    }
    return r;
  }
  virtual void ProcessNonemitting(BaseFloat cutoff) {
    Base::ProcessNonemitting(cutoff);
    if (!collect) return;
    FrameStat s = pending;
    pending = FrameStat();
    s.next_cutoff = cutoff;
    float best = std::numeric_limits<float>::infinity();
    for (const Elem *e = _toks.GetList(); e != NULL; e = e->tail) {
      ++s.n_raw;
      if (e->val->_tot_cost < cutoff) ++s.n_within;
      if (e->val->_tot_cost < best) best = e->val->_tot_cost;
    }
    s.best = best;
    if ((int)stats.size() == dump_frame) {
      fprintf(stderr, "ORDER");
      for (const Elem *e = _toks.GetList(); e != NULL; e = e->tail) fprintf(stderr, " %d", e->key);
      fprintf(stderr, "\n");
    }
    stats.push_back(s);
  }
  int NumToks() const { return _num_toks; }
  int NumLinks() const { return _num_links; }
};

unsigned Bits(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}

struct Result {
  bool ok = false;
  std::vector<int> words, ali;
  float tot = 0, lm = 0;
  std::vector<FrameStat> stats;
  int toks_final = 0, links_final = 0;
  int raw_states = -1, raw_arcs = -1, det_states = -1, det_arcs = -1;
  double seconds = 0;
  long long ll_calls = 0;
};

struct Options {
  std::string graph, hmm, loglikes, out;
  LatticeFasterDecoderConfig cfg;
  int threads = 1;
  bool stats = false;
  bool lattice = false;
  int chunk = 0;  // >0: feed AdvanceDecoding in chunks of this many frames
  int repeat = 1;
  int dump_frame = -1;
};

void DecodeOne(Probe *dec, const Utt &u, const Options &o, Result *r) {
  auto t0 = std::chrono::steady_clock::now();
  dec->collect = o.stats;
  dec->dump_frame = o.dump_frame;
  dec->stats.clear();
  dec->InitDecoding();
  MatrixDecodable decodable(&u, u.T);
  if (o.chunk > 0) {
    for (int ready = 0; ready < u.T;) {
      ready = std::min(u.T, ready + o.chunk);
      decodable.SetReady(ready);
      dec->AdvanceDecoding(&decodable);
    }
  } else if (o.stats) {
    // one frame at a time so LogLikelihood calls can be attributed per frame
    long long prev = 0;
    for (int f = 0; f < u.T; ++f) {
      dec->AdvanceDecoding(&decodable, 1);
      dec->stats.back().ll_calls = decodable.calls - prev;
      prev = decodable.calls;
    }
  } else {
    dec->AdvanceDecoding(&decodable);
  }
  dec->FinalizeDecoding();
  r->toks_final = dec->NumToks();
  r->links_final = dec->NumLinks();
  Lattice best_path;
  r->ok = dec->GetBestPath(&best_path);
  r->words.clear();
  r->ali.clear();
  r->tot = r->lm = 0;
  if (r->ok) r->ok = LatticeToVector(best_path, r->words, r->ali, r->tot, r->lm);
  auto t1 = std::chrono::steady_clock::now();
  r->seconds = std::chrono::duration<double>(t1 - t0).count();
  r->ll_calls = decodable.calls;
  r->stats = dec->stats;
  if (o.lattice) {
    Lattice raw, det;
    if (dec->GetRawLattice(&raw, true)) {
      r->raw_states = raw.NumStates();
      int na = 0;
      for (int s = 0; s < raw.NumStates(); ++s) na += (int)raw.GetState(s)->GetArcSize();
      r->raw_arcs = na;
      bool debug_ptr = false;
      DeterminizeLatticeOptions opts;
      if (LatticeCheckFormat(&raw)) {
        DeterminizeLatticeWrapper(&raw, &det, opts, &debug_ptr);
        r->det_states = det.NumStates();
        na = 0;
        for (int s = 0; s < det.NumStates(); ++s) na += (int)det.GetState(s)->GetArcSize();
        r->det_arcs = na;
      }
    }
  }
}

void PrintResult(FILE *fp, int utt, const Utt &u, const Result &r, bool stats) {
  fprintf(fp, "{\"utt\": %d, \"ok\": %s, \"frames\": %d, \"tot\": %.9g, \"tot_bits\": %u, \"lm\": %.9g, \"lm_bits\": %u, ",
          utt, r.ok ? "true" : "false", u.T, r.tot, Bits(r.tot), r.lm, Bits(r.lm));
  fprintf(fp, "\"seconds\": %.6f, \"ll_calls\": %lld, \"toks_final\": %d, \"links_final\": %d, ",
          r.seconds, r.ll_calls, r.toks_final, r.links_final);
  fprintf(fp, "\"raw_states\": %d, \"raw_arcs\": %d, \"det_states\": %d, \"det_arcs\": %d, ",
          r.raw_states, r.raw_arcs, r.det_states, r.det_arcs);
  fprintf(fp, "\"words\": [");
  for (size_t i = 0; i < r.words.size(); ++i) fprintf(fp, "%s%d", i ? "," : "", r.words[i]);
  fprintf(fp, "], \"ali\": [");
  for (size_t i = 0; i < r.ali.size(); ++i) fprintf(fp, "%s%d", i ? "," : "", r.ali[i]);
  fprintf(fp, "]");
  if (stats) {
#define ARR_U(name, expr)                                             \
  fprintf(fp, ", \"" name "\": [");                                   \
  for (size_t i = 0; i < r.stats.size(); ++i) {                       \
    const FrameStat &s = r.stats[i];                                  \
    fprintf(fp, "%s%llu", i ? "," : "", (unsigned long long)(expr)); \
  }                                                                   \
  fprintf(fp, "]");
    ARR_U("n_in", s.n_in)
    ARR_U("n_raw", s.n_raw)
    ARR_U("n_within", s.n_within)
    ARR_U("cur_cutoff_bits", Bits(s.cur_cutoff))
    ARR_U("abeam_bits", Bits(s.abeam))
    ARR_U("next_cutoff_bits", Bits(s.next_cutoff))
    ARR_U("best_bits", Bits(s.best))
    ARR_U("ll_calls_f", s.ll_calls)
#undef ARR_U
  }
  fprintf(fp, "}\n");
}

bool ReadLoglikes(const std::string &file, std::vector<Utt> *utts) {
  FILE *fp = fopen(file.c_str(), "rb");
  if (!fp) return false;
  int magic = 0, n = 0;
  if (fread(&magic, 4, 1, fp) != 1 || magic != 0x4c4c5341) { fclose(fp); return false; }
  if (fread(&n, 4, 1, fp) != 1) { fclose(fp); return false; }
  utts->resize(n);
  for (int i = 0; i < n; ++i) {
    Utt &u = (*utts)[i];
    if (fread(&u.T, 4, 1, fp) != 1 || fread(&u.P, 4, 1, fp) != 1) { fclose(fp); return false; }
    u.ll.resize((size_t)u.T * u.P);
    if (fread(u.ll.data(), 4, u.ll.size(), fp) != u.ll.size()) { fclose(fp); return false; }
  }
  fclose(fp);
  return true;
}

}  // namespace

int main(int argc, char **argv) {
  Options o;
  o.cfg._beam = 13.0;
  o.cfg._max_active = 7000;
  o.cfg._min_active = 200;
  o.cfg._lattice_beam = 8.0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&](const char *name) -> const char * {
      size_t n = strlen(name);
      if (a.compare(0, n, name) == 0 && a.size() > n && a[n] == '=') return a.c_str() + n + 1;
      return NULL;
    };
    const char *v;
    if ((v = val("--graph"))) o.graph = v;
    else if ((v = val("--hmm"))) o.hmm = v;
    else if ((v = val("--loglikes"))) o.loglikes = v;
    else if ((v = val("--out"))) o.out = v;
    else if ((v = val("--beam"))) o.cfg._beam = atof(v);
    else if ((v = val("--max-active"))) o.cfg._max_active = atoi(v);
    else if ((v = val("--min-active"))) o.cfg._min_active = atoi(v);
    else if ((v = val("--lattice-beam"))) o.cfg._lattice_beam = atof(v);
    else if ((v = val("--prune-interval"))) o.cfg._prune_interval = atoi(v);
    else if ((v = val("--beam-delta"))) o.cfg._beam_delta = atof(v);
    else if ((v = val("--hash-ratio"))) o.cfg._hash_ratio = atof(v);
    else if ((v = val("--threads"))) o.threads = atoi(v);
    else if ((v = val("--chunk"))) o.chunk = atoi(v);
    else if ((v = val("--repeat"))) o.repeat = atoi(v);
    else if ((v = val("--dump-frame"))) o.dump_frame = atoi(v);
    else if (a == "--stats") o.stats = true;
    else if (a == "--lattice") o.lattice = true;
    else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
  }
  if (o.graph.empty() || o.loglikes.empty()) {
    fprintf(stderr, "usage: ref_decode --graph=G --loglikes=L [--out=F] [--beam= --max-active= --min-active= "
                    "--lattice-beam= --prune-interval= --beam-delta= --hash-ratio=] [--threads=N] [--chunk=N] "
                    "[--repeat=N] [--stats] [--lattice]\n");
    return 2;
  }
#ifdef ASRD_REF_CLG
  ClgFst fst;
  if (!fst.Init(o.graph, o.hmm)) return 3;
#else
  Fst fst;
  if (!fst.ReadFst(o.graph.c_str())) return 3;
#endif
  std::vector<Utt> utts;
  if (!ReadLoglikes(o.loglikes, &utts)) { fprintf(stderr, "cannot read %s\n", o.loglikes.c_str()); return 3; }
  int n = (int)utts.size();
  int total = n * o.repeat;
  std::vector<Result> results(n);
  int nthread = std::max(1, std::min(o.threads, total));

  // The reference's deployment model: one decoder object per worker thread, one
  // shared read-only Fst (src/v2-asrbin/v2-asr-service.cc:95-104).
  std::vector<Probe *> decs(nthread);
  for (int t = 0; t < nthread; ++t) decs[t] = new Probe(&fst, o.cfg);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthread; ++t) {
    th.emplace_back([&, t]() {
      for (int j = t; j < total; j += nthread) {  // round-robin
        int i = j % n;
        Result r;
        DecodeOne(decs[t], utts[i], o, &r);
        if (j < n) results[i] = r;
      }
    });
  }
  for (auto &x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  double wall = std::chrono::duration<double>(t1 - t0).count();
  for (int t = 0; t < nthread; ++t) delete decs[t];

  FILE *fp = o.out.empty() ? stdout : fopen(o.out.c_str(), "w");
  if (!fp) return 4;
  long long frames = 0, calls = 0;
  for (int i = 0; i < n; ++i) {
    PrintResult(fp, i, utts[i], results[i], o.stats);
    frames += utts[i].T;
    calls += results[i].ll_calls;
  }
  fprintf(fp, "{\"summary\": true, \"utts\": %d, \"repeat\": %d, \"threads\": %d, \"wall_s\": %.6f, \"frames\": %lld, \"ll_calls\": %lld}\n",
          n, o.repeat, nthread, wall, frames * o.repeat, calls * o.repeat);
  if (fp != stdout) fclose(fp);
  return 0;
}
