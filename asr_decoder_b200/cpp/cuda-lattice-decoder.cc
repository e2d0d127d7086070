#include "cuda-lattice-decoder.h"

#include <algorithm>
#include <cstdio>

namespace asrd_host {

bool CudaLm::Read(const char *file, float scale, int device) {
  FILE *fp = fopen(file, "rb");
  if (!fp) return false;
  int32_t hdr[3];
  size_t n_orders = 0;
  bool ok = fread(hdr, 4, 3, fp) == 3 && fread(&n_orders, sizeof(size_t), 1, fp) == 1 && n_orders < 64;
  std::vector<int32_t> counts(n_orders);
  ok = ok && fread(counts.data(), 4, n_orders, fp) == n_orders;
  int32_t n_states = 0, n_arcs = 0;
  ok = ok && fread(&n_states, 4, 1, fp) == 1 && n_states > 0;
  struct Info { int32_t arc_num; float backoff_prob; int32_t backoff_id; };
  std::vector<Info> info(ok ? n_states : 0);
  ok = ok && fread(info.data(), sizeof(Info), n_states, fp) == (size_t)n_states;
  ok = ok && fread(&n_arcs, 4, 1, fp) == 1 && n_arcs > 0;
  std::vector<asrd_lm_arc> arcs(ok ? n_arcs : 0);
  ok = ok && fread(arcs.data(), sizeof(asrd_lm_arc), n_arcs, fp) == (size_t)n_arcs;
  fclose(fp);
  if (!ok) return false;
  std::vector<int32_t> an(n_states), bi(n_states);
  std::vector<float> bp(n_states);
  for (int32_t i = 0; i < n_states; ++i) {
    an[i] = info[i].arc_num;
    bp[i] = info[i].backoff_prob * scale;  // Fsa::Rescale, arpa2fsa.cc:264-276
    bi[i] = info[i].backoff_id;
  }
  for (int32_t a = 0; a < n_arcs; ++a) arcs[a].weight *= scale;
  return FromArrays(hdr[0], hdr[1], n_states, an.data(), bp.data(), bi.data(), arcs.data(), n_arcs, device);
}

CudaLatticeDecoder::CudaLatticeDecoder(CudaFst *graph, const LatticeFasterDecoderConfig &config, int max_frames,
                                       void *cuda_stream, bool prune_tokens)
    : d_(NULL), stream_(cuda_stream), finalized_(false) {
  Create(graph, config, NULL, NULL, max_frames, prune_tokens);
}

CudaLatticeDecoder::CudaLatticeDecoder(CudaFst *graph, const LatticeFasterDecoderConfig &config, CudaLm *oldlm,
                                       CudaLm *newlm, int max_frames, void *cuda_stream)
    : d_(NULL), stream_(cuda_stream), finalized_(false) {
  Create(graph, config, oldlm, newlm, max_frames, false);
}

void CudaLatticeDecoder::Create(CudaFst *graph, const LatticeFasterDecoderConfig &config, CudaLm *oldlm,
                                CudaLm *newlm, int max_frames, bool prune_tokens) {
  config.Check();
  asrd_config c;
  c.beam = config._beam;
  c.max_active = config._max_active;
  c.min_active = config._min_active;
  c.lattice_beam = config._lattice_beam;
  c.prune_interval = config._prune_interval;
  c.beam_delta = config._beam_delta;
  c.hash_ratio = config._hash_ratio;
  c.prune_scale = config._prune_scale;
  asrd_device_options o = asrd_device_options();
  o.max_frames = max_frames;
  o.collect_stats = 1;
  o.prune_tokens = prune_tokens ? 1 : 0;  // PruneActiveTokens every config._prune_interval frames (inl.h:660-661)
  if (oldlm || newlm)
    Check(asrd_decoder_create_biglm(graph ? graph->handle() : NULL, &c, &o, oldlm ? oldlm->handle() : NULL,
                                    newlm ? newlm->handle() : NULL, &d_), "asrd_decoder_create_biglm");
  else
    Check(asrd_decoder_create(graph ? graph->handle() : NULL, &c, &o, &d_), "asrd_decoder_create");
}

CudaLatticeDecoder::~CudaLatticeDecoder() { asrd_decoder_destroy(d_); }

void CudaLatticeDecoder::Check(int status, const char *what) const {
  if (status != ASRD_OK)  // the reference's LOG_ERR throws std::runtime_error (util/log-message.cc:122-144)
    throw std::runtime_error(std::string(what) + ": " + asrd_strerror(status));
}

void CudaLatticeDecoder::InitDecoding() {
  asrd_decoder *h[1] = {d_};
  Check(asrd_init_decoding(h, 1, stream_), "asrd_init_decoding");
  finalized_ = false;
}

int32 CudaLatticeDecoder::NumFramesDecoded() const { return asrd_num_frames_decoded(d_); }

void CudaLatticeDecoder::Upload(AmInterface *decodable, int32 first, int32 count) {
  asrd_decoder *h[1] = {d_};
  const int32 P = decodable->NumIndices();
  const float *rows;
  int32 stride;
  MatrixDecodableInterface *m = dynamic_cast<MatrixDecodableInterface *>(decodable);
  if (m) {
    rows = m->Data() + (size_t)first * m->Stride();
    stride = m->Stride();
  } else {
    stage_.resize((size_t)count * P);  // pull model: T x P virtual calls (itf/decodable-itf.h:65-104)
    for (int32 f = 0; f < count; ++f)
      for (int32 i = 0; i < P; ++i) stage_[(size_t)f * P + i] = decodable->LogLikelihood(first + f, i + 1);
    rows = stage_.data();
    stride = P;
  }
  const float *ptrs[1] = {rows};
  const int32_t nf[1] = {count}, st[1] = {stride};
  Check(asrd_advance_decoding(h, 1, ptrs, nf, st, P, -1, 0, stream_), "asrd_advance_decoding");
  Check(asrd_synchronize(stream_), "asrd_synchronize");  // rows may be reused by the caller
}

void CudaLatticeDecoder::AdvanceDecoding(AmInterface *decodable, int32 max_num_frames) {
  // inl.h:630-668
  if (finalized_) Check(ASRD_ERR_STATE, "AdvanceDecoding after FinalizeDecoding");
  const int32 decoded = NumFramesDecoded();
  int32 target = decodable->NumFramesReady();
  if (target < decoded) Check(ASRD_ERR_BAD_ARG, "NumFramesReady() decreased");
  if (max_num_frames >= 0) target = std::min(target, decoded + max_num_frames);
  if (target > decoded) Upload(decodable, decoded, target - decoded);
}

BaseFloat CudaLatticeDecoder::ProcessEmitting(AmInterface *decodable) {
  const int32 f = NumFramesDecoded();
  Upload(decodable, f, 1);
  std::vector<asrd_frame_stat> st((size_t)f + 2);
  const int32 n = asrd_frame_stats(d_, st.data(), (int32)st.size(), stream_);
  return n > f + 1 ? st[f + 1].next_cutoff : std::numeric_limits<BaseFloat>::infinity();
}

void CudaLatticeDecoder::ProcessNonemitting(BaseFloat) {}

void CudaLatticeDecoder::FinalizeDecoding() {
  asrd_decoder *h[1] = {d_};
  Check(asrd_finalize_decoding(h, 1, stream_), "asrd_finalize_decoding");
  finalized_ = true;
}

bool CudaLatticeDecoder::Decode(AmInterface *decodable) {
  // Every reference caller uses Init/Advance/Finalize; the reference's own Decode() loop reads one
  // frame past the end (inl.h:615) — not reproduced.
  InitDecoding();
  AdvanceDecoding(decodable);
  FinalizeDecoding();
  return NumFramesDecoded() > 0;
}

bool CudaLatticeDecoder::GetBestPath(Lattice *ofst, bool use_final_probs) {
  // inl.h:1071-1094: a linear lattice, built from the end of the path towards the start
  ofst->DeleteStates();
  asrd_decoder *h[1] = {d_};
  int32_t cap = 4 * std::max(NumFramesDecoded(), 0) + 64;
  std::vector<int32_t> il, ol;
  std::vector<float> gr, ac;
  int32_t n = 0, status = 0;
  for (;;) {
    il.resize(cap); ol.resize(cap); gr.resize(cap); ac.resize(cap);
    Check(asrd_get_best_path(h, 1, use_final_probs ? 1 : 0, cap, il.data(), ol.data(), gr.data(), ac.data(), &n,
                             &status, stream_), "asrd_get_best_path");
    if (status != ASRD_ERR_PATH_OVERFLOW) break;
    cap *= 4;
  }
  if (status == ASRD_ERR_NO_TOKENS) return false;  // the reference warns and returns false (inl.h:1078-1079)
  Check(status, "GetBestPath");
  StateId state = ofst->AddState();
  ofst->SetFinal(state);
  for (int32_t k = n - 1; k >= 0; --k) {  // path order is start -> end; the reference adds end -> start
    LatticeArc arc(il[k], ol[k], state, LatticeWeight(gr[k], ac[k]));
    StateId ns = ofst->AddState();
    ofst->AddArc(ns, arc);
    state = ns;
  }
  ofst->SetStart(state);
  return true;
}

bool CudaLatticeDecoder::GetRawLattice(Lattice *ofst, bool use_final_probs) {
  // inl.h:868-975: one lattice state per surviving token (frame by frame, topologically sorted
  // inside a frame, so state 0 is the start state), one arc per surviving forward link.
  ofst->DeleteStates();
  if (finalized_ && !use_final_probs) {
    fprintf(stderr, "WARNING You cannot call FinalizeDecoding() and then call GetRawLattice() with "
                    "use_final_probs == false\n");  // inl.h:879-884
    return false;
  }
  int64_t tok_cap = 1 << 15, link_cap = 1 << 16, nt = 0, nl = 0;
  std::vector<asrd_lat_token> toks;
  std::vector<asrd_lat_link> links;
  for (;;) {
    toks.resize(tok_cap);
    links.resize(link_cap);
    int rc = asrd_get_raw_lattice(d_, use_final_probs ? 1 : 0, toks.data(), tok_cap, links.data(), link_cap, &nt,
                                  &nl, stream_);
    if (rc == ASRD_ERR_PATH_OVERFLOW) {
      tok_cap = std::max(tok_cap, nt + 16);
      link_cap = std::max(link_cap, nl + 16);
      continue;
    }
    if (rc == ASRD_ERR_NO_TOKENS) return false;  // inl.h:906-911
    Check(rc, "asrd_get_raw_lattice");
    break;
  }
  // Topological order inside every frame (eps links stay inside a frame): Kahn's algorithm over
  // the frame's eps links; tokens arrive sorted by (frame, state).
  std::vector<int32_t> indeg(nt, 0), order;
  std::vector<std::vector<int32_t> > eps_out(nt);
  for (int64_t i = 0; i < nl; ++i)
    if (links[i].ilabel == 0) {
      eps_out[links[i].src].push_back(links[i].dst);
      ++indeg[links[i].dst];
    }
  order.reserve(nt);
  for (int64_t b = 0; b < nt;) {
    int64_t e = b;
    while (e < nt && toks[e].frame == toks[b].frame) ++e;
    std::vector<int32_t> ready;
    for (int64_t i = e; i-- > b;)
      if (indeg[i] == 0) ready.push_back((int32_t)i);
    while (!ready.empty()) {
      const int32_t t = ready.back();
      ready.pop_back();
      order.push_back(t);
      for (size_t k = 0; k < eps_out[t].size(); ++k)
        if (--indeg[eps_out[t][k]] == 0) ready.push_back(eps_out[t][k]);
    }
    b = e;
  }
  if ((int64_t)order.size() != nt) Check(ASRD_ERR_STATE, "epsilon loop in the lattice");  // inl.h:1061-1062
  std::vector<StateId> state_of(nt);
  for (int64_t k = 0; k < nt; ++k) state_of[order[k]] = ofst->AddState();
  ofst->SetStart(0);
  for (int64_t i = 0; i < nt; ++i)
    if (toks[i].is_final) ofst->SetFinal(state_of[i]);
  for (int64_t i = 0; i < nl; ++i)  // final costs are 0 on the single super-final state (inl.h:695)
    ofst->AddArc(state_of[links[i].src], LatticeArc(links[i].ilabel, links[i].olabel, state_of[links[i].dst],
                                                    LatticeWeight(links[i].graph, links[i].acoustic)));
  return ofst->NumStates() > 0;
}

}  // namespace asrd_host
