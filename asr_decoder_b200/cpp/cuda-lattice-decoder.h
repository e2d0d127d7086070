// CudaLatticeDecoder — the drop-in `DecoderItf` implementation backed by the B200 library
// (include/asrd.h).  Same constructor shape as the reference decoders
// (OnlineLatticeDecoderBase(FST*, const LatticeFasterDecoderConfig&),
// my-decoder/online-decoder-base.h:95): the graph handle is shared and NOT owned
// (online-decoder-base-inl.h:24, _delete_fst(false)).
#ifndef ASRD_CPP_CUDA_LATTICE_DECODER_H_
#define ASRD_CPP_CUDA_LATTICE_DECODER_H_

#include <stdexcept>
#include <string>
#include <vector>

#include "asrd.h"

// Two ways to compile this class:
//  * inside the reference tree (-DASRD_REFERENCE_TREE, include root = the reference checkout): it
//    derives from the reference's OWN DecoderItf and fills the reference's own Lattice, so it can
//    be handed to DeterminizeLatticeWrapper / NShortestPath and selected by --graph-type like the
//    CPU decoders; CudaFst::FromFst uploads an already loaded reference Fst.
//  * stand-alone: against decoder-itf.h next to this file, a from-scratch mirror of the same
//    surface (so the class and its tests build where the reference sources are absent).
#ifdef ASRD_REFERENCE_TREE
#include "src/my-decoder/decoder-itf.h"
#include "src/my-decoder/lattice-faster-decoder-conf.h"
#include "src/newfst/lattice-fst.h"
#include "src/newfst/optimize-fst.h"
namespace asrd_host {
#ifdef NAMESPACE
using namespace datemoon;
#endif
typedef int int32;
}  // namespace asrd_host
#else
#include "decoder-itf.h"
#endif

namespace asrd_host {

// A decodable that already holds the matrix: AdvanceDecoding uploads rows without T x P
// virtual calls (SURVEY.md section 8b "input side").  Any other AmInterface is served through
// LogLikelihood(frame, index) calls (the reference's pull model, itf/decodable-itf.h:65-104).
class MatrixDecodableInterface : public AmInterface {
 public:
  virtual const BaseFloat *Data() const = 0;  // row-major, column = index - 1
  virtual int32 Stride() const = 0;           // floats between rows
};

// Device-resident HCLG: what replaces `Fst` (newfst/optimize-fst.h:53-307) for this decoder.
class CudaFst {
 public:
  CudaFst() : g_(NULL) {}
  ~CudaFst() { asrd_graph_destroy(g_); }
  // Fst::ReadFst(const char*), optimize-fst.h:208-219 (same file format)
  bool ReadFst(const char *file, int device = 0) {
    asrd_graph_destroy(g_);
    g_ = NULL;
    return asrd_graph_read(file, device, &g_) == ASRD_OK;
  }
  // ConstFst::Read + Fst(const ConstFst&) (newfst/const-fst.h:189-221, optimize-fst.h:82-134):
  // an OpenFst const file, e.g. a Kaldi HCLG.fst (--fst-type=const in OnlineDecoderInfo,
  // kaldi-nnet3/kaldi-online-nnet3-my-decoder.h:210-220)
  bool ReadConstFst(const char *file, int device = 0) {
    asrd_graph_destroy(g_);
    g_ = NULL;
    return asrd_graph_read_const(file, device, &g_) == ASRD_OK;
  }
  // ClgFst::Init(clgfst, hmmfst) (my-decoder/clg-fst.h:17-74): CLG graph + HMM set, written out as one
  // static device graph; decoders built on it follow the reference's CLG decoder
  // (OnlineClgLatticeDecoderMempool, --graph-type=clg in kaldi-online-nnet3-my-decoder.h:250-283)
  bool ReadClg(const char *clg_file, const char *hmm_file, int device = 0) {
    asrd_graph_destroy(g_);
    g_ = NULL;
    return asrd_graph_read_clg(clg_file, hmm_file, device, &g_) == ASRD_OK;
  }
  // from the arrays an already loaded reference `Fst` holds
  bool FromArrays(const asrd_arc *arcs, const uint32_t *num_arcs, const uint32_t *niepsilons, int32_t states,
                  int64_t n_arcs, int32_t start, int32_t final_state, int device = 0) {
    asrd_graph_destroy(g_);
    g_ = NULL;
    return asrd_graph_create(arcs, num_arcs, niepsilons, states, n_arcs, start, final_state, device, &g_) == ASRD_OK;
  }
#ifdef ASRD_REFERENCE_TREE
  // Upload a graph the host already holds as the reference's Fst (Fst::ReadFst or
  // Fst(const ConstFst&), newfst/optimize-fst.h:82-134,208-280): its two flat arrays are read in place.
  bool FromFst(Fst &fst, int device = 0) {
    const int32_t S = fst.TotState();
    std::vector<uint32_t> na((size_t)S), ne((size_t)S);
    int32_t final_state = -1;
    for (int32_t s = 0; s < S; ++s) {
      na[s] = fst.GetState(s)->_num_arcs;
      ne[s] = fst.GetState(s)->_niepsilons;
      if (fst.IsFinal(s)) final_state = s;
    }
    return FromArrays(reinterpret_cast<const asrd_arc *>(fst.GetArc(0)), na.data(), ne.data(), S, fst.TotArc(),
                      fst.Start(), final_state, device);
  }
#endif
  asrd_graph *handle() const { return g_; }

 private:
  CudaFst(const CudaFst &);
  CudaFst &operator=(const CudaFst &);
  asrd_graph *g_;
};

// Device-resident LM FSA: what replaces `ArpaLm` (newlm/arpa2fsa.h:217-480) for the biglm decoder.
// The caller rescales the OLD LM by -1 before uploading, as the reference bin does
// (kaldi-nnet3bin/kaldi-hclg-my-decoder-biglm.cc:55-60: lm1.Rescale(-1.0)).
class CudaLm {
 public:
  CudaLm() : lm_(NULL) {}
  ~CudaLm() { asrd_lm_destroy(lm_); }
  // the arrays ArpaLm::Read loads (arpa2fsa.h:399-439)
  bool FromArrays(int32_t bos, int32_t eos, int32_t n_states, const int32_t *arc_num, const float *backoff_prob,
                  const int32_t *backoff_id, const asrd_lm_arc *arcs, int64_t n_arcs, int device = 0) {
    asrd_lm_destroy(lm_);
    lm_ = NULL;
    return asrd_lm_create(bos, eos, n_states, arc_num, backoff_prob, backoff_id, arcs, n_arcs, device, &lm_) == ASRD_OK;
  }
  // ArpaLm::Read(const char*) file format; `scale` = the Rescale() factor applied while loading
  bool Read(const char *file, float scale = 1.0f, int device = 0);
  // ARPA text -> that file format: what the reference's arpa2fsa tool writes (newlm/arpa2fsa-bin.cc)
  static bool ConvertArpa(const char *arpa_file, const char *wordlist, const char *out_file) {
    return asrd_lm_convert_arpa(arpa_file, wordlist, out_file) == ASRD_OK;
  }
  asrd_lm *handle() const { return lm_; }

 private:
  CudaLm(const CudaLm &);
  CudaLm &operator=(const CudaLm &);
  asrd_lm *lm_;
};

class CudaLatticeDecoder : public DecoderItf {
 public:
  // prune_tokens: run PruneActiveTokens every config._prune_interval frames like the reference does
  // (inl.h:438-480, :660-661) — the device memory of a long utterance stays bounded by its live
  // tokens, at the price of the prune sweeps; results do not change.
  CudaLatticeDecoder(CudaFst *graph, const LatticeFasterDecoderConfig &config, int max_frames = 0,
                     void *cuda_stream = NULL, bool prune_tokens = false);
  // biglm: OnlineLatticeDecoderMempoolBaseBiglm(fst, config, oldlm, newlm) (…-biglm.h:21-30)
  CudaLatticeDecoder(CudaFst *graph, const LatticeFasterDecoderConfig &config, CudaLm *oldlm, CudaLm *newlm,
                     int max_frames = 0, void *cuda_stream = NULL);
  virtual ~CudaLatticeDecoder();

  virtual void InitDecoding();
  virtual void AdvanceDecoding(AmInterface *decodable, int32 max_num_frames = -1);
  virtual void FinalizeDecoding();
  virtual int32 NumFramesDecoded() const;
  // One frame step on the device (emitting expansion AND the eps closure, which is fused into
  // the same step); returns the next cutoff like the reference (inl.h:349-350).
  virtual BaseFloat ProcessEmitting(AmInterface *decodable);
  // The closure already ran inside the step; kept for interface completeness (decoder-itf.h:20).
  virtual void ProcessNonemitting(BaseFloat cost_cutoff);
  virtual bool Decode(AmInterface *decodable);
  virtual bool GetBestPath(Lattice *ofst, bool use_final_probs = true);
  // Raw lattice after the lattice-beam pruning of FinalizeDecoding; the forward links are
  // regenerated and pruned on the device (asrd_get_raw_lattice), the Lattice is built here.
  virtual bool GetRawLattice(Lattice *ofst, bool use_final_probs = true);

  asrd_decoder *handle() const { return d_; }

 private:
  void Check(int status, const char *what) const;  // LOG_ERR -> throw std::runtime_error
  void Upload(AmInterface *decodable, int32 first, int32 count);
  void Create(CudaFst *graph, const LatticeFasterDecoderConfig &config, CudaLm *oldlm, CudaLm *newlm, int max_frames,
              bool prune_tokens);
  asrd_decoder *d_;
  void *stream_;
  bool finalized_;
  std::vector<float> stage_;
};

}  // namespace asrd_host
#endif
