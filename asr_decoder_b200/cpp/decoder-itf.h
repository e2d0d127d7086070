// Host-side C++ mirror of the reference decoder interface, for the drop-in class
// CudaLatticeDecoder (cuda-lattice-decoder.h).  Written from scratch; names, argument meaning
// and error behaviour follow the reference (paths relative to the reference's src/):
//   DecodableInterface / AmInterface   itf/decodable-itf.h:65-105
//   LatticeWeight, LatticeArc          newfst/weigth.h:192-324, newfst/arc.h:17-149
//   LatticeState, Lattice              newfst/lattice-fst.h:18-346
//   LatticeFasterDecoderConfig         my-decoder/lattice-faster-decoder-conf.h:8-69
//   DecoderItf                         my-decoder/decoder-itf.h:10-25
//   LatticeToVector                    newfst/lattice-functions.cc:179-217
// Inside the reference tree the maintainer deletes this header and includes the reference's own
// (INTEGRATION.md); the class below compiles against either because only this surface is used.
#ifndef ASRD_CPP_DECODER_ITF_H_
#define ASRD_CPP_DECODER_ITF_H_

#include <cassert>
#include <cstdint>
#include <limits>
#include <vector>

namespace asrd_host {

typedef float BaseFloat;
typedef int int32;
typedef int StateId;
typedef int Label;
const int kNoStateId = -1;

class DecodableInterface {
 public:
  virtual BaseFloat LogLikelihood(int32 frame, int32 index) = 0;  // index is 1-based (ilabel)
  virtual bool IsLastFrame(int32 frame) const = 0;
  virtual int32 NumFramesReady() const { return -1; }
  virtual int32 NumIndices() const = 0;
  virtual ~DecodableInterface() {}
};
typedef DecodableInterface AmInterface;

class LatticeWeight {
 public:
  LatticeWeight() : v1_(0.0f), v2_(0.0f) {}
  LatticeWeight(BaseFloat graph, BaseFloat acoustic) : v1_(graph), v2_(acoustic) {}
  BaseFloat Value1() const { return v1_; }
  BaseFloat Value2() const { return v2_; }
  BaseFloat Value() const { return v1_ + v2_; }
  static LatticeWeight One() { return LatticeWeight(0.0f, 0.0f); }
  static LatticeWeight Zero() {
    return LatticeWeight(std::numeric_limits<BaseFloat>::infinity(), std::numeric_limits<BaseFloat>::infinity());
  }

 private:
  BaseFloat v1_, v2_;
};

struct LatticeArc {
  Label _input;
  Label _output;
  LatticeWeight _w;
  StateId _to;
  LatticeArc() : _input(0), _output(0), _to(0) {}
  LatticeArc(Label i, Label o, StateId to, LatticeWeight w) : _input(i), _output(o), _w(w), _to(to) {}
};

class LatticeState {
 public:
  LatticeState() : _final(0) {}
  void SetFinal() { _final = 1; }
  void UnsetFinal() { _final = 0; }
  bool IsFinal() const { return _final != 0; }
  void AddArc(const LatticeArc &arc) { _arc.push_back(arc); }
  LatticeArc *GetArc(size_t i) { return i < _arc.size() ? &_arc[i] : NULL; }
  size_t GetArcSize() const { return _arc.size(); }

 private:
  int _final;
  std::vector<LatticeArc> _arc;
};

class Lattice {
 public:
  Lattice() : _startid(kNoStateId) {}
  ~Lattice() { DeleteStates(); }
  int NumStates() const { return (int)_state.size(); }
  StateId AddState() {
    _state.push_back(new LatticeState);
    return (StateId)_state.size() - 1;
  }
  void AddArc(StateId s, const LatticeArc &arc) {
    assert(s < (StateId)_state.size());
    _state[s]->AddArc(arc);
  }
  void SetStart(StateId s) { _startid = s; }
  StateId Start() const { return _startid; }
  void SetFinal(StateId s, float = 0) {
    assert(s < (StateId)_state.size());
    _state[s]->SetFinal();
  }
  bool Final(StateId s) const { return _state[s]->IsFinal(); }
  LatticeState *GetState(StateId s) {
    assert(s < (StateId)_state.size());
    return _state[s];
  }
  void DeleteStates() {
    for (size_t i = 0; i < _state.size(); ++i) delete _state[i];
    _state.clear();
    _startid = kNoStateId;
  }

 private:
  Lattice(const Lattice &);
  Lattice &operator=(const Lattice &);
  StateId _startid;
  std::vector<LatticeState *> _state;
};

struct LatticeFasterDecoderConfig {
  float _beam;
  int _max_active;
  int _min_active;
  float _lattice_beam;
  int _prune_interval;
  bool _determinize_lattice;
  float _beam_delta;
  float _hash_ratio;
  float _prune_scale;
  LatticeFasterDecoderConfig()
      : _beam(16.0f), _max_active(std::numeric_limits<int>::max()), _min_active(200), _lattice_beam(10.0f),
        _prune_interval(25), _determinize_lattice(true), _beam_delta(0.5f), _hash_ratio(2.0f),
        _prune_scale(0.1f) {}
  void Check() const {
    assert(_beam > 0.0 && _max_active > 1 && _lattice_beam > 0.0 && _prune_interval > 0 &&
           _beam_delta > 0.0 && _hash_ratio >= 1.0 && _prune_scale > 0.0 && _prune_scale < 1.0);
  }
};

class DecoderItf {
 public:
  virtual ~DecoderItf() {}
  virtual void InitDecoding() = 0;
  virtual void AdvanceDecoding(AmInterface *decodable, int32 max_num_frames = -1) = 0;
  virtual void FinalizeDecoding() = 0;
  virtual int32 NumFramesDecoded() const = 0;
  virtual BaseFloat ProcessEmitting(AmInterface *decodable) = 0;
  virtual void ProcessNonemitting(BaseFloat cost_cutoff) = 0;
  virtual bool Decode(AmInterface *decodable) = 0;
  virtual bool GetBestPath(Lattice *ofst, bool use_final_probs = true) = 0;
  virtual bool GetRawLattice(Lattice *ofst, bool use_final_probs = true) = 0;
};

// words = non-zero olabels, alignment = non-zero ilabels, float sums in path order
inline bool LatticeToVector(Lattice &best_path, std::vector<int> &words, std::vector<int> &phones,
                            float &tot_score, float &lm_score) {
  if (best_path.Start() == kNoStateId) return false;
  tot_score = 0;
  lm_score = 0;
  LatticeState *cur = best_path.GetState(best_path.Start());
  while (!cur->IsFinal()) {
    LatticeArc *arc = cur->GetArc(0);
    if (arc->_input != 0) phones.push_back(arc->_input);
    if (arc->_output != 0) words.push_back(arc->_output);
    lm_score += arc->_w.Value1();
    tot_score += arc->_w.Value1() + arc->_w.Value2();
    cur = best_path.GetState(arc->_to);
  }
  return true;
}

}  // namespace asrd_host
#endif
