// Test driver for the C++ drop-in class: decodes a loglikes file through a `DecoderItf*`
// exactly like the reference's offline bin (kaldi-nnet3bin/kaldi-hclg-my-decoder.cc:95-122)
// and prints one JSON line per utterance (same fields as oracle/ref_harness.cc).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cuda-lattice-decoder.h"

using namespace asrd_host;

namespace {
struct Utt {
  int T, P;
  std::vector<float> ll;
};
class PullDecodable : public AmInterface {  // forces the generic LogLikelihood() path
 public:
  explicit PullDecodable(const Utt *u) : u_(u), ready_(u->T) {}
  virtual BaseFloat LogLikelihood(int32 f, int32 i) { return u_->ll[(size_t)f * u_->P + (i - 1)]; }
  virtual bool IsLastFrame(int32 f) const { return f == u_->T - 1; }
  virtual int32 NumFramesReady() const { return ready_; }
  virtual int32 NumIndices() const { return u_->P; }
  void SetReady(int r) { ready_ = r; }
 protected:
  const Utt *u_;
  int ready_;
};
class FastDecodable : public MatrixDecodableInterface {
 public:
  explicit FastDecodable(const Utt *u) : u_(u), ready_(u->T) {}
  virtual BaseFloat LogLikelihood(int32 f, int32 i) { return u_->ll[(size_t)f * u_->P + (i - 1)]; }
  virtual bool IsLastFrame(int32 f) const { return f == u_->T - 1; }
  virtual int32 NumFramesReady() const { return ready_; }
  virtual int32 NumIndices() const { return u_->P; }
  virtual const BaseFloat *Data() const { return u_->ll.data(); }
  virtual int32 Stride() const { return u_->P; }
  void SetReady(int r) { ready_ = r; }
 private:
  const Utt *u_;
  int ready_;
};
unsigned Bits(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
}  // namespace

int main(int argc, char **argv) {
  std::string graph, hmm, loglikes, lm1f, lm2f;
  LatticeFasterDecoderConfig cfg;
  cfg._beam = 13.0f; cfg._max_active = 7000; cfg._min_active = 200; cfg._lattice_beam = 8.0f;
  int chunk = 0;
  bool pull = false, lattice = false, prune = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    size_t eq = a.find('=');
    std::string k = a.substr(0, eq), v = eq == std::string::npos ? "" : a.substr(eq + 1);
    if (k == "--graph") graph = v;
    else if (k == "--hmm") hmm = v;  // --graph is then a CLG graph (ClgFst::Init(clgfst, hmmfst))
    else if (k == "--loglikes") loglikes = v;
    else if (k == "--beam") cfg._beam = atof(v.c_str());
    else if (k == "--max-active") cfg._max_active = atoi(v.c_str());
    else if (k == "--min-active") cfg._min_active = atoi(v.c_str());
    else if (k == "--lattice-beam") cfg._lattice_beam = atof(v.c_str());
    else if (k == "--chunk") chunk = atoi(v.c_str());
    else if (k == "--pull") pull = true;
    else if (k == "--lattice") lattice = true;
    else if (k == "--prune") prune = true;  // PruneActiveTokens every prune_interval frames on the device
    else if (k == "--prune-interval") cfg._prune_interval = atoi(v.c_str());
    else if (k == "--lm1") lm1f = v;
    else if (k == "--lm2") lm2f = v;
    else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
  }
  FILE *fp = fopen(loglikes.c_str(), "rb");
  if (!fp) return 3;
  int magic = 0, n = 0;
  if (fread(&magic, 4, 1, fp) != 1 || magic != 0x4c4c5341 || fread(&n, 4, 1, fp) != 1) return 3;
  std::vector<Utt> utts(n);
  int max_t = 1;
  for (int i = 0; i < n; ++i) {
    if (fread(&utts[i].T, 4, 1, fp) != 1 || fread(&utts[i].P, 4, 1, fp) != 1) return 3;
    utts[i].ll.resize((size_t)utts[i].T * utts[i].P);
    if (fread(utts[i].ll.data(), 4, utts[i].ll.size(), fp) != utts[i].ll.size()) return 3;
    if (utts[i].T > max_t) max_t = utts[i].T;
  }
  fclose(fp);
  try {
    CudaFst fst;
    if (!(hmm.empty() ? fst.ReadFst(graph.c_str()) : fst.ReadClg(graph.c_str(), hmm.c_str()))) { fprintf(stderr, "load fst error.\n"); return 4; }
    CudaLm lm1, lm2;  // kaldi-hclg-my-decoder-biglm.cc:55-60: lm1.Read, lm2.Read, lm1.Rescale(-1.0)
    const bool biglm = !lm1f.empty();
    if (biglm && (!lm1.Read(lm1f.c_str(), -1.0f) || !lm2.Read(lm2f.c_str()))) { fprintf(stderr, "load lm error.\n"); return 4; }
    CudaLatticeDecoder plain(&fst, cfg, max_t + 8, NULL, prune);
    CudaLatticeDecoder rescoring(&fst, cfg, biglm ? &lm1 : NULL, biglm ? &lm2 : NULL, max_t + 8);
    DecoderItf *decode = biglm ? &rescoring : &plain;  // everything below goes through the reference interface
    for (int i = 0; i < n; ++i) {
      decode->InitDecoding();
      PullDecodable pd(&utts[i]);
      FastDecodable fd(&utts[i]);
      AmInterface *dec = pull ? (AmInterface *)&pd : (AmInterface *)&fd;
      if (chunk > 0) {
        for (int ready = 0; ready < utts[i].T;) {
          ready = std::min(utts[i].T, ready + chunk);
          pd.SetReady(ready);
          fd.SetReady(ready);
          decode->AdvanceDecoding(dec);
        }
      } else {
        decode->AdvanceDecoding(dec);
      }
      decode->FinalizeDecoding();
      Lattice best_path;
      std::vector<int> words, ali;
      float tot = 0, lm = 0;
      bool ok = decode->GetBestPath(&best_path);
      if (ok) ok = LatticeToVector(best_path, words, ali, tot, lm);
      int raw_states = -1, raw_arcs = -1, raw_finals = 0;
      if (lattice) {  // kaldi-hclg-my-decoder.cc:134: decode.GetRawLattice(&lat1, true)
        Lattice lat1;
        if (decode->GetRawLattice(&lat1, true)) {
          raw_states = lat1.NumStates();
          raw_arcs = 0;
          for (int st = 0; st < lat1.NumStates(); ++st) {
            raw_arcs += (int)lat1.GetState(st)->GetArcSize();
            raw_finals += lat1.Final(st) ? 1 : 0;
            for (size_t a = 0; a < lat1.GetState(st)->GetArcSize(); ++a)
              if (lat1.GetState(st)->GetArc(a)->_to <= st && lat1.GetState(st)->GetArc(a)->_to != st) raw_arcs = -1000000;  // must be topologically sorted
          }
        }
      }
      printf("{\"utt\": %d, \"ok\": %s, \"frames\": %d, \"tot\": %.9g, \"tot_bits\": %u, \"lm_bits\": %u, "
             "\"raw_states\": %d, \"raw_arcs\": %d, \"raw_finals\": %d, \"words\": [",
             i, ok ? "true" : "false", decode->NumFramesDecoded(), tot, Bits(tot), Bits(lm), raw_states, raw_arcs,
             raw_finals);
      for (size_t k = 0; k < words.size(); ++k) printf("%s%d", k ? "," : "", words[k]);
      printf("], \"ali\": [");
      for (size_t k = 0; k < ali.size(); ++k) printf("%s%d", k ? "," : "", ali[k]);
      printf("]}\n");
    }
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 5;
  }
  return 0;
}
