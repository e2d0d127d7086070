// TEST INFRASTRUCTURE — not product code.
//
// Driver for the UNMODIFIED reference biglm decoder OnlineLatticeDecoderMempoolBiglm
// (src/my-decoder/online-decoder-mempool-base-biglm.h:570), compiled in place from /root/reference.
// Call sequence of src/kaldi-nnet3bin/kaldi-hclg-my-decoder-biglm.cc:55-60 and its decode loop:
//   lm1.Read, lm2.Read, lm1.Rescale(-1.0); decoder(&fst, opt, &lm1, &lm2);
//   InitDecoding -> AdvanceDecoding -> FinalizeDecoding -> GetBestPath -> LatticeToVector
// One JSON object per utterance.  Only tests and bench baselines may execute this binary.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include "src/my-decoder/online-decoder-mempool-base-biglm.h"
#include "src/newfst/lattice-functions.h"

using namespace datemoon;

namespace {
struct Utt {
  int T = 0, P = 0;
  std::vector<float> ll;
};
class MatrixDecodable : public AmInterface {
 public:
  explicit MatrixDecodable(const Utt *u) : u_(u) {}
  virtual BaseFloat LogLikelihood(int32 frame, int32 index) { return u_->ll[(size_t)frame * u_->P + (index - 1)]; }
  virtual bool IsLastFrame(int32 frame) const { return frame == u_->T - 1; }
  virtual int32 NumFramesReady() const { return u_->T; }
  virtual int32 NumIndices() const { return u_->P; }
 private:
  const Utt *u_;
};
struct Stat { unsigned n_raw, n_within; float cutoff; };
class Probe : public OnlineLatticeDecoderMempoolBiglm {
 public:
  typedef OnlineLatticeDecoderMempoolBiglm Base;
  Probe(Fst *fst, const LatticeFasterDecoderConfig &c, ArpaLm *a, ArpaLm *b) : Base(fst, c, a, b) {}
  std::vector<Stat> stats;
  virtual void ProcessNonemitting(BaseFloat cutoff) {
    Base::ProcessNonemitting(cutoff);
    Stat s = {0, 0, cutoff};
    for (const Elem *e = _toks.GetList(); e != NULL; e = e->tail) {
      ++s.n_raw;
      if (e->val->_tot_cost < cutoff) ++s.n_within;
    }
    stats.push_back(s);
  }
};
unsigned Bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
}  // namespace

int main(int argc, char **argv) {
  std::string graph, loglikes, lm1f, lm2f;
  LatticeFasterDecoderConfig cfg;
  cfg._beam = 13.0; cfg._max_active = 7000; cfg._min_active = 200; cfg._lattice_beam = 8.0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    size_t eq = a.find('=');
    std::string k = a.substr(0, eq), v = eq == std::string::npos ? "" : a.substr(eq + 1);
    if (k == "--graph") graph = v;
    else if (k == "--loglikes") loglikes = v;
    else if (k == "--lm1") lm1f = v;
    else if (k == "--lm2") lm2f = v;
    else if (k == "--beam") cfg._beam = atof(v.c_str());
    else if (k == "--max-active") cfg._max_active = atoi(v.c_str());
    else if (k == "--min-active") cfg._min_active = atoi(v.c_str());
    else if (k == "--lattice-beam") cfg._lattice_beam = atof(v.c_str());
    else if (k == "--hash-ratio") cfg._hash_ratio = atof(v.c_str());
    else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
  }
  ArpaLm lm1, lm2;
  // the reference's readers chat on stdout: keep it clean for the JSON lines
  FILE *real_out = fdopen(dup(fileno(stdout)), "w");
  if (!freopen("/dev/null", "w", stdout)) return 3;
  if (!lm1.Read(lm1f.c_str()) || !lm2.Read(lm2f.c_str())) return 3;
  lm1.Rescale(-1.0);
  Fst fst;
  if (!fst.ReadFst(graph.c_str())) return 3;
  FILE *fp = fopen(loglikes.c_str(), "rb");
  if (!fp) return 3;
  int magic = 0, n = 0;
  if (fread(&magic, 4, 1, fp) != 1 || magic != 0x4c4c5341 || fread(&n, 4, 1, fp) != 1) return 3;
  std::vector<Utt> utts(n);
  for (int i = 0; i < n; ++i) {
    if (fread(&utts[i].T, 4, 1, fp) != 1 || fread(&utts[i].P, 4, 1, fp) != 1) return 3;
    utts[i].ll.resize((size_t)utts[i].T * utts[i].P);
    if (fread(utts[i].ll.data(), 4, utts[i].ll.size(), fp) != utts[i].ll.size()) return 3;
  }
  fclose(fp);
  Probe dec(&fst, cfg, &lm1, &lm2);
  for (int i = 0; i < n; ++i) {
    const auto t_begin = std::chrono::steady_clock::now();
    dec.stats.clear();
    dec.InitDecoding();
    MatrixDecodable decodable(&utts[i]);
    dec.AdvanceDecoding(&decodable);
    dec.FinalizeDecoding();
    Lattice best_path;
    std::vector<int> words, ali;
    float tot = 0, lm = 0;
    bool ok = dec.GetBestPath(&best_path);
    if (ok) ok = LatticeToVector(best_path, words, ali, tot, lm);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    fprintf(real_out, "{\"utt\": %d, \"ok\": %s, \"frames\": %d, \"seconds\": %.6f, \"tot\": %.9g, \"tot_bits\": %u, \"lm_bits\": %u, \"words\": [",
            i, ok ? "true" : "false", utts[i].T, secs, tot, Bits(tot), Bits(lm));
    for (size_t k = 0; k < words.size(); ++k) fprintf(real_out, "%s%d", k ? "," : "", words[k]);
    fprintf(real_out, "], \"ali\": [");
    for (size_t k = 0; k < ali.size(); ++k) fprintf(real_out, "%s%d", k ? "," : "", ali[k]);
    fprintf(real_out, "], \"n_raw\": [");
    for (size_t k = 0; k < dec.stats.size(); ++k) fprintf(real_out, "%s%u", k ? "," : "", dec.stats[k].n_raw);
    fprintf(real_out, "], \"n_within\": [");
    for (size_t k = 0; k < dec.stats.size(); ++k) fprintf(real_out, "%s%u", k ? "," : "", dec.stats[k].n_within);
    fprintf(real_out, "], \"next_cutoff_bits\": [");
    for (size_t k = 0; k < dec.stats.size(); ++k) fprintf(real_out, "%s%u", k ? "," : "", Bits(dec.stats[k].cutoff));
    fprintf(real_out, "]}\n");
  }
  fclose(real_out);
  return 0;
}
