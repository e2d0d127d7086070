// TEST INFRASTRUCTURE — not product code.
//
// The drop-in class inside the reference tree.  This file is compiled WITH THE REFERENCE'S OWN
// HEADERS AND SOURCES (oracle/Makefile, outputs only under oracle/_ref/): CudaLatticeDecoder
// (asr_decoder_b200/cpp/cuda-lattice-decoder.{h,cc}, -DASRD_REFERENCE_TREE) derives from the
// reference's DecoderItf (src/my-decoder/decoder-itf.h:10-25), is uploaded from a reference Fst
// (CudaFst::FromFst), and fills the reference's Lattice, which then goes through the reference's
// own post-pass exactly like its offline bin does (src/kaldi-nnet3bin/kaldi-hclg-my-decoder.cc:134-165):
//
//   DecoderItf *dec = --decoder=cuda ? new CudaLatticeDecoder(...) : new OnlineLatticeDecoderMempool(...)
//   dec->InitDecoding(); dec->AdvanceDecoding(&decodable); dec->FinalizeDecoding();
//   dec->GetBestPath(&best) -> LatticeToVector
//   dec->GetRawLattice(&raw) -> LatticeCheckFormat -> DeterminizeLatticeWrapper -> NShortestPath
//                            -> ConvertNbestToVector -> LatticeToVector per path
//
// One JSON line per utterance: one-best, raw / determinised lattice sizes, the n-best list
// (words + costs), and the seconds spent in decode / GetRawLattice / determinise+n-best.
// --decoder=ref needs no GPU; --decoder=cuda needs libasrd_b200.so and a B200.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "src/my-decoder/online-decoder-mempool-base.h"
#include "src/newfst/lattice-determinize-api.h"
#include "src/newfst/lattice-functions.h"
#include "src/newfst/lattice-to-nbest.h"

#include "asr_decoder_b200/cpp/cuda-lattice-decoder.h"

using namespace datemoon;

namespace {

struct Utt {
  int T = 0, P = 0;
  std::vector<float> ll;
};

// the matrix decodable the drop-in class uploads from without T x P virtual calls; the reference
// decoder uses it through LogLikelihood like any other AmInterface
class MatrixDecodable : public asrd_host::MatrixDecodableInterface {
 public:
  explicit MatrixDecodable(const Utt *u) : u_(u) {}
  virtual BaseFloat LogLikelihood(int32 frame, int32 index) { return u_->ll[(size_t)frame * u_->P + (index - 1)]; }
  virtual bool IsLastFrame(int32 frame) const { return frame == u_->T - 1; }
  virtual int32 NumFramesReady() const { return u_->T; }
  virtual int32 NumIndices() const { return u_->P; }
  virtual const BaseFloat *Data() const { return u_->ll.data(); }
  virtual int32 Stride() const { return u_->P; }

 private:
  const Utt *u_;
};

bool ReadLoglikes(const std::string &file, std::vector<Utt> *utts) {
  FILE *fp = fopen(file.c_str(), "rb");
  if (!fp) return false;
  int magic = 0, n = 0;
  bool ok = fread(&magic, 4, 1, fp) == 1 && magic == 0x4c4c5341 && fread(&n, 4, 1, fp) == 1;
  if (ok) utts->resize(n);
  for (int i = 0; ok && i < n; ++i) {
    Utt &u = (*utts)[i];
    ok = fread(&u.T, 4, 1, fp) == 1 && fread(&u.P, 4, 1, fp) == 1;
    if (ok) {
      u.ll.resize((size_t)u.T * u.P);
      ok = fread(u.ll.data(), 4, u.ll.size(), fp) == u.ll.size();
    }
  }
  fclose(fp);
  return ok;
}

unsigned Bits(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}

double Now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void PrintInts(const char *name, const std::vector<int> &v) {
  printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%d", i ? "," : "", v[i]);
  printf("]");
}

int CountArcs(Lattice &l) {
  int n = 0;
  for (int s = 0; s < l.NumStates(); ++s) n += (int)l.GetState(s)->GetArcSize();
  return n;
}

}  // namespace

int main(int argc, char **argv) {
  std::string graph, loglikes, which = "cuda";
  LatticeFasterDecoderConfig cfg;
  cfg._beam = 13.0;
  cfg._max_active = 7000;
  cfg._min_active = 200;
  cfg._lattice_beam = 8.0;
  int nbest = 10, max_frames = 0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&](const char *name) -> const char * {
      size_t n = strlen(name);
      return (a.compare(0, n, name) == 0 && a.size() > n && a[n] == '=') ? a.c_str() + n + 1 : NULL;
    };
    const char *v;
    if ((v = val("--graph"))) graph = v;
    else if ((v = val("--loglikes"))) loglikes = v;
    else if ((v = val("--decoder"))) which = v;
    else if ((v = val("--beam"))) cfg._beam = atof(v);
    else if ((v = val("--max-active"))) cfg._max_active = atoi(v);
    else if ((v = val("--min-active"))) cfg._min_active = atoi(v);
    else if ((v = val("--lattice-beam"))) cfg._lattice_beam = atof(v);
    else if ((v = val("--nbest"))) nbest = atoi(v);
    else if ((v = val("--max-frames"))) max_frames = atoi(v);
    else { fprintf(stderr, "unknown arg %s\n", a.c_str()); return 2; }
  }
  if (graph.empty() || loglikes.empty() || (which != "cuda" && which != "ref")) {
    fprintf(stderr, "usage: dropin_nbest --graph=G --loglikes=L --decoder=cuda|ref [--beam= --max-active= "
                    "--min-active= --lattice-beam= --nbest=N --max-frames=N]\n");
    return 2;
  }
  Fst fst;  // the reference's own loader (src/newfst/optimize-fst.h:208-280)
  if (!fst.ReadFst(graph.c_str())) return 3;
  std::vector<Utt> utts;
  if (!ReadLoglikes(loglikes, &utts)) { fprintf(stderr, "cannot read %s\n", loglikes.c_str()); return 3; }
  int longest = 0;
  for (size_t i = 0; i < utts.size(); ++i) longest = std::max(longest, utts[i].T);

  asrd_host::CudaFst cuda_graph;
  DecoderItf *dec = NULL;  // the one pointer type the reference's callers hold (kaldi-online-nnet3-my-decoder.h:386)
  if (which == "cuda") {
    if (!cuda_graph.FromFst(fst)) { fprintf(stderr, "graph upload failed\n"); return 4; }
    dec = new asrd_host::CudaLatticeDecoder(&cuda_graph, cfg, max_frames > 0 ? max_frames : longest + 8);
  } else {
    dec = new OnlineLatticeDecoderMempool(&fst, cfg);
  }
  for (size_t u = 0; u < utts.size(); ++u) {
    MatrixDecodable decodable(&utts[u]);
    const double t0 = Now();
    dec->InitDecoding();
    dec->AdvanceDecoding(&decodable);
    dec->FinalizeDecoding();
    Lattice best;
    std::vector<int> words, ali;
    float tot = 0, lm = 0;
    bool ok = dec->GetBestPath(&best) && LatticeToVector(best, words, ali, tot, lm);
    const double t1 = Now();
    Lattice raw, det, nb;
    int raw_states = -1, raw_arcs = -1, det_states = -1, det_arcs = -1;
    double t2 = t1, t3 = t1;
    std::vector<Lattice> paths;
    if (ok && dec->GetRawLattice(&raw, true)) {
      t2 = Now();
      raw_states = raw.NumStates();
      raw_arcs = CountArcs(raw);
      bool debug_ptr = false;
      DeterminizeLatticeOptions opts;
      if (LatticeCheckFormat(&raw) && DeterminizeLatticeWrapper(&raw, &det, opts, &debug_ptr)) {
        det_states = det.NumStates();
        det_arcs = CountArcs(det);
        NShortestPath(det, &nb, (size_t)nbest);
        ConvertNbestToVector(nb, &paths);
      }
      t3 = Now();
    }
    printf("{\"utt\": %d, \"decoder\": \"%s\", \"ok\": %s, \"frames\": %d, \"tot\": %.9g, \"tot_bits\": %u, \"lm_bits\": %u, ",
           (int)u, which.c_str(), ok ? "true" : "false", utts[u].T, tot, Bits(tot), Bits(lm));
    PrintInts("words", words);
    printf(", ");
    PrintInts("ali", ali);
    printf(", \"raw_states\": %d, \"raw_arcs\": %d, \"det_states\": %d, \"det_arcs\": %d, ", raw_states, raw_arcs,
           det_states, det_arcs);
    printf("\"decode_s\": %.6f, \"raw_lattice_s\": %.6f, \"determinize_nbest_s\": %.6f, \"nbest\": [", t1 - t0, t2 - t1,
           t3 - t2);
    for (size_t k = 0; k < paths.size(); ++k) {
      std::vector<int> w, p;
      float ptot = 0, plm = 0;
      LatticeToVector(paths[k], w, p, ptot, plm);
      printf("%s{\"tot\": %.9g, \"tot_bits\": %u, \"lm\": %.9g, ", k ? ", " : "", ptot, Bits(ptot), plm);
      PrintInts("words", w);
      printf("}");
    }
    printf("]}\n");
    fflush(stdout);
  }
  delete dec;
  return 0;
}
