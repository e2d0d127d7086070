// TEST INFRASTRUCTURE — not product code.
//
// The reference's own OpenFst-const ingestion, unmodified: ConstFst<StdArc,int>::Read
// (src/newfst/const-fst.h:170-221) followed by Fst(const ConstFst&) (src/newfst/optimize-fst.h:82-134),
// dumped in the flat newfst layout (optimize-fst.h:226-280) through the Fst's public accessors so
// that the product's reader (asrd_graph_read_const / fstio.read_const_fst) can be compared with it
// byte for byte.  Compiled in place from /root/reference by oracle/Makefile.
#include <cstdio>
#include <vector>

#include "src/newfst/const-fst.h"
#include "src/newfst/optimize-fst.h"

using namespace datemoon;

int main(int argc, char **argv) {
  if (argc != 3) {
    fprintf(stderr, "usage: const2flat in.const.fst out.flat.fst\n");
    return 2;
  }
  ConstFst<StdArc, int> cfst;
  if (!cfst.Read(std::string(argv[1]))) return 3;
  Fst fst(cfst);
  FILE *fp = fopen(argv[2], "wb");
  if (!fp) return 4;
  const int S = fst.TotState(), A = fst.TotArc();
  int final_state = -1, nie = 0, noe = 0;
  for (int s = 0; s < S; ++s) {
    if (fst.IsFinal(s)) final_state = s;
    nie += (int)fst.NumInputEpsilons(s);
    noe += (int)fst.NumOutputEpsilons(s);
  }
  int hdr[6] = {fst.Start(), final_state, S, A, nie, noe};
  fwrite(hdr, 4, 6, fp);
  for (int s = 0; s < S; ++s) {
    unsigned info[3] = {fst.GetState(s)->_num_arcs, fst.GetState(s)->_niepsilons, fst.GetState(s)->_noepsilons};
    fwrite(info, 4, 3, fp);
  }
  for (int s = 0; s < S; ++s)
    if (fst.GetState(s)->_num_arcs) fwrite(fst.GetState(s)->_arcs, sizeof(StdArc), fst.GetState(s)->_num_arcs, fp);
  fclose(fp);
  return 0;
}
