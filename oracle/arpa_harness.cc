// TEST INFRASTRUCTURE — not product code.
//
// The reference's own ARPA -> FSA converter, unmodified: Arpa2Fsa (src/newlm/arpa2fsa.{h,cc}) with
// one thread, driven exactly like its tool (src/newlm/arpa2fsa-bin.cc:8-31), so that the product's
// converter (asrd_lm_convert_arpa) can be compared with its output byte for byte.  Compiled in
// place from /root/reference by oracle/Makefile.
#include <cstdio>
#include <string>

#include "src/newlm/arpa2fsa.h"

using namespace datemoon;

int main(int argc, char **argv) {
  if (argc != 4) {
    fprintf(stderr, "usage: arpa2fsa arpafile wordlist outputfile\n");
    return 2;
  }
  Arpa2Fsa conv(1, argv[1], argv[2]);
  if (!conv.ConvertArpa2Fsa()) return 1;
  return conv.Write(argv[3]) ? 0 : 1;
}
