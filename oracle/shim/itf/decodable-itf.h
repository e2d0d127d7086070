// TEST INFRASTRUCTURE (oracle build only).
// Minimal stand-in for Kaldi's itf/decodable-itf.h so that the reference's
// src/itf/decodable-itf.h:55-62 (-DKALDI branch) can alias it.  The reference's
// own non-Kaldi branch is ill-formed C++ (src/itf/decodable-itf.h:105), see
// SURVEY.md Appendix C.  Nothing here is reference code.
#ifndef ASRD_ORACLE_SHIM_DECODABLE_ITF_H_
#define ASRD_ORACLE_SHIM_DECODABLE_ITF_H_
namespace kaldi {
typedef float BaseFloat;
typedef int int32;
class DecodableInterface {
 public:
  virtual BaseFloat LogLikelihood(int32 frame, int32 index) = 0;
  virtual bool IsLastFrame(int32 frame) const = 0;
  virtual int32 NumFramesReady() const { return -1; }
  virtual int32 NumIndices() const = 0;
  virtual ~DecodableInterface() {}
};
}  // namespace kaldi
#endif
