// CLIP ModifiedResNet-50 frame encoder -- not built yet in this round.
#include "rn50.cuh"
#include "api_common.cuh"

namespace spm {
struct Rn50 { int unused; };
int rn50_create(Rn50**, cudaStream_t, int, const WeightGetter&) {
  set_error("the RN50 frame encoder is not implemented yet (ViT-B/16 only)");
  return 1;
}
int rn50_encode(Rn50*, cudaStream_t, const float*, int, float*) {
  set_error("the RN50 frame encoder is not implemented yet (ViT-B/16 only)");
  return 1;
}
void rn50_destroy(Rn50*) {}
}  // namespace spm
