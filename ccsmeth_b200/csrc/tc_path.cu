// tcgen05 tensor-core path (placeholder until the kernels land).
#include "ccsm_internal.h"

namespace ccsm {
struct TcState {};
int tc_upload_weights(ccsm_model*) {
  set_error("tensor-core path not built yet");
  return CCSM_EUNSUPPORTED;
}
void tc_release(ccsm_model*) {}
int tc_forward_att2s(ccsm_model*, int64_t, const ccsm_strand*, const ccsm_strand*, const float*, const float*, float*,
                     float*, cudaStream_t) {
  set_error("tensor-core path not built yet");
  return CCSM_EUNSUPPORTED;
}
}  // namespace ccsm
