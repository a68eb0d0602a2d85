// K-attn, tcgen05 variant (placeholder until the tensor-core kernel lands; reports "unsupported").
#include "mnf_common.cuh"

namespace mnf {
bool window_attn_tc_supports(int, int, int, int, int) { return false; }
int launch_window_attn_tc(const float*, const float*, const float*, float*, int, int, int, int, int, int, cudaStream_t) {
  set_error("tcgen05 attention not built");
  return MNF_EUNSUPPORTED;
}
}  // namespace mnf
