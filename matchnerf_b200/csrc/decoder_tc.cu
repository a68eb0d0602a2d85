// K-mlp-composite, tcgen05 variant (placeholder until the tensor-core kernel lands; reports "unsupported").
#include "decoder_weights.cuh"
#include "mnf_common.cuh"

namespace mnf {
struct DecoderWeightsTC { int unused; };
int decoder_tc_pack(const float*, const ParamOffsets&, DecoderWeightsTC** out) { *out = nullptr; return MNF_OK; }
void decoder_tc_free(DecoderWeightsTC*) {}
bool decoder_tc_supports(const mnf_decoder_cfg&) { return false; }
int launch_decoder_tc(const DevCams&, const DevRays&, const mnf_decoder_cfg&, const DecoderWeightsTC*, const HeadParams*,
                      const __half*, int, float*, float*, float*, float*, cudaStream_t) {
  set_error("tcgen05 decoder not built");
  return MNF_EUNSUPPORTED;
}
}  // namespace mnf
