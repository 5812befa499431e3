// placeholder until the tcgen05 kernel lands
#include "conv_internal.cuh"
namespace sae {
bool tc_available() { return false; }
bool tc_fprop_eligible(const sae_conv_geom*) { return false; }
bool tc_dgrad_eligible(const sae_conv_geom*) { return false; }
bool tc_wgrad_eligible(const sae_conv_geom*) { return false; }
int tc_fprop(const float*, const float*, float*, const sae_conv_geom*, const EpiParams&, cudaStream_t) { return fail(SAE_E_UNSUPPORTED, "tcgen05 path not built"); }
int tc_dgrad(const float*, const float*, float*, const sae_conv_geom*, const EpiParams&, cudaStream_t) { return fail(SAE_E_UNSUPPORTED, "tcgen05 path not built"); }
int tc_wgrad(const float*, const float*, float*, const sae_conv_geom*, cudaStream_t) { return fail(SAE_E_UNSUPPORTED, "tcgen05 path not built"); }
}
