// Internal declarations shared by api.cu, conv_generic.cu and conv_tcgen05.cu.
#pragma once
#include "common.cuh"

namespace sae {

struct GatherParams {
    int N, OH, OW;          // output pixel space, M = N*OH*OW
    int IH, IW, Cs;         // source activation
    int R, S;
    int SY, DY, OFFY, OFFX, DIV;
    int Ncol;               // output channels
    int K;                  // R*S*Cs
    int64_t M;
};

// conv_generic.cu
int conv_gather_dispatch(const float* src, const float* wmat, float* out, const GatherParams& p, const EpiParams& e,
                         cudaStream_t st);
int conv_wgrad_generic(const float* dy, const float* x, float* dw, const sae_conv_geom* g, cudaStream_t st);

// conv_tcgen05.cu
bool tc_available();
bool tc_fprop_eligible(const sae_conv_geom* g);
bool tc_dgrad_eligible(const sae_conv_geom* g);
bool tc_wgrad_eligible(const sae_conv_geom* g);
int tc_fprop(const float* x, const float* w, float* y, const sae_conv_geom* g, const EpiParams& e, cudaStream_t st);
int tc_dgrad(const float* dy, const float* wt, float* dx, const sae_conv_geom* g, const EpiParams& e, cudaStream_t st);
int tc_wgrad(const float* dy, const float* x, float* dw, const sae_conv_geom* g, cudaStream_t st);
// style-modulated convolution: per-sample filters (forward / data gradient) and the matching weight gradient
bool tc_per_sample_eligible(const sae_conv_geom* g, int dgrad);
int tc_conv_per_sample(const float* src, const float* w, float* out, const sae_conv_geom* g, int dgrad, const EpiParams& e,
                       cudaStream_t st);
bool tc_wgrad_modulated_eligible(const sae_conv_geom* g);
int tc_wgrad_modulated(const float* dy, const float* x, const float* s, const float* w_krsc, float* dw, float* ds,
                       const sae_conv_geom* g, cudaStream_t st);

}  // namespace sae
