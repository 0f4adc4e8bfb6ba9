// Definitions shared by the generic and the vectorised fused mean-field step kernels.
#pragma once
#include "common.cuh"

namespace d3p {

constexpr int kStepThreads = 256;
constexpr int kStepWarps = kStepThreads / 32;

struct StepArgs {
  const float* params;
  const float* x;
  size_t x_stride;
  const int32_t* y;
  const int32_t* idx;
  const uint8_t* mask;
  const int32_t* num_valid;
  uint32_t B, pos_begin, pos_end;
  uint32_t k0, k1;
  const uint32_t* key_d;      // *_dk entry points: the Threefry key lives in device memory (else nullptr)
  float inv_S, L, C, N;
  float inv_var, log_norm_lik;
  uint32_t d, n_main, half, P, loc_off, rho_off, b_loc_off, b_rho_off;
  int has_b;
  float* px_norms;
  float* px_grads;
  float* px_loss;
  float* partials;
};

template <int LINK>
D3P_D void link_terms(float rho, float inv_S, float& s, float& a, float& bt, float& log_s) {
  if (LINK == D3P_LINK_EXP) {
    s = expf(rho); a = s; bt = inv_S; log_s = rho;
  } else {
    s = softplus_f(rho); a = sigmoid_f(rho); bt = inv_S * a / s; log_s = logf(s);
  }
}

template <int G>
D3P_D unsigned group_mask(int lane) {
  return G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
}

template <int G>
D3P_D float gsum(float v, unsigned m) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
  return v;
}


}  // namespace d3p
