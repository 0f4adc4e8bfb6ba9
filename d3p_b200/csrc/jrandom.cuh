// jax.random's gamma sampler on the device (Threefry keys): Marsaglia-Tsang with the alpha < 1 boost, per-element keys,
// as jax <= 0.4.10's _gamma_one lowers it ([3P]: restated from the published algorithm, mirrored by oracle/gamma.py).
// Shared by the mixture-model kernels (gmm_step.cu) and the jax.random entry points of jrandom.cu.
#pragma once
#include "common.cuh"

namespace d3p {

// jax.random.split(key, 3): words 0..5 from calls (0,3), (1,4), (2,5)
D3P_D void tf_split3(const TfKey& k, TfKey& a, TfKey& b, TfKey& c) {
  uint32_t a0, a1, b0, b1, c0, c1;
  threefry2x32(k, 0u, 3u, a0, a1);
  threefry2x32(k, 1u, 4u, b0, b1);
  threefry2x32(k, 2u, 5u, c0, c1);
  a = TfKey(a0, b0); b = TfKey(c0, a1); c = TfKey(b1, c1);
}
D3P_D float tf_scalar_bits_normal(const TfKey& k) { uint32_t y0, y1; threefry2x32(k, 0u, 0u, y0, y1); return bits_to_normal<false>(y0); }
D3P_D float tf_scalar_uniform(const TfKey& k) { uint32_t y0, y1; threefry2x32(k, 0u, 0u, y0, y1); return bits_to_unit_float(y0); }

// One pass of the outer loop of jax's _gamma_one: proposes (X, V, U) and advances the key.
D3P_D void mt_propose(TfKey& key, float c, float& X, float& V, float& U) {
  TfKey nk, x_key, U_key;
  tf_split3(key, nk, x_key, U_key);
  key = nk;
  float x = 0.f, v = -1.0f;
  while (v <= 0.f) {
    TfKey xk2, sub;
    tf_split2(x_key, xk2, sub);
    x_key = xk2;
    x = tf_scalar_bits_normal(sub);
    v = 1.0f + x * c;
  }
  X = x * x;
  V = (v * v) * v;
  U = tf_scalar_uniform(U_key);
}
// the loop continues (= the proposal is rejected) while this holds
D3P_D bool mt_reject(float X, float V, float U, float d) {
  return (U >= 1.0f - 0.0331f * (X * X)) && (logf(U) >= X * 0.5f + d * ((1.0f - V) + logf(V)));
}

// jax _gamma_one(key, alpha, log_space = true): log of a Gamma(alpha, 1) draw
D3P_D float loggamma_one(TfKey key, float alpha_orig) {
  const bool boost_mask = alpha_orig >= 1.0f;
  const float alpha = boost_mask ? alpha_orig : alpha_orig + 1.0f;
  const float d = alpha - (1.0f / 3.0f);
  const float c = (1.0f / 3.0f) / sqrtf(d);
  TfKey k2, subkey;
  tf_split2(key, k2, subkey);
  const float u_boost = tf_scalar_uniform(subkey);
  float X = 0.f, V = 1.0f, U = 2.0f;
  while (mt_reject(X, V, U, d)) mt_propose(k2, c, X, V, U);
  const float log_samples = log1pf(-u_boost);
  const float log_boost = (boost_mask || log_samples == 0.f) ? 0.f : log_samples * (1.0f / alpha_orig);
  return (logf(d) + logf(V)) + log_boost;
}


// jax _gamma_one(key, alpha, log_space = false): a Gamma(alpha, 1) draw
D3P_D float gamma_one(TfKey key, float alpha_orig) {
  const bool boost_mask = alpha_orig >= 1.0f;
  const float alpha = boost_mask ? alpha_orig : alpha_orig + 1.0f;
  const float d = alpha - (1.0f / 3.0f);
  const float c = (1.0f / 3.0f) / sqrtf(d);
  TfKey k2, subkey;
  tf_split2(key, k2, subkey);
  const float u_boost = tf_scalar_uniform(subkey);
  float X = 0.f, V = 1.0f, U = 2.0f;
  while (mt_reject(X, V, U, d)) mt_propose(k2, c, X, V, U);
  const float boost = (boost_mask || u_boost == 0.f) ? 1.0f : powf(u_boost, 1.0f / alpha_orig);
  return (d * V) * boost;
}

}  // namespace d3p
