// jax.random with Threefry keys on the device, for the predictive-sampling facade (d3p/modelling.py:39-223 runs
// numpyro models under the `seed` handler, whose sites draw jax.random.normal / bernoulli / gamma / categorical):
// random_bits / uniform / normal in jax's legacy (non-partitionable) layout - n variates from ceil(n / 2) Threefry
// calls, call c yields elements c and c + half - and gamma / loggamma with per-element keys split(key, n).
#include "common.cuh"
#include "jrandom.cuh"
#include "launch.cuh"

namespace d3p {

enum { kTfBits = 0, kTfUniform = 1, kTfNormal = 2 };

template <int kMode>
__global__ void __launch_bounds__(256) threefry_stream_kernel(uint32_t k0, uint32_t k1, void* out, size_t n, float lo,
                                                              float hi) {
  const TfKey K(k0, k1);
  const size_t half = (n + 1) / 2;
  uint32_t* o = static_cast<uint32_t*>(out);
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < half; c += (size_t)gridDim.x * blockDim.x) {
    uint32_t y0, y1;
    const size_t c1 = c + half;
    threefry2x32(K, (uint32_t)c, c1 < n ? (uint32_t)c1 : 0u, y0, y1);
    if (kMode == kTfBits) {
      o[c] = y0;
      if (c1 < n) o[c1] = y1;
    } else if (kMode == kTfUniform) {     // jax.random.uniform: max(lo, f (hi - lo) + lo)
      o[c] = __float_as_uint(bits_to_uniform(y0, lo, hi));
      if (c1 < n) o[c1] = __float_as_uint(bits_to_uniform(y1, lo, hi));
    } else {
      o[c] = __float_as_uint(bits_to_normal<false>(y0));
      if (c1 < n) o[c1] = __float_as_uint(bits_to_normal<false>(y1));
    }
  }
}

// jax.random.gamma / loggamma(key, alpha[n]): element i uses split(key, n)[i]
__global__ void __launch_bounds__(128) threefry_gamma_kernel(uint32_t k0, uint32_t k1, const float* __restrict__ alpha,
                                                             uint32_t alpha_n, uint32_t n, int log_space,
                                                             float* __restrict__ out) {
  const TfKey K(k0, k1);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const TfKey ek(tf_split_word(K, n, 2u * i), tf_split_word(K, n, 2u * i + 1u));
    const float al = alpha[alpha_n == 1 ? 0 : i];
    out[i] = log_space ? loggamma_one(ek, al) : gamma_one(ek, al);
  }
}

template <int kMode>
static int32_t launch_tf(const uint32_t* key_h, void* out_d, size_t n, float lo, float hi, void* stream) {
  if (!key_h || (!out_d && n) || n >= 0x100000000ull) return D3P_ERR_INVALID_ARGUMENT;
  if (n == 0) return D3P_OK;
  size_t grid = ((n + 1) / 2 + 255) / 256;
  if (grid > (size_t)sm_count() * 16) grid = (size_t)sm_count() * 16;
  threefry_stream_kernel<kMode><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(key_h[0], key_h[1], out_d, n, lo, hi);
  return check_launch();
}

}  // namespace d3p

using namespace d3p;

extern "C" {

int32_t d3p_threefry_random_bits(const uint32_t key_h[2], uint32_t* out_d, size_t n_words, void* stream) {
  return launch_tf<kTfBits>(key_h, out_d, n_words, 0.f, 1.f, stream);
}
int32_t d3p_threefry_uniform_f32(const uint32_t key_h[2], float lo, float hi, float* out_d, size_t n, void* stream) {
  return launch_tf<kTfUniform>(key_h, out_d, n, lo, hi, stream);
}
int32_t d3p_threefry_normal_f32(const uint32_t key_h[2], float* out_d, size_t n, void* stream) {
  return launch_tf<kTfNormal>(key_h, out_d, n, 0.f, 1.f, stream);
}
int32_t d3p_threefry_gamma_f32(const uint32_t key_h[2], const float* alpha_d, uint32_t alpha_n, uint32_t n,
                               int32_t log_space, float* out_d, void* stream) {
  if (!key_h || !alpha_d || (!out_d && n) || (alpha_n != 1 && alpha_n != n)) return D3P_ERR_INVALID_ARGUMENT;
  if (n == 0) return D3P_OK;
  unsigned grid = (n + 127) / 128;
  if (grid > (unsigned)sm_count() * 16) grid = (unsigned)sm_count() * 16;
  threefry_gamma_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(key_h[0], key_h[1], alpha_d, alpha_n, n, log_space, out_d);
  return check_launch();
}

}  // extern "C"
