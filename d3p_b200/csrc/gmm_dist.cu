// d3p.gmm.GaussianMixture.log_prob (d3p/gmm.py:71-86) as a stand-alone entry point: for every row x[E] of a batch,
//   log p(x) = logsumexp_k [ log pi_k + sum_e log N(x_e; loc_ke, scale_ke) ]
// (all event dimensions flattened into E, the K components sharing the weights).  One warp per row, the K-way
// logsumexp kept online; the fused mixture step kernel (gmm_step.cu) has its own copy of this arithmetic.
#include "common.cuh"
#include "launch.cuh"

namespace d3p {
namespace {
constexpr float kHalfLog2Pi = 0.9189385332046727f;

__global__ void __launch_bounds__(256) gmm_log_prob_kernel(const float* __restrict__ x, size_t x_stride,
                                                           const float* __restrict__ locs,
                                                           const float* __restrict__ scales,
                                                           const float* __restrict__ pis, uint32_t B, uint32_t K,
                                                           uint32_t E, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (blockDim.x >> 5) * gridDim.x;
  for (uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < B; r += warps) {
    const float* xr = x + (size_t)r * x_stride;
    float m = -INFINITY, s = 0.f;              // online logsumexp over the components
    for (uint32_t k = 0; k < K; ++k) {
      float acc = 0.f;
      for (uint32_t e = lane; e < E; e += 32) {
        const float sc = __ldg(scales + (size_t)k * E + e);
        const float z = (xr[e] - __ldg(locs + (size_t)k * E + e)) / sc;
        acc += -0.5f * z * z - logf(sc) - kHalfLog2Pi;      // numpyro Normal.log_prob
      }
      acc = group_sum<32>(acc) + logf(__ldg(pis + k));
      const float mn = fmaxf(m, acc);
      if (mn == -INFINITY) continue;                         // exp(-inf - -inf): keep the empty state
      s = s * expf(m - mn) + expf(acc - mn);
      m = mn;
    }
    if (lane == 0) out[r] = m + logf(s);
  }
}
}  // namespace
}  // namespace d3p

extern "C" int32_t d3p_gmm_log_prob_f32(const float* x_d, size_t x_row_stride, const float* locs_d, const float* scales_d,
                                        const float* pis_d, uint32_t B, uint32_t K, uint32_t E, float* out_d,
                                        void* stream) {
  using namespace d3p;
  if (K == 0 || E == 0 || ((!x_d || !out_d) && B) || !locs_d || !scales_d || !pis_d || x_row_stride < E)
    return D3P_ERR_INVALID_ARGUMENT;
  if (B == 0) return D3P_OK;
  unsigned grid = (B + 7) / 8;
  const unsigned cap = (unsigned)sm_count() * 8;
  if (grid > cap) grid = cap;
  gmm_log_prob_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_d, x_row_stride, locs_d, scales_d, pis_d, B, K, E, out_d);
  return check_launch();
}
