// C ABI of the standalone 3xTF32 tcgen05 GEMM (d3p_gemm_tf32x3) and of the TF32 hi/lo splitter.
// The GEMM is the building block of the VAE path (K7): per-example gradients of a dense layer are
// a_i (x) delta_i, so the clipped sum  sum_i c_i a_i (x) delta_i = A^T diag(c) Delta  is one GEMM whose
// contraction runs over the batch (d3p/svi.py:310-348 applied to examples/vae.py:65-153).
#include "common.cuh"
#include "launch.cuh"
#include "tc_gemm_kernel.cuh"

#ifdef D3P_GEMM_TRACE
namespace d3p { namespace tc {
unsigned long long* g_trace_buf = nullptr;
unsigned int g_trace_next = 0, g_trace_cap = 0;
} }
// development build only: register a device buffer of `records` x 8 u64; returns the number of records used so far
extern "C" uint32_t d3p_dev_gemm_trace(unsigned long long* buf_d, uint32_t records) {
  const uint32_t used = d3p::tc::g_trace_next;
  d3p::tc::g_trace_buf = buf_d; d3p::tc::g_trace_cap = records; d3p::tc::g_trace_next = 0;
  return used;
}
#endif

namespace d3p {

// hi = x * scale[row] with the low 13 mantissa bits cleared, lo = x * scale[row] - hi (exact in fp32)
__global__ void split_tf32_kernel(const float* __restrict__ x, const float* __restrict__ row_scale, uint32_t cols,
                                  float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = x[i];
    if (row_scale) v *= row_scale[i / cols];
    const float h = tc::tf32_hi(v);
    hi[i] = h;
    if (lo) lo[i] = v - h;
  }
}

template <bool A_MN, bool B_MN, int EW = tc::kGemmEpiWarps>
static int32_t dispatch_bn(int bn, const tc::GemmOperand& A, const tc::GemmOperand& B, uint32_t M, uint32_t N, uint32_t K,
                           uint32_t split_k, const tc::EpiStore::Args& ea, cudaStream_t s) {
  switch (bn) {
    case 128: return tc::launch_tc_gemm<A_MN, B_MN, 128, tc::EpiStore, EW>(A, B, M, N, K, split_k, ea, s);
    case 224: return tc::launch_tc_gemm<A_MN, B_MN, 224, tc::EpiStore, EW>(A, B, M, N, K, split_k, ea, s);
    default: return D3P_ERR_UNSUPPORTED;
  }
}

}  // namespace d3p

using namespace d3p;

extern "C" int32_t d3p_split_tf32(const float* x_d, const float* row_scale_d, uint32_t cols, float* hi_d, float* lo_d,
                                  size_t n, void* stream) {
  if (!x_d || !hi_d || (row_scale_d && cols == 0)) return D3P_ERR_INVALID_ARGUMENT;
  if (n == 0) return D3P_OK;
  unsigned grid = (unsigned)((n + 255) / 256);
  unsigned cap = (unsigned)sm_count() * 16;
  if (grid > cap) grid = cap;
  split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_d, row_scale_d, cols, hi_d, lo_d, n);
  return check_launch();
}

extern "C" int32_t d3p_gemm_tf32x3(const float* a_hi_d, const float* a_lo_d, int32_t a_mn_major, size_t lda,
                                   const float* b_hi_d, const float* b_lo_d, int32_t b_mn_major, size_t ldb, uint32_t M,
                                   uint32_t N, uint32_t K, uint32_t split_k, int32_t tile_n, float* out_d, size_t ldc,
                                   size_t split_stride, int32_t transpose_out, void* stream) {
  if (!out_d) return D3P_ERR_INVALID_ARGUMENT;
  tc::GemmOperand A{a_hi_d, a_lo_d, a_mn_major, lda}, B{b_hi_d, b_lo_d, b_mn_major, ldb};
  tc::EpiStore::Args ea{out_d, ldc, split_stride, transpose_out};
  cudaStream_t s = (cudaStream_t)stream;
  if (!a_mn_major && !b_mn_major) return dispatch_bn<false, false>(tile_n, A, B, M, N, K, split_k, ea, s);
  if (!a_mn_major && b_mn_major) return dispatch_bn<false, true>(tile_n, A, B, M, N, K, split_k, ea, s);
  if (a_mn_major && !b_mn_major) return dispatch_bn<true, false>(tile_n, A, B, M, N, K, split_k, ea, s);
  return dispatch_bn<true, true>(tile_n, A, B, M, N, K, split_k, ea, s);
}

// Same product from unsplit fp32 operands: the kernel splits every stage in shared memory (tc_gemm_kernel.cuh).
extern "C" int32_t d3p_gemm_f32x3(const float* a_d, int32_t a_mn_major, size_t lda, const float* b_d, int32_t b_mn_major,
                                  size_t ldb, uint32_t M, uint32_t N, uint32_t K, uint32_t split_k, int32_t tile_n,
                                  float* out_d, size_t ldc, size_t split_stride, int32_t transpose_out, void* stream) {
  if (!out_d) return D3P_ERR_INVALID_ARGUMENT;
  tc::GemmOperand A{a_d, nullptr, a_mn_major, lda, 1}, B{b_d, nullptr, b_mn_major, ldb, 1};
  tc::EpiStore::Args ea{out_d, ldc, split_stride, transpose_out};
  cudaStream_t s = (cudaStream_t)stream;
  if (!a_mn_major && !b_mn_major) return dispatch_bn<false, false, 16>(tile_n, A, B, M, N, K, split_k, ea, s);
  if (!a_mn_major && b_mn_major) return dispatch_bn<false, true, 16>(tile_n, A, B, M, N, K, split_k, ea, s);
  if (a_mn_major && !b_mn_major) return dispatch_bn<true, false, 16>(tile_n, A, B, M, N, K, split_k, ea, s);
  return dispatch_bn<true, true, 16>(tile_n, A, B, M, N, K, split_k, ea, s);
}
