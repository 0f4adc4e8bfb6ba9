// K2 Feistel sampler, K3 Poisson select + ordered compaction, K4 masked row gather.
// Replace d3p/util.py:216-301 and d3p/minibatch.py:29-39,103-131,210,233,306.
#include "common.cuh"
#include "launch.cuh"

namespace d3p {

// ---------------------------------------------------------------------------------------------
// K2: keyed Feistel bijection with cycle walking (d3p/util.py:249-299)
// ---------------------------------------------------------------------------------------------
struct FeistelParams {
  uint32_t rc[30];
  uint32_t capacity, bits_lower, bits_upper, mask_lower, mask_upper;
};

D3P_D uint32_t feistel_rounds(const FeistelParams& p, uint32_t x) {
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    uint32_t xu = x >> p.bits_lower;
    uint32_t xl = x & p.mask_lower;
    uint32_t f = ((xu * p.rc[3 * j + 1]) >> p.bits_upper) ^ p.rc[3 * j + 2];
    uint32_t yu = (f & p.mask_lower) ^ xl;
    uint32_t yl = (xu * p.rc[3 * j]) & p.mask_upper;
    x = (yu << p.bits_upper) | yl;
  }
  return x;
}

__global__ void __launch_bounds__(256) feistel_kernel(FeistelParams p, uint32_t first_pos, uint32_t n,
                                                      int32_t* __restrict__ idx) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t x = feistel_rounds(p, first_pos + i);
    while (x >= p.capacity) x = feistel_rounds(p, x);
    idx[i] = (int32_t)x;
  }
}

// ---------------------------------------------------------------------------------------------
// K3: Poisson sampling. One thread per ChaCha block (16 records), one CTA per tile of 4096 records.
//   pass A: selectors -> 16-bit masks + per-tile counts
//   pass B: one CTA scans tile counts from the HIGH end (output order is descending index)
//   pass C: ordered scatter of selected (and, for the padding slots, unselected) indices
// ---------------------------------------------------------------------------------------------
constexpr int kPoisThreads = 256;
constexpr uint32_t kTileRecords = kPoisThreads * 16;

__global__ void __launch_bounds__(kPoisThreads) poisson_select_kernel(ChaChaState st, float q, uint32_t n_records,
                                                                      uint16_t* __restrict__ masks,
                                                                      int32_t* __restrict__ tile_counts) {
  uint32_t blk = blockIdx.x * kPoisThreads + threadIdx.x;
  uint32_t n_blocks = (n_records + 15) / 16;
  uint32_t m = 0;
  if (blk < n_blocks) {
    uint32_t ks[16];
    chacha20_block(st.w, st.w[12] + blk, ks);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      // rng_suite.uniform(key, (N,), float32) <= q   (d3p/minibatch.py:34)
      float u = fmaxf(0.0f, bits_to_unit_float(ks[i]));
      bool sel = (u <= q) && (blk * 16u + i < n_records);
      m |= (sel ? 1u : 0u) << i;
    }
    masks[blk] = (uint16_t)m;
  }
  int c = __popc(m);
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int warp_c[kPoisThreads / 32];
  if ((threadIdx.x & 31) == 0) warp_c[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kPoisThreads / 32; ++w) t += warp_c[w];
    tile_counts[blockIdx.x] = t;
  }
}

// tile_off[t] = number of selected records in tiles > t ; counts[0] = total, counts[1] = effective.
__global__ void __launch_bounds__(1024) poisson_scan_kernel(const int32_t* __restrict__ tile_counts,
                                                            int32_t* __restrict__ tile_off, uint32_t n_tiles,
                                                            uint32_t max_b, int suppress, int32_t* __restrict__ counts,
                                                            uint8_t* __restrict__ mask) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n_tiles; base += 1024) {
    uint32_t r = base + threadIdx.x;            // r-th tile counted from the high end
    int v = (r < n_tiles) ? tile_counts[n_tiles - 1 - r] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;   // exclusive
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_tot[warp] + incl - v;
    if (r < n_tiles) tile_off[n_tiles - 1 - r] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  int total = carry_s;
  int eff = suppress ? ((uint32_t)total <= max_b ? total : 0) : ((uint32_t)total < max_b ? total : (int)max_b);
  if (threadIdx.x == 0) { counts[0] = total; counts[1] = eff; }
  if (mask)
    for (uint32_t i = threadIdx.x; i < max_b; i += 1024) mask[i] = (i < (uint32_t)eff) ? 1 : 0;
}

__global__ void __launch_bounds__(kPoisThreads) poisson_scatter_kernel(const uint16_t* __restrict__ masks,
                                                                       const int32_t* __restrict__ tile_off,
                                                                       const int32_t* __restrict__ counts,
                                                                       uint32_t n_records, uint32_t max_b,
                                                                       int32_t* __restrict__ idx) {
  uint32_t blk = blockIdx.x * kPoisThreads + threadIdx.x;
  uint32_t n_blocks = (n_records + 15) / 16;
  uint32_t m = (blk < n_blocks) ? masks[blk] : 0u;
  int c = __popc(m);
  // exclusive scan over threads in DESCENDING thread order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  __shared__ int warp_tot[kPoisThreads / 32];
  if (lane == 0) warp_tot[warp] = incl;
  __syncthreads();
  int above = 0;
#pragma unroll
  for (int w = 0; w < kPoisThreads / 32; ++w)
    if (w > warp) above += warp_tot[w];
  uint32_t s = (uint32_t)(tile_off[blockIdx.x] + above + incl - c);   // selected records with a larger index
  if (blk >= n_blocks) return;
  const uint32_t total = (uint32_t)counts[0];
  const uint32_t rec_hi = blk * 16u + 15u;
  // quick reject: nothing from this chacha block can land in [0, max_b)
  bool any_sel = (m != 0) && (s < max_b);
  bool any_unsel = total < max_b && ((n_records - 1u - min(rec_hi, n_records - 1u)) - s + total < max_b + 16u);
  if (!any_sel && !any_unsel) return;
#pragma unroll
  for (int i = 15; i >= 0; --i) {
    uint32_t rec = blk * 16u + i;
    if (rec >= n_records) continue;
    if ((m >> i) & 1u) {
      if (s < max_b) idx[s] = (int32_t)rec;
      ++s;
    } else {
      uint32_t pos = total + (n_records - 1u - rec) - s;
      if (pos < max_b) idx[pos] = (int32_t)rec;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K4: masked row gather (d3p/minibatch.py:126-131): one warp per row, 16-byte vectors when aligned
// ---------------------------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const V* __restrict__ src, size_t row_vecs,
                                                          const int32_t* __restrict__ idx,
                                                          const int32_t* __restrict__ num_valid, uint32_t b,
                                                          V* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps_per_grid = gridDim.x * (blockDim.x >> 5);
  const uint32_t nv = num_valid ? (uint32_t)max(*num_valid, 0) : b;
  V zero;
  memset(&zero, 0, sizeof(V));
  for (uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < b; r += warps_per_grid) {
    V* d = dst + (size_t)r * row_vecs;
    if (r < nv) {
      const V* s = src + (size_t)idx[r] * row_vecs;
      for (size_t v = lane; v < row_vecs; v += 32) d[v] = __ldg(s + v);
    } else {
      for (size_t v = lane; v < row_vecs; v += 32) d[v] = zero;
    }
  }
}

static ChaChaState load_state(const uint32_t* s) {
  ChaChaState st;
  for (int i = 0; i < 16; ++i) st.w[i] = s[i];
  return st;
}

struct PoissonWs {
  uint16_t* masks;
  int32_t* tile_counts;
  int32_t* tile_off;
  size_t bytes;
};

static PoissonWs carve_poisson_ws(void* ws, uint32_t n_records) {
  size_t n_blocks = ((size_t)n_records + 15) / 16;
  size_t n_tiles = (n_blocks + kPoisThreads - 1) / kPoisThreads;
  size_t o_masks = 0;
  size_t o_counts = align_up(n_blocks * sizeof(uint16_t), 256);
  size_t o_off = o_counts + align_up(n_tiles * sizeof(int32_t), 256);
  PoissonWs w;
  char* base = reinterpret_cast<char*>(ws);
  w.masks = reinterpret_cast<uint16_t*>(base + o_masks);
  w.tile_counts = reinterpret_cast<int32_t*>(base + o_counts);
  w.tile_off = reinterpret_cast<int32_t*>(base + o_off);
  w.bytes = o_off + align_up(n_tiles * sizeof(int32_t), 256);
  return w;
}

}  // namespace d3p

using namespace d3p;

extern "C" {

int32_t d3p_feistel_round_constants_h(const uint32_t state_h[16], uint32_t rc_h[30]) {
  if (!state_h || !rc_h) return D3P_ERR_INVALID_ARGUMENT;
  int32_t rc = d3p_chacha_random_bits_h(state_h, 0, rc_h, 30);
  if (rc != D3P_OK) return rc;
  for (int j = 0; j < 10; ++j) rc_h[3 * j] |= 1u;
  return D3P_OK;
}

int32_t d3p_feistel_sample(const uint32_t rc_h[30], uint32_t capacity, uint32_t first_pos, uint32_t n,
                           int32_t* idx_d, void* stream) {
  if (!rc_h || capacity == 0 || (!idx_d && n)) return D3P_ERR_INVALID_ARGUMENT;
  if ((uint64_t)first_pos + n > (uint64_t)capacity) return D3P_ERR_INVALID_ARGUMENT;
  if (n == 0) return D3P_OK;
  FeistelParams p;
  for (int i = 0; i < 30; ++i) p.rc[i] = rc_h[i];
  uint32_t bits = 0;
  for (uint32_t c = capacity - 1; c; c >>= 1) ++bits;   // (capacity - 1).bit_length()
  p.capacity = capacity;
  p.bits_lower = bits >> 1;
  p.bits_upper = bits - p.bits_lower;
  p.mask_lower = (1u << p.bits_lower) - 1u;
  p.mask_upper = (1u << p.bits_upper) - 1u;
  unsigned grid = (n + 255) / 256;
  unsigned cap = (unsigned)sm_count() * 16;
  if (grid > cap) grid = cap;
  feistel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, first_pos, n, idx_d);
  return check_launch();
}

size_t d3p_poisson_workspace_bytes(uint32_t n_records) { return carve_poisson_ws(nullptr, n_records).bytes; }

int32_t d3p_poisson_sample(const uint32_t state_h[16], float q, uint32_t n_records, uint32_t max_b,
                           int32_t suppress, int32_t* idx_d, int32_t* counts_d, uint8_t* mask_d, void* ws_d,
                           size_t ws_bytes, void* stream) {
  if (!state_h || !counts_d || !ws_d || (!idx_d && max_b) || n_records == 0 || max_b > n_records)
    return D3P_ERR_INVALID_ARGUMENT;
  PoissonWs w = carve_poisson_ws(ws_d, n_records);
  if (ws_bytes < w.bytes) return D3P_ERR_WORKSPACE;
  uint32_t n_blocks = (n_records + 15) / 16;
  uint32_t n_tiles = (n_blocks + kPoisThreads - 1) / kPoisThreads;
  cudaStream_t s = (cudaStream_t)stream;
  poisson_select_kernel<<<n_tiles, kPoisThreads, 0, s>>>(load_state(state_h), q, n_records, w.masks, w.tile_counts);
  poisson_scan_kernel<<<1, 1024, 0, s>>>(w.tile_counts, w.tile_off, n_tiles, max_b, suppress, counts_d, mask_d);
  if (max_b)
    poisson_scatter_kernel<<<n_tiles, kPoisThreads, 0, s>>>(w.masks, w.tile_off, counts_d, n_records, max_b, idx_d);
  return check_launch();
}

int32_t d3p_gather_rows_masked(const void* src_d, size_t row_bytes, const int32_t* idx_d,
                               const int32_t* num_valid_d, uint32_t b, void* dst_d, void* stream) {
  if ((!src_d || !idx_d || !dst_d) && b) return D3P_ERR_INVALID_ARGUMENT;
  if (row_bytes == 0 || (row_bytes & 3)) return D3P_ERR_INVALID_ARGUMENT;
  if (b == 0) return D3P_OK;
  unsigned grid = (b + 7) / 8;
  unsigned cap = (unsigned)sm_count() * 8;
  if (grid > cap) grid = cap;
  cudaStream_t s = (cudaStream_t)stream;
  bool vec16 = (row_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(src_d) | reinterpret_cast<uintptr_t>(dst_d)) % 16 == 0);
  if (vec16)
    gather_rows_kernel<uint4><<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(src_d), row_bytes / 16, idx_d,
                                                   num_valid_d, b, reinterpret_cast<uint4*>(dst_d));
  else
    gather_rows_kernel<uint32_t><<<grid, 256, 0, s>>>(reinterpret_cast<const uint32_t*>(src_d), row_bytes / 4, idx_d,
                                                      num_valid_d, b, reinterpret_cast<uint32_t*>(dst_d));
  return check_launch();
}

}  // extern "C"
