// Host-side launch helpers shared by the translation units of libd3p_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/d3p_b200.h"

namespace d3p {

// Number of SMs of the current device (148 on B200); cached per device id.
inline int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

inline int32_t check_launch() { return cudaGetLastError() == cudaSuccess ? D3P_OK : D3P_ERR_CUDA; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace d3p
