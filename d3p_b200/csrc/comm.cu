// Host side of the NVLink peer-memory window (see comm.cuh): allocation, CUDA-IPC exchange, teardown.
#include <new>
#include <string.h>

#include "comm.cuh"
#include "launch.cuh"

namespace d3p {

// the host-visible time-out counter: written by the waiting kernels with system-scope atomics, read here
// without touching the device
static bool timed_out(const d3p_comm* c) { return *static_cast<volatile const uint32_t*>(c->err_host) != 0u; }

int32_t comm_next(d3p_comm* c, uint32_t n_params, uint32_t n_ctas, CommDev* out) {
  if (!c || !c->connected || n_params > c->max_params || n_ctas == 0 || n_ctas > D3P_COMM_MAX_CTAS)
    return D3P_ERR_INVALID_ARGUMENT;
  if (timed_out(c)) return D3P_ERR_PEER_TIMEOUT;
  c->epoch += 1;
  out->world = c->world;
  out->rank = c->rank;
  out->epoch = c->epoch;
  out->extra_off = c->max_params;
  out->err = reinterpret_cast<uint32_t*>(c->local + c->err_off);
  out->err_host = c->err_host_dev;
  out->timeout_ns = c->timeout_ns;
  out->ll_stride = c->ll_stride;
  const size_t block = (size_t)(c->epoch & 1u) * c->world * c->ll_stride;      // u64 elements
  out->ll_local = reinterpret_cast<unsigned long long*>(c->local + c->ll_off) + block;
  for (int r = 0; r < D3P_COMM_MAX_RANKS; ++r) {
    uint8_t* base = r < c->world ? c->peer[r] : c->local;
    out->ll_peer[r] = reinterpret_cast<unsigned long long*>(base + c->ll_off) + block + (size_t)c->rank * c->ll_stride;
  }
  return D3P_OK;
}

int32_t samp_next(d3p_comm* c, uint32_t n_records, uint32_t n_tiles, SampDev* out) {
  if (!c || !c->connected || c->max_records == 0 || n_records > c->max_records) return D3P_ERR_INVALID_ARGUMENT;
  if (timed_out(c)) return D3P_ERR_PEER_TIMEOUT;
  c->samp_epoch += 1;
  const size_t n_blocks_max = ((size_t)c->max_records + 15) / 16;
  const size_t counts_off = align_up(n_blocks_max * sizeof(uint16_t), 256);
  out->world = c->world; out->rank = c->rank; out->epoch = c->samp_epoch;
  out->n_tiles = n_tiles;
  out->tiles_per_rank = (n_tiles + c->world - 1) / c->world;
  out->err = reinterpret_cast<uint32_t*>(c->local + c->err_off);
  out->err_host = c->err_host_dev;
  out->timeout_ns = c->timeout_ns;
  for (int r = 0; r < D3P_COMM_MAX_RANKS; ++r) {
    uint8_t* base = r < c->world ? c->peer[r] : c->local;
    uint8_t* buf = base + c->samp_off + (size_t)(c->samp_epoch & 1u) * c->samp_stride;
    out->masks_peer[r] = reinterpret_cast<const uint16_t*>(buf);
    out->counts_peer[r] = reinterpret_cast<uint32_t*>(buf + counts_off);
  }
  uint8_t* mine = c->local + c->samp_off + (size_t)(c->samp_epoch & 1u) * c->samp_stride;
  out->masks_local = reinterpret_cast<uint16_t*>(mine);
  out->counts_local = reinterpret_cast<const uint32_t*>(mine + counts_off);
  return D3P_OK;
}

}  // namespace d3p

extern "C" int32_t d3p_comm_create(int32_t rank, int32_t world, uint32_t max_params, uint32_t max_records,
                                   d3p_comm** comm_out, uint8_t handle_out_h[64]) {
  if (!comm_out || !handle_out_h || world < 1 || world > D3P_COMM_MAX_RANKS || rank < 0 || rank >= world)
    return D3P_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  d3p_comm* c = new (std::nothrow) d3p_comm();
  if (!c) return D3P_ERR_CUDA;
  memset(c, 0, sizeof(*c));
  c->rank = rank; c->world = world; c->max_params = max_params;
  c->err_off = 0;
  c->ll_off = 256;
  c->ll_stride = d3p::align_up((size_t)max_params + 2 * (size_t)D3P_COMM_MAX_CTAS, 32);
  c->samp_off = d3p::align_up(c->ll_off + 2 * (size_t)world * c->ll_stride * sizeof(unsigned long long), 256);
  c->max_records = max_records;
  if (max_records) {
    const size_t n_blocks = ((size_t)max_records + 15) / 16, n_tiles = (n_blocks + 255) / 256;
    c->samp_stride = d3p::align_up(n_blocks * sizeof(uint16_t), 256) + d3p::align_up(n_tiles * sizeof(uint32_t), 256);
  }
  c->total = c->samp_off + 2 * c->samp_stride;
  c->timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;      // d3p_comm_set_timeout_ms changes it
  c->sampler_margin = 16;                                   // d3p_comm_set_sampler_margin
  void* eh = nullptr;
  if (cudaHostAlloc(&eh, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { delete c; return D3P_ERR_CUDA; }
  memset(eh, 0, 64);
  c->err_host = static_cast<uint32_t*>(eh);
  void* ehd = nullptr;
  if (cudaHostGetDevicePointer(&ehd, eh, 0) != cudaSuccess) { cudaFreeHost(eh); delete c; return D3P_ERR_CUDA; }
  c->err_host_dev = static_cast<uint32_t*>(ehd);
  void* p = nullptr;
  if (cudaMalloc(&p, c->total) != cudaSuccess) { cudaFreeHost(eh); delete c; return D3P_ERR_CUDA; }
  c->local = static_cast<uint8_t*>(p);
  if (cudaMemset(p, 0, c->total) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(p); cudaFreeHost(eh); delete c; return D3P_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); cudaFreeHost(eh); delete c; return D3P_ERR_CUDA; }
  memcpy(handle_out_h, &h, 64);
  c->peer[rank] = c->local;
  c->connected = world == 1;
  *comm_out = c;
  return D3P_OK;
}

extern "C" int32_t d3p_comm_connect(d3p_comm* c, const uint8_t* handles_h) {
  if (!c || !handles_h) return D3P_ERR_INVALID_ARGUMENT;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles_h + 64 * (size_t)r, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return D3P_ERR_CUDA;
    c->peer[r] = static_cast<uint8_t*>(p);
  }
  c->connected = true;
  c->ipc = true;
  return D3P_OK;
}

// Same-process peers: one host thread (or several) driving all ranks' windows, e.g. one process that owns
// several GPUs with peer access enabled, or - as tests/test_gpu_comm_loopback.py does on a one-GPU box - several
// logical ranks on ONE device, each on its own stream.  windows_h[r] = d3p_comm_window of rank r.
extern "C" int32_t d3p_comm_window(d3p_comm* c, void** window_out_h, size_t* bytes_out_h) {
  if (!c || !window_out_h) return D3P_ERR_INVALID_ARGUMENT;
  *window_out_h = c->local;
  if (bytes_out_h) *bytes_out_h = c->total;
  return D3P_OK;
}

extern "C" int32_t d3p_comm_connect_local(d3p_comm* c, void* const* windows_h) {
  if (!c || !windows_h || c->connected) return D3P_ERR_INVALID_ARGUMENT;
  for (int r = 0; r < c->world; ++r) {
    if (!windows_h[r]) return D3P_ERR_INVALID_ARGUMENT;
    if (r != c->rank) c->peer[r] = static_cast<uint8_t*>(windows_h[r]);
  }
  c->connected = true;
  c->ipc = false;
  return D3P_OK;
}

// Number of exchanges that timed out on this rank so far (sticky: once non-zero, every d3p_* call that takes
// this window returns D3P_ERR_PEER_TIMEOUT).  Reads the host-mapped counter: no device synchronisation, so
// work still queued on a stream is not covered until the caller has synchronised that stream.
extern "C" int32_t d3p_comm_timeouts(d3p_comm* c, uint32_t* count_out_h) {
  if (!c || !count_out_h) return D3P_ERR_INVALID_ARGUMENT;
  *count_out_h = *static_cast<volatile const uint32_t*>(c->err_host);
  return D3P_OK;
}

// out_h[0] = all time-outs, [1] = waits of a finalize kernel for a peer's clipped sums, [2] = waits of the sharded
// sampler for a peer's tile counts, [3] = 0 (reserved)
extern "C" int32_t d3p_comm_timeout_detail(d3p_comm* c, uint32_t out_h[4]) {
  if (!c || !out_h) return D3P_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < 3; ++i) out_h[i] = static_cast<volatile const uint32_t*>(c->err_host)[i];
  out_h[3] = 0;
  return D3P_OK;
}

extern "C" int32_t d3p_comm_set_timeout_ms(d3p_comm* c, uint32_t timeout_ms) {
  if (!c || timeout_ms == 0) return D3P_ERR_INVALID_ARGUMENT;
  c->timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
  return D3P_OK;
}

// Sharded Poisson sampler: every rank draws `tiles` extra tiles (4096 records each) on both sides of the slice it
// owns, so that the tiles its batch positions fall into are local; all ranks must use the same value.  0 forces the
// re-draw path of the compaction kernel (tests).
extern "C" int32_t d3p_comm_set_sampler_margin(d3p_comm* c, uint32_t tiles) {
  if (!c) return D3P_ERR_INVALID_ARGUMENT;
  c->sampler_margin = tiles;
  return D3P_OK;
}

extern "C" int32_t d3p_comm_destroy(d3p_comm* c) {
  if (!c) return D3P_ERR_INVALID_ARGUMENT;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (c->ipc && r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->local) cudaFree(c->local);
  if (c->err_host) cudaFreeHost(c->err_host);
  delete c;
  return D3P_OK;
}
