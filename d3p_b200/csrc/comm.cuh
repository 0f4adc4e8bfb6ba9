// NVLink peer-memory exchange used by the finalize kernels when a minibatch is sharded over the GPUs of
// one box.  Every rank owns one cudaMalloc'ed "window" that all peers map through
// CUDA IPC; a step's clipped sums are exchanged INSIDE the finalize kernel (one-shot all-reduce: publish
// my P + 2 partial sums, raise a flag in every peer's window, wait for the peers' flags, read their sums
// over NVLink and add them in rank order), so the multi-GPU step has no separate reduce kernel and no
// NCCL launch.  All ranks add in the same order => bit-identical replicas.
#pragma once
#include "common.cuh"

#define D3P_COMM_MAX_RANKS 8
#define D3P_COMM_MAX_CTAS 32768u

struct d3p_comm {
  int rank, world;
  uint32_t max_params, epoch;
  uint32_t max_records, samp_epoch;             // sharded Poisson sampler (0 = not provisioned)
  size_t flags_bytes, err_off, data_off, stride_floats, total;
  size_t samp_off, samp_stride;                 // 2 x [masks u16[n_blocks] | tile_counts i32[n_tiles]] by epoch parity
  uint8_t* local;                               // this rank's window
  uint8_t* peer[D3P_COMM_MAX_RANKS];            // mapped windows (peer[rank] == local)
  bool connected;
};

namespace d3p {

// Window layout (identical on every rank):
//   flags  u32 [MAX_RANKS][MAX_CTAS]   slot [r][cta] is written by rank r, read by the owner
//   err    u32 [64]                    [0] = number of spin time-outs seen by this rank's kernels
//   data   f32 [2][stride]             double-buffered by epoch parity: [0, max_params) column sums,
//                                      [max_params + 2 * cta, +2) = (count, loss) as seen by CTA `cta`
struct CommDev {
  int world, rank;
  uint32_t epoch;                 // 1, 2, 3, ... one per exchange; identical on all ranks
  uint32_t extra_off;             // = max_params
  uint32_t* flags_local;
  uint32_t* err;
  uint32_t* flags_peer[D3P_COMM_MAX_RANKS];
  float* data_peer[D3P_COMM_MAX_RANKS];   // already offset to this epoch's buffer
};

D3P_D void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
D3P_D uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
D3P_D float ld_peer(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// Called by the first warp of a CTA after every thread of the CTA has stored its values into the local
// window and the CTA (or warp, if only this warp wrote) has synchronised: the barrier orders those stores
// before the signalling lanes, whose st.release.sys is cumulative, so no per-thread system fence is needed.
// On return the peers' values of this CTA's slots are readable (after the caller's next barrier).
D3P_D void comm_signal_and_wait(const CommDev& c, uint32_t cta, int lane) {
  if (lane < c.world && lane != c.rank) st_release_sys(c.flags_peer[lane] + (size_t)c.rank * D3P_COMM_MAX_CTAS + cta, c.epoch);
  if (lane < c.world && lane != c.rank) {
    const uint32_t* f = c.flags_local + (size_t)lane * D3P_COMM_MAX_CTAS + cta;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - c.epoch) < 0) {
      if (clock64() - t0 > (3LL << 31)) {          // ~3 s: a peer died or diverged; do not hang the GPU
        atomicAdd(c.err, 1u);
        break;
      }
    }
  }
  __syncwarp();
}


// Sharded Poisson sampler (samplers.cu): rank r draws the selectors of its slice of the records and
// publishes the 16-bit selection masks and per-tile counts in its window; every rank scans all counts
// (read from the owners over NVLink) and compacts only the batch positions it owns.
//   err[8 + r]  "select done" flag written by rank r (sampler epoch);  err[32] = local CTA-done counter
struct SampDev {
  int world, rank;
  uint32_t epoch;
  uint32_t tiles_per_rank, n_tiles;
  uint32_t* flags_local;                      // [MAX_RANKS]
  uint32_t* done_counter;                     // local
  uint32_t* err;
  uint32_t* flags_peer[D3P_COMM_MAX_RANKS];
  const uint16_t* masks_peer[D3P_COMM_MAX_RANKS];     // this epoch's buffer in every window
  int32_t* counts_peer[D3P_COMM_MAX_RANKS];           // owners PUSH their tile counts into every window
  uint16_t* masks_local;
  int32_t* counts_local;
};
bool samp_next(d3p_comm* comm, uint32_t n_records, uint32_t n_tiles, SampDev* out);

// Fills the device view for the next exchange (advances the epoch); false if the shapes do not fit.
bool comm_next(d3p_comm* comm, uint32_t n_params, uint32_t n_ctas, CommDev* out);

}  // namespace d3p
