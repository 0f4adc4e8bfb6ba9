// NVLink peer-memory exchange used when a minibatch is sharded over the GPUs of one box (SURVEY.md
// section 8e).  Every rank owns one cudaMalloc'ed "window" that all peers map through CUDA IPC.
//
// Protocol: tagged words pushed into the READER's window (the scheme NCCL calls LL).  A value travels as
// one aligned 8-byte store {payload, epoch}; 8-byte stores are single-copy atomic, so the reader simply
// polls its own memory until the tag equals the current epoch — no flag, no fence, no remote read: the
// cost of an exchange is ONE NVLink traversal.  (Measured on this pool: a fence.sys with remote stores
// outstanding costs ~10 us, which is what a publish -> fence -> flag -> pull scheme pays twice per step.)
//   * finalize kernels: every rank pushes its P + 2 partial sums to all peers and adds the G copies in
//     rank order => one-shot all-reduce inside the kernel, bit-identical on all ranks;
//   * sharded Poisson sampler: tile owners push {count, epoch} words of their slice to all peers.
// Buffers alternate with the epoch parity: a writer can be at most one exchange ahead of a reader, since
// it cannot finish exchange e + 1 before the reader has pushed its own e + 1 data, i.e. finished reading e.
#pragma once
#include "common.cuh"

#define D3P_COMM_MAX_RANKS 8
#define D3P_COMM_MAX_CTAS 32768u

struct d3p_comm {
  int rank, world;
  uint32_t max_params, epoch;
  uint32_t max_records, samp_epoch;             // sharded Poisson sampler (0 = not provisioned)
  size_t err_off, ll_off, ll_stride, total;     // ll: u64 [2][world][ll_stride]
  size_t samp_off, samp_stride;                 // 2 x [masks u16[n_blocks] | tagged tile counts u32[n_tiles]]
  uint8_t* local;                               // this rank's window
  uint32_t* err_host;                           // pinned, device-mapped mirror of the window's time-out counter
  uint32_t* err_host_dev;                       // device address of err_host
  unsigned long long timeout_ns;                // spin time-out of the waiting kernels
  uint32_t sampler_margin;                      // sharded sampler: tiles drawn redundantly on each side of the owned slice
  uint8_t* peer[D3P_COMM_MAX_RANKS];            // mapped windows (peer[rank] == local)
  bool connected;
  bool ipc;                                     // peers were mapped with cudaIpcOpenMemHandle (else: same-process pointers)
};

namespace d3p {

// Window layout (identical on every rank):
//   err   u32 [64]                      [0] = number of spin time-outs seen by this rank's kernels
//   ll    u64 [2][world][ll_stride]     slot [parity][src][j]: written by rank src, read by the owner;
//                                       j < max_params: column sums, max_params + 2 cta (+1): (count, loss)
//   samp  2 x { masks, tagged counts }  see SampDev
struct CommDev {
  int world, rank;
  uint32_t epoch;                 // 1, 2, 3, ... one per exchange; identical on all ranks
  uint32_t extra_off;             // = max_params
  uint32_t* err;                  // window word [0]: number of spin time-outs seen by this rank's kernels (sticky)
  uint32_t* err_host;             // host-mapped mirror: the host polls it without synchronising the device
  unsigned long long timeout_ns;
  unsigned long long* ll_local;                        // this epoch's [world][ll_stride] block of my window
  unsigned long long* ll_peer[D3P_COMM_MAX_RANKS];     // row `rank` of this epoch's block in every peer's window
  size_t ll_stride;
};

D3P_D void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
D3P_D uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
D3P_D uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// push one tagged float to slot j of my row in every peer's window
D3P_D void ll_push(const CommDev& c, size_t j, float v) {
  const unsigned long long w = ((unsigned long long)c.epoch << 32) | (unsigned long long)__float_as_uint(v);
  for (int r = 0; r < c.world; ++r)
    if (r != c.rank) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(c.ll_peer[r] + j), "l"(w) : "memory");
}
D3P_D unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A peer did not deliver within the time-out (it died, diverged or stalled): count it in the window (device
// flag, read by every later finalize kernel) and in the host-mapped mirror (read by the next d3p_* call).
// `which`: 1 = the clipped-sum exchange of a finalize kernel, 2 = the tile counts of the sharded sampler (the host
// mirror keeps one counter per kind at words [1], [2]: d3p_comm_timeout_detail).
// The counters live in the window (device memory: err[0] all, err[1] exchange, err[2] sampler) and are MIRRORED into
// the host-mapped words with plain system-scope stores: an atomic on mapped host memory needs PCIe atomics, which the
// platform may not offer (compute-sanitizer rejects it).  Racing mirrors may land out of order, so the host copy is
// a lower bound of the device count - and non-zero as soon as one wait has given up, which is what the host tests.
D3P_D void comm_flag_timeout(uint32_t* err, uint32_t* err_host, int which) {
  const uint32_t all = atomicAdd(err, 1u) + 1u;
  const uint32_t kind = atomicAdd(err + which, 1u) + 1u;
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(err_host), "r"(all) : "memory");
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(err_host + which), "r"(kind) : "memory");
}
// wait for rank src's slot j of this epoch (spins on LOCAL memory).  A slot whose tag never matches is NOT
// consumed: the result is NaN, which poisons this step's gradient, parameters and loss on this rank and, through
// the next exchange, on every rank; the error is sticky (comm_next refuses further exchanges).
D3P_D float ll_wait(const CommDev& c, int src, size_t j) {
  const unsigned long long* p = c.ll_local + (size_t)src * c.ll_stride + j;
  unsigned long long w;
  unsigned long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    if ((uint32_t)(w >> 32) == c.epoch) return __uint_as_float((uint32_t)w);
    if ((spins & 1023u) == 1023u) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > c.timeout_ns || ld_relaxed_sys_u32(c.err) != 0u) break;   // once flagged, nobody waits long
    }
  }
  comm_flag_timeout(c.err, c.err_host, 1);
  return __uint_as_float(0x7fc00000u);
}
// true when an earlier kernel of this rank has seen a time-out (sampler or exchange): the caller poisons its output
D3P_D bool comm_poisoned(const CommDev& c) { return ld_relaxed_sys_u32(c.err) != 0u; }

// Sharded Poisson sampler (samplers.cu): rank r draws the selectors of its slice of the records, keeps the
// 16-bit selection masks in its window and pushes one tagged word {count (low 16 bits), epoch} per tile to
// every rank (itself included); every rank then adds up all counts and compacts only the batch positions
// it owns, pulling the masks of a tile from its owner (the tagged word is stored with release semantics
// after the tile's masks, the reader fences before the pull).
struct SampDev {
  int world, rank;
  uint32_t epoch;                             // 16-bit tag = epoch & 0xffff
  uint32_t tiles_per_rank, n_tiles;
  uint32_t* err;
  uint32_t* err_host;
  unsigned long long timeout_ns;
  const uint16_t* masks_peer[D3P_COMM_MAX_RANKS];     // this epoch's buffer in every window
  uint32_t* counts_peer[D3P_COMM_MAX_RANKS];          // tagged counts: owners push into every window
  uint16_t* masks_local;
  const uint32_t* counts_local;
};
// D3P_OK, D3P_ERR_INVALID_ARGUMENT (shapes do not fit) or D3P_ERR_PEER_TIMEOUT (sticky: an earlier exchange timed out)
int32_t samp_next(d3p_comm* comm, uint32_t n_records, uint32_t n_tiles, SampDev* out);

// A count that never arrives is flagged (sticky, see ll_wait) and read as 0; the finalize kernel of the step
// sees the flag and poisons the update.
D3P_D uint32_t samp_wait_count(const SampDev& sd, uint32_t tile) {
  const uint32_t tag = sd.epoch & 0xffffu;
  uint32_t w;
  unsigned long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    w = ld_relaxed_sys_u32(sd.counts_local + tile);
    if ((w >> 16) == tag) return w & 0xffffu;
    if ((spins & 1023u) == 1023u) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > sd.timeout_ns || ld_relaxed_sys_u32(sd.err) != 0u) break;
    }
  }
  comm_flag_timeout(sd.err, sd.err_host, 2);
  return 0u;
}

// Fills the device view for the next exchange (advances the epoch).  Same return codes as samp_next.
int32_t comm_next(d3p_comm* comm, uint32_t n_params, uint32_t n_ctas, CommDev* out);

}  // namespace d3p
