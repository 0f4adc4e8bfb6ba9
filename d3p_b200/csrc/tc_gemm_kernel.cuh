// 3xTF32 GEMM on the 5th-generation tensor cores:  D[M, N] = sum_k A[m, k] * B[n, k]  in fp32
// accuracy, with A = a_hi + a_lo and B = b_hi + b_lo pre-split by the producing kernel
// (a_hi = A with the 13 low mantissa bits cleared, exactly representable in TF32):
//     D = a_hi*b_hi + a_hi*b_lo + a_lo*b_hi        (the a_lo*b_lo term is < 2^-22 relative)
// Each term is one tcgen05.mma.kind::tf32.  An operand whose values are exact in TF32 (e.g. binarised pixels)
// passes lo = NULL and its correction MMA is skipped.
//
// TWO TMEM accumulators.  Measured (scripts/gemm_accuracy.py): the fp32 accumulation of tcgen05.mma TRUNCATES —
// with operands exact in TF32 (every product exact) a K = 4096 sum of positive terms comes out 3.0e-5 low, i.e.
// ~2^-24 per UMMA_K = 8 step, and 3 MMAs per step into one accumulator tripled that (8.8e-5).  The small cross
// terms a_hi*b_lo + a_lo*b_hi (2^-10 of the product) therefore go to a second accumulator (columns [BN, 2 BN)),
// where their truncation is 2^-10 times smaller, and the epilogue adds the two in fp32 (round to nearest): the
// main accumulator sees one addition per k-step instead of three.
//
// In-kernel split (GemmOperand::split): the operand arrives as ONE fp32 array.  kind::tf32 reads the top 19 bits of
// a 32-bit operand word, i.e. it TRUNCATES (measured: scripts/tf32_operand_rounding.py — raw x as the hi operand
// with lo = x - trunc(x) gives bit-identical products to a masked hi), so the raw tile IS the hi operand and only
// lo = x - trunc(x) has to be made: the epilogue warps, idle during the main loop, compute it for every landed stage
// (an element-wise pass over the stage, so the TMA swizzle is irrelevant) into a separate 2-slot lo ring, then
// fence.proxy.async and arrive on the slot's barrier the MMA issuer waits on.  A raw stage is half the size of a
// hi + lo stage, so the TMA ring is 3 deep instead of 2 for a 128 x 224 tile and half the operand bytes cross
// L2 -> shared memory.
//
// Structure (one CTA per 128 x BN output tile and K split, 32 EW threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes of 32 fp32 (one 128 B swizzle span) into
//               a STAGES-deep shared-memory ring, mbarrier complete_tx
//   warp 1      MMA issuer (one thread): 4 k-steps of UMMA_K = 8 per 32-deep k-block, 1-3 MMAs each;
//               tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..   main loop: split landed stages into hi / lo when an operand arrives unsplit
//   all EW warps (0 and 1 join when their loops are done): epilogue, tcgen05.ld 32 lanes x 32 columns -> registers ->
//               Epi::tile().  EW = 16 means 16 warps in the CTA, 4 per scheduler, i.e. 128 registers per thread
//               (the former 2 + 16 warps put 5 warps on two schedulers: a 96-register cap and a spilling epilogue)
// Operand major-ness (template): K-major = the contraction index is contiguous in memory,
// MN-major = the M (or N) index is contiguous (needed for A^T * diag(c) * Delta, where the
// contraction runs over the batch, the slow axis of both activations).
#pragma once
#include "tc_gemm.cuh"

namespace d3p {
namespace tc {

constexpr int kBM = 128;          // UMMA M
constexpr int kKB = 32;           // fp32 per k-block = one 128-byte swizzle span
constexpr int kUK = 8;            // UMMA K for tf32
constexpr int kGemmEpiWarps = 4;   // default number of epilogue warps (store epilogues)

struct GemmMaps { CUtensorMap a_hi, a_lo, b_hi, b_lo; };

struct GemmShape {
  uint32_t M, N, K;
  uint32_t num_k_blocks;          // ceil(K / 32)
  uint32_t split_k;               // gridDim.z
  int has_a_lo, has_b_lo;
  const int* a_lo_flag;           // optional device flag: 0 = a_lo is all zeros, skip its MMA (and its loads)
  int a_split, b_split;           // operand given as one fp32 array: hi / lo are produced in shared memory
#ifdef D3P_GEMM_TRACE
  unsigned long long* trace;      // development build only: 8 x u64 per CTA (see trace_stamp)
  uint32_t trace_id;
#endif
};

#ifdef D3P_GEMM_TRACE
// Development build (D3P_NVCC_DEFINES=D3P_GEMM_TRACE): %globaltimer stamps of one CTA's life, written to a buffer the
// host registered with d3p_dev_gemm_trace.  Never compiled into the product library.
extern unsigned long long* g_trace_buf;
extern unsigned int g_trace_next, g_trace_cap;
__device__ __forceinline__ void trace_stamp(const GemmShape& g, int slot) {
  if (!g.trace) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const size_t cta = (size_t)blockIdx.x + (size_t)gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z);
  g.trace[cta * 8 + slot] = t;
}
#define D3P_TRACE(slot) trace_stamp(g, slot)
#else
#define D3P_TRACE(slot)
#endif

template <int BN>
struct GemmCfg {
  static constexpr uint32_t kABytes = kBM * 128;
  static constexpr uint32_t kBBytes = BN * 128;
  static constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;       // with both lo parts present
  static constexpr int kMaxStages = 6;
  // ring capacity: as many stages as fit 220 KB; the kernel re-derives the stage count at run time from the
  // operands that are really present (an absent / all-zero lo part frees its slots for deeper pipelining)
  static constexpr uint32_t kRingBytes = 220u * 1024u;
  static constexpr uint32_t kSmemBytes = kRingBytes + 1024;
  // hi*hi accumulator in columns [0, BN), cross-term accumulator in [BN, 2 BN); allocations are powers of two
  static constexpr uint32_t kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32, 256]");
  static_assert(kRingBytes >= 2 * kStageBytes, "tile too large for a 2-stage pipeline");
  static_assert(kRingBytes >= 16 * 32 * 36 * 4, "the epilogue scratch (16 warps) lives in the ring");
};

// ---- coalesced row I/O for the epilogues ---------------------------------------------------------------------------
// tcgen05.ld 32x32b hands lane t the 32 columns of ROW t, so a direct store (or load) of those values touches 32
// different rows per instruction: 32 half-filled sectors, and the epilogues of the forward / backward GEMMs were
// bound by exactly that (traced: G5's epilogue 21 us of a 41 us kernel, ~1 sector per clock).  These helpers move a
// 32 x 32 chunk through the warp's private scratch (32 rows x 36 floats of the operand ring, which is idle once the
// accumulator is complete) so that every global access is a 16-byte piece of a full 128-byte row segment:
// 8 lanes x 16 B = one row, 4 rows per instruction.  Row stride 36 floats keeps both phases bank-conflict free.
constexpr int kEpiScratchLd = 36;
constexpr int kEpiScratchFloats = 32 * kEpiScratchLd;

// vals[j] = src[(lane) * ld + j] for the warp's 32 rows x 32 columns; rows >= n_rows / cols >= n_cols read as 0.
// `src` points at (first row of the warp, first column of the chunk), 16-byte aligned, ld % 4 == 0, n_cols % 4 == 0.
__device__ __forceinline__ void warp_load_rows(float* scratch, const float* __restrict__ src, size_t ld, uint32_t n_rows,
                                               uint32_t n_cols, float (&vals)[32]) {
  const int lane = threadIdx.x & 31, sub = lane >> 3, c4 = (lane & 7) * 4;
  float4 t[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t r = (uint32_t)(sub + 4 * i);
    t[i] = (r < n_rows && (uint32_t)c4 < n_cols) ? __ldg(reinterpret_cast<const float4*>(src + (size_t)r * ld + c4))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(scratch + (sub + 4 * i) * kEpiScratchLd + c4) = t[i];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 q = *reinterpret_cast<const float4*>(scratch + lane * kEpiScratchLd + 4 * j);
    vals[4 * j] = q.x; vals[4 * j + 1] = q.y; vals[4 * j + 2] = q.z; vals[4 * j + 3] = q.w;
  }
  __syncwarp();
}

// dst[(lane) * ld + j] = vals[j], same geometry; rows >= n_rows and columns >= n_cols are not written.
__device__ __forceinline__ void warp_store_rows(float* scratch, float* __restrict__ dst, size_t ld, uint32_t n_rows,
                                                uint32_t n_cols, const float (&vals)[32]) {
  const int lane = threadIdx.x & 31, sub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(scratch + lane * kEpiScratchLd + 4 * j) =
        make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t r = (uint32_t)(sub + 4 * i);
    const float4 q = *reinterpret_cast<const float4*>(scratch + r * kEpiScratchLd + c4);
    if (r < n_rows && (uint32_t)c4 < n_cols) *reinterpret_cast<float4*>(dst + (size_t)r * ld + c4) = q;
  }
  __syncwarp();
}

// EW = number of warps (multiple of 4), all of which run the epilogue: the EW / 4 warps that share a TMEM lane quarter
// take the 32-column chunks round-robin; Epi::begin/end see `part` in [0, EW / 4) for their row reductions.
template <bool A_MN, bool B_MN, int BN, class Epi, int EW = kGemmEpiWarps>
__global__ void __launch_bounds__(32 * EW, 1)
tc_gemm_kernel(const __grid_constant__ GemmMaps maps, const GemmShape g, const typename Epi::Args ea) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::kMaxStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_lo_full[2];    // in-kernel split: lo slot written by the epilogue warps
  __shared__ __align__(8) uint64_t bar_lo_empty[2];   // ... and read by the MMAs that were committed
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t m_tile = blockIdx.x, n_tile = blockIdx.y, split = blockIdx.z;
  if (threadIdx.x == 0) D3P_TRACE(0);                                  // CTA start
  // balanced k-block ranges: the first (num_k_blocks % split_k) splits get one block more
  const uint32_t base = g.num_k_blocks / g.split_k, rem = g.num_k_blocks % g.split_k;
  const uint32_t kb_begin = split * base + (split < rem ? split : rem);
  const uint32_t kb_end = kb_begin + base + (split < rem ? 1u : 0u);
  const bool has_a_lo = (g.has_a_lo || g.a_split) && (g.a_lo_flag == nullptr || __ldg(g.a_lo_flag) != 0);
  const bool has_b_lo = g.has_b_lo != 0 || g.b_split != 0;
  const bool a_conv = g.a_split && has_a_lo, b_conv = g.b_split != 0;      // lo parts made here, not loaded
  const bool any_conv = a_conv || b_conv;
  const bool load_a_lo = has_a_lo && !a_conv, load_b_lo = has_b_lo && !b_conv;     // lo parts that arrive by TMA
  // TMA ring: [A | A lo (loaded)] [B | B lo (loaded)] per stage; in-kernel split: 2 more slots of [A lo | B lo] made here
  const uint32_t stage_bytes = Cfg::kABytes * (load_a_lo ? 2u : 1u) + Cfg::kBBytes * (load_b_lo ? 2u : 1u);
  const uint32_t lo_slot_bytes = (a_conv ? Cfg::kABytes : 0u) + (b_conv ? Cfg::kBBytes : 0u);
  const uint32_t ring_avail = Cfg::kRingBytes - 2u * lo_slot_bytes;
  const uint32_t stages = ring_avail / stage_bytes > (uint32_t)STAGES ? (uint32_t)STAGES : ring_avail / stage_bytes;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a_hi);
    tma_prefetch_desc(&maps.b_hi);
    if (has_a_lo) tma_prefetch_desc(&maps.a_lo);
    if (has_b_lo) tma_prefetch_desc(&maps.b_lo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (uint32_t s = 0; s < stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
      for (uint32_t l = 0; l < 2; ++l) { mbar_init(&bar_lo_full[l], EW - 2); mbar_init(&bar_lo_empty[l], 1); }
      mbar_init(&bar_acc, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::kTmemCols>(&tmem_base_s);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) D3P_TRACE(1);                                  // barriers + TMEM ready

  auto stage_ptr = [&](uint32_t s, int which) -> uint8_t* {   // which: 0 a (hi), 1 a_lo (loaded), 2 b (hi), 3 b_lo (loaded)
    uint8_t* p = smem + (size_t)s * stage_bytes;
    if (which >= 1) p += Cfg::kABytes;
    if (which >= 2 && load_a_lo) p += Cfg::kABytes;
    if (which >= 3) p += Cfg::kBBytes;
    return p;
  };
  auto lo_ptr = [&](uint32_t l, int which) -> uint8_t* {      // in-kernel split: which 0 a_lo, 1 b_lo
    uint8_t* p = smem + (size_t)stages * stage_bytes + (size_t)l * lo_slot_bytes;
    if (which >= 1 && a_conv) p += Cfg::kABytes;
    return p;
  };

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = Cfg::kABytes * (1 + (load_a_lo ? 1 : 0)) + Cfg::kBBytes * (1 + (load_b_lo ? 1 : 0));
      uint32_t it = 0;
      for (uint32_t kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const uint32_t s = it % stages;
        const uint32_t ph = (it / stages) & 1u;
        mbar_wait(&bar_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&bar_full[s], tx);
        const int32_t k0 = (int32_t)(kb * kKB);
        const int32_t m0 = (int32_t)(m_tile * kBM), n0 = (int32_t)(n_tile * BN);
        if (!A_MN) {
          tma_load_2d(&maps.a_hi, &bar_full[s], stage_ptr(s, 0), k0, m0);
          if (load_a_lo) tma_load_2d(&maps.a_lo, &bar_full[s], stage_ptr(s, 1), k0, m0);
        } else {
#pragma unroll
          for (int c = 0; c < kBM / 32; ++c) {
            tma_load_2d(&maps.a_hi, &bar_full[s], stage_ptr(s, 0) + c * 4096, m0 + c * 32, k0);
            if (load_a_lo) tma_load_2d(&maps.a_lo, &bar_full[s], stage_ptr(s, 1) + c * 4096, m0 + c * 32, k0);
          }
        }
        if (!B_MN) {
          tma_load_2d(&maps.b_hi, &bar_full[s], stage_ptr(s, 2), k0, n0);
          if (load_b_lo) tma_load_2d(&maps.b_lo, &bar_full[s], stage_ptr(s, 3), k0, n0);
        } else {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) {
            tma_load_2d(&maps.b_hi, &bar_full[s], stage_ptr(s, 2) + c * 4096, n0 + c * 32, k0);
            if (load_b_lo) tma_load_2d(&maps.b_lo, &bar_full[s], stage_ptr(s, 3) + c * 4096, n0 + c * 32, k0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? 4096u : 16u, b_lbo = B_MN ? 4096u : 16u;
      constexpr uint32_t a_step = A_MN ? 1024u : 32u, b_step = B_MN ? 1024u : 32u;
      constexpr uint32_t a_sbo = A_MN ? 512u : 1024u, b_sbo = B_MN ? 512u : 1024u;
      constexpr uint32_t a_lt = A_MN ? kLayoutSw128Base32 : kLayoutSw128, b_lt = B_MN ? kLayoutSw128Base32 : kLayoutSw128;
      uint32_t it = 0, accumulate = 0, accumulate_x = 0;
      const uint32_t tmem_x = tmem_base + (uint32_t)BN;      // cross-term accumulator
      for (uint32_t kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const uint32_t s = it % stages;
        const uint32_t ph = (it / stages) & 1u;
        const uint32_t l = it & 1u, lph = (it >> 1) & 1u;
        mbar_wait(any_conv ? &bar_lo_full[l] : &bar_full[s], any_conv ? lph : ph);
        tc_fence_after();
        if (it == 0) D3P_TRACE(2);                                     // first stage landed (and converted)
        const uint32_t a_hi = smem_u32(stage_ptr(s, 0)), a_lo = smem_u32(a_conv ? lo_ptr(l, 0) : stage_ptr(s, 1));
        const uint32_t b_hi = smem_u32(stage_ptr(s, 2)), b_lo = smem_u32(b_conv ? lo_ptr(l, 1) : stage_ptr(s, 3));
#pragma unroll
        for (int ks = 0; ks < kKB / kUK; ++ks) {
          const uint64_t da_hi = make_smem_desc(a_hi + ks * a_step, a_lbo, a_sbo, a_lt);
          const uint64_t db_hi = make_smem_desc(b_hi + ks * b_step, b_lbo, b_sbo, b_lt);
          umma_tf32(tmem_base, da_hi, db_hi, idesc, accumulate);
          accumulate = 1;
          if (has_b_lo) {
            umma_tf32(tmem_x, da_hi, make_smem_desc(b_lo + ks * b_step, b_lbo, b_sbo, b_lt), idesc, accumulate_x);
            accumulate_x = 1;
          }
          if (has_a_lo) {
            umma_tf32(tmem_x, make_smem_desc(a_lo + ks * a_step, a_lbo, a_sbo, a_lt), db_hi, idesc, accumulate_x);
            accumulate_x = 1;
          }
        }
        umma_commit(&bar_empty[s]);     // frees the stage once these MMAs have read it
        if (any_conv) umma_commit(&bar_lo_empty[l]);
      }
      umma_commit(&bar_acc);            // accumulator complete
      D3P_TRACE(3);                                                    // last MMA issued
    }
    __syncwarp();
  } else {
    if (any_conv) {
      // main-loop duty of warps 2..: split the operands of every landed stage (see the file header)
      const uint32_t ct = (uint32_t)(warp - 2) * 32u + (uint32_t)lane, nct = (uint32_t)(EW - 2) * 32u;
      auto convert = [&](const uint8_t* raw_p, uint8_t* lo_p, uint32_t bytes) {
#pragma unroll 4
        for (uint32_t off = ct * 16u; off < bytes; off += nct * 16u) {
          const float4 v = *reinterpret_cast<const float4*>(raw_p + off);
          *reinterpret_cast<float4*>(lo_p + off) =
              make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
        }
      };
      uint32_t it = 0;
      for (uint32_t kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const uint32_t s = it % stages;
        const uint32_t ph = (it / stages) & 1u;
        const uint32_t l = it & 1u, lph = (it >> 1) & 1u;
        mbar_wait(&bar_lo_empty[l], lph ^ 1u);        // the MMAs that read this lo slot two k-blocks ago are done
        mbar_wait(&bar_full[s], ph);
        if (a_conv) convert(stage_ptr(s, 0), lo_ptr(l, 0), Cfg::kABytes);
        if (b_conv) convert(stage_ptr(s, 2), lo_ptr(l, 1), Cfg::kBBytes);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> tensor-core reads
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&bar_lo_full[l])) : "memory");
      }
    }
  }
  {
    // ---- epilogue: every warp (the producer and the MMA issuer have nothing left to do) ----------------------
    static_assert(EW % 4 == 0 && EW >= 4, "EW must be a multiple of 4");
    constexpr int PARTS = EW / 4;
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int part = warp >> 2;         // which of the PARTS warps of that quarter
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    // The scratch below reuses the operand ring.  Its last readers (the split pass of the final k-block) are ordered
    // before this point through bar_lo_full -> MMA -> tcgen05.commit -> bar_acc; the CTA barrier states the same
    // ordering in a form compute-sanitizer's racecheck can follow (once per CTA, every warp is here anyway).
    __syncthreads();
    if (warp == 2 && lane == 0) D3P_TRACE(4);                          // accumulator complete
    const uint32_t row = m_tile * kBM + q * 32 + lane;
    const uint32_t slot = n_tile * PARTS + part;     // row-reduction slot of this (tile, warp)
    typename Epi::RowState rs;
    // per-warp transpose scratch for coalesced epilogue I/O: the operand ring is idle now (every MMA has completed)
    float* scratch = reinterpret_cast<float*>(smem) + (size_t)warp * kEpiScratchFloats;
    Epi::begin(ea, g, row, slot, split, rs);
#pragma unroll 1
    for (int c = part; c < BN / 32; c += PARTS) {
      if (n_tile * BN + c * 32 >= g.N) break;             // column chunks past N (warp-uniform)
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      if (has_a_lo || has_b_lo) {                     // + the cross-term accumulator (fp32, round to nearest)
        uint32_t x[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c * 32), x);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(x[j]));
      }
      Epi::tile(ea, g, row, n_tile * BN + c * 32, split, v, rs, scratch);
    }
    Epi::end(ea, g, row, slot, split, rs);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) D3P_TRACE(5);                                  // epilogue done
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// 8-byte variant of warp_store_rows for destinations that are only 8-byte aligned (the partial rows of the clipped
// sums are P + 2 floats apart, so every other K split sits 8 bytes off a 16-byte boundary): 16 lanes x 8 B = one row,
// 2 rows per instruction.
__device__ __forceinline__ void warp_store_rows8(float* scratch, float* __restrict__ dst, size_t ld, uint32_t n_rows,
                                                 uint32_t n_cols, const float (&vals)[32]) {
  const int lane = threadIdx.x & 31, sub = lane >> 4, c2 = (lane & 15) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(scratch + lane * kEpiScratchLd + 4 * j) =
        make_float4(vals[4 * j], vals[4 * j + 1], vals[4 * j + 2], vals[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t r = (uint32_t)(sub + 2 * i);
    const float2 q = *reinterpret_cast<const float2*>(scratch + r * kEpiScratchLd + c2);
    if (r < n_rows && (uint32_t)c2 < n_cols) *reinterpret_cast<float2*>(dst + (size_t)r * ld + c2) = q;
  }
  __syncwarp();
}

// ---- plain store epilogue: out[split][m, n] (or transposed) --------------------------------------
struct EpiStore {
  struct Args {
    float* out;
    size_t ldc;            // row stride of the stored matrix (floats)
    size_t split_stride;   // floats between the outputs of consecutive K splits
    int transpose;         // 0: out[m * ldc + n]; 1: out[n * ldc + m]
  };
  struct RowState {};
  __device__ static void begin(const Args&, const GemmShape&, uint32_t, uint32_t, uint32_t, RowState&) {}
  __device__ static void end(const Args&, const GemmShape&, uint32_t, uint32_t, uint32_t, RowState&) {}
  __device__ static void tile(const Args& a, const GemmShape& g, uint32_t row, uint32_t col0, uint32_t split,
                              const uint32_t (&v)[32], RowState&, float*) {
    float* out = a.out + (size_t)split * a.split_stride;
    if (!a.transpose) {
      if (row >= g.M) return;
      float* p = out + (size_t)row * a.ldc + col0;
      if (col0 + 32 <= g.N && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(p + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                          __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < g.N) p[j] = __uint_as_float(v[j]);
      }
    } else {
      if (row >= g.M) return;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < g.N) out[(size_t)(col0 + j) * a.ldc + row] = __uint_as_float(v[j]);
    }
  }
};

// ---- host launcher ------------------------------------------------------------------------------------
struct GemmOperand {
  const float* hi;
  const float* lo;     // may be NULL (operand exact in TF32, or split == 1)
  int mn_major;        // 0: stored [rows = M or N, K] (K contiguous); 1: stored [K, M or N]
  size_t ld;           // row stride in floats (multiple of 4)
  int split = 0;       // 1: `hi` holds the full fp32 values, the kernel makes hi / lo in shared memory (lo == NULL)
};

template <bool A_MN, bool B_MN, int BN, class Epi, int EW = kGemmEpiWarps>
int32_t launch_tc_gemm(const GemmOperand& A, const GemmOperand& B, uint32_t M, uint32_t N, uint32_t K, uint32_t split_k,
                       const typename Epi::Args& ea, cudaStream_t stream, const int* a_lo_flag = nullptr) {
  using Cfg = GemmCfg<BN>;
  if (!A.hi || !B.hi || M == 0 || N == 0 || K == 0 || split_k == 0) return D3P_ERR_INVALID_ARGUMENT;
  if ((A.ld & 3) || (B.ld & 3) || (reinterpret_cast<uintptr_t>(A.hi) & 15) || (reinterpret_cast<uintptr_t>(B.hi) & 15) ||
      (A.lo && (reinterpret_cast<uintptr_t>(A.lo) & 15)) || (B.lo && (reinterpret_cast<uintptr_t>(B.lo) & 15)))
    return D3P_ERR_INVALID_ARGUMENT;
  GemmShape g;
  g.M = M; g.N = N; g.K = K;
  g.num_k_blocks = (K + kKB - 1) / kKB;
  g.split_k = split_k > g.num_k_blocks ? g.num_k_blocks : split_k;
  if ((A.split && A.lo) || (B.split && B.lo)) return D3P_ERR_INVALID_ARGUMENT;
  g.has_a_lo = A.lo ? 1 : 0;
  g.has_b_lo = B.lo ? 1 : 0;
  g.a_lo_flag = a_lo_flag;
  g.a_split = A.split ? 1 : 0;
  g.b_split = B.split ? 1 : 0;
#ifdef D3P_GEMM_TRACE
  {
    const unsigned ctas = ((M + kBM - 1) / kBM) * ((N + BN - 1) / BN) * g.split_k;
    g.trace = nullptr; g.trace_id = 0;
    if (g_trace_buf && g_trace_next + ctas + 1 <= g_trace_cap) {
      // record header: {id = launch sequence, ctas, M, N, K, BN, split_k, EW} then one 8-word record per CTA
      unsigned long long hdr[8] = {g_trace_next, ctas, M, N, K, (unsigned long long)BN, g.split_k, (unsigned long long)EW};
      cudaMemcpyAsync(g_trace_buf + (size_t)g_trace_next * 8, hdr, sizeof(hdr), cudaMemcpyHostToDevice, stream);
      g.trace = g_trace_buf + (size_t)(g_trace_next + 1) * 8;
      g_trace_next += ctas + 1;
    }
  }
#endif
  GemmMaps maps;
  bool ok = true;
  if (!A_MN) {
    ok = ok && make_map_2d(&maps.a_hi, A.hi, M, K, A.ld, kBM, false);
    ok = ok && make_map_2d(&maps.a_lo, A.lo ? A.lo : A.hi, M, K, A.ld, kBM, false);
  } else {
    ok = ok && make_map_2d(&maps.a_hi, A.hi, K, M, A.ld, kKB, true);
    ok = ok && make_map_2d(&maps.a_lo, A.lo ? A.lo : A.hi, K, M, A.ld, kKB, true);
  }
  if (!B_MN) {
    ok = ok && make_map_2d(&maps.b_hi, B.hi, N, K, B.ld, BN, false);
    ok = ok && make_map_2d(&maps.b_lo, B.lo ? B.lo : B.hi, N, K, B.ld, BN, false);
  } else {
    ok = ok && make_map_2d(&maps.b_hi, B.hi, K, N, B.ld, kKB, true);
    ok = ok && make_map_2d(&maps.b_lo, B.lo ? B.lo : B.hi, K, N, B.ld, kKB, true);
  }
  if (!ok) return D3P_ERR_CUDA;
  auto kern = tc_gemm_kernel<A_MN, B_MN, BN, Epi, EW>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes) != cudaSuccess)
    return D3P_ERR_CUDA;
  dim3 grid((M + kBM - 1) / kBM, (N + BN - 1) / BN, g.split_k);
  kern<<<grid, 32 * EW, Cfg::kSmemBytes, stream>>>(maps, g, ea);
  return cudaGetLastError() == cudaSuccess ? D3P_OK : D3P_ERR_CUDA;
}

}  // namespace tc
}  // namespace d3p
