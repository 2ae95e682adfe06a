// Diagnostics exported through the C ABI (developer tools, not on the render path):
//   nrf_bench_umma  -- issue-rate probe of tcgen05.mma from shared memory on every SM, with and
//                      without a concurrent TMA weight stream; answers "what bounds the MMA issuer:
//                      the tensor pipe, the smem operand reads, or the L2 -> SMEM weight stream?"
#include <cuda_runtime.h>

#include "nrf_plan.h"
#include "nrf_ptx.cuh"

namespace nrf {

// mode bit0: N = 256 per instruction (else 128);  bit1: concurrent TMA stream of `stage_bytes` per K=64 step
// bit2: two A passes per B stage (the parity pattern: a_hi, a_lo against the same weights)
__global__ void __launch_bounds__(128, 1) bench_umma_kernel(int mode, int iters, const uint8_t* __restrict__ wsrc,
                                                             uint32_t wsrc_bytes, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  // layout: A hi 16K | A lo 16K | B ring 3 x 32K | barriers
  const uint32_t a_hi = smem_u32(sm), a_lo = a_hi + 16384, ring = a_hi + 32768;
  const uint32_t bars = ring + 3 * 32768;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + 32768 + 3 * 32768 + 128);
  const int warp = threadIdx.x >> 5;
  const bool n256 = mode & 1, tma = mode & 2, two_a = mode & 4;
  const uint32_t stage_bytes = n256 ? 32768u : 16384u;
  for (uint32_t i = threadIdx.x; i < (32768 + 3 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 24 + 8 * s, 1); }
    mbar_init(bars + 48, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 1 && (threadIdx.x & 31) == 0 && tma) {
    uint32_t stage = 0, phase = 0, ofs = (blockIdx.x * 65536u) % wsrc_bytes;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bars + 24 + 8 * stage, phase ^ 1);
      mbar_arrive_expect_tx(bars + 8 * stage, stage_bytes);
      bulk_g2s(ring + stage * 32768u, wsrc + ofs, stage_bytes, bars + 8 * stage);
      ofs += stage_bytes; if (ofs + stage_bytes > wsrc_bytes) ofs = 0;
      if (++stage == 3) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 0) {
    const uint32_t n = n256 ? 256u : 128u;
    const uint32_t idesc = umma_idesc_f16(128, n);
    uint32_t stage = 0, phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (tma) { mbar_wait(bars + 8 * stage, phase); tc_fence_after_sync(); }
      const uint64_t bdesc = umma_desc_sw128(ring + stage * 32768u);
      for (int ap = 0; ap < (two_a ? 2 : 1); ++ap) {
        const uint64_t adesc = umma_desc_sw128(ap ? a_lo : a_hi);
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ss(tmem + (it & 1) * 256u, adesc + 2u * ks, bdesc + 2u * ks, idesc, (it > 1 || ks || ap) ? 1u : 0u);
      }
      umma_commit(bars + 24 + 8 * stage);
      if (++stage == 3) { stage = 0; phase ^= 1; }
    }
    umma_commit(bars + 48);
    mbar_wait(bars + 48, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// CTA-pair variant (cta_group::2, M = 256, N = 256): each CTA holds its 16 KB half of every [256 x 64]
// weight stage.  mode bit1: stream the halves through an n_slots-deep TMA ring, the odd CTA relaying its
// "landed" to the issuer's full barrier exactly as the renderer does (bit4: release.cluster arrive
// instead of the relaxed one); bit2: two A passes per stage.
// One step = one stage = 4 (or 8) K=16 instructions; floor 512 (1024) cycles.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1)
bench_umma2_kernel(int mode, int iters, int n_slots, const uint8_t* __restrict__ wsrc, uint32_t wsrc_bytes,
                   long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const uint32_t a_hi = smem_u32(sm), a_lo = a_hi + 16384, ring = a_hi + 32768;
  const uint32_t bars = ring + 4 * 16384;           // full[8] | empty[8] | done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + 32768 + 4 * 16384 + 192);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool tma = mode & 2, two_a = mode & 4, release = mode & 16, random = mode & 32, waiters = mode & 64, n128 = mode & 1;
  for (uint32_t i = threadIdx.x; i < (32768 + 4 * 16384) / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;   // fp16 1.0 pairs, or pseudo-random halves in [-2, 2)
    if (random) { uint32_t h = (i + 977u * blockIdx.x) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; v = (h & 0x83FF83FFu) | 0x3C003C00u; }
    reinterpret_cast<uint32_t*>(sm)[i] = v;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(bars + 8 * s, rank == 0 ? 2 : 1); mbar_init(bars + 64 + 8 * s, 1); }
    mbar_init(bars + 128, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc2<512>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t ns = static_cast<uint32_t>(n_slots);
  if (warp == 1 && (threadIdx.x & 31) == 0 && tma) {          // producer, both CTAs
    uint32_t stage = 0, phase = 0, ofs = ((blockIdx.x >> 1) * 65536u) % wsrc_bytes + rank * 16384u;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bars + 64 + 8 * stage, phase ^ 1);
      mbar_arrive_expect_tx(bars + 8 * stage, 16384);
      bulk_g2s(ring + stage * 16384u, wsrc + ofs, 16384, bars + 8 * stage);
      ofs += 32768; if (ofs + 32768 > wsrc_bytes) ofs = rank * 16384u;
      if (++stage == ns) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 2 && (threadIdx.x & 31) == 0 && tma && rank == 1) {   // relay
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bars + 8 * stage, phase);
      if (release) mbar_arrive_cluster(bars + 8 * stage, 0); else mbar_arrive_cluster_relaxed(bars + 8 * stage, 0);
      if (++stage == ns) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 0 && rank == 0) {                  // issuer
    const uint32_t idesc = umma_idesc_f16(256, n128 ? 128 : 256);
    uint32_t stage = 0, phase = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (tma) { mbar_wait(bars + 8 * stage, phase); tc_fence_after_sync(); }
      const uint64_t bdesc = umma_desc_sw128(ring + stage * 16384u);
      for (int ap = 0; ap < (two_a ? 2 : 1); ++ap) {
        const uint64_t adesc = umma_desc_sw128(ap ? a_lo : a_hi);
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) umma2_f16_ss(tmem + ((it >> 3) & 1) * 256u, adesc + 2u * ks, bdesc + 2u * ks, idesc, (it > 15 || ks || ap) ? 1u : 0u);
      }
      umma2_commit(bars + 64 + 8 * stage);
      if (++stage == ns) { stage = 0; phase ^= 1; }
    }
    umma2_commit(bars + 128);
    mbar_wait(bars + 128, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x >> 1] = t1 - t0;
  } else if (warp >= 4 && waiters) {
    mbar_wait(bars + 128, 0);     // 16 warps parked on an mbarrier, like the renderer's epilogue warps
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2<512>(tmem);
}

}  // namespace nrf

using namespace nrf;

extern "C" int nrf_bench_umma2(int mode, int iters, int n_slots, const void* wsrc, size_t wsrc_bytes, long long* cycles,
                               int n_pairs, void* stream) {
  if (!cycles || iters < 1 || n_pairs < 1 || n_slots < 1 || n_slots > 4) { set_error("bench_umma2: bad arguments"); return NRF_E_INVALID; }
  if ((mode & 2) && (!wsrc || wsrc_bytes < 131072)) { set_error("bench_umma2: TMA mode needs a >= 128 KiB source buffer"); return NRF_E_INVALID; }
  const int smem = 32768 + 4 * 16384 + 256;
  cudaError_t e = cudaFuncSetAttribute(bench_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  bench_umma2_kernel<<<2 * n_pairs, (mode & 64) ? 640 : 128, smem, static_cast<cudaStream_t>(stream)>>>(mode, iters, n_slots, static_cast<const uint8_t*>(wsrc),
                                                                                    static_cast<uint32_t>(wsrc_bytes), cycles);
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "bench_umma2_kernel launch");
}


extern "C" int nrf_bench_umma(int mode, int iters, const void* wsrc, size_t wsrc_bytes, long long* cycles, int n_ctas,
                              void* stream) {
  if (!cycles || iters < 1 || n_ctas < 1) { set_error("bench_umma: bad arguments"); return NRF_E_INVALID; }
  if ((mode & 2) && (!wsrc || wsrc_bytes < 65536)) { set_error("bench_umma: TMA mode needs a >= 64 KiB source buffer"); return NRF_E_INVALID; }
  const int smem = 32768 + 3 * 32768 + 256;
  cudaError_t e = cudaFuncSetAttribute(bench_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  bench_umma_kernel<<<n_ctas, 128, smem, static_cast<cudaStream_t>(stream)>>>(mode, iters, static_cast<const uint8_t*>(wsrc),
                                                                               static_cast<uint32_t>(wsrc_bytes), cycles);
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "bench_umma_kernel launch");
}
