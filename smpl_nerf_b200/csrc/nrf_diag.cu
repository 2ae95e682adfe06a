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

}  // namespace nrf

using namespace nrf;

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
