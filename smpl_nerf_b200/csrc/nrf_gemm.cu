// tcgen05 GEMM kernels of the training path -- see nrf_gemm.cuh for the three products they serve.
//
//   tile_gemm_kernel   C[S, N] = A[S, K] . op(B): the weight operand (K <= 320) stays RESIDENT in shared memory, the CTA (or CTA
//                      pair: cta_group::2, M = 256, N = 256) streams sample tiles through a TMA ring and double-buffers the
//                      accumulator in TMEM, so the epilogue of tile i (bias / ReLU / ReLU' bit mask / hi-lo split -> planes through
//                      per-warp staging blocks and TMA stores) overlaps the MMAs of tile i + 1.  Roofline: HBM -- a 256 x 256
//                      layer in parity mode is 194 flop/B (measured 80-82 % of the copy bandwidth).
//   dw_gemm_kernel     dW[128, N] += A[s, m]^T B[s, n] over this CTA's share of the samples (split-K), both operands
//                      MN-major straight out of the row-major planes; partial sums to HBM, reduced by dw_reduce_kernel.
//                      Roofline: HBM (each sample row of dY and X is read once per 128-wide M half: ~3 KB per sample per
//                      layer in parity mode against 2 x 3 x 256 x 256 flop = 131 flop/B).
//
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (converged warp, elected lane), the rest = epilogue
// (TMEM lane quarter = warp % 4).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda_bf16.h>

#include "nrf_gemm.cuh"
#include "nrf_plan.h"
#include "nrf_ptx.cuh"

namespace nrf {

constexpr uint32_t kGemmSmemLimit = 232448;
constexpr int kCT = 32;                                  // accumulator columns a thread drains per 64-column unit
constexpr int kEpiWarps = 4 * (64 / kCT);                // 4 TMEM lane quarters x 2 column groups
constexpr int kTileThreads = 32 * (2 + kEpiWarps);       // TMA warp, MMA warp, epilogue warps
constexpr int kDwThreads = 192;       // TMA warp, MMA warp, 4 epilogue warps
constexpr uint32_t kATile = 16384;    // [128 x 64] fp16

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t cvt_f16x2_satfinite(float x0, float x1) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
  return r;
}

struct GemmBars { uint64_t b_full, a_full[4], a_empty[4], acc_full[2], acc_empty[2]; uint32_t tmem; uint32_t pad; };

// ------------------------------------------------------------------------------------------------ tile GEMM
// kExact: bf16 x 3 planes, six passes, per-chunk accumulators (see below).  kPair: the CTA runs as one half of a CTA PAIR
// (cluster of 2, tcgen05 cta_group::2): every MMA is M = 256 -- 128 sample rows of each CTA -- by N = n_tile (128 or 256), each CTA
// keeps only ITS half of the weight slice resident and streams only its own rows of A.  Per byte of A streamed the pair does twice
// the MMA work of a single CTA with a 128-column slice (1,536 instead of 768 tensor cycles per 32 KB chunk), which is what lets a
// 2-stage ring keep up with the L2 latency; A is also read once instead of once per 128-column slice.
template <bool kExact, bool kPair>
__global__ void __launch_bounds__(kTileThreads, 1) tile_gemm_kernel(const __grid_constant__ TileGemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int getenv_pf_dist = P.pf_dist;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t planes = P.passes == 6 ? 3u : (P.passes == 3 ? 2u : 1u);
  const int nk = P.kc[0] + (P.n_src > 1 ? P.kc[1] : 0);
  const int n_cta = kPair ? P.n_tile / 2 : P.n_tile;                    // weight rows (output features) staged by THIS CTA
  const uint32_t b_chunk = static_cast<uint32_t>(n_cta) * 128u;
  const uint32_t b_bytes = P.b_stream ? 0u : static_cast<uint32_t>(nk) * planes * b_chunk;
  const uint32_t a_stage = planes * kATile + (P.b_stream ? planes * b_chunk : 0u);      // streaming mode: [A planes | B planes] per stage
  const uint32_t smem_b = smem_u32(smem), smem_a = smem_b + b_bytes;
  const uint32_t out_bytes = P.staged ? 32768u : 0u;                    // staging: [128 x 64] fp16 hi block | lo block of one 64-column unit
  const uint32_t smem_o = smem_a + static_cast<uint32_t>(P.n_stages) * a_stage;
  GemmBars* bars = reinterpret_cast<GemmBars*>(smem + b_bytes + static_cast<uint32_t>(P.n_stages) * a_stage + out_bytes);
  const int n0 = blockIdx.y * P.n_tile;
  const int rows_tile = kPair ? 256 : 128;
  const int64_t n_mt = (P.S + rows_tile - 1) / rows_tile;
  const int64_t mt0 = kPair ? (blockIdx.x >> 1) : blockIdx.x, mt_step = kPair ? (gridDim.x >> 1) : gridDim.x;
  // tile order: all CTAs sweep the samples together, front to back or (P.reverse) back to front -- a kernel that starts where the previous one
  // ended finds the last ~L2-sized part of that kernel's output (or input) still in L2 instead of re-reading it from HBM
  auto tile_of = [&](int64_t m) { return P.reverse ? n_mt - 1 - m : m; };

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->b_full), 1);
    for (int s = 0; s < 4; ++s) { mbar_init(smem_u32(&bars->a_full[s]), 1); mbar_init(smem_u32(&bars->a_empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&bars->acc_full[b]), 1); mbar_init(smem_u32(&bars->acc_empty[b]), kExact ? 8 : (kPair ? 2 * kEpiWarps : kEpiWarps)); }
    fence_mbar_init();
  }
  if (warp == 1) { if (kPair) tmem_alloc2<512>(smem_u32(&bars->tmem)); else tmem_alloc<512>(smem_u32(&bars->tmem)); }
  tc_fence_before_sync();
  if (kPair) cluster_sync_all(); else __syncthreads();       // pair: both CTAs' barriers exist before any remote arrival
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem;

  if (warp == 0) {
    if (lane == 0) {
      // pair: both CTAs' bytes are credited to the barrier of the ISSUING (rank 0) CTA, whose producer announces the total
      auto bar_of = [&](uint64_t* b) { return kPair ? mapa_shared(smem_u32(b), 0) : smem_u32(b); };
      auto load = [&](uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
        if (kPair) tma2_load_2d(dst, map, c0, c1, bar); else tma_load_2d(dst, map, c0, c1, bar);
      };
      const int nb0 = n0 + static_cast<int>(rank) * n_cta;             // first output feature of this CTA's weight rows
      auto load_b = [&](int j, int lc, uint32_t dst0, uint32_t bar) {
        for (uint32_t p = 0; p < planes; ++p) {
          const CUtensorMap* map = &P.b_map[j][p];
          const uint32_t dst = dst0 + p * b_chunk;
          if (!P.b_mn) load(dst, map, 64 * lc, nb0, bar);
          else for (int g = 0; g < n_cta / 64; ++g) load(dst + g * 8192u, map, nb0 + 64 * g, 64 * lc, bar);
        }
      };
      const uint32_t mult = kPair ? 2u : 1u;
      // ---- resident weight slice
      if (!P.b_stream) {
        if (rank == 0) mbar_arrive_expect_tx(smem_u32(&bars->b_full), mult * b_bytes);
        int c = 0;
        for (int j = 0; j < P.n_src; ++j)
          for (int lc = 0; lc < P.kc[j]; ++lc, ++c) load_b(j, lc, smem_b + static_cast<uint32_t>(c) * planes * b_chunk, bar_of(&bars->b_full));
      }
      // ---- sample tiles
      uint32_t stage = 0, phase = 0;
      for (int64_t mt = mt0; mt < n_mt; mt += mt_step)
        for (int j = 0; j < P.n_src; ++j)
          for (int lc = 0; lc < P.kc[j]; ++lc) {
            mbar_wait(smem_u32(&bars->a_empty[stage]), phase ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(smem_u32(&bars->a_full[stage]), mult * a_stage);
            const uint32_t dst = smem_a + stage * a_stage;
            const int32_t r0 = static_cast<int32_t>(tile_of(mt) * rows_tile + rank * 128);
            for (uint32_t p = 0; p < planes; ++p) load(dst + p * kATile, &P.a_map[j][p], 64 * lc, r0, bar_of(&bars->a_full[stage]));
            // the activation planes stream from HBM (hundreds of MB per layer): pull this CTA's NEXT tile into L2 now, so the 2-3 deep
            // ring is refilled at L2 latency (two tiles ahead was measured worse: 290 MB of a 402 MB operand were evicted by the
            // output stream before use and read twice)
            const int64_t pf = mt + (getenv_pf_dist > 0 ? getenv_pf_dist : 1) * mt_step;
            if (pf < n_mt) for (uint32_t p = 0; p < planes; ++p) tma_prefetch_2d(&P.a_map[j][p], 64 * lc, static_cast<int32_t>(tile_of(pf) * rows_tile + rank * 128));
            if (P.b_stream) load_b(j, lc, dst + planes * kATile, bar_of(&bars->a_full[stage]));
            if (++stage == static_cast<uint32_t>(P.n_stages)) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    if (rank == 0) {      // (the issuer warp of the odd CTA of a pair idles)
    const uint32_t idesc = umma_idesc_f16_major(kPair ? 256u : 128u, static_cast<uint32_t>(P.n_tile), 0u, P.b_mn ? 1u : 0u) | (kExact ? ((1u << 7) | (1u << 10)) : 0u);
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc_flag) {
      if (kPair) umma2_f16_ss_warp(d, ad, bd, idesc, acc_flag); else umma_f16_ss_warp(d, ad, bd, idesc, acc_flag);
    };
    auto commit = [&](uint64_t* b) { if (kPair) umma2_commit_warp(smem_u32(b)); else umma_commit_warp(smem_u32(b)); };
    if (!P.b_stream) mbar_wait(smem_u32(&bars->b_full), 0);
    tc_fence_after_sync();
    uint32_t stage = 0, phase = 0, it = 0;      // it: accumulator hand-offs so far (one per tile; one per K-chunk in the exact mode)
    for (int64_t mt = mt0; mt < n_mt; mt += mt_step) {
      uint32_t buf = it & 1u, acc = 0, accumulate = 0;
      if (!kExact) {
        mbar_wait(smem_u32(&bars->acc_empty[buf]), ((it >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        acc = tmem + buf * static_cast<uint32_t>(P.n_tile);
      }
      for (int c = 0; c < nk; ++c) {
        if (kExact) {
          // exact mode: the tensor core ROUNDS TOWARD ZERO when it adds into an fp32 accumulator (measured: ~2^-24.5 relative per
          // accumulating instruction, tests/test_gpu_train.py::test_gemm_exact_mode), so long chains drift.  Every K-chunk gets
          // fresh accumulators -- `main` for the 4 hi*hi instructions, `corr` for the 20 small correction products -- and the
          // epilogue warps add the chunks up in registers with round-to-nearest fp32 adds.
          buf = it & 1u;
          mbar_wait(smem_u32(&bars->acc_empty[buf]), ((it >> 1) & 1u) ^ 1u);
          tc_fence_after_sync();
          acc = tmem + buf * 2u * static_cast<uint32_t>(P.n_tile);
        }
        mbar_wait(smem_u32(&bars->a_full[stage]), phase);
        tc_fence_after_sync();
        const uint32_t a_hi = smem_a + stage * a_stage;
        const uint32_t b_hi = P.b_stream ? a_hi + planes * kATile : smem_b + static_cast<uint32_t>(c) * planes * b_chunk;
        for (uint32_t pass = 0; pass < static_cast<uint32_t>(P.passes); ++pass) {
          // (A plane, B plane): hi*hi, lo*hi, hi*lo [, ll*hi, lo*lo, hi*ll]: every product of combined order <= planes - 1
          const uint32_t pa = (0x012010u >> (4 * pass)) & 0xFu, pb = (0x210100u >> (4 * pass)) & 0xFu;
          const uint32_t a = a_hi + pa * kATile, b = b_hi + pb * b_chunk;
          const uint64_t ad = umma_desc_sw128(a);
          const uint32_t dst = (kExact && pass > 0) ? acc + static_cast<uint32_t>(P.n_tile) : acc;
          if (kExact && pass <= 1) accumulate = 0;       // first instruction into main (pass 0) / corr (pass 1)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks) {
            const uint64_t bd = P.b_mn ? umma_desc_mn_sw128(b + ks * 2048u, 8192u, 1024u) : umma_desc_sw128(b) + 2u * ks;
            mma(dst, ad + 2u * ks, bd, accumulate);
            accumulate = 1;
          }
        }
        commit(&bars->a_empty[stage]);
        if (++stage == static_cast<uint32_t>(P.n_stages)) { stage = 0; phase ^= 1; }
        if (kExact) { commit(&bars->acc_full[buf]); ++it; }
      }
      if (!kExact) { commit(&bars->acc_full[buf]); ++it; }
    }
    }
  } else if (!kExact) {
    // ------------------------------------------------------------ epilogue warps (parity / fast modes)
    // The accumulator is drained in UNITS of 64 columns: warp = 32 rows (TMEM lane quarter q) x kCT columns (cg), so a thread holds only
    // kCT values at a time.  Every warp owns a PRIVATE 4 KB staging block ([32 x 32] hi | lo, 64-byte rows, SWIZZLE_64B; or one
    // [32 x 32] fp32 block, SWIZZLE_128B) and issues its own TMA stores: lane 0 waits for the warp's previous stores to have read the
    // block just before the next unit overwrites it (~a unit later, so the wait is free) and NO barrier joins the epilogue warps.
    // (CTA-wide staging with two named barriers per unit kept all warps in lock step -- TMEM read, arithmetic and st.shared phases
    // queued on the same unit one after the other instead of overlapping; a unit took 2.3k cycles whatever the instruction count.)
    const int q = warp & 3, cg = (warp - 2) >> 2;
    const int n_units = P.n_tile / 64;
    const float s_out = P.sc_out ? __ldg(P.sc_out) : 1.f, inv_out = P.sc_out ? __ldg(P.sc_out + 1) : 1.f;
    const float ratio = (P.epi == GEPI_F32 ? 1.f : s_out) * (P.sc_in ? __ldg(P.sc_in + 1) : 1.f);     // stored-in -> stored-out (or real)
    const bool rescale = P.sc_in != nullptr || P.sc_out != nullptr;
    float l1_run = 0.f;
    bool saturated = false;
    uint32_t it = 0;
    const bool e0 = warp == 2 && lane == 0;
    const int r_t = 32 * q + lane;                        // row inside this CTA's 128 rows of the tile
    int tr_n = 0;
#ifdef NRF_GEMM_TRACE_BUILD      // developer timeline (build with -DNRF_GEMM_TRACE_BUILD, run with NRF_GEMM_TRACE=n): ~70 issued instructions per unit
    auto TR = [&](int ev) { if (P.trace && blockIdx.x == 0 && blockIdx.y == 0 && e0 && tr_n < 600) { P.trace[2 * tr_n] = ev; P.trace[2 * tr_n + 1] = clock64(); ++tr_n; } };
#else
    auto TR = [&](int) {};
    (void)e0; (void)tr_n;
#endif
    // bias and ReLU' mask of a unit are global loads with an L2 round trip: they are issued ONE UNIT AHEAD (for the first unit of a
    // tile: during the last unit of the previous tile) into the registers the current unit has just finished with, so the round trip
    // hides behind the conversion + staging of the current unit instead of stalling every unit (profiles/r2 timeline).
    float bias[kCT];
    uint32_t mk = 0u;                                     // ReLU' bit mask of the unit's kCT = 32 columns (one word per thread)
#pragma unroll
    for (int i = 0; i < kCT; ++i) bias[i] = 0.f;
    auto aux_load = [&](int64_t mt_l, int u_l) {
      const int64_t row_l = tile_of(mt_l) * rows_tile + rank * 128 + r_t;
      const bool ok = row_l < P.S;
      const int c0 = n0 + 64 * u_l + kCT * cg;
      if (P.bias) {
        const float4* bp = reinterpret_cast<const float4*>(P.bias + (P.bias_ld ? (ok ? row_l / P.rows_per_ray : 0) * P.bias_ld : 0) + c0);
#pragma unroll
        for (int i = 0; i < kCT / 4; ++i) { const float4 b = __ldg(bp + i); bias[4 * i] = b.x; bias[4 * i + 1] = b.y; bias[4 * i + 2] = b.z; bias[4 * i + 3] = b.w; }
      }
      if (P.mask_bits && ok) mk = __ldg(P.mask_bits + row_l * P.bits_ld + (c0 >> 5));
    };
    if (mt0 < n_mt) aux_load(mt0, 0);
    for (int64_t mt = mt0; mt < n_mt; mt += mt_step, ++it) {
      const uint32_t buf = it & 1u;
      TR(0);
      const int64_t row = tile_of(mt) * rows_tile + rank * 128 + r_t;
      const bool row_ok = row < P.S;
      const float rs = (P.row_scale && row_ok) ? P.row_scale[row * P.row_scale_ld] * s_out : 0.f;
      float l1 = 0.f;
      for (int u = 0; u < n_units; ++u) {
        const int col0 = n0 + 64 * u + kCT * cg;          // this thread's kCT columns of the unit
        if (u == 0) {
          mbar_wait(smem_u32(&bars->acc_full[buf]), (it >> 1) & 1u);
          tc_fence_after_sync();
          TR(1);
        }
        const uint32_t taddr = tmem + buf * static_cast<uint32_t>(P.n_tile) + (static_cast<uint32_t>(32 * q) << 16) + static_cast<uint32_t>(64 * u + kCT * cg);
        uint32_t v[kCT];
#pragma unroll
        for (int g = 0; g < kCT / 16; ++g) tmem_ld16(taddr + 16u * g, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * g]));
        tmem_ld_wait();
        TR(5);
        if (u == n_units - 1) {                              // the accumulator has been read completely: hand it back to the issuer
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) { if (kPair && rank != 0) mbar_arrive_cluster_relaxed(smem_u32(&bars->acc_empty[buf]), 0); else mbar_arrive(smem_u32(&bars->acc_empty[buf])); }
        }
        float x[kCT];
#pragma unroll
        for (int i = 0; i < kCT; ++i) x[i] = __uint_as_float(v[i]);
        if (rescale) {
#pragma unroll
          for (int i = 0; i < kCT; ++i) x[i] *= ratio;
        }
#pragma unroll
        for (int i = 0; i < kCT; ++i) x[i] += bias[i];
        if (P.row_scale) {
#pragma unroll
          for (int i = 0; i < kCT; ++i) x[i] = fmaf(rs, __ldg(P.col_vec + col0 + i), x[i]);
        }
        if (P.relu) {
#pragma unroll
          for (int i = 0; i < kCT; ++i) x[i] = fmaxf(x[i], 0.f);
        }
        if (P.mask_bits && row_ok) {
#pragma unroll
          for (int i = 0; i < kCT; ++i) x[i] = (mk >> i) & 1u ? x[i] : 0.f;
        }
        if (P.bits_out && row_ok) {
          // ReLU' mask for the backward: bit i <=> the fp16 hi half of column i is non-zero (the value the next layer's GEMM sees as
          // "active"); round-to-nearest-even sends exactly the values <= 2^-25 to zero.  32 x fewer bytes for dX to read than the hi plane.
          uint32_t wd = 0u;
#pragma unroll
          for (int i = 0; i < kCT; ++i) wd |= x[i] > 2.98023223876953125e-8f ? (1u << i) : 0u;
          P.bits_out[row * P.bits_ld + (col0 >> 5)] = wd;
        }
        TR(6);
        if (u + 1 < n_units) aux_load(mt, u + 1);          // the next unit's bias / mask (see above)
        else if (mt + mt_step < n_mt) aux_load(mt + mt_step, 0);
        if (P.l1max && row_ok) {
#pragma unroll
          for (int i = 0; i < kCT; ++i) l1 += fabsf(x[i]);
        }
        if (P.epi == GEPI_F32) {
          if (row_ok) {
            float4* op = reinterpret_cast<float4*>(P.out_f32 + row * P.out_f32_ld + col0);
#pragma unroll
            for (int i = 0; i < kCT / 4; ++i) {
              float4 o = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
              if (P.accumulate) { const float4 t = op[i]; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
              op[i] = o;
            }
          }
          continue;
        }
        const uint32_t st_w = smem_o + static_cast<uint32_t>(warp - 2) * 4096u;           // this warp's staging block
        const int32_t r0w = static_cast<int32_t>(tile_of(mt) * rows_tile + rank * 128 + 32 * q);    // first row of the warp's 32
        if (P.f32_staged) {                                // fp32 copy of the unit: [32 rows x 32 floats], 128-byte rows, SWIZZLE_128B
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          const uint32_t rowf = st_w + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowf + ((static_cast<uint32_t>(i) ^ (lane & 7u)) << 4)), "f"(x[4 * i]), "f"(x[4 * i + 1]),
                         "f"(x[4 * i + 2]), "f"(x[4 * i + 3]) : "memory");
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&P.f_map, 2 * col0, r0w, st_w); tma_store_commit(); }
        }
        uint32_t h[kCT / 2], l[kCT / 2];
        __half2 amax2 = __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kCT / 2; ++i) {
          h[i] = cvt_f16x2_satfinite(x[2 * i], x[2 * i + 1]);
          const __half2 hh = *reinterpret_cast<const __half2*>(&h[i]);
          amax2 = __hmax2(amax2, __habs2(hh));
          const float2 f = __half22float2(hh);
          l[i] = cvt_f16x2_satfinite(x[2 * i] - f.x, x[2 * i + 1] - f.y);
        }
        {
          const uint32_t am = *reinterpret_cast<const uint32_t*>(&amax2);
          saturated |= (am & 0xFFFFu) >= 0x7BFFu || (am >> 16) >= 0x7BFFu;
        }
        if (P.out_f32 && !P.f32_staged && row_ok) {
          float4* op = reinterpret_cast<float4*>(P.out_f32 + row * P.out_f32_ld + col0);
#pragma unroll
          for (int i = 0; i < kCT / 4; ++i) op[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        }
        if (!P.staged) {
          if (row_ok) {
            uint4* ph = reinterpret_cast<uint4*>(P.out_hi + row * P.out_ld + col0);
#pragma unroll
            for (int i = 0; i < kCT / 8; ++i) ph[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
            if (P.out_lo) {
              uint4* pl = reinterpret_cast<uint4*>(P.out_lo + row * P.out_ld + col0);
#pragma unroll
              for (int i = 0; i < kCT / 8; ++i) pl[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
            }
          }
          continue;
        }
        TR(2);
        if (lane == 0) tma_store_wait_read();              // the warp's previous stores have read its staging block ...
        __syncwarp();
        TR(3);
        const uint32_t rowb = st_w + static_cast<uint32_t>(lane) * 64u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {                      // ... this unit takes its place: 64-byte rows, 16-byte chunks swizzled (SWIZZLE_64B)
          const uint32_t ofs = (static_cast<uint32_t>(i) ^ ((lane >> 1) & 3u)) << 4;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ofs), "r"(h[4 * i]), "r"(h[4 * i + 1]), "r"(h[4 * i + 2]), "r"(h[4 * i + 3]) : "memory");
          if (P.out_lo)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + 2048u + ofs), "r"(l[4 * i]), "r"(l[4 * i + 1]), "r"(l[4 * i + 2]), "r"(l[4 * i + 3]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&P.o_map[0], col0, r0w, st_w);
          if (P.out_lo) tma_store_2d(&P.o_map[1], col0, r0w, st_w + 2048u);
          tma_store_commit();
        }
        TR(4);
      }
      l1_run = fmaxf(l1_run, l1);
    }
    if (P.l1max) {
      // real units; a row's L1 norm is bounded by (number of column segments it is split into) x the largest segment sum
      l1_run *= (P.epi == GEPI_F32 ? 1.f : inv_out) * static_cast<float>(gridDim.y * (64 / kCT));       // 64 / kCT column groups (warp sets) per slice
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) l1_run = fmaxf(l1_run, __shfl_xor_sync(0xffffffffu, l1_run, o));
      if (lane == 0 && isfinite(l1_run)) atomicMax(P.l1max, __float_as_uint(l1_run));
    }
    if (saturated && P.status) atomicOr(P.status, 2);
    if (P.staged && lane == 0) tma_store_wait_all();
  } else {
    // ------------------------------------------------------------ epilogue warps, exact mode (inference only: single CTA, one slice of
    // <= 128 columns, direct stores; no ReLU' mask, no gradient scaling).  The tensor core ROUNDS TOWARD ZERO when it adds into an fp32
    // accumulator, so every K-chunk arrives in fresh `main` / `corr` accumulators and is added here with round-to-nearest fp32 adds,
    // starting from the bias.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int cols_w = P.n_tile / 2;                      // 64 or 32 columns per thread
    const int r_t = 32 * q + lane;
    uint32_t it = 0;
    for (int64_t mt = mt0; mt < n_mt; mt += mt_step) {
      const int64_t row = tile_of(mt) * rows_tile + r_t;
      const bool row_ok = row < P.S;
      const int col0 = n0 + half * cols_w;
      const float* bias_row = (P.bias && row_ok) ? P.bias + (P.bias_ld ? (row / P.rows_per_ray) * P.bias_ld : 0) + col0 : nullptr;
      float sum[64];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 b = (bias_row && 4 * i < cols_w) ? __ldg(reinterpret_cast<const float4*>(bias_row) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        sum[4 * i] = b.x; sum[4 * i + 1] = b.y; sum[4 * i + 2] = b.z; sum[4 * i + 3] = b.w;
      }
      for (int c = 0; c < nk; ++c, ++it) {
        const uint32_t buf = it & 1u;
        mbar_wait(smem_u32(&bars->acc_full[buf]), (it >> 1) & 1u);
        tc_fence_after_sync();
        const uint32_t ta = tmem + buf * 2u * static_cast<uint32_t>(P.n_tile) + (static_cast<uint32_t>(32 * q) << 16) + static_cast<uint32_t>(half * cols_w);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (16 * g < cols_w) {
            uint32_t vm[16], vc[16];
            tmem_ld16(ta + 16u * g, vm);
            tmem_ld16(ta + static_cast<uint32_t>(P.n_tile) + 16u * g, vc);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) sum[16 * g + i] = __fadd_rn(sum[16 * g + i], __fadd_rn(__uint_as_float(vm[i]), __uint_as_float(vc[i])));
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[buf]));
      }
      if (!row_ok) continue;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (16 * g >= cols_w) continue;
        const int col = col0 + 16 * g;
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = P.relu ? fmaxf(sum[16 * g + i], 0.f) : sum[16 * g + i];
        if (P.epi == GEPI_F32) {
          float4* op = reinterpret_cast<float4*>(P.out_f32 + row * P.out_f32_ld + col);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 o = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
            if (P.accumulate) { const float4 t = op[i]; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
            op[i] = o;
          }
          continue;
        }
        uint32_t h[8], l[8], ll[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {      // x = hi + lo + ll exactly (3 x 8 significant bits), fp32 exponent range
          float r0 = x[2 * i], r1 = x[2 * i + 1];
          const __nv_bfloat162 b0 = __floats2bfloat162_rn(r0, r1);
          r0 -= __low2float(b0); r1 -= __high2float(b0);
          const __nv_bfloat162 b1 = __floats2bfloat162_rn(r0, r1);
          r0 -= __low2float(b1); r1 -= __high2float(b1);
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(r0, r1);
          h[i] = *reinterpret_cast<const uint32_t*>(&b0); l[i] = *reinterpret_cast<const uint32_t*>(&b1); ll[i] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        uint4* ph = reinterpret_cast<uint4*>(P.out_hi + row * P.out_ld + col);
        ph[0] = make_uint4(h[0], h[1], h[2], h[3]); ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
        if (P.out_lo) {
          uint4* pl = reinterpret_cast<uint4*>(P.out_lo + row * P.out_ld + col);
          pl[0] = make_uint4(l[0], l[1], l[2], l[3]); pl[1] = make_uint4(l[4], l[5], l[6], l[7]);
        }
        if (P.out_ll) {
          uint4* pq = reinterpret_cast<uint4*>(P.out_ll + row * P.out_ld + col);
          pq[0] = make_uint4(ll[0], ll[1], ll[2], ll[3]); pq[1] = make_uint4(ll[4], ll[5], ll[6], ll[7]);
        }
        if (P.out_f32) {
          float4* op = reinterpret_cast<float4*>(P.out_f32 + row * P.out_f32_ld + col);
#pragma unroll
          for (int i = 0; i < 4; ++i) op[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        }
      }
    }
  }
  tc_fence_before_sync();
  if (kPair) cluster_sync_all(); else __syncthreads();       // pair: neither CTA may exit (or free TMEM) while its peer can still address it
  if (warp == 1) { if (kPair) tmem_dealloc2<512>(tmem); else tmem_dealloc<512>(tmem); }
}

// ------------------------------------------------------------------------------------------------ dW GEMM (split-K over samples)
__global__ void __launch_bounds__(kDwThreads, 1) dw_gemm_kernel(const __grid_constant__ DwGemmParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t planes = P.passes == 3 ? 2u : 1u;
  const uint32_t a_bytes = 16384u, b_bytes = static_cast<uint32_t>(P.N) * 128u;       // one plane of one 64-sample stage
  const uint32_t stage_bytes = planes * (a_bytes + b_bytes);
  const uint32_t ones_bytes = P.colsum_partial ? 8192u : 0u;                            // [64 samples x 128 B] of fp16 1.0
  const uint32_t smem_ones = smem_u32(smem) + static_cast<uint32_t>(P.n_stages) * stage_bytes;
  GemmBars* bars = reinterpret_cast<GemmBars*>(smem + static_cast<uint32_t>(P.n_stages) * stage_bytes + ones_bytes);
  const int64_t n_chunks = (P.S + 63) / 64;
  const int m_tile = P.m0 + 128 * blockIdx.x;
  if (P.colsum_partial) {      // every 16-byte chunk is the same, so the 128-byte swizzle leaves the tile unchanged
    for (uint32_t i = threadIdx.x; i < 8192u / 16u; i += blockDim.x)
      *reinterpret_cast<uint4*>(smem + static_cast<uint32_t>(P.n_stages) * stage_bytes + 16u * i) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(smem_u32(&bars->a_full[s]), 1); mbar_init(smem_u32(&bars->a_empty[s]), 1); }
    mbar_init(smem_u32(&bars->acc_full[0]), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&bars->tmem));      // columns [0, N): dW tile; [256, 272): column sums
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t c = blockIdx.y; c < n_chunks; c += gridDim.y) {
        mbar_wait(smem_u32(&bars->a_empty[stage]), phase ^ 1);
        const uint32_t full = smem_u32(&bars->a_full[stage]);
        mbar_arrive_expect_tx(full, stage_bytes);
        const uint32_t base = smem_u32(smem) + stage * stage_bytes;
        const int32_t s0 = static_cast<int32_t>((P.reverse ? n_chunks - 1 - c : c) * 64);
        for (uint32_t p = 0; p < planes; ++p) {
          const uint32_t da = base + p * a_bytes, db = base + planes * a_bytes + p * b_bytes;
          for (int g = 0; g < 2; ++g) tma_load_2d(da + g * 8192u, p ? &P.a_lo : &P.a_hi, m_tile + 64 * g, s0, full);
          for (int g = 0; g < P.N / 64; ++g) tma_load_2d(db + g * 8192u, p ? &P.b_lo : &P.b_hi, P.n0 + 64 * g, s0, full);
        }
        const int64_t pf = c + 3 * static_cast<int64_t>(gridDim.y);       // three chunks ahead of the 2-deep ring: into L2
        if (pf < n_chunks)
          for (uint32_t p = 0; p < planes; ++p) {
            for (int g = 0; g < 2; ++g) tma_prefetch_2d(p ? &P.a_lo : &P.a_hi, m_tile + 64 * g, static_cast<int32_t>((P.reverse ? n_chunks - 1 - pf : pf) * 64));
            for (int g = 0; g < P.N / 64; ++g) tma_prefetch_2d(p ? &P.b_lo : &P.b_hi, P.n0 + 64 * g, static_cast<int32_t>((P.reverse ? n_chunks - 1 - pf : pf) * 64));
          }
        if (++stage == static_cast<uint32_t>(P.n_stages)) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16_major(128, static_cast<uint32_t>(P.N), 1u, 1u);
    const uint32_t idesc_cs = umma_idesc_f16_major(128, 16u, 1u, 1u);
    uint32_t stage = 0, phase = 0, accumulate = 0;
    for (int64_t c = blockIdx.y; c < n_chunks; c += gridDim.y) {
      mbar_wait(smem_u32(&bars->a_full[stage]), phase);
      tc_fence_after_sync();
      const uint32_t base = smem_u32(smem) + stage * stage_bytes;
      const uint32_t a_hi = base, a_lo = base + a_bytes, b_hi = base + planes * a_bytes, b_lo = b_hi + b_bytes;
      for (uint32_t pass = 0; pass < static_cast<uint32_t>(P.passes); ++pass) {
        const uint32_t a = pass == 1 ? a_lo : a_hi, b = pass == 2 ? b_lo : b_hi;
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          umma_f16_ss_warp(tmem, umma_desc_mn_sw128(a + ks * 2048u, 8192u, 1024u), umma_desc_mn_sw128(b + ks * 2048u, 8192u, 1024u), idesc, accumulate);
          accumulate = 1;
        }
      }
      if (P.colsum_partial) {      // bias gradient: (dY_hi + dY_lo)^T . ones, N = 16 (only column 0 is read back)
        for (uint32_t pl = 0; pl < planes; ++pl) {
          const uint32_t a = pl ? a_lo : a_hi;
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma_f16_ss_warp(tmem + 256u, umma_desc_mn_sw128(a + ks * 2048u, 8192u, 1024u), umma_desc_mn_sw128(smem_ones + ks * 2048u, 8192u, 1024u), idesc_cs,
                             (c != static_cast<int64_t>(blockIdx.y) || pl || ks) ? 1u : 0u);
        }
      }
      umma_commit_warp(smem_u32(&bars->a_empty[stage]));
      if (++stage == static_cast<uint32_t>(P.n_stages)) { stage = 0; phase ^= 1; }
    }
    umma_commit_warp(smem_u32(&bars->acc_full[0]));
  } else {
    const int q = warp & 3;
    const bool any = static_cast<int64_t>(blockIdx.y) < n_chunks;
    mbar_wait(smem_u32(&bars->acc_full[0]), 0);
    tc_fence_after_sync();
    const int m = 128 * blockIdx.x + 32 * q + lane;
    float* dst = P.partial + (static_cast<size_t>(blockIdx.y) * P.M_total + m) * P.N;
    if (P.colsum_partial) {
      uint32_t v[16];
      tmem_ld16(tmem + (static_cast<uint32_t>(32 * q) << 16) + 256u, v);
      tmem_ld_wait();
      P.colsum_partial[static_cast<size_t>(blockIdx.y) * P.M_total + m] = any ? __uint_as_float(v[0]) : 0.f;
    }
    for (int g = 0; g < P.N / 16; ++g) {
      uint32_t v[16];
      tmem_ld16(tmem + (static_cast<uint32_t>(32 * q) << 16) + 16u * g, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i)
        reinterpret_cast<float4*>(dst + 16 * g)[i] = any ? make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_planes_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NRF_E_CUDA; }
    fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  if (!base || (reinterpret_cast<uintptr_t>(base) & 15u) || (ld_elems & 7u)) { set_error("plane tensors must be 16-byte aligned with a row pitch that is a multiple of 8 elements"); return NRF_E_INVALID; }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld_elems * 2u};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for a [%llu x %llu] plane", static_cast<int>(r), (unsigned long long)rows, (unsigned long long)cols); return NRF_E_CUDA; }
  return NRF_OK;
}

static int device_sms(int n_sms, int* out) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (n_sms > 0 && n_sms < sms) sms = n_sms;
  *out = sms;
  return NRF_OK;
}

int launch_tile_gemm(const TileGemmArgs& a, int n_sms, cudaStream_t stream) {
  static thread_local TileGemmParams P;
  memset(&P, 0, sizeof(P));
  if (a.n_src < 1 || a.n_src > 2) { set_error("tile_gemm: n_src %d", a.n_src); return NRF_E_INVALID; }
  if (a.passes != 1 && a.passes != 3 && a.passes != 6) { set_error("tile_gemm: passes must be 1, 3 or 6"); return NRF_E_INVALID; }
  const int n_planes = a.passes == 6 ? 3 : (a.passes == 3 ? 2 : 1);
  P.bf16 = a.passes == 6 ? 1 : 0;
  if (a.N < 64 || (a.N & 63)) { set_error("tile_gemm: N = %d must be a multiple of 64", a.N); return NRF_E_INVALID; }
  const int64_t S = a.a[0].rows;
  if (S <= 0) return NRF_OK;
  int rc, sms;
  if ((rc = device_sms(n_sms, &sms)) != NRF_OK) return rc;
  int nk = 0;
  for (int j = 0; j < a.n_src; ++j) {
    if (a.a[j].cols < 64 || (a.a[j].cols & 63) || a.a[j].rows != S) { set_error("tile_gemm: A source %d has %d columns / %lld rows", j, a.a[j].cols, (long long)a.a[j].rows); return NRF_E_INVALID; }
    nk += a.a[j].cols / 64;
  }
  const uint32_t planes = static_cast<uint32_t>(n_planes);
  const uint32_t a_plane_stage = planes * kATile;
  // CTA pairs (cta_group::2, M = 256) whenever the output is 128 / 256 columns wide (or a multiple of 256), the weight half-slice
  // fits beside a 2-stage ring, and this is not the exact mode (whose per-chunk accumulators need all of TMEM for 128 columns)
  // (same-box A/B of the whole training step, 20 steps each, twice: 13.59 ms never paired, 13.01 ms large launches only, 12.95 ms whenever
  //  possible; a 393k x 256 x 256 layer alone: 187 us against 203 us, with half the L2 -> SMEM traffic)
  static int pair_mode = -1;      // NRF_GEMM_PAIR=0 / 1 / 2: never / large launches only / whenever possible (default); developer A/B
  if (pair_mode < 0) pair_mode = getenv("NRF_GEMM_PAIR") ? atoi(getenv("NRF_GEMM_PAIR")) : 2;
  bool pair = pair_mode > 0 && a.passes != 6 && sms >= 2 && (a.N == 128 || a.N % 256 == 0) && (pair_mode == 2 || (nk >= 4 && S >= 65536));
  if (pair) {
    const int nt = a.N == 128 ? 128 : 256;
    if (static_cast<uint32_t>(nk) * planes * (nt / 2) * 128u + 2 * a_plane_stage + 256 > kGemmSmemLimit) pair = false;
    else P.n_tile = nt;
  }
  if (!pair) P.n_tile = (a.N % 128 == 0) ? 128 : 64;
  const int n_cta = pair ? P.n_tile / 2 : P.n_tile;
  P.n_src = a.n_src; P.b_mn = a.b_mn ? 1 : 0; P.passes = a.passes; P.S = S;
  for (int j = 0; j < a.n_src; ++j) {
    const Planes& A = a.a[j]; const Planes& B = a.b[j];
    const __half* ap[3] = {A.hi, A.lo, A.ll};
    const __half* bp[3] = {B.hi, B.lo, B.ll};
    for (int q = 0; q < n_planes; ++q) if (!ap[q] || !bp[q]) { set_error("tile_gemm: %d passes need %d planes per operand", a.passes, n_planes); return NRF_E_INVALID; }
    P.kc[j] = A.cols / 64;
    if (!a.b_mn) {      // B[N, K]: rows = output features
      if (B.rows != a.N || B.cols != A.cols) { set_error("tile_gemm: B source %d is [%lld x %d], expected [%d x %d]", j, (long long)B.rows, B.cols, a.N, A.cols); return NRF_E_INVALID; }
    } else {            // B[K, N]: rows = reduction index
      if (B.rows != A.cols || B.cols != a.N) { set_error("tile_gemm: B source %d is [%lld x %d], expected [%d x %d]", j, (long long)B.rows, B.cols, A.cols, a.N); return NRF_E_INVALID; }
    }
    for (int q = 0; q < n_planes; ++q) {
      if ((rc = encode_planes_map(&P.a_map[j][q], ap[q], S, A.cols, A.ld, 64, 128)) != NRF_OK) return rc;
      if ((rc = encode_planes_map(&P.b_map[j][q], bp[q], B.rows, B.cols, B.ld, 64, a.b_mn ? 64 : n_cta)) != NRF_OK) return rc;
    }
  }
  uint32_t b_bytes = static_cast<uint32_t>(nk) * planes * n_cta * 128u, a_stage = a_plane_stage;
  if (b_bytes + 2 * a_stage + 256 > kGemmSmemLimit) {      // K too large for a resident weight slice: stream the B chunks with the A chunks
    P.b_stream = 1;
    a_stage += planes * n_cta * 128u;
    b_bytes = 0;
  }
  // staged epilogue (fp16 planes out through shared memory + TMA stores) whenever the ring keeps >= 2 stages beside it
  const uint32_t stage_out = 32768u;       // 8 epilogue warps x 4 KB private staging blocks ([32 x 32] hi | lo, or one [32 x 32] fp32 block)
  uint32_t out_bytes = 0;
  if (a.epi == GEPI_PLANES && a.passes != 6 && a.out.hi && !(reinterpret_cast<uintptr_t>(a.out.hi) & 15u) && !(a.out.ld & 7) &&
      (!a.out.lo || !(reinterpret_cast<uintptr_t>(a.out.lo) & 15u)) && b_bytes + 2 * a_stage + stage_out + 256 <= kGemmSmemLimit) {
    P.staged = 1;
    out_bytes = stage_out;
    if ((rc = encode_planes_map(&P.o_map[0], a.out.hi, S, a.N, a.out.ld, 32, 32, 64)) != NRF_OK) return rc;
    if (a.out.lo && (rc = encode_planes_map(&P.o_map[1], a.out.lo, S, a.N, a.out.ld, 32, 32, 64)) != NRF_OK) return rc;
    if (a.out_f32 && !(reinterpret_cast<uintptr_t>(a.out_f32) & 15u) && !(a.out_f32_ld & 3)) {
      // the fp32 copy (read by the scalar heads) takes the same route: a thread's 32 floats are one 128-byte row of a [128 x 32] box;
      // written straight from registers they were 8 x 16-byte stores into 32 different lines per warp (+130 us on a 250 us layer)
      P.f32_staged = 1;
      if ((rc = encode_planes_map(&P.f_map, a.out_f32, S, 2 * static_cast<uint64_t>(a.N), 2 * static_cast<uint64_t>(a.out_f32_ld), 64, 32)) != NRF_OK) return rc;
    }
  }
  int stages = static_cast<int>((kGemmSmemLimit - 256 - b_bytes - out_bytes) / a_stage);
  P.n_stages = stages > 4 ? 4 : stages;
  P.epi = a.epi; P.relu = a.relu; P.bias = a.bias; P.bias_ld = a.bias_ld; P.rows_per_ray = a.rows_per_ray > 0 ? a.rows_per_ray : 1;
  P.out_hi = a.out.hi; P.out_lo = a.out.lo; P.out_ll = a.out.ll; P.out_ld = a.out.ld; P.out_f32 = a.out_f32; P.out_f32_ld = a.out_f32_ld; P.accumulate = a.accumulate;
  P.mask_bits = a.mask_bits; P.bits_out = a.bits_out; P.bits_ld = a.bits_ld; P.row_scale = a.row_scale; P.row_scale_ld = a.row_scale_ld; P.col_vec = a.col_vec;
  P.sc_in = a.sc_in; P.sc_out = a.sc_out; P.l1max = a.l1max; P.status = a.status; P.reverse = a.reverse ? 1 : 0;
  if (a.epi == GEPI_PLANES && !a.out.hi) { set_error("tile_gemm: planes epilogue without an output"); return NRF_E_INVALID; }
  if (a.epi == GEPI_F32 && !a.out_f32) { set_error("tile_gemm: fp32 epilogue without an output"); return NRF_E_INVALID; }
  const int slices = a.N / P.n_tile;
  const int rows_tile = pair ? 256 : 128;
  const int64_t n_mt = (S + rows_tile - 1) / rows_tile;
  int gx = sms / slices;
  if (pair) gx /= 2;                                      // pairs per slice
  if (gx < 1) gx = 1;
  if (gx > n_mt) gx = static_cast<int>(n_mt);
  if (pair) gx *= 2;
  const uint32_t smem_bytes = b_bytes + static_cast<uint32_t>(P.n_stages) * a_stage + out_bytes + 256;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(gx, slices); cfg.blockDim = dim3(kTileThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  static int pf_dist = -1;
  if (pf_dist < 0) pf_dist = getenv("NRF_GEMM_PF") ? atoi(getenv("NRF_GEMM_PF")) : 1;
  P.pf_dist = pf_dist;
  static long long* dbg_trace = nullptr;
  static int dbg_left = -1;
  if (dbg_left < 0) dbg_left = getenv("NRF_GEMM_TRACE") ? atoi(getenv("NRF_GEMM_TRACE")) : 0;
  const bool tracing = dbg_left > 0 && S > 300000 && a.N == 256 && a.epi == GEPI_PLANES;
  if (tracing) {
    if (!dbg_trace) cudaMalloc(&dbg_trace, 1200 * sizeof(long long));
    cudaMemsetAsync(dbg_trace, 0, 1200 * sizeof(long long), stream);
    P.trace = dbg_trace;
  }
  cudaError_t e;
  if (P.bf16) {
    e = cudaFuncSetAttribute(tile_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmemLimit));
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, tile_gemm_kernel<true, false>, P);
  } else if (pair) {
    e = cudaFuncSetAttribute(tile_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmemLimit));
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, tile_gemm_kernel<false, true>, P);
  } else {
    e = cudaFuncSetAttribute(tile_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmemLimit));
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, tile_gemm_kernel<false, false>, P);
  }
  ++g_train_launches;
  if (tracing && e == cudaSuccess) {       // developer tap: print CTA 0's epilogue timeline of this launch (synchronises!)
    --dbg_left;
    static long long host[1200];
    cudaStreamSynchronize(stream);
    cudaMemcpy(host, dbg_trace, sizeof(host), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tile_gemm trace] pair=%d staged=%d stages=%d mask=%d b_mn=%d\n", pair ? 1 : 0, P.staged, P.n_stages, a.mask_bits ? 1 : 0, P.b_mn);
    long long t0 = host[1];
    for (int i = 0; i < 600 && host[2 * i + 1]; ++i) { fprintf(stderr, "%lld:%lld ", host[2 * i], host[2 * i + 1] - t0); if (host[2 * i] == 8 || (i + 1 < 600 && host[2 * (i + 1)] == 0)) fprintf(stderr, "\n"); }
    fprintf(stderr, "\n");
  }
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "tile_gemm_kernel launch");
}

int dw_gemm_max_split(int n_sms, int M) {
  int sms = 148;
  if (device_sms(n_sms, &sms) != NRF_OK) sms = 148;
  const int mt = M / 128 > 0 ? M / 128 : 1;
  return sms / mt > 0 ? sms / mt : 1;
}

int launch_dw_gemm(const Planes& a, int m0, int M, const Planes& b, int n0, int N, int passes, float* partial, int max_split,
                   int* n_split_out, int n_sms, cudaStream_t stream, float* colsum_partial, int reverse) {
  static thread_local DwGemmParams P;
  memset(&P, 0, sizeof(P));
  if (passes != 1 && passes != 3) { set_error("dw_gemm: passes must be 1 or 3"); return NRF_E_INVALID; }
  if (M < 128 || (M & 127) || N < 64 || N > 256 || (N & 63)) { set_error("dw_gemm: M = %d (multiple of 128) / N = %d (64..256, multiple of 64) unsupported", M, N); return NRF_E_INVALID; }
  if (a.rows != b.rows || a.rows <= 0) { set_error("dw_gemm: operand row counts differ (%lld vs %lld)", (long long)a.rows, (long long)b.rows); return NRF_E_INVALID; }
  if (m0 + M > ((a.cols + 127) & ~127) || n0 + N > b.cols) { set_error("dw_gemm: column window outside the planes"); return NRF_E_INVALID; }     // dY columns past a.cols are zero-filled by the TMA
  if (passes == 3 && (!a.lo || !b.lo)) { set_error("dw_gemm: 3 passes need the lo planes"); return NRF_E_INVALID; }
  int rc;
  if ((rc = encode_planes_map(&P.a_hi, a.hi, a.rows, a.cols, a.ld, 64, 64)) != NRF_OK) return rc;
  if ((rc = encode_planes_map(&P.b_hi, b.hi, b.rows, b.cols, b.ld, 64, 64)) != NRF_OK) return rc;
  if (passes == 3) {
    if ((rc = encode_planes_map(&P.a_lo, a.lo, a.rows, a.cols, a.ld, 64, 64)) != NRF_OK) return rc;
    if ((rc = encode_planes_map(&P.b_lo, b.lo, b.rows, b.cols, b.ld, 64, 64)) != NRF_OK) return rc;
  }
  P.m0 = m0; P.n0 = n0; P.N = N; P.passes = passes; P.S = a.rows; P.partial = partial; P.M_total = M; P.colsum_partial = colsum_partial; P.reverse = reverse ? 1 : 0;
  const uint32_t planes = passes == 3 ? 2u : 1u;
  const uint32_t stage_bytes = planes * (16384u + static_cast<uint32_t>(N) * 128u);
  const uint32_t ones_bytes = colsum_partial ? 8192u : 0u;
  int stages = static_cast<int>((kGemmSmemLimit - 256 - ones_bytes) / stage_bytes);
  P.n_stages = stages > 4 ? 4 : stages;
  const int64_t n_chunks = (a.rows + 63) / 64;
  int split = dw_gemm_max_split(n_sms, M);
  if (split > max_split) split = max_split;
  if (split > n_chunks) split = static_cast<int>(n_chunks);
  if (split < 1) split = 1;
  *n_split_out = split;
  cudaError_t e = cudaFuncSetAttribute(dw_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmemLimit));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dw_gemm)");
  dw_gemm_kernel<<<dim3(M / 128, split), kDwThreads, static_cast<uint32_t>(P.n_stages) * stage_bytes + ones_bytes + 256, stream>>>(P);
  ++g_train_launches;
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "dw_gemm_kernel launch");
}

}  // namespace nrf
