// The fused SMPL-NeRF forward: ONE persistent kernel per ray batch.
//
// Replaces (file:line in HannesStark/SMPL-NeRF)
//   models/nerf_pipeline.py:26-67, models/smpl_nerf_pipeline.py:27-100,
//   models/append_to_nerf_pipeline.py:25-90      (orchestration)
//   utils.py:114-131 (positional encoding), models/render_ray_net.py:42-61 and
//   models/warp_field_net.py:17-21 (MLPs), utils.py:134-191 (compositing),
//   utils.py:194-264 + torchsearchsorted (inverse-CDF sampling, sort-merge).
//
// Work item = a group of G rays (G * n_coarse <= 128) owned by one CTA from the coarse pass to the
// final colour, so no [rays x samples x features] tensor ever exists in HBM.
//
// CTAs run as PAIRS (cluster of 2 on one TPC, tcgen05 cta_group::2): every MMA is M = 256 -- 128 sample
// rows of each CTA's own tile -- and each CTA stages only its half of the weight (B) operand, so the
// L2 -> SMEM weight stream per SM is halved (a 48 KB ring of single-CTA stages could only keep
// ~27 B/clk/SM in flight against the ~43 B/clk/SM the 3-pass MMA consumes; profiles/r1).
//
// CTA = 18 warps, warp-specialised:
//   warp 0      weight producer: streams this CTA's half of the pre-swizzled fp16 hi/lo weight stages
//               L2 -> SMEM ring with 2-D tensor TMA (cp.async.bulk.tensor.2d.cta_group::2); both CTAs' halves
//               credit their bytes to the ONE full-barrier of the issuing CTA (no relay hop)
//   warp 1      even CTA: MMA issuer -- one thread issues tcgen05.mma.cta_group::2 (M=256, N=n_out,
//               fp16 x fp16 -> fp32 in both CTAs' TMEM); fp32-level accuracy comes from the split
//               x*w ~= xh*wh + xl*wh + xh*wl (3 passes).  odd CTA: idle
//   warps 2..17 "epilogue" warps (thread = one sample row x 16 columns of every 64-feature chunk):
//               positional encoding -> SMEM operand tiles, TMEM -> bias/ReLU/hi-lo split -> next
//               layer's A operand (in place), sigma / rgb / warp heads as fp32 dot products, then per
//               ray: alpha compositing, inverse-CDF sampling, sorted merge.
// Two 256-column TMEM accumulators ping-pong between consecutive layers, and activations are handed
// to the MMA issuer per 64-feature K-chunk, so layer l+1's MMAs start while layer l's epilogue is
// still draining.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "nrf_plan.h"
#include "nrf_ptx.cuh"
#include "nrf_stages.cuh"

namespace nrf {

constexpr int kMaxStages = 3;                      // weight ring slots (2 when the per-ray tables need the room)
constexpr uint32_t kSlotBytes = 16384;             // this CTA's half of a [256 x 64] fp16 stage: 128 rows x 128 B
constexpr uint32_t kChunkBytes = 16384;            // one [128 x 64] fp16 operand tile
constexpr uint32_t kOffA = 0;                      // 4 chunks x (hi, lo)
constexpr uint32_t kOffAux = 4 * 2 * kChunkBytes;  // 131072
constexpr uint32_t kOffRing = kOffAux + 2 * kChunkBytes;
constexpr uint32_t kOffXchg = kOffA + 3 * 2 * kChunkBytes;   // head-partial exchange: aliases A chunk 3 (hi) while it is dead
constexpr uint32_t kSmemLimit = 232448;            // 227 KB opt-in maximum per CTA
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kRayVec = 8 + kMaxRayFeat + kChunkK; // offset of unit direction[3], raw pose pair[2]: behind o, d, |d|, valid (8), pose feats (<= kMaxRayFeat), dir feats (<= kChunkK, what plan_raynet admits)
constexpr int kRayFloats = kRayVec + 8;
constexpr int kRayScratchFloats = 2 * kMaxFineRows + (kTileRows + 16) + kTileRows + kMaxFineRows + kTeamScratch;   // tf, tt, cdf, pdf, z_samples, team partials of one ray
static_assert((kTileRows / 16) * kRayScratchFloats * 4 <= 2 * 2 * 16384, "per-ray scratch must stay inside activation chunks 0 and 1 (chunk 2 stages the warp encoding, chunk 3 the exchange slots)");            // o[3] d[3] |d| valid | pose feats[64] | dir feats[64] | unit d, pose

// barrier slots inside the misc area (8 bytes each)
enum { BAR_FULL = 0, BAR_EMPTY = kMaxStages, BAR_ACC = 2 * kMaxStages, BAR_AREADY = 2 * kMaxStages + 2,
       BAR_COUNT = 2 * kMaxStages + 2 + 5 };
constexpr uint32_t kBarBytes = 256;                // barriers + TMEM base slot
static_assert(8 * BAR_COUNT + 8 <= kBarBytes, "barrier area too small");

struct RenderParams {
  CUtensorMap tmap[3];   // weight streams of blob[0..2] as [rows, 128 B] uint8 tensors, box = 64 rows (8 KB)
  NetPlan net[2];
  NetPlan warp;
  const uint8_t* blob[3];
  NrfRenderIO io;
  int64_t n_rays;
  int32_t kind, n_coarse, n_fine, n_all, run_fine, white_bkgd, fast;
  int32_t pose_freqs, pose_identity, pose_encoded, pose_stride, pose_col0, pose_col1, pose_dim;
  int32_t G, tiles_f, n_groups, n_pairs, n_stages;
  uint32_t off_misc;   // byte offset of the barrier + per-ray area (behind the weight ring)
  // float offsets inside the misc area (after the barriers)
  uint32_t o_ray, o_rb, o_rbw, o_raw, o_zc, o_zf, o_u, o_hw, o_hr, o_hs;
};

struct Smem {
  uint8_t* base;
  float* misc;
  uint32_t bar0;
  uint32_t n_stages;
  uint32_t rank;       // rank of this CTA in its pair (0 = issuer)
  __device__ uint32_t bar(int i) const { return bar0 + 8u * i; }
};

// ---------------------------------------------------------------------------------- small helpers
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// {x0 -> low half, x1 -> high half}, round-to-nearest, clamped to +-65504 (SASS: F2FP.SATFINITE.F16.F32.PACK_AB)
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float x0, float x1) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
  return r;
}
// fp32 pair -> packed fp16 hi pair and packed fp16 residual pair (x ~= hi + lo, ~22 significant bits)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& h, uint32_t& l) {
  h = cvt_f16x2_sat(x0, x1);
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
  l = cvt_f16x2_sat(x0 - f.x, x1 - f.y);
}

// Write 16 consecutive features (16-byte swizzle chunks cc0, cc0+1 of row `row`; cc0 even) as hi/lo fp16
// into an operand tile pair (hi tile at `tile`, lo tile at `tile + kChunkBytes`).  amax2 accumulates
// max |hi| (packed halves) for the fp16-range status flag.
__device__ __forceinline__ void store_feat16(uint32_t tile, int row, int cc0, const float (&x)[16], bool fast, __half2& amax2) {
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = cvt_f16x2_sat(x[2 * i], x[2 * i + 1]);
    amax2 = __hmax2(amax2, __habs2(*reinterpret_cast<const __half2*>(&h[i])));
  }
  const uint32_t rofs = static_cast<uint32_t>(row) * 128u;
  const uint32_t o0 = rofs + (static_cast<uint32_t>((cc0 ^ row) & 7) << 4);
  const uint32_t o1 = rofs + (static_cast<uint32_t>(((cc0 + 1) ^ row) & 7) << 4);
  st_shared_v4(tile + o0, h[0], h[1], h[2], h[3]);
  st_shared_v4(tile + o1, h[4], h[5], h[6], h[7]);
  if (!fast) {      // residuals only in parity mode (the fast mode is epilogue-bound: skipping them there is worth ~35% of it)
    uint32_t l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
      l[i] = cvt_f16x2_sat(x[2 * i] - f.x, x[2 * i + 1] - f.y);
    }
    st_shared_v4(tile + kChunkBytes + o0, l[0], l[1], l[2], l[3]);
    st_shared_v4(tile + kChunkBytes + o1, l[4], l[5], l[6], l[7]);
  }
}

// Encoder features [16*cg, 16*cg+16) of vector v in ENGINE order (see enc_ref_col) -> aux tile:
// pair p = comp * L + k owns features 2p (sin) and 2p+1 (cos) of v[comp] * 2^k; identity components follow.
// Within a run of consecutive frequencies of one component only every third pair is evaluated directly; the two
// after it come from the double-angle formulas (sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a): 4.9e-7 max abs
// error after two doublings against 6e-8 for a correctly rounded sin -- far below what the 1e-4 sigma bar needs --
// for about half the instructions.
__device__ __forceinline__ void write_encoding(uint32_t aux_tile, int row, int cg, float vx, float vy, float vz, int freqs,
                                               int identity, bool fast) {
  float f[16];
  int p = 8 * cg, comp = freqs > 0 ? p / freqs : 3, k = freqs > 0 ? p - comp * freqs : 0;
  float sv = 0.f, cv = 1.f;
  int depth = 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float v = comp == 0 ? vx : (comp == 1 ? vy : vz);
    if (p < 3 * freqs) {
      if (depth >= 2 || k == 0) { sincos_pe(v * __int_as_float((127 + k) << 23), sv, cv); depth = 0; }
      else { const float s2 = 2.f * sv * cv; cv = fmaf(-2.f * sv, sv, 1.f); sv = s2; ++depth; }
      f[2 * i] = sv; f[2 * i + 1] = cv;
    } else {
      const int c0 = 2 * p - 6 * freqs, c1 = c0 + 1;   // identity components follow the sin/cos block
      f[2 * i] = (identity && c0 < 3) ? (c0 == 0 ? vx : (c0 == 1 ? vy : vz)) : 0.f;
      f[2 * i + 1] = (identity && c1 < 3) ? (c1 == 1 ? vy : vz) : 0.f;
    }
    ++p;
    if (++k == freqs) { k = 0; ++comp; }
  }
  __half2 dummy = __floats2half2_rn(0.f, 0.f);
  store_feat16(aux_tile, row, 2 * cg, f, fast, dummy);
}

// Developer tap: CTA 0 appends (event, counter, SM clock) to io.trace ([0] = capacity, [1] = count).  The slot
// counter lives in shared memory (a global atomic's round trip would inflate every traced segment by ~1000 cycles)
// and is copied to trace[1] when the kernel ends.
constexpr uint32_t kTraceCtrOfs = 128;   // inside the barrier area
__device__ __forceinline__ uint32_t trace_slots(const RenderParams& P, uint32_t n) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  return atomicAdd(reinterpret_cast<uint32_t*>(smem_raw + P.off_misc + kTraceCtrOfs), n);
}
__device__ __forceinline__ void trace_ev(const RenderParams& P, int ev, uint32_t ctr) {
  if (P.io.trace && blockIdx.x == 0) {
    const long long t = clock64();
    const uint32_t i = trace_slots(P, 1);
    if (static_cast<long long>(i) < P.io.trace[0]) { long long* e = P.io.trace + 2 + 3 * i; e[0] = ev; e[1] = ctr; e[2] = t; }
  }
}

// ---------------------------------------------------------------------------------- roles
struct RingState { uint32_t stage = 0, phase = 0; __device__ void advance(uint32_t n) { if (++stage == n) { stage = 0; phase ^= 1; } } };

__device__ __forceinline__ void producer_layer(const Smem& sm, const CUtensorMap* tmap, const Layer& L, RingState& rs, bool fast) {
  const uint32_t half_bytes = L.n_out * 64u;       // this CTA's [n_out/2 x 64] fp16 half of a stage
  const int half_rows = L.n_out >> 1;              // 128-byte rows of that half
  int row = static_cast<int>(L.stream_ofs >> 7) + static_cast<int>(sm.rank) * half_rows;
  for (int kc = 0; kc < L.nk; ++kc) {
    for (int is_lo = 0; is_lo < 2; ++is_lo, row += 2 * half_rows) {
      if (fast && is_lo) continue;                 // lo stages are not streamed in fast mode
      mbar_wait(sm.bar(BAR_EMPTY + rs.stage), rs.phase ^ 1);
      // both CTAs' halves are credited to the ISSUER's full barrier (its own producer announces the total)
      const uint32_t full0 = mapa_shared(sm.bar(BAR_FULL + rs.stage), 0);
      if (sm.rank == 0) mbar_arrive_expect_tx(sm.bar(BAR_FULL + rs.stage), 2u * half_bytes);
      const uint32_t dst = smem_u32(sm.base + kOffRing) + rs.stage * kSlotBytes;
      for (int r0 = 0; r0 < half_rows; r0 += 64) tma2_load_2d(dst + static_cast<uint32_t>(r0) * 128u, tmap, 0, row + r0, full0);
      rs.advance(sm.n_stages);
    }
  }
}

struct MmaState { RingState rs; uint32_t a_phase = 0; uint32_t layer_ctr = 0; };

// One layer: for every 64-feature K-chunk, a hi and a lo weight stage; the hi stage multiplies both the
// hi and the lo activations, the lo stage only the hi activations:  x*w ~= xh*wh + xl*wh + xh*wl.
// Every instruction is M=256 (pair) x N=n_out x K=16; 8 + 4 of them per chunk, one commit per stage.
// Executed by all 32 lanes of warp 1 (warp-uniform); an elected lane issues.
__device__ __forceinline__ void mma_layer(const Smem& sm, const RenderParams& P, uint32_t tmem_base, const Layer& L, MmaState& st, bool fast) {
  const uint32_t acc = tmem_base + (st.layer_ctr & 1u) * 256u;
  const uint32_t idesc = umma_idesc_f16(256, L.n_out);
  const uint32_t a_base = smem_u32(sm.base);
  const bool tracing = P.io.trace != nullptr && blockIdx.x == 0;
  long long w_ops = 0, w_full = 0, t_in = 0;
  if (tracing) t_in = clock64();
  uint32_t accumulate = 0;
  for (int kc = 0; kc < L.nk; ++kc) {
    const int src = L.ksrc[kc];
    if (src != kSrcAux || (L.flags & LF_AUX_WAIT)) {
      const long long t0 = tracing ? clock64() : 0;
      mbar_wait(sm.bar(BAR_AREADY + src), (st.a_phase >> src) & 1u);
      if (tracing) w_ops += clock64() - t0;
      st.a_phase ^= 1u << src;
      tc_fence_after_sync();
    }
    const uint32_t a_hi = a_base + (src == kSrcAux ? kOffAux : kOffA + static_cast<uint32_t>(src) * 2u * kChunkBytes);
    const uint32_t a_lo = a_hi + kChunkBytes;
    for (int is_lo = 0; is_lo < 2; ++is_lo) {
      if (fast && is_lo) continue;
      const long long t0 = tracing ? clock64() : 0;
      mbar_wait(sm.bar(BAR_FULL + st.rs.stage), st.rs.phase);
      if (tracing) w_full += clock64() - t0;
      tc_fence_after_sync();
      const uint64_t bdesc = umma_desc_sw128(a_base + kOffRing + st.rs.stage * kSlotBytes);
      const int n_apass = (is_lo || fast) ? 1 : 2;
      for (int ap = 0; ap < n_apass; ++ap) {
        const uint64_t adesc = umma_desc_sw128(ap == 0 ? a_hi : a_lo);
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          umma2_f16_ss_warp(acc, adesc + 2u * ks, bdesc + 2u * ks, idesc, accumulate);   // +32 bytes per K=16 step
          accumulate = 1;
        }
      }
      umma2_commit_warp(sm.bar(BAR_EMPTY + st.rs.stage));   // frees the ring slot in both CTAs when these MMAs are done
      st.rs.advance(sm.n_stages);
    }
  }
  umma2_commit_warp(sm.bar(BAR_ACC + (st.layer_ctr & 1u)));   // accumulators complete -> both CTAs' epilogues
  if (tracing && (threadIdx.x & 31) == 0) {   // one record per layer: issue window, cycles blocked on operands / on the weight ring
    const long long t_out = clock64();
    const uint32_t i = trace_slots(P, 4);
    if (static_cast<long long>(i) + 4 <= P.io.trace[0]) {
      long long* e = P.io.trace + 2 + 3 * i;
      e[0] = 0; e[1] = st.layer_ctr; e[2] = t_in;  e[3] = 8; e[4] = st.layer_ctr; e[5] = t_out;
      e[6] = 20; e[7] = st.layer_ctr; e[8] = w_ops; e[9] = 21; e[10] = st.layer_ctr; e[11] = w_full;
    }
  }
  st.layer_ctr++;
}

// Per-thread state of an epilogue thread: row = 32*q + lane (q = warp % 4 selects the TMEM lane
// quarter this warp may read), cg = column group: the thread owns columns 64j + 16cg .. +16 of every
// 64-feature chunk j, so chunk j of the next layer's A operand completes after 1/nch of the epilogue.
struct EpiCtx {
  int warp, lane, q, cg, row, tid;     // tid: 0..511 within the epilogue group
  uint32_t acc_phase = 0;              // bit b: parity to wait for on accumulator b
  uint32_t layer_ctr = 0;
  uint32_t tmem_base;
  uint32_t lane_taddr;                 // TMEM lane field for this warp's quarter
  __half2 amax2;                       // running max |activation| (fp16-range status flag)
  const float* head_s;                 // smem copy of the 3-row head weights of the layer being drained ([3][n_out])
  const float* sigma_s;                // smem copy of the sigma head weights ([256])
};

// Publish an operand tile (A chunk `which` or the aux tile) this CTA's 16 epilogue warps just wrote:
// generic-proxy stores -> async proxy, CTA-wide named barrier (keeps the warps in lockstep so chunk j is
// ready after (j+1)/nch of the epilogue even though the warp scheduler is not fair), then ONE arrival
// on the issuer CTA's barrier (expected count 2: one per CTA of the pair).
__device__ __forceinline__ void epi_publish(const Smem& sm, const EpiCtx& c, int which) {
  fence_proxy_async_smem();
  tc_fence_before_sync();
  named_bar_sync(2, kEpiThreads);
  if (c.tid == 0) {
    if (sm.rank == 0) mbar_arrive(sm.bar(BAR_AREADY + which));
    else mbar_arrive_cluster_relaxed(sm.bar(BAR_AREADY + which), 0);
  }
}

struct HeadOut { float h0, h1, h2, sig; };

__device__ __forceinline__ void ld_f16(const float* p, float (&b)[16]) {   // 16 consecutive floats, 16-byte aligned
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void ldg_f16(const float* p, float (&b)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
    b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ float dot16(const float (&x)[16], const float (&w)[16], float acc) {
#pragma unroll
  for (int i = 0; i < 16; ++i) acc = fmaf(x[i], w[i], acc);
  return acc;
}

// Epilogue of one MMA layer for this thread's row and its 16-column slice of every chunk.
//   kRelu: out = relu(acc + bias) (else acc + bias)     kWriteA: out -> next layer's A operand chunks
//   kHead3: 3-row head dot products (rgb / warp)        kSigma: sigma head dot product
//   g: ray index inside the group of this thread's row (selects the per-ray bias vector)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

template <bool kRelu, bool kWriteA, bool kHead3, bool kSigma>
__device__ __forceinline__ void epilogue_layer(const Smem& sm, const RenderParams& P, const NetPlan& net, const float* f32,
                                               const float* rb_base, const Layer& L, EpiCtx& c, int g, HeadOut& ho) {
  const uint32_t buf = c.layer_ctr & 1u;
  const uint32_t acc = c.tmem_base + buf * 256u + c.lane_taddr + 16u * static_cast<uint32_t>(c.cg);
  const bool fast = P.fast != 0;
  const int nch = L.n_out >> 6;
  const float* bias_g = f32 + L.bias_ofs + 16 * c.cg;
  const float* bias_s = (L.ray_slot >= 0) ? rb_base + (static_cast<int>(L.ray_slot) * P.G + g) * kWidth + 16 * c.cg : nullptr;
  const float* wh = c.head_s + 16 * c.cg;     // [3][n_out]   (shared memory: a global load here is pure latency)
  const float* ws = c.sigma_s + 16 * c.cg;    // [256]
  float h0 = 0.f, h1 = 0.f, h2 = 0.f, sg = 0.f;

  // warm L1 with this thread's bias lines while the accumulator is still being produced
  if (!bias_s) {
#pragma unroll 1
    for (int j = 0; j < nch; ++j) asm volatile("prefetch.global.L1 [%0];" ::"l"(bias_g + 64 * j));
  }

  mbar_wait(sm.bar(BAR_ACC + buf), (c.acc_phase >> buf) & 1u);
  c.acc_phase ^= 1u << buf;
  tc_fence_after_sync();
  if (c.tid == 0) trace_ev(P, 10, c.layer_ctr);

#pragma unroll 1
  for (int j = 0; j < nch; ++j) {
    // issue the TMEM load, then fetch the biases (and the head weights) while it is in flight.  (Hoisting the
    // next chunk's TMEM load / bias above the publish of this one was measured slower: tcgen05.fence::
    // before_thread_sync waits for the load, and 32 more live registers across the loop cost more than they hide.)
    uint32_t v[16];
    tmem_ld16_issue(acc + 64u * static_cast<uint32_t>(j), v);
    float b[16];
    if (bias_s) ld_f16(bias_s + 64 * j, b);
    else ldg_f16(bias_g + 64 * j, b);
    float w[16];
    if (kSigma) ld_f16(ws + 64 * j, w);
    if (kHead3) ld_f16(wh + 64 * j, w);
    tmem_ld_wait();
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float t = __uint_as_float(v[i]) + b[i];
      x[i] = kRelu ? fmaxf(t, 0.f) : t;
    }
    if (kSigma) sg = dot16(x, w, sg);
    if (kHead3) {
      ld_f16(wh + L.n_out + 64 * j, b);      h0 = dot16(x, w, h0);
      ld_f16(wh + 2 * L.n_out + 64 * j, w);  h1 = dot16(x, b, h1);
      h2 = dot16(x, w, h2);
    }
    if (kWriteA) {
      const uint32_t tile = smem_u32(sm.base) + kOffA + static_cast<uint32_t>(j) * 2u * kChunkBytes;
      store_feat16(tile, c.row, 2 * c.cg, x, fast, c.amax2);
      epi_publish(sm, c, j);
      if (c.tid == 0) trace_ev(P, 11 + j, c.layer_ctr);
    }
  }
  if (!kWriteA) { tc_fence_before_sync(); if (c.tid == 0) trace_ev(P, 15, c.layer_ctr); }
  ho.h0 = h0; ho.h1 = h1; ho.h2 = h2; if (kSigma) ho.sig = sg;
  c.layer_ctr++;
}

__device__ __forceinline__ void epilogue_dispatch(const Smem& sm, const RenderParams& P, const NetPlan& net, const float* f32,
                                                  const float* rb_base, const Layer& L, EpiCtx& c, int g, HeadOut& ho) {
  if (L.epi == EPI_RELU) {
    if (L.flags & LF_SIGMA_HEAD) epilogue_layer<true, true, false, true>(sm, P, net, f32, rb_base, L, c, g, ho);   // folded additional_linear_layer
    else epilogue_layer<true, true, false, false>(sm, P, net, f32, rb_base, L, c, g, ho);
  } else if (L.epi == EPI_LINEAR) {
    if (L.flags & LF_SIGMA_HEAD) epilogue_layer<false, true, false, true>(sm, P, net, f32, rb_base, L, c, g, ho);
    else epilogue_layer<false, true, false, false>(sm, P, net, f32, rb_base, L, c, g, ho);
  } else epilogue_layer<true, false, true, false>(sm, P, net, f32, rb_base, L, c, g, ho);
}

// ---------------------------------------------------------------------------------- the kernel
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) nrf_fused_kernel(const __grid_constant__ RenderParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem sm;
  sm.base = smem_raw;
  sm.misc = reinterpret_cast<float*>(smem_raw + P.off_misc + kBarBytes);
  sm.bar0 = smem_u32(smem_raw + P.off_misc);
  sm.n_stages = static_cast<uint32_t>(P.n_stages);
  sm.rank = cluster_ctarank();
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + P.off_misc + 8 * BAR_COUNT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool fast = P.fast != 0;
  const bool smpl = P.kind == NRF_KIND_SMPL;

  if (threadIdx.x == 0) {
    *reinterpret_cast<uint32_t*>(smem_raw + P.off_misc + kTraceCtrOfs) = 0u;
    // a stage is full when the issuer CTA's producer has arrived (announcing the bytes of BOTH halves) and both CTAs' TMA bytes have landed
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(sm.bar(BAR_FULL + s), 1); mbar_init(sm.bar(BAR_EMPTY + s), 1); }
    mbar_init(sm.bar(BAR_ACC + 0), 1); mbar_init(sm.bar(BAR_ACC + 1), 1);
    for (int j = 0; j < 5; ++j) mbar_init(sm.bar(BAR_AREADY + j), 2);   // one arrival per CTA of the pair
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2<512>(smem_u32(tmem_slot));
  tc_fence_before_sync();
  cluster_sync_all();          // barriers of both CTAs are initialised before any remote arrival
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int n_pass = P.run_fine ? 2 : 1;

  if (warp == 0) {
    // =========================== weight producer (both CTAs: own half of every stage) ===========================
    if (lane == 0) {
      RingState rs;
      for (int pr = blockIdx.x >> 1; pr < P.n_pairs; pr += gridDim.x >> 1)
        for (int pass = 0; pass < n_pass; ++pass) {
          const int tiles = pass == 0 ? 1 : P.tiles_f;
          for (int t = 0; t < tiles; ++t) {
            if (smpl) producer_layer(sm, &P.tmap[2], P.warp.layers[0], rs, fast);
            for (int l = 0; l < P.net[pass].n_layers; ++l) producer_layer(sm, &P.tmap[pass], P.net[pass].layers[l], rs, fast);
          }
        }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (even CTA; this warp idles in the odd CTA) ===========================
    // the whole warp runs these loops convergently; one elected lane issues each instruction
    {
      MmaState st;
      for (int pr = blockIdx.x >> 1; pr < P.n_pairs; pr += gridDim.x >> 1)
        for (int pass = 0; pass < n_pass; ++pass) {
          const int tiles = pass == 0 ? 1 : P.tiles_f;
          for (int t = 0; t < tiles; ++t) {
            if (sm.rank == 0) {
              if (smpl) mma_layer(sm, P, tmem_base, P.warp.layers[0], st, fast);
              for (int l = 0; l < P.net[pass].n_layers; ++l) mma_layer(sm, P, tmem_base, P.net[pass].layers[l], st, fast);
            }
          }
        }
    }
  } else {
    // =========================== epilogue warps ===========================
    EpiCtx c;
    c.warp = warp; c.lane = lane; c.q = warp & 3; c.cg = (warp - 2) >> 2; c.row = 32 * c.q + lane;
    c.tid = threadIdx.x - 64; c.tmem_base = tmem_base; c.lane_taddr = static_cast<uint32_t>(32 * c.q) << 16;
    c.amax2 = __floats2half2_rn(0.f, 0.f);
    c.head_s = nullptr; c.sigma_s = nullptr;
    const int ew = warp - 2;   // 0..15
    float* ray = sm.misc + P.o_ray;          // [G][kRayFloats]
    float* rb = sm.misc + P.o_rb;            // [slots][G][256]
    float* rbw = sm.misc + P.o_rbw;          // warp net: [G][256]
    float4* raw4 = reinterpret_cast<float4*>(sm.misc + P.o_raw);   // [G * n_all]
    float* zc = sm.misc + P.o_zc;            // [G][n_coarse]
    float* zf = sm.misc + P.o_zf;            // [G][n_all]
    float* u_s = sm.misc + P.o_u;            // [n_fine] the sampler's u = linspace(0, 1, n_fine)
    float* hw_s = sm.misc + P.o_hw;          // warp net head weights [3][256] (smpl)
    float* hr_s = sm.misc + P.o_hr;          // rgb head weights [3][128] of the current net
    float* hs_s = sm.misc + P.o_hs;          // sigma head weights [256] of the current net
    // head-partial exchange: a [column group][row] float4 table inside A chunk 3 (dead whenever it is used)
    float4* xchg = reinterpret_cast<float4*>(sm.base + kOffXchg) + c.row;   // slot of column group k: xchg[kTileRows * k] (lanes contiguous: no bank conflicts)
    float4* xchg_w = xchg + 4 * kTileRows;                                   // second table (same dead chunk) for the warp-net head
    const uint32_t aux_tile = smem_u32(sm.base) + kOffAux;
    const int G = P.G, nc = P.n_coarse, nf = P.n_fine, na = P.n_all;
    // head biases live in registers (a dependent global load here sits on the tile-to-tile critical path)
    float wb0 = 0.f, wb1 = 0.f, wb2 = 0.f;
    if (smpl) {
      const float* b2 = reinterpret_cast<const float*>(P.blob[2] + P.warp.f32_ofs) + P.warp.head_ofs + 3 * kWidth;
      wb0 = __ldg(b2 + 0); wb1 = __ldg(b2 + 1); wb2 = __ldg(b2 + 2);
      for (int i = c.tid; i < 3 * kWidth; i += kEpiThreads) hw_s[i] = __ldg(b2 - 3 * kWidth + i);
    }

    if (P.run_fine && !P.io.z_all_in) for (int i = c.tid; i < nf; i += kEpiThreads) u_s[i] = __ldg(P.io.u_fine + i);

    for (int pr = blockIdx.x >> 1; pr < P.n_pairs; pr += gridDim.x >> 1) {
      // the pair renders ray groups 2*pr and 2*pr+1 in lockstep; a group past the end has no valid
      // rays (every load/store is guarded) but still runs its MMAs, which the pair shares
      const int64_t ray0 = (static_cast<int64_t>(pr) * 2 + sm.rank) * G;
      // ---- per-ray constants: origin, direction, |d|, raw pose pair
      if (c.tid < G) {
        const int g = c.tid;
        float* r = ray + g * kRayFloats;
        const int64_t ri = ray0 + g;
        const bool valid = ri < P.n_rays;
        float o[3] = {0, 0, 0}, d[3] = {0, 0, 1};
        if (valid) for (int k = 0; k < 3; ++k) { o[k] = P.io.ray_origin[ri * 3 + k]; d[k] = P.io.ray_dir[ri * 3 + k]; }
        const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
        for (int k = 0; k < 3; ++k) { r[k] = o[k]; r[3 + k] = d[k]; r[kRayVec + k] = __fdiv_rn(d[k], nrm); }
        r[6] = nrm; r[7] = valid ? 1.f : 0.f;
        float pz0 = 0.f, pz1 = 0.f;
        if (valid && P.pose_dim > 0) { pz0 = P.io.goal_pose[ri * P.pose_stride + P.pose_col0]; pz1 = P.io.goal_pose[ri * P.pose_stride + P.pose_col1]; }
        r[kRayVec + 3] = pz0; r[kRayVec + 4] = pz1;
      }
      named_bar_sync(1, kEpiThreads);
      // ---- per-ray feature vectors in REFERENCE column order, one sin/cos per thread
      {
        const int n_dirf = smpl ? 0 : enc_dim(P.net[0].dir_freqs, P.net[0].dir_identity);
        const int per_ray = P.pose_dim + n_dirf;
        for (int idx = c.tid; idx < G * per_ray; idx += kEpiThreads) {
          const int g = idx / per_ray;
          int i = idx - g * per_ray;
          float* r = ray + g * kRayFloats;
          const bool is_dir = i >= P.pose_dim;
          if (is_dir) i -= P.pose_dim;
          const int ncomp = is_dir ? 3 : 2;
          const float* v = r + kRayVec + (is_dir ? 0 : 3);
          const int ident = is_dir ? P.net[0].dir_identity : (P.pose_encoded ? P.pose_identity : 1);
          float val;
          if (ident && i < ncomp) val = v[i];
          else {
            const int j = i - (ident ? ncomp : 0);
            const int k = j / (2 * ncomp), rem = j - k * 2 * ncomp;
            float sv, cv;
            sincos_pe(v[rem % ncomp] * __int_as_float((127 + k) << 23), sv, cv);
            val = rem < ncomp ? sv : cv;
          }
          r[8 + (is_dir ? kMaxRayFeat : 0) + i] = val;
        }
      }
      for (int i = c.tid; i < G * nc; i += kEpiThreads) {
        const int64_t ri = ray0 + i / nc;
        zc[i] = ri < P.n_rays ? P.io.z_vals[ri * nc + (i % nc)] : static_cast<float>(i % nc);
      }
      named_bar_sync(1, kEpiThreads);
      if (smpl && P.warp.layers[0].ray_slot >= 0) {   // warp net pose bias, shared by both passes
        const Layer& L = P.warp.layers[0];
        const float* f32 = reinterpret_cast<const float*>(P.blob[2] + P.warp.f32_ofs);
        for (int i = c.tid; i < G * kWidth; i += kEpiThreads) {
          const int g = i >> 8, col = i & 255;
          float acc = f32[L.bias_ofs + col];
          const float* feat = ray + g * kRayFloats + 8;
          for (int k = 0; k < L.ray_k; ++k) acc = fmaf(f32[L.rayw_ofs + k * kWidth + col], feat[k], acc);
          rbw[i] = acc;
        }
      }

      for (int pass = 0; pass < n_pass; ++pass) {
        const NetPlan& net = P.net[pass];
        const float* f32 = reinterpret_cast<const float*>(P.blob[pass] + net.f32_ofs);
        const bool last_pass = (pass == n_pass - 1);
        const int n = pass == 0 ? nc : na;
        const int tiles = pass == 0 ? 1 : P.tiles_f;
        const float hb0 = __ldg(f32 + net.head_ofs + 3 * (kWidth / 2) + 0), hb1 = __ldg(f32 + net.head_ofs + 3 * (kWidth / 2) + 1),
                    hb2 = __ldg(f32 + net.head_ofs + 3 * (kWidth / 2) + 2), sb = __ldg(f32 + net.sigma_ofs + kWidth);
        for (int i = c.tid; i < 3 * (kWidth / 2); i += kEpiThreads) hr_s[i] = __ldg(f32 + net.head_ofs + i);
        for (int i = c.tid; i < kWidth; i += kEpiThreads) hs_s[i] = __ldg(f32 + net.sigma_ofs + i);
        c.sigma_s = hs_s;
        // ---- per-ray bias vectors of this net (pose / direction contributions)
        for (int l = 0; l < net.n_layers; ++l) {
          const Layer& L = net.layers[l];
          if (L.ray_slot < 0) continue;
          if (L.ray_src == RAY_POSE_EXT) {      // computed per ray by nrf_ray_bias (bias included)
            const float* src = pass == 0 ? P.io.ray_bias_coarse : P.io.ray_bias_fine;
            const bool shared_pose = P.io.ray_bias_nonuniform && __ldg(P.io.ray_bias_nonuniform) == 0;   // then only row 0 exists
            for (int i = c.tid; i < G * kWidth; i += kEpiThreads) {
              const int g = i >> 8, col = i & 255;
              const int64_t ri = ray0 + g;
              rb[(static_cast<int>(L.ray_slot) * G + g) * kWidth + col] =
                  ri < P.n_rays ? __ldg(src + ((shared_pose ? 0 : ri) * net.n_ext_slots + L.ext_idx) * kWidth + col) : 0.f;
            }
            continue;
          }
          for (int i = c.tid; i < G * L.n_out; i += kEpiThreads) {
            const int g = i / L.n_out, col = i % L.n_out;
            float acc = f32[L.bias_ofs + col];
            const float* feat = ray + g * kRayFloats + 8 + (L.ray_src == RAY_DIR ? kMaxRayFeat : 0);
            for (int k = 0; k < L.ray_k; ++k) acc = fmaf(f32[L.rayw_ofs + k * L.n_out + col], feat[k], acc);
            rb[(static_cast<int>(L.ray_slot) * G + g) * kWidth + col] = acc;
          }
        }
        named_bar_sync(1, kEpiThreads);

        // sample point of this thread's row in tile t (coarse: as given; fine: o + d * z with separate mul / add)
        auto tile_point = [&](int t, bool store, float& x, float& y, float& z) {
          const int R = t * kTileRows + c.row;
          const bool in_rows = R < G * n;
          const int g = in_rows ? R / n : 0;
          const int s = in_rows ? R - g * n : 0;
          const int64_t ri = ray0 + g;
          x = y = z = 0.f;
          if (!(in_rows && ri < P.n_rays)) return;
          if (pass == 0) {
            const float* ps = P.io.ray_samples + (ri * nc + s) * 3;
            x = __ldcs(ps); y = __ldcs(ps + 1); z = __ldcs(ps + 2);
          } else {
            const float* r = ray + g * kRayFloats;
            const float zz = zf[g * na + s];
            x = __fadd_rn(r[0], __fmul_rn(r[3], zz)); y = __fadd_rn(r[1], __fmul_rn(r[4], zz)); z = __fadd_rn(r[2], __fmul_rn(r[5], zz));
            if (store && c.cg == 0 && P.io.samples_out) { float* po = P.io.samples_out + (ri * na + s) * 3; __stcs(po, x); __stcs(po + 1, y); __stcs(po + 2, z); }
          }
        };
        // smpl: the warp net's input encoding of tile t is written one step EARLY -- for tile 0 here, for tile t + 1
        // between the last two layers of tile t -- into activation chunk 2 (dead from the dir layer's MMAs on), so the
        // warp MMA of the next tile runs under the rgb head of this one instead of after it
        const uint32_t wpe_tile = smem_u32(sm.base) + kOffA + 2u * 2u * kChunkBytes;
        float px = 0.f, py = 0.f, pz = 0.f;     // this row's point in the tile whose warp input was encoded last
        auto warp_encode = [&](int t) {
          float x, y, z;
          if (c.tid == 0) trace_ev(P, 31, c.layer_ctr);
          tile_point(t, true, x, y, z);
          px = x; py = y; pz = z;
          write_encoding(wpe_tile, c.row, c.cg, x, y, z, P.warp.in_freqs, P.warp.in_identity, fast);
          if (c.tid == 0) trace_ev(P, 32, c.layer_ctr);
          epi_publish(sm, c, 2);
          if (c.tid == 0) trace_ev(P, 33, c.layer_ctr);
        };
        // nerf / append kinds: the same one step early for the xyz encoding that feeds the first layer (the aux tile's
        // last reader, a skip layer, is long done by then)
        auto pos_encode = [&](int t) {
          float x, y, z;
          if (c.tid == 0) trace_ev(P, 36, c.layer_ctr);
          tile_point(t, true, x, y, z);
          write_encoding(aux_tile, c.row, c.cg, x, y, z, net.in_freqs, net.in_identity, fast);
          if (c.tid == 0) trace_ev(P, 37, c.layer_ctr);
          epi_publish(sm, c, kSrcAux);
          if (c.tid == 0) trace_ev(P, 38, c.layer_ctr);
        };
        if (smpl) warp_encode(0);
        else pos_encode(0);

        for (int t = 0; t < tiles; ++t) {
          const int R = t * kTileRows + c.row;
          const bool in_rows = R < G * n;
          const int g = in_rows ? R / n : 0;
          const int s = in_rows ? R - g * n : 0;
          const int64_t ri = ray0 + g;
          const bool valid = in_rows && ri < P.n_rays;
          const float* r = ray + g * kRayFloats;
          float x, y, z;
          if (c.tid == 0) trace_ev(P, 30, c.layer_ctr);
          x = px; y = py; z = pz;                    // smpl: computed by warp_encode(t) one step earlier (else unused)
          float ux = 0.f, uy = 0.f, uz = 1.f;   // unit view direction of this sample (smpl)
          float dnorm_s = r[6];                 // |direction| that scales this sample's delta (utils.py:165-167)
          HeadOut ho = {0.f, 0.f, 0.f, 0.f};
          if (smpl) {
            // ---- warp field: x -> x + W2 relu(W1 [enc(x), pose] + b1) + b2   (its input was encoded one step earlier)
            const float* wf32 = reinterpret_cast<const float*>(P.blob[2] + P.warp.f32_ofs);
            c.head_s = hw_s;
            epilogue_layer<true, false, true, false>(sm, P, P.warp, wf32, rbw, P.warp.layers[0], c, g, ho);
            // all four threads of a row need the full 256-column dot products: exchange the column-group
            // partials through smem and add them in the same order -> identical bits in every thread
            // (its own table, xchg_w: since the next tile's warp input is encoded early there is no CTA-wide barrier
            //  any more between the tile-end reads of xchg and this write)
            xchg_w[kTileRows * c.cg] = make_float4(ho.h0, ho.h1, ho.h2, 0.f);
            named_bar_sync(1, kEpiThreads);
            if (c.tid == 0) trace_ev(P, 35, c.layer_ctr);
            float w0, w1, w2;
            {
              const float4 p0 = xchg_w[0], p1 = xchg_w[kTileRows], p2 = xchg_w[2 * kTileRows], p3 = xchg_w[3 * kTileRows];
              w0 = __fadd_rn(__fadd_rn(__fadd_rn(p0.x, p1.x), __fadd_rn(p2.x, p3.x)), wb0);
              w1 = __fadd_rn(__fadd_rn(__fadd_rn(p0.y, p1.y), __fadd_rn(p2.y, p3.y)), wb1);
              w2 = __fadd_rn(__fadd_rn(__fadd_rn(p0.z, p1.z), __fadd_rn(p2.z, p3.z)), wb2);
            }
            const float wx = __fadd_rn(x, w0), wy = __fadd_rn(y, w1), wz = __fadd_rn(z, w2);
            if (valid && last_pass && c.cg == 1) {
              if (P.io.warp_out) { float* po = P.io.warp_out + (ri * n + s) * 3; __stcs(po, w0); __stcs(po + 1, w1); __stcs(po + 2, w2); }
              if (P.io.warped_out) { float* po = P.io.warped_out + (ri * n + s) * 3; __stcs(po, wx); __stcs(po + 1, wy); __stcs(po + 2, wz); }
            }
            x = wx; y = wy; z = wz;
            const float dx = __fsub_rn(x, r[0]), dy = __fsub_rn(y, r[1]), dz = __fsub_rn(z, r[2]);
            const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
            if (in_rows) { ux = __fdiv_rn(dx, nrm); uy = __fdiv_rn(dy, nrm); uz = __fdiv_rn(dz, nrm); }
            if (pass == 0) dnorm_s = nrm;       // coarse pass: per-sample |warped - o| (smpl_nerf_pipeline.py:52,63); fine: |ray_direction| (:95-98)
            // (no barrier needed before the exchange slots are reused: every later writer of A chunk 3
            //  sits behind an mbarrier that all 16 warps arrive on only after these reads)
          }
          if (smpl) {
            // ---- encoded (warped) position -> aux (first layer and skip layer read it)
            if (c.tid == 0) trace_ev(P, 36, c.layer_ctr);
            write_encoding(aux_tile, c.row, c.cg, x, y, z, net.in_freqs, net.in_identity, fast);
            if (c.tid == 0) trace_ev(P, 37, c.layer_ctr);
            epi_publish(sm, c, kSrcAux);
            if (c.tid == 0) trace_ev(P, 38, c.layer_ctr);
          }

          ho = {0.f, 0.f, 0.f, 0.f};
          float sigma_part = 0.f;
          c.head_s = hr_s;
          for (int l = 0; l < net.n_layers; ++l) {
            const Layer& L = net.layers[l];
            HeadOut hl = {0.f, 0.f, 0.f, 0.f};
            if (l == net.n_layers - 1 && t + 1 < tiles) {
              if (smpl) warp_encode(t + 1);      // chunks 2, 3 are dead: the rgb layer reads 0 and 1
              else pos_encode(t + 1);
            }
            epilogue_dispatch(sm, P, net, f32, rb, L, c, g, hl);
            if (L.flags & LF_SIGMA_HEAD) sigma_part = hl.sig;
            if (L.epi == EPI_RGB) ho = hl;
            if (L.flags & LF_WRITE_DIRPE) {
              // this layer was the last reader of the xyz encoding and its MMAs are done: aux takes the per-sample
              // direction encoding now (the dir layer consumes it several layers later)
              write_encoding(aux_tile, c.row, c.cg, ux, uy, uz, net.dir_freqs, net.dir_identity, fast);
              epi_publish(sm, c, kSrcAux);
            }
          }
          // ---- combine the four column-group partials of the heads: raw = (rgb_raw, sigma_raw)
          xchg[kTileRows * c.cg] = make_float4(ho.h0, ho.h1, ho.h2, sigma_part);
          named_bar_sync(1, kEpiThreads);
          if (c.tid == 0) trace_ev(P, 39, c.layer_ctr);
          {
            // the four threads of a row split the per-sample half of raw2outputs: thread cg owns component cg of
            // (rgb_raw[3], sigma_raw) -- sums the four partials in a fixed order, adds the bias, applies the
            // sigmoid (cg < 3) or the alpha formula (cg = 3) and writes its scalar of raw4[R]
            const float* qx = reinterpret_cast<const float*>(xchg) + c.cg;
            const float q0 = qx[0], q1 = qx[4 * kTileRows], q2 = qx[8 * kTileRows], q3 = qx[12 * kTileRows];
            const float bias = c.cg == 0 ? hb0 : (c.cg == 1 ? hb1 : (c.cg == 2 ? hb2 : sb));
            const float val = __fadd_rn(__fadd_rn(__fadd_rn(q0, q1), __fadd_rn(q2, q3)), bias);
            if (valid) {
              float* tap = pass == 0 ? P.io.raw_coarse : P.io.raw_fine;
              if (tap) tap[(ri * n + s) * 4 + c.cg] = val;
            }
            if (in_rows) {
              float out;
              if (c.cg < 3) out = sigmoidf_ref(val);
              else {
                const float* zz = pass == 0 ? zc + g * nc : zf + g * na;
                const float dz = (s < n - 1) ? __fsub_rn(zz[s + 1], zz[s]) : 1e10f;
                const float* nb = pass == 0 ? P.io.noise_coarse : P.io.noise_fine;
                out = alpha_sample(val, dz, dnorm_s, (valid && nb) ? nb + ri * n + s : nullptr);
                if (valid && last_pass && P.io.alpha_out) __stcs(P.io.alpha_out + ri * n + s, out);
              }
              reinterpret_cast<float*>(raw4 + R)[c.cg] = out;
            }
          }
          // (xchg is next written at the next tile's end, or as part of activation chunk 3 by the first layer's epilogue:
          //  both sit behind CTA-wide publish barriers that every warp reaches only after these reads)
        }
        named_bar_sync(1, kEpiThreads);   // raw4 of every tile of this pass is complete
        if (c.tid == 0) trace_ev(P, 40, c.layer_ctr);

        // ---- per-ray: compositing (+ sampling after the coarse pass) by a team of 16 / G warps per ray.
        //      Scratch lives in the activation operand region, which is dead between the tiles of two passes.
        {
          const int W = kEpiWarps / G;
          if (ew < G * W) {
            const int g = ew / W;
            const RayTeam tm = {32 * (ew - g * W) + lane, 32 * W, ew - g * W, W, lane, static_cast<uint32_t>(3 + g)};
            const int64_t ri = ray0 + g;
            const bool valid = ri < P.n_rays;
            float* scr = reinterpret_cast<float*>(sm.base + kOffA) + g * kRayScratchFloats;
            float* tf = scr, *tt = scr + kMaxFineRows, *cdfx = scr + 2 * kMaxFineRows, *pd = cdfx + kTileRows + 16, *zs = pd + kTileRows;
            float* ts = zs + kMaxFineRows;
            float* rgb_dst = valid ? (pass == 0 ? P.io.rgb : P.io.rgb_fine) : nullptr;
            if (rgb_dst) rgb_dst += ri * 3;
            float* w_dst = (valid && pass == 0 && P.io.weights_coarse) ? P.io.weights_coarse + ri * nc : nullptr;
            composite_ray_activated(raw4 + g * n, n, P.white_bkgd, rgb_dst, w_dst, tf, tt, ts, tm);
            if (pass == 0 && P.run_fine) {
              float* zfg = zf + g * na;
              if (P.io.z_all_in) {      // teacher forcing; padding rays get a sorted ramp (nothing of theirs is stored)
                for (int i = tm.t; i < na; i += tm.T) zfg[i] = valid ? P.io.z_all_in[ri * na + i] : static_cast<float>(i);
                tm.sync();
              } else {
                sample_ray(&raw4[g * nc].w, 4, zc + g * nc, nc, nf, u_s, cdfx, pd, zs, zfg,
                           (valid && P.io.z_new) ? P.io.z_new + ri * nf : nullptr, ts, tm);
              }
              if (valid && P.io.z_all) for (int i = tm.t; i < na; i += tm.T) P.io.z_all[ri * na + i] = zfg[i];
            }
          }
        }
        named_bar_sync(1, kEpiThreads);
        if (c.tid == 0) trace_ev(P, 41, c.layer_ctr);
      }
    }
    const uint32_t am = *reinterpret_cast<const uint32_t*>(&c.amax2);
    if (((am & 0xFFFFu) >= 0x7BFFu || (am >> 16) >= 0x7BFFu) && P.io.status) atomicOr(P.io.status, 1);
  }

  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0 && P.io.trace && blockIdx.x == 0) P.io.trace[1] = *reinterpret_cast<uint32_t*>(smem_raw + P.off_misc + kTraceCtrOfs);
  if (threadIdx.x == 0 && blockIdx.x == 0 && P.io.status) {      // a weight that left the fp16 range at pack time is as fatal as an activation that does
    int bad = 0;
    if (P.blob[0]) bad |= *(reinterpret_cast<const int32_t*>(P.blob[0] + P.net[0].f32_ofs) + P.net[0].flag_ofs);
    if (P.blob[1]) bad |= *(reinterpret_cast<const int32_t*>(P.blob[1] + P.net[1].f32_ofs) + P.net[1].flag_ofs);
    if (P.blob[2]) bad |= *(reinterpret_cast<const int32_t*>(P.blob[2] + P.warp.f32_ofs) + P.warp.flag_ofs);
    if (bad) atomicOr(P.io.status, 1);
  }
  cluster_sync_all();          // neither CTA may exit (or free TMEM) while its peer can still address it
  if (warp == 1) tmem_dealloc2<512>(tmem_base);
}

// ---------------------------------------------------------------------------------- host launcher
// The weight stream of a packed net as a 2-D uint8 tensor [rows, 128 B] (rows = 128-byte swizzle rows of the
// pre-swizzled stages), box = 64 rows x 128 B, no swizzle/interleave: a plain strided copy of 8 KB per request.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_stream_map(CUtensorMap* map, const void* blob, uint32_t stream_bytes) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NRF_E_CUDA; }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t dims[2] = {128, stream_bytes / 128u};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t box[2] = {128, 64};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(blob), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r)); return NRF_E_CUDA; }
  return NRF_OK;
}

static int check_ptr(const void* p, const char* name) {
  if (!p) { set_error("%s is NULL", name); return NRF_E_INVALID; }
  return NRF_OK;
}

}  // namespace nrf

using namespace nrf;

extern "C" int nrf_render_launches(void) { return 1; }

extern "C" int nrf_render(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const void* packed_coarse,
                          const NrfRayNetDesc* fine, const void* packed_fine, const NrfWarpNetDesc* warp,
                          const void* packed_warp, const NrfRenderIO* io, int64_t n_rays, int n_sms, void* stream) {
  if (!pipe || !coarse || !io) { set_error("pipe/coarse/io is NULL"); return NRF_E_INVALID; }
  if (n_rays < 0) { set_error("n_rays < 0"); return NRF_E_INVALID; }
  if (n_rays == 0) return NRF_OK;
  static thread_local RenderParams P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = plan_raynet(coarse, &P.net[0])) != NRF_OK) return rc;
  if ((rc = check_ptr(packed_coarse, "packed_coarse")) != NRF_OK) return rc;
  P.blob[0] = static_cast<const uint8_t*>(packed_coarse);
  const bool smpl = pipe->kind == NRF_KIND_SMPL;
  if (pipe->kind != NRF_KIND_NERF && pipe->kind != NRF_KIND_SMPL && pipe->kind != NRF_KIND_APPEND) { set_error("unknown pipeline kind %d", pipe->kind); return NRF_E_INVALID; }
  if (pipe->run_fine) {
    if (!fine) { set_error("run_fine=1 needs the fine net"); return NRF_E_INVALID; }
    if ((rc = plan_raynet(fine, &P.net[1])) != NRF_OK) return rc;
    if ((rc = check_ptr(packed_fine, "packed_fine")) != NRF_OK) return rc;
    P.blob[1] = static_cast<const uint8_t*>(packed_fine);
  }
  if (smpl) {
    if (!warp) { set_error("smpl pipeline needs the warp net"); return NRF_E_INVALID; }
    if ((rc = plan_warpnet(warp, &P.warp)) != NRF_OK) return rc;
    if ((rc = check_ptr(packed_warp, "packed_warp")) != NRF_OK) return rc;
    P.blob[2] = static_cast<const uint8_t*>(packed_warp);
    if (!coarse->per_sample_dirs || (pipe->run_fine && !fine->per_sample_dirs)) { set_error("smpl pipeline needs per_sample_dirs=1 nets"); return NRF_E_INVALID; }
    if (pipe->run_fine && !pipe->pose_encoded) { set_error("smpl pipeline with run_fine=1 requires human_pose_encoding=1 (the reference feeds the warp net encoded inputs in the fine pass, smpl_nerf_pipeline.py:71-77)"); return NRF_E_INVALID; }
  } else if (coarse->per_sample_dirs) { set_error("per_sample_dirs=1 is only valid for the smpl pipeline"); return NRF_E_INVALID; }
  for (int i = 0; i < 3; ++i) if (P.blob[i] && (reinterpret_cast<uintptr_t>(P.blob[i]) & 1023u)) { set_error("packed buffers must be 1024-byte aligned"); return NRF_E_INVALID; }
  for (int i = 0; i < 3; ++i) {
    if (!P.blob[i]) continue;
    const NetPlan& np = i < 2 ? P.net[i] : P.warp;
    if ((rc = encode_stream_map(&P.tmap[i], P.blob[i], np.stream_bytes)) != NRF_OK) return rc;
  }

  const int nc = pipe->n_coarse, nf = pipe->run_fine ? pipe->n_fine : 0;
  if (nc < 16 || nc > kTileRows) { set_error("n_coarse %d unsupported (16..%d)", nc, kTileRows); return NRF_E_INVALID; }
  if (pipe->run_fine && nf < 1) { set_error("n_fine %d invalid", nf); return NRF_E_INVALID; }
  if (pipe->run_fine && nc < 3) { set_error("hierarchical sampling needs n_coarse >= 3"); return NRF_E_INVALID; }
  const int G = kTileRows / nc, na = nc + nf;
  if (G * na > kMaxFineRows) { set_error("n_coarse + n_fine = %d too large for %d rays per group (max %d rows)", na, G, kMaxFineRows); return NRF_E_INVALID; }
  P.io = *io;
  P.n_rays = n_rays;
  P.kind = pipe->kind; P.n_coarse = nc; P.n_fine = nf; P.n_all = na; P.run_fine = pipe->run_fine ? 1 : 0;
  P.white_bkgd = pipe->white_background ? 1 : 0; P.fast = pipe->precision == 1 ? 1 : 0;
  P.pose_freqs = pipe->pose_freqs; P.pose_identity = pipe->pose_identity; P.pose_encoded = pipe->pose_encoded ? 1 : 0;
  P.pose_stride = pipe->pose_stride; P.pose_col0 = pipe->pose_col0; P.pose_col1 = pipe->pose_col1;
  P.pose_dim = 0;
  const bool ext_pose = pipe->kind == NRF_KIND_APPEND && P.net[0].n_ext_slots > 0;
  if (ext_pose) {
    if (pipe->run_fine && P.net[1].n_ext_slots != P.net[0].n_ext_slots) { set_error("coarse/fine nets disagree on ext_pose_bias"); return NRF_E_INVALID; }
    if ((rc = check_ptr(io->ray_bias_coarse, "ray_bias_coarse")) != NRF_OK) return rc;
    if (pipe->run_fine && (rc = check_ptr(io->ray_bias_fine, "ray_bias_fine")) != NRF_OK) return rc;
  } else if (pipe->kind != NRF_KIND_NERF) {
    P.pose_dim = pipe->pose_encoded ? 2 * (2 * pipe->pose_freqs + (pipe->pose_identity ? 1 : 0)) : 2;
    const int want = smpl ? warp->pose_dim : coarse->additional_input_dim;
    if (P.pose_dim != want) { set_error("pose feature count %d does not match the net's pose input dim %d", P.pose_dim, want); return NRF_E_INVALID; }
    if (P.pose_dim > kMaxRayFeat) { set_error("pose feature count %d > %d", P.pose_dim, kMaxRayFeat); return NRF_E_INVALID; }
    if ((rc = check_ptr(io->goal_pose, "goal_pose")) != NRF_OK) return rc;
    if (pipe->pose_col0 < 0 || pipe->pose_col1 < 0 || pipe->pose_col0 >= pipe->pose_stride || pipe->pose_col1 >= pipe->pose_stride) { set_error("pose columns out of range"); return NRF_E_INVALID; }
  } else if (coarse->additional_input_dim != 0) { set_error("nerf pipeline with additional_input_dim != 0"); return NRF_E_INVALID; }
  if (pipe->run_fine && (fine->additional_input_dim != coarse->additional_input_dim)) { set_error("coarse/fine additional_input_dim differ"); return NRF_E_INVALID; }
  if ((rc = check_ptr(io->ray_samples, "ray_samples")) || (rc = check_ptr(io->ray_origin, "ray_origin")) ||
      (rc = check_ptr(io->ray_dir, "ray_dir")) || (rc = check_ptr(io->z_vals, "z_vals")) || (rc = check_ptr(io->rgb, "rgb"))) return rc;
  if (pipe->run_fine) {
    if ((rc = check_ptr(io->rgb_fine, "rgb_fine")) != NRF_OK) return rc;
    if (!io->z_all_in && (rc = check_ptr(io->u_fine, "u_fine")) != NRF_OK) return rc;
  }
  P.G = G;
  P.tiles_f = (G * na + kTileRows - 1) / kTileRows;
  P.n_groups = static_cast<int32_t>((n_rays + G - 1) / G);
  P.n_pairs = (P.n_groups + 1) / 2;

  // shared-memory layout of the misc area (floats, 16-byte aligned pieces)
  uint32_t f = 0;
  auto take = [&](uint32_t n) { uint32_t o = f; f = (f + n + 3u) & ~3u; return o; };
  const int slots = P.net[0].n_ray_slots > P.net[1].n_ray_slots ? P.net[0].n_ray_slots : P.net[1].n_ray_slots;
  P.o_ray = take(G * kRayFloats);
  P.o_rb = take(slots * G * kWidth);
  P.o_rbw = take(smpl ? G * kWidth : 0);
  P.o_raw = take(G * na * 4);
  P.o_zc = take(G * nc);
  P.o_zf = take(G * na);
  P.o_u = take(nf > 0 ? nf : 4);
  P.o_hw = take(smpl ? 3 * kWidth : 0);
  P.o_hr = take(3 * (kWidth / 2));
  P.o_hs = take(kWidth);
  P.n_stages = kMaxStages;
  while (P.n_stages > 2 && kOffRing + static_cast<uint32_t>(P.n_stages) * kSlotBytes + kBarBytes + f * 4 > kSmemLimit) --P.n_stages;   // big per-ray tables: shorter weight ring
  P.off_misc = kOffRing + static_cast<uint32_t>(P.n_stages) * kSlotBytes;
  const uint32_t smem_bytes = P.off_misc + kBarBytes + f * 4;
  if (smem_bytes > kSmemLimit) { set_error("configuration needs %u bytes of shared memory per CTA (limit %u)", smem_bytes, kSmemLimit); return NRF_E_INVALID; }

  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (n_sms > 0 && n_sms < sms) sms = n_sms;
  const int pairs = P.n_pairs < sms / 2 ? P.n_pairs : sms / 2;
  if (pairs < 1) { set_error("the engine needs at least 2 SMs (CTA pairs)"); return NRF_E_INVALID; }
  const int grid = 2 * pairs;
  e = cudaFuncSetAttribute(nrf_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(max dynamic smem)");
  nrf_fused_kernel<<<grid, kThreads, smem_bytes, static_cast<cudaStream_t>(stream)>>>(P);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "nrf_fused_kernel launch");
  return NRF_OK;
}
