// Stand-alone ops with the same arithmetic as the fused kernel's stages, exported so that the
// reference's utils.py functions and torchsearchsorted can be replaced one by one:
//   nrf_positional_encoding  <- utils.py:127-131  PositionalEncoder.encode
//   nrf_raw2outputs          <- utils.py:134-191  raw2outputs
//   nrf_sample_pdf           <- utils.py:194-228  sample_pdf
//   nrf_searchsorted         <- torchsearchsorted/src/cuda/searchsorted_cuda_kernel.cu:83-142
// plus nrf_selftest_umma, a one-tile tcgen05 GEMM through the renderer's operand layout.
// These are HBM-bound streaming kernels: coalesced loads, one warp per ray where a scan is needed.
#include <cuda_runtime.h>

#include "nrf_plan.h"
#include "nrf_ptx.cuh"
#include "nrf_stages.cuh"

namespace nrf {

// ------------------------------------------------------------------------------ positional encoding
// out[i, :] = [x_i ?] ++ for k: sin(2^k x_i) ++ cos(2^k x_i)    (x_i has c components)
__global__ void pe_kernel(const float* __restrict__ x, int64_t n, int c, int freqs, int identity, float* __restrict__ out) {
  const int width = c * (2 * freqs + (identity ? 1 : 0));
  const int64_t total = n * width;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / width;
    int col = static_cast<int>(idx - row * width);
    float val;
    if (identity && col < c) val = x[row * c + col];
    else {
      col -= identity ? c : 0;
      const int k = col / (2 * c), rem = col - k * 2 * c;
      const float a = x[row * c + rem % c] * static_cast<float>(1u << k);
      val = rem < c ? sinf(a) : cosf(a);
    }
    out[idx] = val;
  }
}

// backward of the encoding: g_x[i, j] = [identity] g[i, j] + sum_k 2^k (cos(2^k x) g_sin[k] - sin(2^k x) g_cos[k])
__global__ void pe_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, int64_t n, int c, int freqs, int identity,
                              float* __restrict__ gx) {
  const int width = c * (2 * freqs + (identity ? 1 : 0));
  const int64_t total = n * c;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = idx / c;
    const int j = static_cast<int>(idx - row * c);
    const float v = x[idx];
    const float* gr = g + row * width;
    float acc = identity ? gr[j] : 0.f;
    const int base = identity ? c : 0;
    for (int k = 0; k < freqs; ++k) {
      const float f = static_cast<float>(1u << k), a = v * f;
      acc = fmaf(f, fmaf(cosf(a), gr[base + k * 2 * c + j], -sinf(a) * gr[base + k * 2 * c + c + j]), acc);
    }
    gx[idx] = acc;
  }
}

// ------------------------------------------------------------------------------ raw2outputs
// one warp per ray; raw staged in smem as float4
__global__ void raw2outputs_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ dirs,
                                   const float* __restrict__ noise, int64_t B, int n, int white, float* __restrict__ rgb,
                                   float* __restrict__ weights, float* __restrict__ alpha) {
  extern __shared__ __align__(16) float smem[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n4 = (n + 3) & ~3;
  float4* raw4 = reinterpret_cast<float4*>(smem) + static_cast<size_t>(w) * n;
  float* tf = smem + static_cast<size_t>(wpb) * n * 4 + static_cast<size_t>(w) * (4 * n4 + kTeamScratch);   // 16-byte aligned scan scratch
  float* tt = tf + n4;
  float* ts = tf + 4 * n4;
  float* zs = tt + n4;
  float* dn = zs + n4;
  for (int64_t ray = blockIdx.x * static_cast<int64_t>(wpb) + w; ray < B; ray += static_cast<int64_t>(gridDim.x) * wpb) {
    for (int i = lane; i < n; i += 32) {
      raw4[i] = *reinterpret_cast<const float4*>(raw + (ray * n + i) * 4);
      zs[i] = z[ray * n + i];
      const float* d = dirs + (ray * n + i) * 3;
      dn[i] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    }
    __syncwarp();
    composite_ray(raw4, zs, dn, 0.f, n, noise ? noise + ray * n : nullptr, white, rgb + ray * 3, alpha + ray * n,
                  weights + ray * n, tf, tt, ts, lane);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------ raw2outputs backward
// d(loss)/d(raw) of utils.py:134-191 given d(loss)/d(rgb) [B,3] and, optionally, d(loss)/d(weights) [B,n] and
// d(loss)/d(alpha) [B,n] (the reference's GMM density loss reads the densities).  With c = sigmoid(rgb_raw),
// a = 1 - exp(-relu(s) delta), keep = 1 - a + 1e-10, T_i = prod_{j<i} keep_j, w = a T:
//   gw_i = g_rgb . c_i  - [white] sum(g_rgb) + g_weights_i
//   d/dc_i = w_i g_rgb                          d/da_i = gw_i T_i - (sum_{k>i} gw_k w_k) / keep_i + g_alpha_i
//   d/ds_i = d/da_i * delta_i (1 - a_i) [s_i > 0]          d/drgb_raw = d/dc * c (1 - c)
// One warp per ray; the forward quantities are recomputed (nothing is saved by the forward), the transmittance with
// the same sequential product as the forward, the suffix sum sequentially from the far end.
__global__ void raw2outputs_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ dirs,
                                       const float* __restrict__ noise, int64_t B, int n, int white, const float* __restrict__ g_rgb,
                                       const float* __restrict__ g_w, const float* __restrict__ g_a, float* __restrict__ g_raw) {
  extern __shared__ __align__(16) float smem[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n4 = (n + 3) & ~3;
  float4* act4 = reinterpret_cast<float4*>(smem) + static_cast<size_t>(w) * n;
  float* base = smem + static_cast<size_t>(wpb) * n * 4 + static_cast<size_t>(w) * 5 * n4;
  float* keep = base; float* T = keep + n4; float* gw = T + n4; float* suf = gw + n4; float* dsig = suf + n4;
  for (int64_t ray = blockIdx.x * static_cast<int64_t>(wpb) + w; ray < B; ray += static_cast<int64_t>(gridDim.x) * wpb) {
    const float gr = g_rgb[ray * 3], gg = g_rgb[ray * 3 + 1], gb = g_rgb[ray * 3 + 2];
    for (int i = lane; i < n; i += 32) {
      const float4 r = *reinterpret_cast<const float4*>(raw + (ray * n + i) * 4);
      const float* d = dirs + (ray * n + i) * 3;
      const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
      const float dz = (i < n - 1) ? __fsub_rn(z[ray * n + i + 1], z[ray * n + i]) : 1e10f;
      const float delta = __fmul_rn(dz, nrm);
      const float s = noise ? __fadd_rn(r.w, noise[ray * n + i]) : r.w;
      const float e = expf(-__fmul_rn(fmaxf(s, 0.f), delta));          // = 1 - alpha
      const float a = __fsub_rn(1.f, e);
      act4[i] = make_float4(sigmoidf_ref(r.x), sigmoidf_ref(r.y), sigmoidf_ref(r.z), a);
      keep[i] = __fadd_rn(__fsub_rn(1.f, a), 1e-10f);
      dsig[i] = s > 0.f ? delta * e : 0.f;
    }
    __syncwarp();
    if (lane == 0) serial_scan<true>(keep, T, n);
    __syncwarp();
    const float bg = white ? (gr + gg + gb) : 0.f;
    for (int i = lane; i < n; i += 32) {
      const float4 c = act4[i];
      const float g = gr * c.x + gg * c.y + gb * c.z - bg + (g_w ? g_w[ray * n + i] : 0.f);
      gw[i] = g;
      suf[i] = g * (c.w * T[i]);
    }
    __syncwarp();
    if (lane == 0) {          // exclusive suffix sum, from the far end
      float run = 0.f;
      for (int i = n - 1; i >= 0; --i) { const float v = suf[i]; suf[i] = run; run += v; }
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const float4 c = act4[i];
      const float wi = c.w * T[i];
      const float ga = gw[i] * T[i] - suf[i] / keep[i] + (g_a ? g_a[ray * n + i] : 0.f);
      float4 o;
      o.x = wi * gr * c.x * (1.f - c.x);
      o.y = wi * gg * c.y * (1.f - c.y);
      o.z = wi * gb * c.z * (1.f - c.z);
      o.w = ga * dsig[i];
      *reinterpret_cast<float4*>(g_raw + (ray * n + i) * 4) = o;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------ sample_pdf
// one warp per ray: pdf -> sequential cdf -> right-sided bisection -> guarded lerp
__global__ void sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, const float* __restrict__ u,
                                  int64_t B, int m, int nf, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* cdf = smem + static_cast<size_t>(w) * 2 * m;
  float* bn = cdf + m;
  for (int64_t ray = blockIdx.x * static_cast<int64_t>(wpb) + w; ray < B; ray += static_cast<int64_t>(gridDim.x) * wpb) {
    float part = 0.f;
    for (int i = lane; i < m; i += 32) bn[i] = bins[ray * m + i];
    for (int i = 1 + lane; i < m; i += 32) { const float v = __fadd_rn(weights[ray * (m - 1) + i - 1], 1e-5f); cdf[i] = v; part += v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __syncwarp();
    for (int i = 1 + lane; i < m; i += 32) cdf[i] = __fdiv_rn(cdf[i], part);
    __syncwarp();
    if (lane == 0) { float run = 0.f; cdf[0] = 0.f; for (int i = 1; i < m; ++i) { run = __fadd_rn(run, cdf[i]); cdf[i] = run; } }
    __syncwarp();
    for (int j = lane; j < nf; j += 32) {
      const float uu = u[j];
      int lo = 0, hi = m;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= uu) lo = mid + 1; else hi = mid; }
      const int below = max(0, lo - 1), above = min(m - 1, lo);
      const float c0 = cdf[below], c1 = cdf[above], b0 = bn[below], b1 = bn[above];
      float denom = __fsub_rn(c1, c0);
      if (denom < 1e-5f) denom = 1.f;
      const float t = __fdiv_rn(__fsub_rn(uu, c0), denom);
      out[ray * nf + j] = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------ fine_sampling
// utils.py:231-264: z_mid, sample_pdf, sort(cat(z, z_samples)), points = o + d * z   (one warp per ray)
__global__ void fine_sampling_kernel(const float* __restrict__ origin, const float* __restrict__ dir, const float* __restrict__ z,
                                     const float* __restrict__ weights, const float* __restrict__ u, int64_t B, int nc, int nf,
                                     float* __restrict__ z_all, float* __restrict__ pts) {
  extern __shared__ __align__(16) float smem[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int na = nc + nf;
  const int nc4 = (nc + 3) & ~3, nf4 = (nf + 3) & ~3;
  float* base = smem + static_cast<size_t>(w) * (5 * nc4 + 2 * nf4 + 4 + kTeamScratch);
  float* ts = base + 5 * nc4 + 2 * nf4 + 4;
  const RayTeam tm = {lane, 32, 0, 1, lane, 0u};
  float* cdfx = base; float* pd = cdfx + nc4 + 4; float* zc = pd + nc4; float* wc = zc + nc4; float* zs = wc + nc4; float* zf = zs + nf4;
  for (int64_t ray = blockIdx.x * static_cast<int64_t>(wpb) + w; ray < B; ray += static_cast<int64_t>(gridDim.x) * wpb) {
    for (int i = lane; i < nc; i += 32) { zc[i] = z[ray * nc + i]; wc[i] = weights[ray * nc + i]; }
    __syncwarp();
    sample_ray(wc, 1, zc, nc, nf, u, cdfx, pd, zs, zf, nullptr, ts, tm);
    const float ox = origin[ray * 3], oy = origin[ray * 3 + 1], oz = origin[ray * 3 + 2];
    const float dx = dir[ray * 3], dy = dir[ray * 3 + 1], dz = dir[ray * 3 + 2];
    for (int i = lane; i < na; i += 32) {
      const float zz = zf[i];
      z_all[ray * na + i] = zz;
      float* po = pts + (ray * na + i) * 3;
      po[0] = __fadd_rn(ox, __fmul_rn(dx, zz)); po[1] = __fadd_rn(oy, __fmul_rn(dy, zz)); po[2] = __fadd_rn(oz, __fmul_rn(dz, zz));
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------ searchsorted
// res[r, c] = #{ j : a[r, j] < v[r, c] } (side left)  or  #{ j : a[r, j] <= v[r, c] } (side right).
// Same results as the reference's bisection for sorted rows; one thread per query, queries of a row
// are contiguous so loads of v and stores of res coalesce.
__global__ void searchsorted_kernel(const float* __restrict__ a, int64_t rows_a, int64_t na, const float* __restrict__ v,
                                    int64_t rows_v, int64_t nv, int64_t* __restrict__ res, int side_left, int64_t rows) {
  const int64_t total = rows * nv;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = idx / nv, c = idx - r * nv;
    const float* ar = a + (rows_a == 1 ? 0 : r) * na;
    const float val = v[(rows_v == 1 ? 0 : r) * nv + c];
    int64_t lo = 0, hi = na;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const float am = ar[mid];
      const bool go_right = side_left ? (am < val) : (am <= val);
      if (go_right) lo = mid + 1; else hi = mid;
    }
    res[idx] = lo;
  }
}

// ------------------------------------------------------------------------------ tcgen05 self-test
// D[128,256] = A[128,64] x B[256,64]^T: A one [128 x 64] SWIZZLE_128B tile, B one [256 x 64] SWIZZLE_128B
// tile, four K=16 tcgen05.mma with N = 256, accumulator read back with tcgen05.ld.
__global__ void __launch_bounds__(128, 1) selftest_umma_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                float* __restrict__ d) {
  extern __shared__ __align__(1024) uint8_t smem_st[];
  uint8_t* sa = smem_st;             // 16 KB
  uint8_t* sb = smem_st + 16384;     // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_st + 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_st + 49152 + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int idx = threadIdx.x; idx < 128 * 64; idx += 128) {
    const int r = idx >> 6, k = idx & 63;
    *reinterpret_cast<__half*>(sa + sw128_offset(r, k)) = __float2half_rn(a[idx]);
  }
  for (int idx = threadIdx.x; idx < 256 * 64; idx += 128) {
    const int r = idx >> 6, k = idx & 63;
    *reinterpret_cast<__half*>(sb + sw128_offset(r, k)) = __float2half_rn(b[idx]);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 256);
    const uint64_t ad = umma_desc_sw128(smem_u32(sa)), bd = umma_desc_sw128(smem_u32(sb));
    for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ss(tmem, ad + 2u * ks, bd + 2u * ks, idesc, ks ? 1u : 0u);
    umma_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after_sync();
  const int row = 32 * (warp & 3) + lane;
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[row * 256 + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// Same through a CTA pair: D[256,256] = A[256,64] x B[256,64]^T with tcgen05.mma.cta_group::2 (M = 256).
// CTA r holds A rows [128r, 128r+128) and B rows (output features) [128r, 128r+128) as one [128 x 64]
// SWIZZLE_128B half-stage -- the renderer's layout; the even CTA issues, the commit multicasts.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
selftest_umma2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d) {
  extern __shared__ __align__(1024) uint8_t smem_st[];
  uint8_t* sa = smem_st;             // 16 KB
  uint8_t* sb = smem_st + 16384;     // 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_st + 32768);       // [0] accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_st + 32768 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int idx = threadIdx.x; idx < 128 * 64; idx += 128) {
    const int r = idx >> 6, k = idx & 63;
    *reinterpret_cast<__half*>(sa + sw128_offset(r, k)) = __float2half_rn(a[(128 * rank + r) * 64 + k]);
    *reinterpret_cast<__half*>(sb + sw128_offset(r, k)) = __float2half_rn(b[(128 * rank + r) * 64 + k]);
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); mbar_init(smem_u32(bar + 1), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc2<256>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(256, 256);
    const uint64_t ad = umma_desc_sw128(smem_u32(sa)), bd = umma_desc_sw128(smem_u32(sb));
    for (uint32_t ks = 0; ks < 4; ++ks) umma2_f16_ss(tmem, ad + 2u * ks, bd + 2u * ks, idesc, ks ? 1u : 0u);
    umma2_commit(smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after_sync();
  const int row = 32 * (warp & 3) + lane;
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[(128 * rank + row) * 256 + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2<256>(tmem);
}

static int grid_for(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  return static_cast<int>(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}


// ------------------------------------------------------------------------------ per-ray pose bias: row-uniformity probe
// (the GEMM itself is tcgen05: nrf_ray_bias in nrf_train.cu)
// flag <- 1 if any feature row differs (bitwise) from row 0
__global__ void rows_differ_kernel(const float* __restrict__ feats, int64_t B, int A, int32_t* __restrict__ flag) {
  const int64_t total = B * A;
  bool differ = false;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x)
    differ |= __float_as_uint(feats[idx]) != __float_as_uint(__ldg(feats + idx % A));
  if (__any_sync(0xffffffffu, differ) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}
int launch_rows_differ(const float* feats, int64_t B, int A, int32_t* flag, cudaStream_t stream) {
  cudaError_t e0 = cudaMemsetAsync(flag, 0, sizeof(int32_t), stream);
  if (e0 != cudaSuccess) return cuda_fail(e0, "cudaMemsetAsync(nonuniform)");
  rows_differ_kernel<<<grid_for(B * A, 256), 256, 0, stream>>>(feats, B, A, flag);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "rows_differ_kernel launch");
}


// ------------------------------------------------------------------------------ SSIM (util/scores.py:88-173)
// Per image plane (one channel of one image, [H, W]): Gaussian-window means / variances / covariance with a valid (no padding)
// ks x ks window, cs = (2 s12 + c2) / (s1 + s2 + c2), ssim = (2 m1 m2 + c1) / (m1^2 + m2^2 + c1) * cs, averaged over the
// (H - ks + 1) x (W - ks + 1) positions.  One 16 x 16 output tile per CTA (inputs staged in shared memory); per-tile sums go
// to `partial` and ssim_reduce_kernel adds them in a fixed order (deterministic, no atomics).
constexpr int kSsimTile = 16, kSsimMaxK = 15;
__global__ void __launch_bounds__(kSsimTile * kSsimTile) ssim_tile_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                                                                           const float* __restrict__ kern, int ks, float c1, float c2,
                                                                           float* __restrict__ partial) {
  __shared__ float xs[(kSsimTile + kSsimMaxK - 1) * (kSsimTile + kSsimMaxK - 1)], ys[(kSsimTile + kSsimMaxK - 1) * (kSsimTile + kSsimMaxK - 1)];
  __shared__ float kw[kSsimMaxK * kSsimMaxK];
  __shared__ float red[2][kSsimTile * kSsimTile / 32];
  const int Ho = H - ks + 1, Wo = W - ks + 1;
  const int tiles_x = (Wo + kSsimTile - 1) / kSsimTile;
  const int ty0 = (blockIdx.x / tiles_x) * kSsimTile, tx0 = (blockIdx.x % tiles_x) * kSsimTile;
  const float* xp = x + static_cast<size_t>(blockIdx.y) * H * W;
  const float* yp = y + static_cast<size_t>(blockIdx.y) * H * W;
  const int span = kSsimTile + ks - 1;
  for (int i = threadIdx.x; i < span * span; i += blockDim.x) {
    const int r = ty0 + i / span, c = tx0 + i % span;
    const bool ok = r < H && c < W;
    xs[i] = ok ? xp[r * W + c] : 0.f;
    ys[i] = ok ? yp[r * W + c] : 0.f;
  }
  for (int i = threadIdx.x; i < ks * ks; i += blockDim.x) kw[i] = kern[i];
  __syncthreads();
  const int oy = threadIdx.x / kSsimTile, ox = threadIdx.x % kSsimTile;
  float sv = 0.f, cv = 0.f;
  if (ty0 + oy < Ho && tx0 + ox < Wo) {
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
    for (int i = 0; i < ks; ++i)
      for (int j = 0; j < ks; ++j) {
        const float w = kw[i * ks + j], a = xs[(oy + i) * span + ox + j], b = ys[(oy + i) * span + ox + j];
        m1 = fmaf(w, a, m1); m2 = fmaf(w, b, m2);
        e11 = fmaf(w, a * a, e11); e22 = fmaf(w, b * b, e22); e12 = fmaf(w, a * b, e12);
      }
    const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    const float s1 = e11 - m11, s2 = e22 - m22, s12 = e12 - m12;
    cv = (2.f * s12 + c2) / (s1 + s2 + c2);
    sv = ((2.f * m12 + c1) / (m11 + m22 + c1)) * cv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sv += __shfl_xor_sync(0xffffffffu, sv, o); cv += __shfl_xor_sync(0xffffffffu, cv, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv; red[1][threadIdx.x >> 5] = cv; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < kSsimTile * kSsimTile / 32; ++i) { a += red[0][i]; b += red[1][i]; }
    float* dst = partial + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 2;
    dst[0] = a; dst[1] = b;
  }
}
__global__ void ssim_reduce_kernel(const float* __restrict__ partial, int n_tiles, float inv_count, float* __restrict__ ssim, float* __restrict__ cs) {
  const int plane = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n_tiles; i += 32) { a += partial[(static_cast<size_t>(plane) * n_tiles + i) * 2]; b += partial[(static_cast<size_t>(plane) * n_tiles + i) * 2 + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (threadIdx.x == 0) { ssim[plane] = static_cast<float>(a * inv_count); if (cs) cs[plane] = static_cast<float>(b * inv_count); }
}

// inference.py:260-262: clip to [0, 1], * 255, truncate to uint8, RGB -> BGR (the reference flips channels before writing)
__global__ void quantize_bgr_kernel(const float* __restrict__ rgb, int64_t n_pix, uint8_t* __restrict__ out, int flip) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_pix; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = fminf(fmaxf(rgb[i * 3 + c], 0.f), 1.f) * 255.f;
      out[i * 3 + (flip ? 2 - c : c)] = static_cast<uint8_t>(v);
    }
  }
}

// ------------------------------------------------------------------------------ ray generation + coarse sampling
// utils.py:26-54 get_rays + datasets/transforms.py:82-89 CoarseSampling + :13-19 ToTensor, for one view:
//   local = ((i - W/2) / focal, -(j - H/2) / focal, -1)       i, j float32 pixel indices, the rest float64
//   dir   = sum_c local[c] * R[k][c]   (products, then (p0 + p1) + p2: numpy's elementwise form, no FMA)
//   z     = lower + (upper - lower) * jitter[ray]               one jitter scalar per ray
//   pts   = origin + dir * z                                     float64, rounded ONCE to fp32 on store
// One thread per (ray, sample); HBM-bound (12 B/sample written).
struct CamParams { double r[9]; double t[3]; double focal; int32_t H, W, n; int64_t ray0, n_rays; };   // rays [ray0, ray0 + n_rays) of the H x W view

__global__ void generate_rays_kernel(const __grid_constant__ CamParams cam, const double* __restrict__ lower, const double* __restrict__ span,
                                     const double* __restrict__ jitter, float* __restrict__ samples, float* __restrict__ origin,
                                     float* __restrict__ dir, float* __restrict__ z_vals) {
  const int64_t total = cam.n_rays * cam.n;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t ray = idx / cam.n;                  // index inside the window (outputs, jitter)
    const int s = static_cast<int>(idx - ray * cam.n);
    const int64_t pix = cam.ray0 + ray;               // pixel of the full view
    const int j = static_cast<int>(pix / cam.W), i = static_cast<int>(pix - static_cast<int64_t>(j) * cam.W);
    const float fi = __fsub_rn(static_cast<float>(i), static_cast<float>(cam.W * .5));
    const float fj = __fsub_rn(static_cast<float>(j), static_cast<float>(cam.H * .5));
    const double l0 = __ddiv_rn(static_cast<double>(fi), cam.focal);
    const double l1 = __ddiv_rn(static_cast<double>(-fj), cam.focal);
    const double l2 = -1.0;
    double d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      d[k] = __dadd_rn(__dadd_rn(__dmul_rn(l0, cam.r[3 * k]), __dmul_rn(l1, cam.r[3 * k + 1])), __dmul_rn(l2, cam.r[3 * k + 2]));
    const double z = __dadd_rn(lower[s], __dmul_rn(span[s], jitter[ray]));
    float* p = samples + idx * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = static_cast<float>(__dadd_rn(cam.t[k], __dmul_rn(d[k], z)));
    z_vals[idx] = static_cast<float>(z);
    if (s == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { origin[ray * 3 + k] = static_cast<float>(cam.t[k]); dir[ray * 3 + k] = static_cast<float>(d[k]); }
    }
  }
}

}  // namespace nrf

using namespace nrf;

extern "C" int nrf_positional_encoding(const float* x, int64_t n, int32_t c, int32_t freqs, int32_t identity, float* out,
                                       void* stream) {
  if (!x || !out) { set_error("x/out is NULL"); return NRF_E_INVALID; }
  if (n < 0 || c < 1 || freqs < 0 || freqs > 24 || (freqs == 0 && !identity)) { set_error("bad positional-encoding shape (n=%lld c=%d L=%d id=%d)", (long long)n, c, freqs, identity); return NRF_E_INVALID; }
  if (n == 0) return NRF_OK;
  const int64_t total = n * c * (2 * freqs + (identity ? 1 : 0));
  pe_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, c, freqs, identity, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "pe_kernel launch");
}

extern "C" int nrf_raw2outputs(const float* raw, const float* z, const float* dirs, const float* noise, int64_t B, int32_t n,
                               int32_t white_background, float* rgb, float* weights, float* alpha, void* stream) {
  if (!raw || !z || !dirs || !rgb || !weights || !alpha) { set_error("raw2outputs: NULL argument"); return NRF_E_INVALID; }
  if (B < 0 || n < 2 || n > 1024) { set_error("raw2outputs: unsupported shape B=%lld n=%d (n in 2..1024)", (long long)B, n); return NRF_E_INVALID; }
  if (B == 0) return NRF_OK;
  const int wpb = 4;
  const size_t smem = static_cast<size_t>(wpb) * (4 * n + 4 * ((n + 3) & ~3) + kTeamScratch) * sizeof(float);
  const int grid = static_cast<int>(B / wpb + 1 > 148 * 8 ? 148 * 8 : B / wpb + 1);
  cudaError_t e = cudaFuncSetAttribute(raw2outputs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  raw2outputs_kernel<<<grid, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(raw, z, dirs, noise, B, n, white_background, rgb, weights, alpha);
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "raw2outputs_kernel launch");
}

extern "C" int nrf_sample_pdf(const float* bins, const float* weights, const float* u, int64_t B, int32_t m, int32_t n_fine,
                              float* samples, void* stream) {
  if (!bins || !weights || !u || !samples) { set_error("sample_pdf: NULL argument"); return NRF_E_INVALID; }
  if (B < 0 || m < 2 || m > 4096 || n_fine < 1) { set_error("sample_pdf: unsupported shape B=%lld m=%d n_fine=%d", (long long)B, m, n_fine); return NRF_E_INVALID; }
  if (B == 0) return NRF_OK;
  const int wpb = 4;
  const size_t smem = static_cast<size_t>(wpb) * 2 * m * sizeof(float);
  const int grid = static_cast<int>(B / wpb + 1 > 148 * 8 ? 148 * 8 : B / wpb + 1);
  cudaError_t e = cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  sample_pdf_kernel<<<grid, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(bins, weights, u, B, m, n_fine, samples);
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "sample_pdf_kernel launch");
}

extern "C" int nrf_fine_sampling(const float* origin, const float* dir, const float* z, const float* weights, const float* u,
                                 int64_t B, int32_t n_coarse, int32_t n_fine, float* z_all, float* pts, void* stream) {
  if (!origin || !dir || !z || !weights || !u || !z_all || !pts) { set_error("fine_sampling: NULL argument"); return NRF_E_INVALID; }
  if (B < 0 || n_coarse < 3 || n_coarse > 1024 || n_fine < 1 || n_fine > 4096) { set_error("fine_sampling: unsupported shape B=%lld n_coarse=%d n_fine=%d", (long long)B, n_coarse, n_fine); return NRF_E_INVALID; }
  if (B == 0) return NRF_OK;
  const int wpb = 4;
  const size_t smem = static_cast<size_t>(wpb) * (5 * ((n_coarse + 3) & ~3) + 2 * ((n_fine + 3) & ~3) + 4 + kTeamScratch) * sizeof(float);
  const int grid = static_cast<int>(B / wpb + 1 > 148 * 8 ? 148 * 8 : B / wpb + 1);
  cudaError_t e = cudaFuncSetAttribute(fine_sampling_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  fine_sampling_kernel<<<grid, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(origin, dir, z, weights, u, B, n_coarse, n_fine, z_all, pts);
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "fine_sampling_kernel launch");
}

extern "C" int nrf_searchsorted(const float* a, int64_t rows_a, int64_t na, const float* v, int64_t rows_v, int64_t nv,
                                int64_t* res, int32_t side_left, void* stream) {
  if (!a || !v || !res) { set_error("searchsorted: NULL argument"); return NRF_E_INVALID; }
  if (rows_a < 1 || rows_v < 1 || na < 0 || nv < 0) { set_error("searchsorted: bad shape"); return NRF_E_INVALID; }
  if (rows_a != rows_v && rows_a != 1 && rows_v != 1) { set_error("searchsorted: `a` and `v` must have the same number of rows or one of them must have only one"); return NRF_E_INVALID; }
  const int64_t rows = rows_a > rows_v ? rows_a : rows_v;
  if (rows * nv == 0) return NRF_OK;
  searchsorted_kernel<<<grid_for(rows * nv, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, rows_a, na, v, rows_v, nv, res, side_left, rows);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "searchsorted_kernel launch");
}

extern "C" int nrf_selftest_umma2(const float* a, const float* b, float* d, void* stream) {
  if (!a || !b || !d) { set_error("selftest: NULL argument"); return NRF_E_INVALID; }
  const int smem = 32768 + 64;
  cudaError_t e = cudaFuncSetAttribute(selftest_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  selftest_umma2_kernel<<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(a, b, d);
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "selftest_umma2_kernel launch");
}

extern "C" int nrf_selftest_umma(const float* a, const float* b, float* d, void* stream) {
  if (!a || !b || !d) { set_error("selftest: NULL argument"); return NRF_E_INVALID; }
  const int smem = 49152 + 64;
  cudaError_t e = cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  selftest_umma_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a, b, d);
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "selftest_umma_kernel launch");
}

extern "C" int nrf_raynet_ext_slots(const NrfRayNetDesc* d) {
  NetPlan p;
  return plan_raynet(d, &p) == NRF_OK ? p.n_ext_slots : -1;
}

extern "C" int nrf_generate_rays(int32_t H, int32_t W, double focal, const double* camera_transform_host, const double* lower,
                                 const double* span, const double* jitter, int32_t n_coarse, float* ray_samples, float* ray_origin,
                                 float* ray_dir, float* z_vals, void* stream) {
  return nrf_generate_rays_range(H, W, focal, camera_transform_host, lower, span, jitter, n_coarse, 0, static_cast<int64_t>(H) * W, ray_samples,
                                 ray_origin, ray_dir, z_vals, stream);
}

extern "C" int nrf_generate_rays_range(int32_t H, int32_t W, double focal, const double* camera_transform_host, const double* lower,
                                       const double* span, const double* jitter, int32_t n_coarse, int64_t ray0, int64_t n_rays,
                                       float* ray_samples, float* ray_origin, float* ray_dir, float* z_vals, void* stream) {
  if (!camera_transform_host || !lower || !span || !jitter || !ray_samples || !ray_origin || !ray_dir || !z_vals) { set_error("generate_rays: NULL argument"); return NRF_E_INVALID; }
  if (H < 1 || W < 1 || n_coarse < 1 || !(focal > 0.0)) { set_error("generate_rays: bad shape H=%d W=%d n_coarse=%d focal=%g", H, W, n_coarse, focal); return NRF_E_INVALID; }
  CamParams cam;
  for (int k = 0; k < 3; ++k) {
    for (int c = 0; c < 3; ++c) cam.r[3 * k + c] = camera_transform_host[4 * k + c];
    cam.t[k] = camera_transform_host[4 * k + 3];
  }
  if (ray0 < 0 || n_rays < 0 || ray0 + n_rays > static_cast<int64_t>(H) * W) { set_error("generate_rays: ray window [%lld, +%lld) outside the %d x %d view", (long long)ray0, (long long)n_rays, H, W); return NRF_E_INVALID; }
  if (n_rays == 0) return NRF_OK;
  cam.focal = focal; cam.H = H; cam.W = W; cam.n = n_coarse; cam.ray0 = ray0; cam.n_rays = n_rays;
  const int64_t total = n_rays * n_coarse;
  generate_rays_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cam, lower, span, jitter, ray_samples, ray_origin,
                                                                                           ray_dir, z_vals);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "generate_rays_kernel launch");
}

extern "C" int nrf_raw2outputs_backward(const float* raw, const float* z, const float* dirs, const float* noise, int64_t B, int32_t n,
                                        int32_t white_background, const float* grad_rgb, const float* grad_weights,
                                        const float* grad_alpha, float* grad_raw, void* stream) {
  if (!raw || !z || !dirs || !grad_rgb || !grad_raw) { set_error("raw2outputs_backward: NULL argument"); return NRF_E_INVALID; }
  if (B < 0 || n < 2 || n > 1024) { set_error("raw2outputs_backward: unsupported shape B=%lld n=%d (n in 2..1024)", (long long)B, n); return NRF_E_INVALID; }
  if (B == 0) return NRF_OK;
  const int wpb = 4;
  const size_t smem = static_cast<size_t>(wpb) * (4 * n + 5 * ((n + 3) & ~3)) * sizeof(float);
  const int grid = static_cast<int>(B / wpb + 1 > 148 * 8 ? 148 * 8 : B / wpb + 1);
  cudaError_t e = cudaFuncSetAttribute(raw2outputs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
  raw2outputs_bwd_kernel<<<grid, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(raw, z, dirs, noise, B, n, white_background, grad_rgb,
                                                                                    grad_weights, grad_alpha, grad_raw);
  e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "raw2outputs_bwd_kernel launch");
}

extern "C" int nrf_positional_encoding_backward(const float* x, const float* grad_out, int64_t n, int32_t c, int32_t freqs, int32_t identity,
                                                float* grad_x, void* stream) {
  if (!x || !grad_out || !grad_x) { set_error("positional_encoding_backward: NULL argument"); return NRF_E_INVALID; }
  if (n < 0 || c < 1 || freqs < 0 || freqs > 30) { set_error("positional_encoding_backward: bad shape n=%lld c=%d freqs=%d", (long long)n, c, freqs); return NRF_E_INVALID; }
  if (n == 0) return NRF_OK;
  pe_bwd_kernel<<<grid_for(n * c, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, grad_out, n, c, freqs, identity, grad_x);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "pe_bwd_kernel launch");
}

extern "C" int nrf_ssim(const float* x, const float* y, int64_t n_planes, int32_t H, int32_t W, const float* kernel2d, int32_t ks, float c1,
                        float c2, float* partial, float* ssim_out, float* cs_out, void* stream) {
  if (!x || !y || !kernel2d || !partial || !ssim_out) { set_error("ssim: NULL argument"); return NRF_E_INVALID; }
  if (ks < 1 || ks > kSsimMaxK || !(ks & 1)) { set_error("ssim: kernel size %d unsupported (odd, <= %d)", ks, kSsimMaxK); return NRF_E_INVALID; }
  if (H < ks || W < ks) { set_error("ssim: Kernel size can't be greater than actual input size (%d x %d, kernel %d)", H, W, ks); return NRF_E_INVALID; }
  if (n_planes < 0 || n_planes > 65535) { set_error("ssim: %lld image planes unsupported (max 65535)", (long long)n_planes); return NRF_E_INVALID; }
  if (n_planes == 0) return NRF_OK;
  const int Ho = H - ks + 1, Wo = W - ks + 1;
  const int tiles = ((Ho + kSsimTile - 1) / kSsimTile) * ((Wo + kSsimTile - 1) / kSsimTile);
  ssim_tile_kernel<<<dim3(tiles, static_cast<unsigned>(n_planes)), kSsimTile * kSsimTile, 0, static_cast<cudaStream_t>(stream)>>>(x, y, H, W, kernel2d, ks, c1, c2, partial);
  ssim_reduce_kernel<<<static_cast<unsigned>(n_planes), 32, 0, static_cast<cudaStream_t>(stream)>>>(partial, tiles, 1.f / (static_cast<float>(Ho) * static_cast<float>(Wo)), ssim_out, cs_out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "ssim kernels launch");
}
extern "C" int64_t nrf_ssim_partial_floats(int64_t n_planes, int32_t H, int32_t W, int32_t ks) {
  if (H < ks || W < ks || ks < 1) return 0;
  const int Ho = H - ks + 1, Wo = W - ks + 1;
  return n_planes * 2 * ((Ho + kSsimTile - 1) / kSsimTile) * ((Wo + kSsimTile - 1) / kSsimTile);
}

extern "C" int nrf_quantize_image(const float* rgb, int64_t n_pixels, uint8_t* out, int32_t to_bgr, void* stream) {
  if (!rgb || !out || n_pixels < 0) { set_error("quantize_image: bad arguments"); return NRF_E_INVALID; }
  if (n_pixels == 0) return NRF_OK;
  quantize_bgr_kernel<<<grid_for(n_pixels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(rgb, n_pixels, out, to_bgr);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? NRF_OK : cuda_fail(e, "quantize_bgr_kernel launch");
}
