// CUDA-core kernels of the training path: everything around the tcgen05 GEMMs of nrf_gemm.cu.
// All of them are HBM-bound streaming / reduction kernels (one pass over a [samples x features] plane or a [rays x n] table).
//
// Reference code being differentiated: models/render_ray_net.py:42-61, models/warp_field_net.py:17-21, utils.py:114-131
// (encoding), utils.py:134-191 (compositing), models/*_pipeline.py (orchestration); the loss side is
// solver/nerf_solver.py:48-51 (MSE on rgb and rgb_fine).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "nrf_plan.h"
#include "nrf_stages.cuh"

namespace nrf {

__device__ __forceinline__ void split_store(float x, __half* hi, __half* lo) {
  const __half h = __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f));
  *hi = h;
  if (lo) *lo = __float2half_rn(x - __half2float(h));
}

// exact mode: x = hi + lo + ll exactly, three bfloat16 values (3 x 8 significant bits = fp32's 24, fp32's exponent range);
// the bit patterns are stored through __half pointers (planes are "16-bit storage")
__device__ __forceinline__ void split_store_bf16x3(float x, __half* hi, __half* lo, __half* ll) {
  const __nv_bfloat16 b0 = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(b0);
  const __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 b2 = __float2bfloat16_rn(r1 - __bfloat162float(b1));
  *reinterpret_cast<__nv_bfloat16*>(hi) = b0; *reinterpret_cast<__nv_bfloat16*>(lo) = b1; *reinterpret_cast<__nv_bfloat16*>(ll) = b2;
}

// ---------------------------------------------------------------------------------- fp32 matrix -> hi/lo planes (weights)
struct SplitJob { const float* src; int32_t rows, cols, ld, col0; __half* hi; __half* lo; int32_t ld_dst, cols_pad; unsigned int* wmax; __half* ll; };   // wmax: max |src| (float bits), optional; ll: third plane (exact mode)
struct SplitTable { int32_t n; int32_t bf16; SplitJob j[36]; };

__global__ void split_planes_kernel(const __grid_constant__ SplitTable t) {
  const SplitJob& j = t.j[blockIdx.y];
  const int total = j.rows * j.cols_pad;
  float m = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / j.cols_pad, c = idx - r * j.cols_pad;
    const float v = c < j.cols ? j.src[static_cast<size_t>(r) * j.ld + j.col0 + c] : 0.f;
    const size_t o = static_cast<size_t>(r) * j.ld_dst + c;
    if (t.bf16) split_store_bf16x3(v, j.hi + o, j.lo + o, j.ll + o);
    else split_store(v, j.hi + o, j.lo ? j.lo + o : nullptr);
    m = fmaxf(m, fabsf(v));
  }
  if (j.wmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && isfinite(m)) atomicMax(j.wmax, __float_as_uint(m));
  }
}
// max |x| of a small fp32 tensor (head weights) into float bits
__global__ void absmax_kernel(const float* __restrict__ x, int n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && isfinite(m)) atomicMax(out, __float_as_uint(m));
}

// ---------------------------------------------------------------------------------- positional encoding -> planes
// utils.py:127-131 in REFERENCE feature order ([x?] ++ for k: sin(2^k x), cos(2^k x), each over the 3 components), zero
// padded to 64 features, as fp16 hi/lo planes [S, 64]: the K-chunk an MLP layer multiplies with its xyz / direction columns.
// A block encodes 32 samples: one work item = (sample, frequency, component) -> ONE sincos whose two results are the sin and the cos
// feature of that (frequency, component) (a thread per feature evaluated every angle twice; the kernel is bound by the range-reduced
// sincos, 64 us for 393k samples); the [32 x 64] tiles are assembled in shared memory and leave as 16-byte row-contiguous stores.
constexpr int kEncRows = 32;
__global__ void __launch_bounds__(256) encode_planes_kernel(const float* __restrict__ x, int64_t S, int freqs, int identity, __half* __restrict__ hi,
                                                            __half* __restrict__ lo, __half* __restrict__ ll) {
  __shared__ __align__(16) __half th[kEncRows][64];
  __shared__ __align__(16) __half tl[kEncRows][64];
  __shared__ __align__(16) __half tq[kEncRows][64];
  __shared__ float xs[kEncRows][3];
  const int64_t s0 = static_cast<int64_t>(blockIdx.x) * kEncRows;
  const int tid = threadIdx.x;
  if (tid < kEncRows * 3) { const int64_t i = s0 * 3 + tid; xs[tid / 3][tid % 3] = i < S * 3 ? x[i] : 0.f; }
  for (int i = tid; i < kEncRows * 64 / 8; i += 256) {      // zero: padding features (and rows past S, never stored)
    reinterpret_cast<uint4*>(&th[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(&tl[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4*>(&tq[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int n_id = identity ? 3 : 0;
  auto put = [&](int r, int f, float val) {
    if (ll) split_store_bf16x3(val, &th[r][f], &tl[r][f], &tq[r][f]);
    else split_store(val, &th[r][f], &tl[r][f]);
  };
  const int per = 3 * freqs + n_id;                         // work items per sample: (k, comp) pairs, then the identity components
  for (int w = tid; w < kEncRows * per; w += 256) {
    const int r = w / per, j = w - r * per;
    if (j < 3 * freqs) {
      const int k = j / 3, comp = j - 3 * k;
      float sv, cv;
      sincos_pe(xs[r][comp] * __int_as_float((127 + k) << 23), sv, cv);
      put(r, n_id + 6 * k + comp, sv);
      put(r, n_id + 6 * k + 3 + comp, cv);
    } else {
      const int comp = j - 3 * freqs;
      put(r, comp, xs[r][comp]);
    }
  }
  __syncthreads();
  {
    const int r = tid >> 3, c8 = (tid & 7) * 8;             // 256 threads = 32 rows x 8 chunks of 8 features
    const int64_t srow = s0 + r;
    if (srow < S) {
      *reinterpret_cast<uint4*>(hi + srow * 64 + c8) = *reinterpret_cast<const uint4*>(&th[r][c8]);
      if (lo) *reinterpret_cast<uint4*>(lo + srow * 64 + c8) = *reinterpret_cast<const uint4*>(&tl[r][c8]);
      if (ll) *reinterpret_cast<uint4*>(ll + srow * 64 + c8) = *reinterpret_cast<const uint4*>(&tq[r][c8]);
    }
  }
}

// ---------------------------------------------------------------------------------- per-ray features
// pose features [B, A]: encode(goal_pose[:, cols]) in the reference's order (models/append_to_nerf_pipeline.py:26-37,
// models/append_smpl_params_pipeline.py:30-37, models/smpl_nerf_pipeline.py:28-30); direction features [B, D]:
// encode(ray_direction / |ray_direction|) (models/nerf_pipeline.py:30-35); ray_norm [B] = |ray_direction|.
__global__ void ray_feats_kernel(const float* __restrict__ goal_pose, int pose_stride, int n_sel, int col0, int col1, int pose_all,
                                 int pose_freqs, int pose_identity, int pose_encoded, const float* __restrict__ ray_dir,
                                 int dir_freqs, int dir_identity, int64_t B, int A, int D, float* __restrict__ pose_feat,
                                 float* __restrict__ dir_feat, float* __restrict__ ray_norm) {
  const int per = A + D + 1;
  const int64_t total = B * per;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = idx / per;
    int i = static_cast<int>(idx - b * per);
    if (i < A) {
      auto comp = [&](int c) { return goal_pose[b * pose_stride + (pose_all ? c : (c == 0 ? col0 : col1))]; };
      float val;
      if (!pose_encoded) val = comp(i);
      else if (pose_identity && i < n_sel) val = comp(i);
      else {
        const int j = i - (pose_identity ? n_sel : 0), k = j / (2 * n_sel), rem = j - k * 2 * n_sel;
        const float a = comp(rem % n_sel) * __int_as_float((127 + k) << 23);
        float sv, cv;
        sincos_pe(a, sv, cv);
        val = rem < n_sel ? sv : cv;
      }
      pose_feat[b * A + i] = val;
      continue;
    }
    i -= A;
    const float d0 = ray_dir[b * 3], d1 = ray_dir[b * 3 + 1], d2 = ray_dir[b * 3 + 2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
    if (i == D) { ray_norm[b] = nrm; continue; }
    const float u[3] = {__fdiv_rn(d0, nrm), __fdiv_rn(d1, nrm), __fdiv_rn(d2, nrm)};
    float val;
    if (dir_identity && i < 3) val = u[i];
    else {
      const int j = i - (dir_identity ? 3 : 0), k = j / 6, rem = j - 6 * k;
      float sv, cv;
      sincos_pe(u[rem % 3] * __int_as_float((127 + k) << 23), sv, cv);
      val = rem < 3 ? sv : cv;
    }
    dir_feat[b * D + i] = val;
  }
}

// out[b, n] = bias[n] + sum_k W[n, col0 + k] * feat[b, k]: the per-ray constant inputs of a layer folded into a per-ray bias
// (8 rays per CTA, features staged in shared memory, thread = output feature).
__global__ void ray_bias2_kernel(const float* __restrict__ W, int ld, int col0, int K, const float* __restrict__ bias,
                                 const float* __restrict__ feat, int64_t B, int n_out, float* __restrict__ out) {
  extern __shared__ float fs[];     // [8][K]
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * 8;
  for (int i = threadIdx.x; i < 8 * K; i += blockDim.x) {
    const int64_t b = b0 + i / K;
    fs[i] = b < B ? feat[b * K + i % K] : 0.f;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < n_out; n += blockDim.x) {
    float acc[8];
    const float bn = bias[n];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = bn;
    const float* w = W + static_cast<size_t>(n) * ld + col0;
    for (int k = 0; k < K; ++k) {
      const float wk = __ldg(w + k);
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = fmaf(wk, fs[r * K + k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) if (b0 + r < B) out[(b0 + r) * n_out + n] = acc[r];
  }
}

// ---------------------------------------------------------------------------------- small heads (N = 1 or 3: CUDA cores)
// out[s, c0 + c] = b[c] + sum_k W[c, k] x[s, k]     (sigma_out_layer / rgb_out_layer / WarpFieldNet.linear2)
__global__ void heads_kernel(const float* __restrict__ x, int64_t S, int K, const float* __restrict__ W, const float* __restrict__ b, int nh,
                             float* __restrict__ out, int out_ld, int c0) {
  // 8 lanes per row, 4 rows per warp: 16-byte loads (128 contiguous bytes per row and instruction), 3 shuffle steps per accumulator
  // (a warp per row with 4-byte loads spent its time in 15 shuffles per row: 4.1 TB/s)
  extern __shared__ __align__(16) float ws[];     // [nh][K]
  for (int i = threadIdx.x; i < nh * K; i += blockDim.x) ws[i] = W[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, sub = lane & 7, rw = lane >> 3;
  const int64_t n4 = (S + 3) / 4;                 // groups of 4 rows
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); q < n4; q += static_cast<int64_t>(gridDim.x) * wpb) {
    const int64_t s = q * 4 + rw;
    const bool ok = s < S;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (ok) {
      for (int k = 4 * sub; k < K; k += 32) {
        const float4 v = *reinterpret_cast<const float4*>(x + s * K + k);
        const float4 w0 = *reinterpret_cast<const float4*>(ws + k);
        a0 = fmaf(v.x, w0.x, a0); a0 = fmaf(v.y, w0.y, a0); a0 = fmaf(v.z, w0.z, a0); a0 = fmaf(v.w, w0.w, a0);
        if (nh > 1) {
          const float4 w1 = *reinterpret_cast<const float4*>(ws + K + k), w2 = *reinterpret_cast<const float4*>(ws + 2 * K + k);
          a1 = fmaf(v.x, w1.x, a1); a1 = fmaf(v.y, w1.y, a1); a1 = fmaf(v.z, w1.z, a1); a1 = fmaf(v.w, w1.w, a1);
          a2 = fmaf(v.x, w2.x, a2); a2 = fmaf(v.y, w2.y, a2); a2 = fmaf(v.z, w2.z, a2); a2 = fmaf(v.w, w2.w, a2);
        }
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (sub == 0 && ok) {
      out[s * out_ld + c0] = a0 + b[0];
      if (nh > 1) { out[s * out_ld + c0 + 1] = a1 + b[1]; out[s * out_ld + c0 + 2] = a2 + b[2]; }
    }
  }
}

// Backward of a head over a block of rows.  thread = 4 consecutive input features of one row group (K / 4 threads per row,
// 1024 / K rows in flight per CTA): 16-byte loads of x, 8-byte stores of the gradient planes.
//   dW[c, k] += sum_s g[s, c] x[s, k]      db[c] += sum_s g[s, c]      dY[s, k] = sum_c g[s, c] W[c, k]   (-> planes, masked by x > 0)
// g is the upstream gradient in REAL units (fp32); the planes are written as real * sc_out[0].  l1max: see TileGemmParams (here the
// row segment is a warp's 128 columns).
__global__ void __launch_bounds__(256, 3) head_bwd_kernel(const float* __restrict__ x, int64_t S, int K, const float* __restrict__ g, int g_ld, int c0, int nh,
                                                       const float* __restrict__ W, const float* __restrict__ sc_out, int relu_mask,
                                                       float* __restrict__ dW, float* __restrict__ db, __half* __restrict__ dy_hi,
                                                       __half* __restrict__ dy_lo, int dy_ld, int rows_per_block, unsigned int* __restrict__ l1max) {
  __shared__ float red[3][1024];
  const int tpr = K >> 2;                                 // threads per row
  const int rg = threadIdx.x / tpr, n_rg = blockDim.x / tpr, k4 = (threadIdx.x - rg * tpr) * 4;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const float sc = sc_out ? __ldg(sc_out) : 1.f;
  float4 w0 = *reinterpret_cast<const float4*>(W + k4), w1 = make_float4(0.f, 0.f, 0.f, 0.f), w2 = w1;
  if (nh > 1) { w1 = *reinterpret_cast<const float4*>(W + K + k4); w2 = *reinterpret_cast<const float4*>(W + 2 * K + k4); }
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
  float b0 = 0.f, b1 = 0.f, b2 = 0.f, l1_run = 0.f;
  struct Row { int64_t s; bool ok; float g0, g1, g2; float4 xv; };
  auto fetch = [&](int it) {                                // all loads of a row: issued for TWO rows before either is consumed
    Row r;
    r.s = r0 + static_cast<int64_t>(it) * n_rg + rg;
    r.ok = it * n_rg < rows_per_block && r.s < r0 + rows_per_block && r.s < S;
    r.g0 = r.g1 = r.g2 = 0.f;
    r.xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r.ok) {
      r.g0 = g[r.s * g_ld + c0];
      if (nh > 1) { r.g1 = g[r.s * g_ld + c0 + 1]; r.g2 = g[r.s * g_ld + c0 + 2]; }
      r.xv = *reinterpret_cast<const float4*>(x + r.s * K + k4);
    }
    return r;
  };
  auto consume = [&](const Row& r) {
    const float g0 = r.g0, g1 = r.g1, g2 = r.g2;
    const float4 xv = r.xv;
    b0 += g0; b1 += g1; b2 += g2;
    a0.x = fmaf(g0, xv.x, a0.x); a0.y = fmaf(g0, xv.y, a0.y); a0.z = fmaf(g0, xv.z, a0.z); a0.w = fmaf(g0, xv.w, a0.w);
    if (nh > 1) {
      a1.x = fmaf(g1, xv.x, a1.x); a1.y = fmaf(g1, xv.y, a1.y); a1.z = fmaf(g1, xv.z, a1.z); a1.w = fmaf(g1, xv.w, a1.w);
      a2.x = fmaf(g2, xv.x, a2.x); a2.y = fmaf(g2, xv.y, a2.y); a2.z = fmaf(g2, xv.z, a2.z); a2.w = fmaf(g2, xv.w, a2.w);
    }
    if (dy_hi) {
      float d[4] = {fmaf(g0, w0.x, fmaf(g1, w1.x, g2 * w2.x)), fmaf(g0, w0.y, fmaf(g1, w1.y, g2 * w2.y)),
                    fmaf(g0, w0.z, fmaf(g1, w1.z, g2 * w2.z)), fmaf(g0, w0.w, fmaf(g1, w1.w, g2 * w2.w))};
      if (relu_mask) {
        if (!(xv.x > 0.f)) d[0] = 0.f;
        if (!(xv.y > 0.f)) d[1] = 0.f;
        if (!(xv.z > 0.f)) d[2] = 0.f;
        if (!(xv.w > 0.f)) d[3] = 0.f;
      }
      __align__(8) __half h[4];
      __align__(8) __half l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_store(d[i] * sc, &h[i], &l[i]);
      if (r.ok) {
        *reinterpret_cast<uint2*>(dy_hi + r.s * dy_ld + k4) = *reinterpret_cast<const uint2*>(h);
        if (dy_lo) *reinterpret_cast<uint2*>(dy_lo + r.s * dy_ld + k4) = *reinterpret_cast<const uint2*>(l);
      }
      if (l1max) {          // L1 norm of this warp's 128 columns of the row (real units); a warp never straddles two rows (K >= 128) ...
        float a = fabsf(d[0]) + fabsf(d[1]) + fabsf(d[2]) + fabsf(d[3]);
        const int span = tpr < 32 ? tpr : 32;          // ... and for K = 64 a row is 16 lanes
        for (int o = span >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        l1_run = fmaxf(l1_run, a);
      }
    }
  };
  for (int it = 0; it * n_rg < rows_per_block; it += 4) {   // uniform trip count: the warp shuffles need every lane; 4 rows in flight
    const Row ra = fetch(it), rb = fetch(it + 1), rc = fetch(it + 2), rd = fetch(it + 3);
    consume(ra);
    consume(rb);
    consume(rc);
    consume(rd);
  }
  if (l1max && isfinite(l1_run)) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l1_run = fmaxf(l1_run, __shfl_xor_sync(0xffffffffu, l1_run, o));
    if ((threadIdx.x & 31) == 0) atomicMax(l1max, __float_as_uint(l1_run * static_cast<float>((K + 127) / 128)));
  }
  // fold the row groups of the CTA, then one atomic per (head, feature)
  float* ra[3] = {red[0], red[1], red[2]};
  const float4 acc[3] = {a0, a1, a2};
  for (int c = 0; c < nh; ++c) *reinterpret_cast<float4*>(ra[c] + rg * K + k4) = acc[c];
  __syncthreads();
  for (int i = threadIdx.x; i < nh * K; i += blockDim.x) {
    const int c = i / K, k = i - c * K;
    float t = 0.f;
    for (int r = 0; r < n_rg; ++r) t += ra[c][r * K + k];
    atomicAdd(dW + c * K + k, t);
  }
  __syncthreads();
  if (k4 == 0) { red[0][rg] = b0; red[1][rg] = b1; red[2][rg] = b2; }
  __syncthreads();
  if (threadIdx.x < nh) {
    float t = 0.f;
    for (int r = 0; r < n_rg; ++r) t += red[threadIdx.x][r];
    atomicAdd(db + threadIdx.x, t);
  }
}

// ---------------------------------------------------------------------------------- compositing (forward / backward)
// One warp per ray; dnorm is per sample ([B, n], the SMPL coarse pass: |warped - o|, smpl_nerf_pipeline.py:52,63) or per ray ([B]).
__global__ void composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ dnorm, int per_sample,
                                     const float* __restrict__ noise, int64_t B, int n, int white, float* __restrict__ rgb,
                                     float* __restrict__ weights, float* __restrict__ alpha) {
  extern __shared__ __align__(16) float smem[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n4 = (n + 3) & ~3;
  float4* raw4 = reinterpret_cast<float4*>(smem) + static_cast<size_t>(w) * n;
  float* tf = smem + static_cast<size_t>(wpb) * n * 4 + static_cast<size_t>(w) * (4 * n4 + kTeamScratch);
  float* tt = tf + n4; float* zs = tt + n4; float* dn = zs + n4; float* ts = tf + 4 * n4;
  for (int64_t ray = blockIdx.x * static_cast<int64_t>(wpb) + w; ray < B; ray += static_cast<int64_t>(gridDim.x) * wpb) {
    for (int i = lane; i < n; i += 32) {
      raw4[i] = *reinterpret_cast<const float4*>(raw + (ray * n + i) * 4);
      zs[i] = z[ray * n + i];
      dn[i] = per_sample ? dnorm[ray * n + i] : dnorm[ray];
    }
    __syncwarp();
    composite_ray(raw4, zs, dn, 0.f, n, noise ? noise + ray * n : nullptr, white, rgb + ray * 3, alpha ? alpha + ray * n : nullptr,
                  weights ? weights + ray * n : nullptr, tf, tt, ts, lane);
    __syncwarp();
  }
}

// d(loss)/d(raw) [B, n, 4] from d(loss)/d(rgb) [B, 3] (see raw2outputs_bwd_kernel in nrf_ops.cu for the formulas), plus
// d(loss)/d(dnorm) [B, n] when the per-sample norm is itself a function of the nets (SMPL coarse pass), and the running
// maximum of |d raw| (as float bits: non-negative floats order like unsigned integers) for the gradient scale.
__global__ void composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ dnorm, int per_sample,
                                     const float* __restrict__ noise, int64_t B, int n, int white, const float* __restrict__ g_rgb,
                                     float* __restrict__ g_raw, float* __restrict__ g_dnorm, unsigned int* __restrict__ gmax_bits) {
  extern __shared__ __align__(16) float smem[];
  const int wpb = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n4 = (n + 3) & ~3;
  float4* act4 = reinterpret_cast<float4*>(smem) + static_cast<size_t>(w) * n;
  float* base = smem + static_cast<size_t>(wpb) * n * 4 + static_cast<size_t>(w) * 6 * n4;
  float* keep = base; float* T = keep + n4; float* gw = T + n4; float* suf = gw + n4; float* dsig = suf + n4; float* ddel = dsig + n4;
  float gmax = 0.f;
  for (int64_t ray = blockIdx.x * static_cast<int64_t>(wpb) + w; ray < B; ray += static_cast<int64_t>(gridDim.x) * wpb) {
    const float gr = g_rgb[ray * 3], gg = g_rgb[ray * 3 + 1], gb = g_rgb[ray * 3 + 2];
    for (int i = lane; i < n; i += 32) {
      const float4 r = *reinterpret_cast<const float4*>(raw + (ray * n + i) * 4);
      const float nrm = per_sample ? dnorm[ray * n + i] : dnorm[ray];
      const float dz = (i < n - 1) ? __fsub_rn(z[ray * n + i + 1], z[ray * n + i]) : 1e10f;
      const float delta = __fmul_rn(dz, nrm);
      const float s = noise ? __fadd_rn(r.w, noise[ray * n + i]) : r.w;
      const float e = expf(-__fmul_rn(fmaxf(s, 0.f), delta));          // = 1 - alpha
      const float a = __fsub_rn(1.f, e);
      act4[i] = make_float4(sigmoidf_ref(r.x), sigmoidf_ref(r.y), sigmoidf_ref(r.z), a);
      keep[i] = __fadd_rn(__fsub_rn(1.f, a), 1e-10f);
      dsig[i] = s > 0.f ? delta * e : 0.f;                              // d alpha / d sigma
      ddel[i] = s > 0.f ? s * e * dz : 0.f;                             // d alpha / d dnorm
    }
    __syncwarp();
    if (lane == 0) serial_scan<true>(keep, T, n);
    __syncwarp();
    const float bg = white ? (gr + gg + gb) : 0.f;
    for (int i = lane; i < n; i += 32) {
      const float4 c = act4[i];
      const float g = gr * c.x + gg * c.y + gb * c.z - bg;
      gw[i] = g;
      suf[i] = g * (c.w * T[i]);
    }
    __syncwarp();
    if (lane == 0) { float run = 0.f; for (int i = n - 1; i >= 0; --i) { const float v = suf[i]; suf[i] = run; run += v; } }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const float4 c = act4[i];
      const float wi = c.w * T[i];
      const float ga = gw[i] * T[i] - suf[i] / keep[i];
      float4 o;
      o.x = wi * gr * c.x * (1.f - c.x);
      o.y = wi * gg * c.y * (1.f - c.y);
      o.z = wi * gb * c.z * (1.f - c.z);
      o.w = ga * dsig[i];
      *reinterpret_cast<float4*>(g_raw + (ray * n + i) * 4) = o;
      if (g_dnorm) g_dnorm[ray * n + i] = ga * ddel[i];
      gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  if (lane == 0 && isfinite(gmax)) atomicMax(gmax_bits, __float_as_uint(gmax));
}

// sc = {s, 1/s} with s the largest power of two such that s * bound <= 2^15, where
//   bound = lmax * mul * wmax + ea * eb
// is an upper bound of the magnitudes about to be written into fp16 gradient planes (lmax: largest row L1 norm of the incoming
// gradient or max |d raw|; wmax: max |W| of the layer it is multiplied with; ea * eb: the sigma head's rank-1 term).  The
// bound is typically 30-150x above the actual maximum, which then sits around 2^8..2^10: ~24 octaves of normal fp16 range
// below it, 5 above -- per LAYER, so neither growth nor decay along the backward chain can leave the range.
__global__ void scale_from_bound_kernel(const unsigned int* __restrict__ lmax, float mul, const unsigned int* __restrict__ wmax,
                                        const unsigned int* __restrict__ ea, const unsigned int* __restrict__ eb, float* __restrict__ sc) {
  float bound = __uint_as_float(*lmax) * mul * (wmax ? __uint_as_float(*wmax) : 1.f);
  if (ea && eb) bound += __uint_as_float(*ea) * __uint_as_float(*eb);
  float s = 1.f;
  if (bound > 0.f && isfinite(bound)) {
    int e;
    frexpf(bound, &e);                  // bound = f * 2^e, f in [0.5, 1)  ->  bound <= 2^e
    e = 15 - e;
    e = e > 120 ? 120 : (e < -120 ? -120 : e);
    s = ldexpf(1.f, e);
  }
  sc[0] = s; sc[1] = 1.f / s;
}

// ---------------------------------------------------------------------------------- reductions for bias / per-ray-input grads
// dysum[b, f] = sum over the n samples of ray b of (hi + lo)[s, f]      (block = ray, thread = feature)
__global__ void ray_colsum_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, int ld, int F, int n, float* __restrict__ dysum) {
  const int64_t b = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) {
      const int64_t o = (b * n + i) * ld + f;
      acc += __half2float(hi[o]) + (lo ? __half2float(lo[o]) : 0.f);
    }
    dysum[b * F + f] = acc;
  }
}
// db[f] += inv_scale * sum_b dysum[b, f]      (block = 32 features x 32 ray lanes, fixed summation order)
__global__ void __launch_bounds__(1024) bias_grad_kernel(const float* __restrict__ dysum, int64_t B, int F, const float* __restrict__ scale2,
                                                         float* __restrict__ db) {
  __shared__ float red[32][33];
  const int fx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + fx;
  float a0 = 0.f, a1 = 0.f;
  if (f < F) {
    int64_t b = ry;
    for (; b + 32 < B; b += 64) { a0 += dysum[b * F + f]; a1 += dysum[(b + 32) * F + f]; }
    if (b < B) a0 += dysum[b * F + f];
  }
  red[ry][fx] = a0 + a1;
  __syncthreads();
  if (ry == 0 && f < F) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) t += red[r][fx];
    db[f] += t * __ldg(scale2 + 1);
  }
}
// partial[z][n][k] = sum over the rays of slice z of dysum[b, n] feat[b, k]      (16 x 16 output tile per CTA, rays in steps of 64 via smem;
// gridDim.z ray slices -- 48 CTAs walking all rays were latency bound -- summed in a fixed order by dw_reduce_kernel, which also divides
// the gradient scale out and accumulates into dW[n, col0 + k])
constexpr int kRayFeatSlices = 8;
__global__ void __launch_bounds__(256) rayfeat_dw_kernel(const float* __restrict__ dysum, const float* __restrict__ feat, int64_t B, int n_out, int K,
                                                         float* __restrict__ partial) {
  constexpr int R = 64;
  __shared__ float ys[R][17], fs[R][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int k = blockIdx.x * 16 + tx, n = blockIdx.y * 16 + ty;
  const int nl = blockIdx.y * 16 + tx;                      // the dysum column this thread loads
  const int64_t per = (B + gridDim.z - 1) / gridDim.z;
  const int64_t b_lo = per * blockIdx.z, b_hi = b_lo + per < B ? b_lo + per : B;
  float acc0 = 0.f, acc1 = 0.f;
  for (int64_t b0 = b_lo; b0 < b_hi; b0 += R) {
#pragma unroll
    for (int j = 0; j < R / 16; ++j) {
      const int64_t b = b0 + ty + 16 * j;
      ys[ty + 16 * j][tx] = (b < b_hi && nl < n_out) ? dysum[b * n_out + nl] : 0.f;
      fs[ty + 16 * j][tx] = (b < b_hi && k < K) ? feat[b * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; r += 2) { acc0 = fmaf(ys[r][ty], fs[r][tx], acc0); acc1 = fmaf(ys[r + 1][ty], fs[r + 1][tx], acc1); }
    __syncthreads();
  }
  if (k < K && n < n_out) partial[(static_cast<size_t>(blockIdx.z) * n_out + n) * K + k] = acc0 + acc1;
}
// dst[m, col0 + c] += inv_scale * sum_split partial[split][m][c]        (m < M <= Mp rows of the partials, c < cols <= N)
// and, in the blocks past the first `main_blocks`, db[m] += inv_scale * sum_split cs_partial[split][m] (the ones-column of dw_gemm).
// A block owns 64 consecutive outputs; its 4 thread groups each add every 4th split (independent, unrolled loads -- one thread walking
// all ~74 splits serially was latency bound: 25-45 us for 20 MB) and the 4 sums are combined in a fixed order (deterministic).
constexpr int kDwReduceThreads = 256;
__global__ void __launch_bounds__(kDwReduceThreads) dw_reduce_kernel(const float* __restrict__ partial, int n_split, int Mp, int M, int N, int cols,
                                                                      const float* __restrict__ scale2, float* __restrict__ dst, int ld, int col0,
                                                                      int main_blocks, const float* __restrict__ cs_partial, float* __restrict__ db) {
  __shared__ float part[4][64];
  const int e = threadIdx.x & 63, g = threadIdx.x >> 6;
  const bool bias_part = static_cast<int>(blockIdx.x) >= main_blocks;
  const int idx = (bias_part ? blockIdx.x - main_blocks : blockIdx.x) * 64 + e;
  const bool ok = idx < (bias_part ? M : M * cols);
  const int m = ok ? (bias_part ? idx : idx / cols) : 0, c = (ok && !bias_part) ? idx - m * cols : 0;
  const float* src = bias_part ? cs_partial + m : partial + static_cast<size_t>(m) * N + c;
  const size_t stride = bias_part ? static_cast<size_t>(Mp) : static_cast<size_t>(Mp) * N;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int s = g;
  if (ok) {
    for (; s + 12 < n_split; s += 16) {
      a0 += __ldg(src + s * stride); a1 += __ldg(src + (s + 4) * stride); a2 += __ldg(src + (s + 8) * stride); a3 += __ldg(src + (s + 12) * stride);
    }
    for (; s < n_split; s += 4) a0 += __ldg(src + s * stride);
  }
  part[g][e] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (g == 0 && ok) {
    const float v = ((part[0][e] + part[1][e]) + (part[2][e] + part[3][e])) * __ldg(scale2 + 1);
    if (bias_part) db[m] += v; else dst[static_cast<size_t>(m) * ld + col0 + c] += v;
  }
}

// pts[b, i, :] = o[b] + d[b] * z[b, i] with separate multiply and add (utils.py:262); also copies z (teacher-forced fine depths)
__global__ void points_from_z_kernel(const float* __restrict__ origin, const float* __restrict__ dir, const float* __restrict__ z_in, int64_t B, int n,
                                     float* __restrict__ z_out, float* __restrict__ pts) {
  const int64_t total = B * n;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = idx / n;
    const float zz = z_in[idx];
    z_out[idx] = zz;
#pragma unroll
    for (int k = 0; k < 3; ++k) pts[idx * 3 + k] = __fadd_rn(origin[b * 3 + k], __fmul_rn(dir[b * 3 + k], zz));
  }
}

// ---------------------------------------------------------------------------------- SMPL: warp -> warped points -> view directions
// warped = pts + warp; v = warped - o; dnorm = |v|; u = v / |v|      (models/smpl_nerf_pipeline.py:48-56)
__global__ void smpl_points_kernel(const float* __restrict__ pts, const float* __restrict__ warp, const float* __restrict__ origin, int64_t S,
                                   int n, float* __restrict__ warped, float* __restrict__ u, float* __restrict__ dnorm,
                                   float* __restrict__ warp_out, float* __restrict__ warped_out) {
  for (int64_t s = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; s < S; s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = s / n;
    float wv[3], v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float w = warp[s * 3 + k];
      wv[k] = __fadd_rn(pts[s * 3 + k], w);
      v[k] = __fsub_rn(wv[k], origin[b * 3 + k]);
      warped[s * 3 + k] = wv[k];
      if (warp_out) warp_out[s * 3 + k] = w;
      if (warped_out) warped_out[s * 3 + k] = wv[k];
    }
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
    dnorm[s] = nrm;
#pragma unroll
    for (int k = 0; k < 3; ++k) u[s * 3 + k] = __fdiv_rn(v[k], nrm);
  }
}

// Backward of the chain above: g_warp[s, :] from the gradients of the two encodings (fp32 [S, 64], reference feature order)
// and of dnorm (may be NULL) -- everything in REAL units; gmax_bits receives max |g_warp| for the warp head's plane scale.
//   d enc(x)/dx_j: sum_k 2^k (cos(2^k x_j) g_sin[k, j] - sin(2^k x_j) g_cos[k, j]) (+ identity)
//   u = v / |v|:   g_v = (g_u - u (u . g_u)) / |v|         dnorm = |v|:  g_v += g_dnorm u
__global__ void smpl_points_bwd_kernel(const float* __restrict__ g_encx, int xf, int xid, const float* __restrict__ g_encd, int df, int did,
                                       const float* __restrict__ warped, const float* __restrict__ u, const float* __restrict__ dnorm,
                                       const float* __restrict__ g_dnorm, int64_t S, float* __restrict__ g_warp,
                                       unsigned int* __restrict__ gmax_bits) {
  // 4 lanes per sample: lanes 0..2 own one component each (the sincos chains of the three components run side by side), lane 3 idles;
  // the dot product gu . u is folded with two shuffles inside the group
  float gmax = 0.f;
  const int64_t total = (S + 7) / 8 * 8 * 4;                // whole warps take part in the shuffles
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t s = idx >> 2;
    const int j = static_cast<int>(idx & 3);
    const bool ok = s < S && j < 3;
    float gx = 0.f, gu = 0.f, uj = 0.f;
    if (ok) {
      const float* g = g_encx + s * 64;
      const float x = warped[s * 3 + j];
      gx = xid ? g[j] : 0.f;
      const int base = xid ? 3 : 0;
      for (int k = 0; k < xf; ++k) {
        const float f = __int_as_float((127 + k) << 23);
        float sv, cv;
        sincos_pe(x * f, sv, cv);
        gx = fmaf(f, fmaf(cv, g[base + 6 * k + j], -sv * g[base + 6 * k + 3 + j]), gx);
      }
      const float* gd = g_encd + s * 64;
      uj = u[s * 3 + j];
      gu = did ? gd[j] : 0.f;
      const int based = did ? 3 : 0;
      for (int k = 0; k < df; ++k) {
        const float f = __int_as_float((127 + k) << 23);
        float sv, cv;
        sincos_pe(uj * f, sv, cv);
        gu = fmaf(f, fmaf(cv, gd[based + 6 * k + j], -sv * gd[based + 6 * k + 3 + j]), gu);
      }
    }
    // dot = gu0 u0 + gu1 u1 + gu2 u2 in the order of the one-thread-per-sample formulation: (p0 + p1) + p2
    const float pj = gu * uj;
    const unsigned lane = threadIdx.x & 31u, l0 = lane & ~3u;
    const float p0 = __shfl_sync(0xffffffffu, pj, l0), p1 = __shfl_sync(0xffffffffu, pj, l0 + 1), p2 = __shfl_sync(0xffffffffu, pj, l0 + 2);
    const float dot = p0 + p1 + p2;
    if (ok) {
      const float inv = 1.f / dnorm[s];
      const float gdn = g_dnorm ? g_dnorm[s] : 0.f;
      const float out = gx + (gu - uj * dot) * inv + gdn * uj;
      g_warp[s * 3 + j] = out;
      gmax = fmaxf(gmax, fabsf(out));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  if ((threadIdx.x & 31) == 0 && isfinite(gmax)) atomicMax(gmax_bits, __float_as_uint(gmax));
}

}  // namespace nrf
