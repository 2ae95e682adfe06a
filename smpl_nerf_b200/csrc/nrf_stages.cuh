// Per-ray stages shared by the fused kernel and the stand-alone ops: alpha compositing
// (utils.py:134-191) and inverse-CDF sampling + sorted merge (utils.py:194-264, torchsearchsorted).
// Each function is executed by ONE warp for ONE ray on shared-memory arrays.
#pragma once
#include <cuda_runtime.h>

namespace nrf {

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// ---------------------------------------------------------------------------------- per-ray stages
// Alpha compositing of one ray by one warp (utils.py:134-191).  raw4: [n] (rgb_raw, sigma_raw) in
// smem; on return raw4[i].w holds the weight of sample i.  Lane 0 runs the exclusive transmittance
// product sequentially, exactly like torch.cumprod on the CPU.
__device__ inline void composite_ray(float4* raw4, const float* z, const float* dnorm, float ray_norm, int n, const float* noise,
                              int white, float* rgb_out, float* alpha_out, float* weights_out, int lane) {
  for (int i = lane; i < n; i += 32) {
    float4 r = raw4[i];
    const float nrm = dnorm ? dnorm[i] : ray_norm;
    const float dz = (i < n - 1) ? __fsub_rn(z[i + 1], z[i]) : 1e10f;
    const float delta = __fmul_rn(dz, nrm);
    float s = r.w;
    if (noise) s = __fadd_rn(s, noise[i]);
    const float a = __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(s, 0.f), delta)));
    r.x = sigmoidf_ref(r.x); r.y = sigmoidf_ref(r.y); r.z = sigmoidf_ref(r.z);
    r.w = a;
    raw4[i] = r;
    if (alpha_out) alpha_out[i] = a;
  }
  __syncwarp();
  if (lane == 0) {
    float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f;
    for (int i = 0; i < n; ++i) {
      const float4 r = raw4[i];
      const float w = __fmul_rn(r.w, T);
      cr = fmaf(w, r.x, cr); cg = fmaf(w, r.y, cg); cb = fmaf(w, r.z, cb);
      acc = __fadd_rn(acc, w);
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, r.w), 1e-10f));
      raw4[i].w = w;
    }
    if (white) { const float bg = __fsub_rn(1.f, acc); cr = __fadd_rn(cr, bg); cg = __fadd_rn(cg, bg); cb = __fadd_rn(cb, bg); }
    if (rgb_out) { rgb_out[0] = cr; rgb_out[1] = cg; rgb_out[2] = cb; }
  }
  __syncwarp();
  if (weights_out) for (int i = lane; i < n; i += 32) weights_out[i] = raw4[i].w;
}

// Inverse-CDF sampling + sorted merge of one ray by one warp (utils.py:194-264, torchsearchsorted
// side='right').  w[i*wstride] = coarse weights; zc[nc] coarse depths; writes zf[nc+nf].
__device__ inline void sample_ray(const float* w, int wstride, const float* zc, int nc, int nf, const float* u_fine, float* cdf, float* zs,
                           float* zf, float* z_new_out, int lane) {
  const int m = nc - 1;     // bins = midpoints (m of them); cdf has m entries; m-1 weights
  // pdf numerators into cdf[1..m-1]; their sum with a butterfly so every lane agrees
  float part = 0.f;
  for (int i = 1 + lane; i < m; i += 32) { const float wi = __fadd_rn(w[i * wstride], 1e-5f); cdf[i] = wi; part += wi; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncwarp();
  for (int i = 1 + lane; i < m; i += 32) cdf[i] = __fdiv_rn(cdf[i], part);
  __syncwarp();
  if (lane == 0) {          // sequential cumsum, like torch.cumsum on the CPU
    float run = 0.f;
    cdf[0] = 0.f;
    for (int i = 1; i < m; ++i) { run = __fadd_rn(run, cdf[i]); cdf[i] = run; }
  }
  __syncwarp();
  bool sorted = true;
  for (int j = lane; j < nf; j += 32) {
    const float u = u_fine[j];
    int lo = 0, hi = m;     // upper bound: number of cdf entries <= u
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    const int below = max(0, lo - 1), above = min(m - 1, lo);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = __fmul_rn(.5f, __fadd_rn(zc[below + 1], zc[below]));
    const float b1 = __fmul_rn(.5f, __fadd_rn(zc[above + 1], zc[above]));
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
    const float s = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
    zs[j] = s;
    if (z_new_out) z_new_out[j] = s;
  }
  __syncwarp();
  for (int j = lane; j < nf - 1; j += 32) sorted = sorted && (zs[j] <= zs[j + 1]);
  for (int i = lane; i < nc - 1; i += 32) sorted = sorted && (zc[i] <= zc[i + 1]);
  sorted = __all_sync(0xffffffffu, sorted);
  if (sorted) {
    // stable two-way merge by rank: coarse element i lands at i + #{zs < zc[i]},
    // new sample j at j + #{zc <= zs[j]}
    for (int i = lane; i < nc; i += 32) {
      const float v = zc[i];
      int lo = 0, hi = nf;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (zs[mid] < v) lo = mid + 1; else hi = mid; }
      zf[i + lo] = v;
    }
    for (int j = lane; j < nf; j += 32) {
      const float v = zs[j];
      int lo = 0, hi = nc;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (zc[mid] <= v) lo = mid + 1; else hi = mid; }
      zf[j + lo] = v;
    }
  } else {
    // general case (unsorted input depths): rank every element by counting
    const int n = nc + nf;
    for (int e = lane; e < n; e += 32) {
      const float v = e < nc ? zc[e] : zs[e - nc];
      int rank = 0;
      for (int o = 0; o < n; ++o) {
        const float w = o < nc ? zc[o] : zs[o - nc];
        rank += (w < v || (w == v && o < e)) ? 1 : 0;
      }
      zf[rank] = v;
    }
  }
  __syncwarp();
}


}  // namespace nrf
