// Per-ray stages shared by the fused kernel and the stand-alone ops: alpha compositing
// (utils.py:134-191) and inverse-CDF sampling + sorted merge (utils.py:194-264, torchsearchsorted).
// Each function is executed by ONE warp for ONE ray on shared-memory arrays.
#pragma once
#include <cuda_runtime.h>

namespace nrf {

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }

// sin and cos of a (|a| < ~1e5) to ~1.2 ulp: 2-term Cody-Waite reduction by pi/2 (exact under FMA for
// the encoder's arguments x * 2^k) and the Cephes sinf/cosf minimax polynomials on [-pi/4, pi/4].
// Replaces libdevice sincosf, whose slow path (Payne-Hanek) bloats the kernel when inlined 100x.
__device__ __forceinline__ void sincos_pe(float a, float& s_out, float& c_out) {
  const float t = fmaf(a, 0.636619747f, 12582912.f);
  const int qi = __float_as_int(t);
  const float q = t - 12582912.f;
  float r = fmaf(q, -1.57079637050628662109375f, a);
  r = fmaf(q, 4.37113900018624283e-8f, r);
  const float z = r * r;
  float s = fmaf(fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f), z, -1.6666654611e-1f);
  s = fmaf(s * z, r, r);
  float c = fmaf(fmaf(fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z, -0.5f);
  c = fmaf(c, z, 1.0f);
  float ss = (qi & 1) ? c : s, cc = (qi & 1) ? s : c;
  if (qi & 2) ss = -ss;
  if ((qi + 1) & 2) cc = -cc;
  s_out = ss; c_out = cc;
}

// ---------------------------------------------------------------------------------- per-sample stage
// utils.py:161-175 for one sample: raw = (rgb_raw, sigma_raw) -> (sigmoid(rgb_raw), alpha) with
// alpha = 1 - exp(-relu(sigma_raw [+ noise]) * dz * |dir|); dz is z[i+1] - z[i], or 1e10 for the last sample.
__device__ __forceinline__ float alpha_sample(float sigma_raw, float dz, float nrm, const float* noise) {
  const float delta = __fmul_rn(dz, nrm);
  float s = sigma_raw;
  if (noise) s = __fadd_rn(s, *noise);
  return __fsub_rn(1.f, expf(-__fmul_rn(fmaxf(s, 0.f), delta)));
}
__device__ __forceinline__ float4 activate_sample(float4 r, float dz, float nrm, const float* noise) {
  r.w = alpha_sample(r.w, dz, nrm, noise);
  r.x = sigmoidf_ref(r.x); r.y = sigmoidf_ref(r.y); r.z = sigmoidf_ref(r.z);
  return r;
}

// ---------------------------------------------------------------------------------- per-ray stages
// Sequential running product / sum by ONE lane, in the element order of torch.cumprod / torch.cumsum on
// the CPU (so the bits match), but with the loads of the next block issued ahead of the dependent
// chain: src and dst are distinct arrays (16-byte aligned, padded to a multiple of 4).
//   kProd: dst[i] = prod_{j<i} src[j]  (exclusive, starts at 1)      else: dst[i] = sum_{j<=i} src[j]
template <bool kProd>
__device__ __forceinline__ void serial_scan(const float* __restrict__ src, float* __restrict__ dst, int n) {
  float run = kProd ? 1.f : 0.f;
  int i = 0;
  float4 f = n >= 4 ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (; i + 4 <= n; i += 4) {
    const float4 cur = f;
    if (i + 8 <= n) f = *reinterpret_cast<const float4*>(src + i + 4);
    float4 o;
    if (kProd) {
      o.x = run; run = __fmul_rn(run, cur.x); o.y = run; run = __fmul_rn(run, cur.y);
      o.z = run; run = __fmul_rn(run, cur.z); o.w = run; run = __fmul_rn(run, cur.w);
    } else {
      run = __fadd_rn(run, cur.x); o.x = run; run = __fadd_rn(run, cur.y); o.y = run;
      run = __fadd_rn(run, cur.z); o.z = run; run = __fadd_rn(run, cur.w); o.w = run;
    }
    *reinterpret_cast<float4*>(dst + i) = o;
  }
  for (; i < n; ++i) {
    if (kProd) { dst[i] = run; run = __fmul_rn(run, src[i]); }
    else { run = __fadd_rn(run, src[i]); dst[i] = run; }
  }
}

// A ray is processed by a TEAM of W warps (W = 1 in the stand-alone ops; 16 / G in the fused kernel):
// thread t of T = 32 W, warp tw; sync() is __syncwarp for a one-warp team, a named barrier otherwise.
struct RayTeam {
  int t, T, tw, W, lane;
  uint32_t bar_id;
  __device__ __forceinline__ void sync() const {
    if (W == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(T) : "memory");
  }
};
constexpr int kTeamScratch = 4 * 16 + 16 + 16;   // floats: part4[W], psum[W], flags[W]  (W <= 16)

// Alpha compositing of one ray (utils.py:176-189) on ACTIVATED samples: act4[i] = (sigmoid rgb, alpha).
// On return act4[i].w holds the weight of sample i.  tf, tt: scratch [n padded to 4]; ts: [kTeamScratch].
__device__ inline void composite_ray_activated(float4* act4, int n, int white, float* rgb_out, float* weights_out, float* tf, float* tt,
                                               float* ts, const RayTeam& tm) {
  for (int i = tm.t; i < n; i += tm.T) tf[i] = __fadd_rn(__fsub_rn(1.f, act4[i].w), 1e-10f);
  tm.sync();
  if (tm.t == 0) serial_scan<true>(tf, tt, n);     // exclusive transmittance product, in torch.cumprod's order
  tm.sync();
  float cr = 0.f, cg = 0.f, cb = 0.f, acc = 0.f;
  for (int i = tm.t; i < n; i += tm.T) {
    const float4 r = act4[i];
    const float w = __fmul_rn(r.w, tt[i]);
    cr = fmaf(w, r.x, cr); cg = fmaf(w, r.y, cg); cb = fmaf(w, r.z, cb);
    acc = __fadd_rn(acc, w);
    act4[i].w = w;
    if (weights_out) weights_out[i] = w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cr += __shfl_xor_sync(0xffffffffu, cr, o); cg += __shfl_xor_sync(0xffffffffu, cg, o);
    cb += __shfl_xor_sync(0xffffffffu, cb, o); acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  float4* part4 = reinterpret_cast<float4*>(ts);
  if (tm.W > 1) {
    if (tm.lane == 0) part4[tm.tw] = make_float4(cr, cg, cb, acc);
    tm.sync();
    if (tm.t == 0) {
      cr = cg = cb = acc = 0.f;
      for (int k = 0; k < tm.W; ++k) { const float4 p = part4[k]; cr += p.x; cg += p.y; cb += p.z; acc += p.w; }
    }
  }
  if (tm.t == 0 && rgb_out) {
    if (white) { const float bg = __fsub_rn(1.f, acc); cr = __fadd_rn(cr, bg); cg = __fadd_rn(cg, bg); cb = __fadd_rn(cb, bg); }
    rgb_out[0] = cr; rgb_out[1] = cg; rgb_out[2] = cb;
  }
  tm.sync();
}

// raw2outputs of one ray by one warp (utils.py:134-191): activation + compositing.  raw4: [n] (rgb_raw,
// sigma_raw) in smem; on return raw4[i].w holds the weight of sample i.
__device__ inline void composite_ray(float4* raw4, const float* z, const float* dnorm, float ray_norm, int n, const float* noise,
                              int white, float* rgb_out, float* alpha_out, float* weights_out, float* tf, float* tt, float* ts, int lane) {
  for (int i = lane; i < n; i += 32) {
    const float dz = (i < n - 1) ? __fsub_rn(z[i + 1], z[i]) : 1e10f;
    const float4 r = activate_sample(raw4[i], dz, dnorm ? dnorm[i] : ray_norm, noise ? noise + i : nullptr);
    raw4[i] = r;
    if (alpha_out) alpha_out[i] = r.w;
  }
  __syncwarp();
  const RayTeam tm = {lane, 32, 0, 1, lane, 0u};
  composite_ray_activated(raw4, n, white, rgb_out, weights_out, tf, tt, ts, tm);
}

// Inverse-CDF sampling + sorted merge of one ray by a team (utils.py:194-264, torchsearchsorted
// side='right').  w[i*wstride] = coarse weights; zc[nc] coarse depths; writes zf[nc+nf].
// cdfx, pd: scratch [nc + 4 padded to 4] (16-byte aligned); zs: scratch [nf]; ts: [kTeamScratch].
__device__ inline void sample_ray(const float* w, int wstride, const float* zc, int nc, int nf, const float* u_fine, float* cdfx, float* pd,
                           float* zs, float* zf, float* z_new_out, float* ts, const RayTeam& tm) {
  const int m = nc - 1;     // bins = midpoints (m of them); cdf has m entries; m-1 weights
  float* psum = ts + 64;
  float* flags = ts + 80;
  // pdf numerators into pd[0..m-2]; their sum: butterfly per warp, then the warps' partials in a fixed order,
  // so every thread holds the same bits
  float part = 0.f;
  for (int i = tm.t; i < m - 1; i += tm.T) { const float wi = __fadd_rn(w[(i + 1) * wstride], 1e-5f); pd[i] = wi; part += wi; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (tm.W > 1) {
    if (tm.lane == 0) psum[tm.tw] = part;
    tm.sync();
    part = 0.f;
    for (int k = 0; k < tm.W; ++k) part += psum[k];
  } else {
    __syncwarp();
  }
  for (int i = tm.t; i < m - 1; i += tm.T) pd[i] = __fdiv_rn(pd[i], part);
  tm.sync();
  // cdf[0] = 0, cdf[1 + i] = inclusive sum i: the scan writes at a 16-byte aligned address, so cdf = cdfx + 3
  if (tm.t == 0) { cdfx[3] = 0.f; serial_scan<false>(pd, cdfx + 4, m - 1); }   // sequential, like torch.cumsum on the CPU
  tm.sync();
  const float* cdf = cdfx + 3;
  // right-sided bisection, branch-free with a fixed step count
  const int top_m = 1 << (31 - __clz(m));
  for (int j = tm.t; j < nf; j += tm.T) {
    const float u = u_fine[j];
    int lo = 0;
    for (int step = top_m; step > 0; step >>= 1) { const int p = lo + step; if (p <= m && cdf[p - 1] <= u) lo = p; }   // #{cdf <= u}
    const int below = max(0, lo - 1), above = min(m - 1, lo);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = __fmul_rn(.5f, __fadd_rn(zc[below + 1], zc[below]));
    const float b1 = __fmul_rn(.5f, __fadd_rn(zc[above + 1], zc[above]));
    float denom = __fsub_rn(c1, c0);
    if (denom < 1e-5f) denom = 1.f;
    const float tt = __fdiv_rn(__fsub_rn(u, c0), denom);
    const float sv = __fadd_rn(b0, __fmul_rn(tt, __fsub_rn(b1, b0)));
    zs[j] = sv;
    if (z_new_out) z_new_out[j] = sv;
  }
  tm.sync();
  bool sorted = true;
  for (int j = tm.t; j < nf - 1; j += tm.T) sorted = sorted && (zs[j] <= zs[j + 1]);
  for (int i = tm.t; i < nc - 1; i += tm.T) sorted = sorted && (zc[i] <= zc[i + 1]);
  sorted = __all_sync(0xffffffffu, sorted);
  if (tm.W > 1) {
    if (tm.lane == 0) flags[tm.tw] = sorted ? 1.f : 0.f;
    tm.sync();
    for (int k = 0; k < tm.W; ++k) sorted = sorted && (flags[k] != 0.f);
  }
  const int n = nc + nf;
  if (sorted) {
    // stable two-way merge by rank: coarse element i lands at i + #{zs < zc[i]},
    // new sample j at j + #{zc <= zs[j]}
    const int top_f = 1 << (31 - __clz(nf)), top_c = 1 << (31 - __clz(nc));
    for (int e = tm.t; e < n; e += tm.T) {
      int lo = 0;
      if (e < nc) {
        const float v = zc[e];
        for (int step = top_f; step > 0; step >>= 1) { const int p = lo + step; if (p <= nf && zs[p - 1] < v) lo = p; }
        zf[e + lo] = v;
      } else {
        const float v = zs[e - nc];
        for (int step = top_c; step > 0; step >>= 1) { const int p = lo + step; if (p <= nc && zc[p - 1] <= v) lo = p; }
        zf[e - nc + lo] = v;
      }
    }
  } else {
    // general case (unsorted input depths): rank every element by counting
    for (int e = tm.t; e < n; e += tm.T) {
      const float v = e < nc ? zc[e] : zs[e - nc];
      int rank = 0;
      for (int o = 0; o < n; ++o) {
        const float x = o < nc ? zc[o] : zs[o - nc];
        rank += (x < v || (x == v && o < e)) ? 1 : 0;
      }
      zf[rank] = v;
    }
  }
  tm.sync();
}

}  // namespace nrf
