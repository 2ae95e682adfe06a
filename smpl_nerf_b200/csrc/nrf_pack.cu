// Host planner (net description -> MMA layer plan) and the weight packer kernels.
//
// Reference parameter shapes being lowered: models/render_ray_net.py:19-40 (RenderRayNet) and
// models/warp_field_net.py:14-15 (WarpFieldNet); nn.Linear stores W as [out, in] row-major and
// computes y = x W^T + b, which is exactly the "B operand N x K, K-major" form tcgen05.mma wants.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#include "nrf_plan.h"
#include "nrf_ptx.cuh"

namespace nrf {

// ------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
std::atomic<long long> g_train_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(int err, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(static_cast<cudaError_t>(err)));
  return NRF_E_CUDA;
}
const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------ planning
static uint32_t align4(uint32_t x) { return (x + 3u) & ~3u; }

static void finish_plan(NetPlan* p, uint32_t f32_floats) {
  p->flag_ofs = f32_floats; f32_floats += 4;       // range flag of the packed weights
  uint32_t ofs = 0;
  for (int i = 0; i < p->n_layers; ++i) {
    Layer& L = p->layers[i];
    L.stream_ofs = ofs;
    ofs += static_cast<uint32_t>(L.nk) * 4u * (L.n_out / 2u) * 128u;
  }
  p->stream_bytes = ofs;
  p->f32_ofs = (ofs + 1023u) & ~1023u;
  p->total_bytes = ((p->f32_ofs + f32_floats * 4u) + 1023u) & ~1023u;
}

int plan_raynet(const NrfRayNetDesc* d, NetPlan* p) {
  memset(p, 0, sizeof(*p));
  if (!d) { set_error("RayNet desc is NULL"); return NRF_E_INVALID; }
  const int P = enc_dim(d->pos_freqs, d->pos_identity), D = enc_dim(d->dir_freqs, d->dir_identity);
  const int A = d->additional_input_dim;
  if (d->width != kWidth) { set_error("RenderRayNet width %d unsupported (this build: %d)", d->width, kWidth); return NRF_E_INVALID; }
  if (d->positions_dim != P || P > kChunkK || P <= 0) { set_error("positions_dim %d does not match encoder (3*(id+2L)=%d, max %d)", d->positions_dim, P, kChunkK); return NRF_E_INVALID; }
  if (d->use_directional_input && (d->directions_dim != D || D > kChunkK || D <= 0)) { set_error("directions_dim %d does not match encoder (%d, max %d)", d->directions_dim, D, kChunkK); return NRF_E_INVALID; }
  const bool ext = d->ext_pose_bias != 0 && A > 0;
  if (A < 0 || (!ext && A > kMaxRayFeat) || A > 65535) { set_error("additional_input_dim %d unsupported (max %d in-kernel; set ext_pose_bias for more)", A, kMaxRayFeat); return NRF_E_INVALID; }
  if (d->n_layers < 2 || d->n_layers + 3 > kMaxLayers) { set_error("n_layers %d unsupported (2..%d)", d->n_layers, kMaxLayers - 3); return NRF_E_INVALID; }
  if (d->n_skips < 0 || d->n_skips > NRF_MAX_SKIPS) { set_error("n_skips %d unsupported", d->n_skips); return NRF_E_INVALID; }
  if (d->per_sample_dirs && A > 0) { set_error("per-sample directions with additional inputs is not a reference pipeline"); return NRF_E_INVALID; }
  auto is_skip = [&](int i) { for (int s = 0; s < d->n_skips; ++s) if (d->skips[s] == i) return true; return false; };

  p->in_freqs = d->pos_freqs; p->in_identity = d->pos_identity;
  p->dir_freqs = d->dir_freqs; p->dir_identity = d->dir_identity;
  uint32_t f = 0;   // float cursor in the fp32 section
  int slots = 0, n = 0, n_ext = 0;
  const int nl = d->n_layers;
  const bool fold = d->fold_linear != 0;
  auto add = [&](int n_out, int epi, int flags, int role, int pidx) -> Layer& {
    Layer& L = p->layers[n++];
    L.n_out = static_cast<uint16_t>(n_out); L.epi = static_cast<uint8_t>(epi); L.flags = static_cast<uint8_t>(flags);
    L.role = static_cast<uint8_t>(role); L.pidx = static_cast<uint8_t>(pidx);
    L.ray_slot = -1; L.ray_src = RAY_NONE; L.ray_k = 0; L.nk = 0;
    L.bias_ofs = f; f = align4(f + n_out);
    return L;
  };
  auto add_ray = [&](Layer& L, int src, int k) {
    L.ray_slot = static_cast<int8_t>(slots++);
    if (src == RAY_POSE && ext) { L.ray_src = RAY_POSE_EXT; L.ray_k = 0; L.ext_idx = static_cast<uint8_t>(n_ext++); return; }
    L.ray_src = static_cast<uint8_t>(src); L.ray_k = static_cast<uint16_t>(k);
    L.rayw_ofs = f; f = align4(f + k * L.n_out);
  };
  // first layer: xyz encoding from aux (+ pose features as a per-ray bias)
  { Layer& L = add(kWidth, EPI_RELU, LF_AUX_WAIT, ROLE_FIRST, 0); L.ksrc[L.nk++] = kSrcAux; if (A > 0) add_ray(L, RAY_POSE, A); }
  for (int i = 0; i < d->n_layers - 1; ++i) {
    // folded: the sigma head reads the last trunk layer's output directly (through w_sigma W_add)
    Layer& L = add(kWidth, EPI_RELU, (fold && i == d->n_layers - 2) ? LF_SIGMA_HEAD : 0, ROLE_TRUNK, 2 * (i + 1));
    if (is_skip(i)) { L.ksrc[L.nk++] = kSrcAux; if (A > 0) add_ray(L, RAY_POSE, A); }
    for (int j = 0; j < 4; ++j) L.ksrc[L.nk++] = static_cast<uint8_t>(j);
  }
  const bool dir_aux = d->use_directional_input && d->per_sample_dirs;
  if (dir_aux) {
    // the per-sample direction encoding replaces the xyz encoding in the aux tile as soon as its last reader (the
    // last skip layer, or the first layer) has finished its MMAs: written behind THAT layer's epilogue, where the
    // epilogue warps have slack, instead of on the critical path in front of the dir layer
    int last_aux = 0;
    for (int i = 0; i < d->n_layers - 1; ++i) if (is_skip(i)) last_aux = i + 1;
    p->layers[last_aux].flags |= LF_WRITE_DIRPE;
  }
  if (!fold) { Layer& L = add(kWidth, EPI_LINEAR, LF_SIGMA_HEAD, ROLE_LINEAR, 2 * nl); for (int j = 0; j < 4; ++j) L.ksrc[L.nk++] = static_cast<uint8_t>(j); }
  { Layer& L = add(kWidth / 2, EPI_LINEAR, dir_aux ? LF_AUX_WAIT : 0, ROLE_DIR, 2 * nl + 4);
    for (int j = 0; j < 4; ++j) L.ksrc[L.nk++] = static_cast<uint8_t>(j);
    if (dir_aux) L.ksrc[L.nk++] = kSrcAux;
    else if (d->use_directional_input) add_ray(L, RAY_DIR, D); }
  { Layer& L = add(kWidth / 2, EPI_RGB, 0, ROLE_RGB, 2 * nl + 6); L.ksrc[L.nk++] = 0; L.ksrc[L.nk++] = 1; }
  if (slots > kMaxRaySlots) { set_error("too many per-ray bias layers (%d > %d): at most one skip layer when additional_input_dim > 0", slots, kMaxRaySlots); return NRF_E_INVALID; }
  p->n_layers = n; p->n_ray_slots = slots; p->n_ext_slots = n_ext;
  p->sigma_ofs = f; f = align4(f + kWidth + 1);
  p->head_ofs = f;  f = align4(f + 3 * (kWidth / 2) + 3);
  p->folded = fold ? 1 : 0;
  if (fold) { p->fold_ofs = f; f = align4(f + (kWidth / 2) * (kWidth + (d->use_directional_input ? D : 0)) + kWidth / 2 + kWidth + 1); }
  finish_plan(p, f);
  return NRF_OK;
}

int plan_warpnet(const NrfWarpNetDesc* d, NetPlan* p) {
  memset(p, 0, sizeof(*p));
  if (!d) { set_error("WarpNet desc is NULL"); return NRF_E_INVALID; }
  const int P = enc_dim(d->in_freqs, d->in_identity);
  if (d->width != kWidth) { set_error("WarpFieldNet width %d unsupported (this build: %d)", d->width, kWidth); return NRF_E_INVALID; }
  if (d->positions_dim != P || P > kChunkK || P <= 0) { set_error("warp positions_dim %d does not match encoder (%d)", d->positions_dim, P); return NRF_E_INVALID; }
  if (d->pose_dim < 0 || d->pose_dim > kMaxRayFeat) { set_error("warp pose_dim %d unsupported (max %d)", d->pose_dim, kMaxRayFeat); return NRF_E_INVALID; }
  p->in_freqs = d->in_freqs; p->in_identity = d->in_identity;
  uint32_t f = 0;
  Layer& L = p->layers[0];
  // its one K-chunk (the xyz encoding) is staged in activation chunk 2, not in the aux tile: see warp_encode in nrf_fused.cu
  L.n_out = kWidth; L.epi = EPI_WARP; L.flags = 0; L.nk = 1; L.ksrc[0] = 2; L.role = ROLE_WARP; L.pidx = 0;
  L.bias_ofs = f; f = align4(f + kWidth);
  L.ray_slot = -1;
  if (d->pose_dim > 0) { L.ray_src = RAY_POSE; L.ray_k = static_cast<uint16_t>(d->pose_dim); L.ray_slot = 0; L.rayw_ofs = f; f = align4(f + d->pose_dim * kWidth); }
  p->n_layers = 1; p->n_ray_slots = d->pose_dim > 0 ? 1 : 0;
  p->head_ofs = f; f = align4(f + 3 * kWidth + 3);
  p->sigma_ofs = 0;
  finish_plan(p, f);
  return NRF_OK;
}

// ------------------------------------------------------------------------------ pack kernels
struct PackChunk {        // one K-chunk of one layer: four stages (hi/lo x 2 halves)
  const float* w;         // [n_out, ld]
  int32_t ld, n_out;
  int32_t aux;            // 0: columns col0 + k ; 1: columns col0 + enc_ref_col(k)
  int32_t col0;
  int32_t freqs, identity;
  uint32_t dst;           // byte offset of the chunk's first stage in the blob
};
struct PackTable { int32_t n; PackChunk c[kMaxLayers * kMaxK]; };

struct CopyJob { const float* src; int32_t rows, cols, ld, col0, transpose; uint32_t dst; };  // dst: float offset
struct CopyTable { int32_t n; CopyJob j[4 * kMaxLayers + 8]; };

__device__ __forceinline__ int dev_enc_ref_col(int f, int freqs, int identity) {
  if (f < 6 * freqs) { int p = f >> 1, s = f & 1, comp = p / freqs, k = p % freqs; return (identity ? 3 : 0) + k * 6 + s * 3 + comp; }
  int c = f - 6 * freqs;
  return (identity && c < 3) ? c : -1;
}

__global__ void pack_stream_kernel(const __grid_constant__ PackTable t, uint8_t* __restrict__ blob, int32_t* __restrict__ range_flag) {
  const PackChunk& c = t.c[blockIdx.x];
  const int part = blockIdx.y, half = part & 1, is_lo = part >> 1;   // stage order: hi/half0, hi/half1, lo/half0, lo/half1
  const int rows = c.n_out / 2;
  uint8_t* stage = blob + c.dst + static_cast<uint32_t>(part) * rows * 128u;
  for (int idx = threadIdx.x; idx < rows * kChunkK; idx += blockDim.x) {
    const int n = idx >> 6, k = idx & 63;
    int col = c.aux ? dev_enc_ref_col(k, c.freqs, c.identity) : k;
    float w = 0.f;
    if (col >= 0) w = c.w[static_cast<size_t>(half * rows + n) * c.ld + c.col0 + col];
    if (!(fabsf(w) <= 65504.f)) *range_flag = 1;       // (also catches NaN) the fp16 hi/lo split cannot hold this weight: the renderer raises status bit 0
    __half hi, lo;
    split_f16(w, hi, lo);
    *reinterpret_cast<__half*>(stage + sw128_offset(n, k)) = is_lo ? lo : hi;
  }
}

__global__ void pack_f32_kernel(const __grid_constant__ CopyTable t, float* __restrict__ f32) {
  const CopyJob& j = t.j[blockIdx.x];
  const int total = j.rows * j.cols;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int r = idx / j.cols, c = idx % j.cols;
    const float v = j.src[static_cast<size_t>(r) * j.ld + j.col0 + c];
    f32[j.dst + (j.transpose ? c * j.rows + r : idx)] = v;
  }
}

// additional_linear_layer folded into its two consumers (models/render_ray_net.py:51-57: no activation in between):
//   Wdir'[n, k] = sum_j Wdir[n, j] Wadd[j, k]   (k < 256; the direction columns are copied)     bdir' = bdir + Wdir[:, :256] badd
//   wsig'[k]    = sum_j wsig[j] Wadd[j, k]                                                       bsig' = bsig + wsig . badd
// fp64 accumulation, rounded ONCE to fp32.  out: Wdir'[half][ld] | bdir'[half] | wsig'[256] | bsig'[1]
__global__ void fold_linear_kernel(const float* __restrict__ w_add, const float* __restrict__ b_add, const float* __restrict__ w_sig,
                                   const float* __restrict__ b_sig, const float* __restrict__ w_dir, const float* __restrict__ b_dir,
                                   int ld, float* __restrict__ out) {
  constexpr int half = kWidth / 2;
  float* wd = out; float* bd = out + half * ld; float* ws = bd + half; float* bs = ws + kWidth;
  const int total = (half + 1) * kWidth;              // rows 0..half-1: Wdir', row half: wsig'
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx / kWidth, k = idx - n * kWidth;
    const float* row = n < half ? w_dir + static_cast<size_t>(n) * ld : w_sig;
    double acc = 0.0;
    for (int j = 0; j < kWidth; ++j) acc = fma(static_cast<double>(row[j]), static_cast<double>(w_add[j * kWidth + k]), acc);
    if (n < half) wd[n * ld + k] = static_cast<float>(acc); else ws[k] = static_cast<float>(acc);
  }
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < half * (ld - kWidth); idx += gridDim.x * blockDim.x) {
    const int n = idx / (ld - kWidth), k = kWidth + idx % (ld - kWidth);
    wd[n * ld + k] = w_dir[static_cast<size_t>(n) * ld + k];
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n <= half; n += gridDim.x * blockDim.x) {
    const float* row = n < half ? w_dir + static_cast<size_t>(n) * ld : w_sig;
    double acc = n < half ? static_cast<double>(b_dir[n]) : static_cast<double>(b_sig[0]);
    for (int j = 0; j < kWidth; ++j) acc = fma(static_cast<double>(row[j]), static_cast<double>(b_add[j]), acc);
    if (n < half) bd[n] = static_cast<float>(acc); else bs[0] = static_cast<float>(acc);
  }
}

static int launch_pack(const NetPlan& plan, PackTable& pt, CopyTable& ct, void* packed, cudaStream_t s, bool zero = true) {
  if ((reinterpret_cast<uintptr_t>(packed) & 1023u) != 0) { set_error("packed buffer must be 1024-byte aligned"); return NRF_E_INVALID; }
  cudaError_t e = cudaSuccess;
  if (zero) e = cudaMemsetAsync(packed, 0, plan.total_bytes, s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(packed)");
  pack_stream_kernel<<<dim3(pt.n, 4), 256, 0, s>>>(pt, static_cast<uint8_t*>(packed),
                                                   reinterpret_cast<int32_t*>(static_cast<uint8_t*>(packed) + plan.f32_ofs) + plan.flag_ofs);
  pack_f32_kernel<<<ct.n, 256, 0, s>>>(ct, reinterpret_cast<float*>(static_cast<uint8_t*>(packed) + plan.f32_ofs));
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pack kernels");
  return NRF_OK;
}

}  // namespace nrf

using namespace nrf;

extern "C" const char* nrf_last_error(void) { return nrf::last_error(); }
extern "C" int nrf_abi_version(void) { return NRF_ABI_VERSION; }

extern "C" int nrf_device_supported(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor); return NRF_E_UNSUPPORTED; }
  return NRF_OK;
}

extern "C" size_t nrf_raynet_packed_bytes(const NrfRayNetDesc* d) {
  NetPlan p;
  return plan_raynet(d, &p) == NRF_OK ? p.total_bytes : 0;
}
extern "C" size_t nrf_warpnet_packed_bytes(const NrfWarpNetDesc* d) {
  NetPlan p;
  return plan_warpnet(d, &p) == NRF_OK ? p.total_bytes : 0;
}

extern "C" int nrf_pack_raynet(const NrfRayNetDesc* d, const float* const* params, int n_params, void* packed, void* stream) {
  NetPlan plan;
  int rc = plan_raynet(d, &plan);
  if (rc != NRF_OK) return rc;
  const int nl = d->n_layers;
  if (!params || n_params != 2 * (nl + 5)) { set_error("RenderRayNet expects %d parameter tensors, got %d", 2 * (nl + 5), n_params); return NRF_E_INVALID; }
  if (!packed) { set_error("packed is NULL"); return NRF_E_INVALID; }
  for (int i = 0; i < n_params; ++i) if (!params[i]) { set_error("parameter %d is NULL", i); return NRF_E_INVALID; }
  const int A = d->additional_input_dim, P = d->positions_dim, D = d->directions_dim;
  static thread_local PackTable pt; static thread_local CopyTable ct;
  pt.n = 0; ct.n = 0;
  auto copy = [&](const float* src, int rows, int cols, int ld, int col0, int tr, uint32_t dst) {
    CopyJob& j = ct.j[ct.n++]; j.src = src; j.rows = rows; j.cols = cols; j.ld = ld; j.col0 = col0; j.transpose = tr; j.dst = dst;
  };
  float* f32 = reinterpret_cast<float*>(static_cast<uint8_t*>(packed) + plan.f32_ofs);
  const int ld_dir = d->use_directional_input ? kWidth + D : kWidth;
  const float* w_dir = params[2 * nl + 4], *b_dir = params[2 * nl + 5], *w_sig = params[2 * nl + 2], *b_sig = params[2 * nl + 3];
  bool zeroed = false;
  if (plan.folded) {
    // folded tensors live in the blob's fp32 section; the pack kernels below read them from there (same stream)
    if ((reinterpret_cast<uintptr_t>(packed) & 1023u) != 0) { set_error("packed buffer must be 1024-byte aligned"); return NRF_E_INVALID; }
    cudaError_t e = cudaMemsetAsync(packed, 0, plan.total_bytes, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(packed)");
    zeroed = true;
    float* fo = f32 + plan.fold_ofs;
    fold_linear_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(params[2 * nl], params[2 * nl + 1], w_sig, b_sig, w_dir, b_dir, ld_dir, fo);
    w_dir = fo; b_dir = fo + (kWidth / 2) * ld_dir; w_sig = b_dir + kWidth / 2; b_sig = w_sig + kWidth;
  }
  for (int li = 0; li < plan.n_layers; ++li) {
    const Layer& L = plan.layers[li];
    // which nn.Linear feeds this MMA layer, and where its input column blocks start
    int ld, act0 = 0, aux0 = 0, ray0 = 0, aux_fr = d->pos_freqs, aux_id = d->pos_identity;
    const float* W = params[L.pidx];
    const float* b = params[L.pidx + 1];
    if (L.role == ROLE_FIRST) { ld = A + P; ray0 = 0; aux0 = A; }
    else if (L.role == ROLE_TRUNK) {
      const bool skip = (L.ksrc[0] == kSrcAux);
      ld = skip ? kWidth + A + P : kWidth; ray0 = kWidth; aux0 = kWidth + A;
    } else if (L.role == ROLE_LINEAR) { ld = kWidth; }
    else if (L.role == ROLE_DIR) { W = w_dir; b = b_dir; ld = ld_dir; ray0 = kWidth; aux0 = kWidth; aux_fr = d->dir_freqs; aux_id = d->dir_identity; }
    else { ld = kWidth / 2; }
    for (int kc = 0; kc < L.nk; ++kc) {
      PackChunk& c = pt.c[pt.n++];
      c.w = W; c.ld = ld; c.n_out = L.n_out; c.dst = L.stream_ofs + static_cast<uint32_t>(kc) * 4u * (L.n_out / 2u) * 128u;
      if (L.ksrc[kc] == kSrcAux) { c.aux = 1; c.col0 = aux0; c.freqs = aux_fr; c.identity = aux_id; }
      else { c.aux = 0; c.col0 = act0 + kChunkK * L.ksrc[kc]; c.freqs = 0; c.identity = 0; }
    }
    copy(b, 1, L.n_out, L.n_out, 0, 0, L.bias_ofs);
    if (L.ray_src == RAY_POSE || L.ray_src == RAY_DIR) copy(W, L.n_out, L.ray_k, ld, ray0, 1, L.rayw_ofs);
  }
  copy(w_sig, 1, kWidth, kWidth, 0, 0, plan.sigma_ofs);
  copy(b_sig, 1, 1, 1, 0, 0, plan.sigma_ofs + kWidth);
  copy(params[2 * nl + 8], 3, kWidth / 2, kWidth / 2, 0, 0, plan.head_ofs);
  copy(params[2 * nl + 9], 1, 3, 3, 0, 0, plan.head_ofs + 3 * (kWidth / 2));
  (void)f32;
  return launch_pack(plan, pt, ct, packed, static_cast<cudaStream_t>(stream), !zeroed);
}

extern "C" int nrf_pack_warpnet(const NrfWarpNetDesc* d, const float* const* params, int n_params, void* packed, void* stream) {
  NetPlan plan;
  int rc = plan_warpnet(d, &plan);
  if (rc != NRF_OK) return rc;
  if (!params || n_params != 4) { set_error("WarpFieldNet expects 4 parameter tensors, got %d", n_params); return NRF_E_INVALID; }
  if (!packed) { set_error("packed is NULL"); return NRF_E_INVALID; }
  for (int i = 0; i < 4; ++i) if (!params[i]) { set_error("parameter %d is NULL", i); return NRF_E_INVALID; }
  static thread_local PackTable pt; static thread_local CopyTable ct;
  pt.n = 0; ct.n = 0;
  const Layer& L = plan.layers[0];
  const int ld = d->positions_dim + d->pose_dim;
  PackChunk& c = pt.c[pt.n++];
  c.w = params[0]; c.ld = ld; c.n_out = kWidth; c.aux = 1; c.col0 = 0; c.freqs = d->in_freqs; c.identity = d->in_identity; c.dst = L.stream_ofs;
  auto copy = [&](const float* src, int rows, int cols, int ld2, int col0, int tr, uint32_t dst) {
    CopyJob& j = ct.j[ct.n++]; j.src = src; j.rows = rows; j.cols = cols; j.ld = ld2; j.col0 = col0; j.transpose = tr; j.dst = dst;
  };
  copy(params[1], 1, kWidth, kWidth, 0, 0, L.bias_ofs);
  if (d->pose_dim > 0) copy(params[0], kWidth, d->pose_dim, ld, d->positions_dim, 1, L.rayw_ofs);
  copy(params[2], 3, kWidth, kWidth, 0, 0, plan.head_ofs);
  copy(params[3], 1, 3, 3, 0, 0, plan.head_ofs + 3 * kWidth);
  return launch_pack(plan, pt, ct, packed, static_cast<cudaStream_t>(stream));
}
