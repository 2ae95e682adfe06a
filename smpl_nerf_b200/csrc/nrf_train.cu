// Training path: the pipelines' forward computed LAYER BY LAYER with everything the backward needs kept in a caller-owned
// workspace, and the backward of solver/nerf_solver.py:81-87 (`loss = MSE(rgb) + MSE(rgb_fine); loss.backward()`) down to
// every nn.Linear parameter of the coarse, fine and warp-field nets.
//
//   forward (per pass):  encodings -> planes | per-ray inputs -> per-ray bias vectors | one tcgen05 GEMM per nn.Linear
//                        (tile_gemm: bias / ReLU / hi-lo split fused in the epilogue) | sigma / rgb / warp heads |
//                        compositing | inverse-CDF sampling (the sampler is detached, utils.py:260)
//   backward (per pass): compositing backward -> d raw | heads | per layer: dW = dY^T X (dw_gemm, split-K), db, per-ray-input
//                        columns, dX = dY W with ReLU' fused (tile_gemm, N-major weights) | SMPL: encodings' backward ->
//                        warped points -> warp net
// Gradients travel as fp16 hi/lo planes multiplied by ONE power-of-two scale derived on the device from max |d raw| and
// divided out where a parameter gradient is written.  precision 0 = 3 MMA passes everywhere (fp32-equivalent), 1 = 1 pass.
#include <cstdio>
#include <cstring>

#include "nrf_gemm.cuh"
#include "nrf_plan.h"
#include "nrf_ptx.cuh"
#include "nrf_train_kernels.cuh"

namespace nrf {

struct TLayer {
  int role, pidx, n_out, in_act, aux /*0 none, 1 xyz encoding, 2 direction encoding*/, aux_cols;
  int ray_src /*0, 1 pose, 2 dir*/, ray_k, ld, col_act, col_ray, col_aux, relu;
};
struct TNet { int n; int nl; int W; TLayer L[kMaxLayers]; int A, P, D; };      // W: hidden width of this net (128 / 256 / 512)

static int plan_train_net(const NrfRayNetDesc* d, TNet* t) {
  NetPlan tmp;
  NrfRayNetDesc dd = *d;
  dd.fold_linear = 0;
  dd.ext_pose_bias = d->additional_input_dim > 0 ? 1 : 0;     // any A is fine here: pose inputs are always hoisted per ray
  dd.width = kWidth;                                           // the layer-by-layer path is not tied to the fused kernel's width ...
  int rc = plan_raynet(&dd, &tmp);                             // ... every other shape check is shared with the renderer
  if (rc != NRF_OK) return rc;
  if (d->width != 128 && d->width != 256 && d->width != 512) { set_error("RenderRayNet width %d unsupported (128, 256 or 512)", d->width); return NRF_E_INVALID; }
  memset(t, 0, sizeof(*t));
  const int nl = d->n_layers, W = d->width, A = d->additional_input_dim, P = d->positions_dim;
  t->W = W;
  const int D = d->use_directional_input ? d->directions_dim : 0;
  t->nl = nl; t->A = A; t->P = P; t->D = D;
  auto is_skip = [&](int i) { for (int s = 0; s < d->n_skips; ++s) if (d->skips[s] == i) return true; return false; };
  int n = 0;
  { TLayer& L = t->L[n++]; L.role = ROLE_FIRST; L.pidx = 0; L.n_out = W; L.aux = 1; L.aux_cols = P; L.ray_src = A > 0 ? 1 : 0; L.ray_k = A;
    L.ld = A + P; L.col_ray = 0; L.col_aux = A; L.relu = 1; }
  for (int i = 0; i < nl - 1; ++i) {
    TLayer& L = t->L[n++]; L.role = ROLE_TRUNK; L.pidx = 2 * (i + 1); L.n_out = W; L.in_act = W; L.relu = 1; L.ld = W;
    if (is_skip(i)) { L.aux = 1; L.aux_cols = P; L.ray_src = A > 0 ? 1 : 0; L.ray_k = A; L.ld = W + A + P; L.col_ray = W; L.col_aux = W + A; }
  }
  { TLayer& L = t->L[n++]; L.role = ROLE_LINEAR; L.pidx = 2 * nl; L.n_out = W; L.in_act = W; L.ld = W; }
  { TLayer& L = t->L[n++]; L.role = ROLE_DIR; L.pidx = 2 * nl + 4; L.n_out = W / 2; L.in_act = W; L.ld = W + D;
    if (D > 0) { if (d->per_sample_dirs) { L.aux = 2; L.aux_cols = D; L.col_aux = W; } else { L.ray_src = 2; L.ray_k = D; L.col_ray = W; } } }
  { TLayer& L = t->L[n++]; L.role = ROLE_RGB; L.pidx = 2 * nl + 6; L.n_out = W / 2; L.in_act = W / 2; L.ld = W / 2; L.relu = 1; }
  t->n = n;
  return NRF_OK;
}

// ------------------------------------------------------------------------------ workspace (bump allocator: same walk for sizing and use)
struct Bump {
  uint8_t* base; size_t off;
  template <typename T> T* take(size_t count) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  Planes planes(int64_t rows, int cols, int n_planes) {      // n_planes: 1 (hi), 2 (hi, lo) or 3 (exact mode: bf16 hi, lo, ll)
    Planes p; p.rows = rows; p.cols = cols; p.ld = cols;
    p.hi = take<__half>(static_cast<size_t>(rows) * cols);
    p.lo = n_planes >= 2 ? take<__half>(static_cast<size_t>(rows) * cols) : nullptr;
    p.ll = n_planes >= 3 ? take<__half>(static_cast<size_t>(rows) * cols) : nullptr;
    return p;
  }
};

struct NetWeights { Planes act[kMaxLayers], aux[kMaxLayers]; };
constexpr int kScaleSlots = 64, kWmaxSlots = 48;
// wmax slots: [p * 20 + l] activation block of layer l of net p; [p * 20 + 16] rgb head, [p * 20 + 17] sigma head; [40] warp head
static inline int wslot(int p, int l) { return p * 20 + l; }
struct PassWs {
  int64_t S; int n;
  Planes encx, encd, wpe, warph, act[kMaxLayers];
  uint32_t* bits[kMaxLayers];      // ReLU' bit masks of the ReLU layers ([S, n_out / 32] words): what the backward reads instead of the hi plane
  float *a_f32, *h2_f32, *warph_f32, *raw, *dnorm, *warp_raw, *warped, *u, *rb[kMaxLayers], *rbw, *g_raw, *g_dnorm, *z;
  const float* pts;
};
struct TrainWs {
  NetWeights w[2]; Planes warp_w;
  PassWs pass[2];
  float *pose_feat, *dir_feat, *ray_norm, *weights_c, *alpha_c, *z_all, *pts_fine, *warp_pose_feat;
  Planes dy[2];
  float *dysum, *g_encx, *g_encd, *g_warp, *partial, *cs_partial;
  float* sc;               // [kScaleSlots][2]: {s, 1/s} of every gradient plane tensor of the backward chain
  unsigned int* mx;        // [kScaleSlots]: row-L1 bounds (float bits); mx[0] = max |d raw| over both passes
  unsigned int* wmax;      // [kWmaxSlots]: max |W| per (net, layer) activation block and per head (float bits; written by the forward)
  size_t bytes;
};

struct TrainCfg {
  int kind, nc, nf, na, run_fine, white, passes, smpl, A_pose /*per-ray pose features of the pipeline*/, n_sel;
  int Ww /*warp net width*/, Wmax /*largest hidden width of any net*/;
  int64_t B;
};

static void layout_ws(Bump& m, const TrainCfg& c, const TNet net[2], TrainWs* ws) {
  const int lo = c.passes == 6 ? 3 : (c.passes == 3 ? 2 : 1);      // planes per tensor
  const int n_pass = c.run_fine ? 2 : 1;
  for (int p = 0; p < n_pass; ++p)
    for (int l = 0; l < net[p].n; ++l) {
      const TLayer& L = net[p].L[l];
      if (L.in_act) ws->w[p].act[l] = m.planes(L.n_out, L.in_act, lo);
      if (L.aux) ws->w[p].aux[l] = m.planes(L.n_out, 64, lo);
    }
  if (c.smpl) ws->warp_w = m.planes(c.Ww, 64, lo);
  ws->pose_feat = m.take<float>(static_cast<size_t>(c.B) * (c.A_pose > 0 ? c.A_pose : 1));
  ws->dir_feat = m.take<float>(static_cast<size_t>(c.B) * 64);
  ws->ray_norm = m.take<float>(c.B);
  ws->weights_c = m.take<float>(static_cast<size_t>(c.B) * c.nc);
  ws->alpha_c = m.take<float>(static_cast<size_t>(c.B) * c.nc);
  ws->z_all = m.take<float>(static_cast<size_t>(c.B) * c.na);
  ws->pts_fine = m.take<float>(static_cast<size_t>(c.B) * c.na * 3);
  int64_t Smax = 0;
  for (int p = 0; p < n_pass; ++p) {
    PassWs& w = ws->pass[p];
    w.n = p == 0 ? c.nc : c.na;
    w.S = c.B * w.n;
    Smax = w.S > Smax ? w.S : Smax;
    w.encx = m.planes(w.S, 64, lo);
    if (c.smpl) { w.encd = m.planes(w.S, 64, lo); w.wpe = m.planes(w.S, 64, lo); w.warph = m.planes(w.S, c.Ww, lo);
                  w.warph_f32 = m.take<float>(static_cast<size_t>(w.S) * c.Ww); w.warp_raw = m.take<float>(static_cast<size_t>(w.S) * 3);
                  w.warped = m.take<float>(static_cast<size_t>(w.S) * 3); w.u = m.take<float>(static_cast<size_t>(w.S) * 3);
                  w.rbw = m.take<float>(static_cast<size_t>(c.B) * c.Ww); w.g_dnorm = m.take<float>(w.S); }
    w.dnorm = m.take<float>(w.S);
    for (int l = 0; l < net[p].n; ++l) {
      const TLayer& L = net[p].L[l];
      w.act[l] = m.planes(w.S, L.n_out, lo);
      w.bits[l] = L.relu ? m.take<uint32_t>(static_cast<size_t>(w.S) * (L.n_out / 32)) : nullptr;
      if (L.ray_src) w.rb[l] = m.take<float>(static_cast<size_t>(c.B) * L.n_out);
    }
    w.a_f32 = m.take<float>(static_cast<size_t>(w.S) * net[p].W);
    w.h2_f32 = m.take<float>(static_cast<size_t>(w.S) * (net[p].W / 2));
    w.raw = m.take<float>(static_cast<size_t>(w.S) * 4);
    w.g_raw = m.take<float>(static_cast<size_t>(w.S) * 4);
  }
  ws->dy[0] = m.planes(Smax, c.Wmax, 2);
  ws->dy[1] = m.planes(Smax, c.Wmax, 2);
  ws->dysum = m.take<float>(static_cast<size_t>(c.B) * c.Wmax);
  if (c.smpl) { ws->g_encx = m.take<float>(static_cast<size_t>(Smax) * 64); ws->g_encd = m.take<float>(static_cast<size_t>(Smax) * 64);
                ws->g_warp = m.take<float>(static_cast<size_t>(Smax) * 3); }
  ws->partial = m.take<float>(static_cast<size_t>(148) * 128 * 256);     // split x M x N <= (148 / (M / 128)) x M x 256
  ws->cs_partial = m.take<float>(static_cast<size_t>(148) * 512);
  ws->sc = m.take<float>(2 * kScaleSlots);
  ws->mx = reinterpret_cast<unsigned int*>(m.take<float>(kScaleSlots));
  ws->wmax = reinterpret_cast<unsigned int*>(m.take<float>(kWmaxSlots));
  ws->bytes = (m.off + 255) & ~static_cast<size_t>(255);
}

static int make_cfg(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const NrfRayNetDesc* fine, const NrfWarpNetDesc* warp, int64_t B,
                    TrainCfg* c, TNet net[2]) {
  if (!pipe || !coarse) { set_error("train: pipe/coarse is NULL"); return NRF_E_INVALID; }
  if (pipe->kind != NRF_KIND_NERF && pipe->kind != NRF_KIND_SMPL && pipe->kind != NRF_KIND_APPEND) { set_error("train: unknown pipeline kind %d", pipe->kind); return NRF_E_INVALID; }
  if (B < 0) { set_error("train: B < 0"); return NRF_E_INVALID; }
  memset(c, 0, sizeof(*c));
  c->kind = pipe->kind; c->nc = pipe->n_coarse; c->run_fine = pipe->run_fine ? 1 : 0; c->nf = c->run_fine ? pipe->n_fine : 0;
  c->na = c->nc + c->nf; c->white = pipe->white_background ? 1 : 0; c->passes = pipe->precision == 1 ? 1 : (pipe->precision == 2 ? 6 : 3); c->B = B;
  c->smpl = pipe->kind == NRF_KIND_SMPL;
  if (c->nc < 2 || c->nc > 1024 || (c->run_fine && (c->nc < 3 || c->nf < 1 || c->na > 1024))) { set_error("train: unsupported sample counts %d + %d", c->nc, c->nf); return NRF_E_INVALID; }
  int rc;
  if ((rc = plan_train_net(coarse, &net[0])) != NRF_OK) return rc;
  if (c->run_fine) {
    if (!fine) { set_error("train: run_fine=1 needs the fine net"); return NRF_E_INVALID; }
    if ((rc = plan_train_net(fine, &net[1])) != NRF_OK) return rc;
    if (fine->additional_input_dim != coarse->additional_input_dim) { set_error("train: coarse/fine additional_input_dim differ"); return NRF_E_INVALID; }
  }
  c->Wmax = net[0].W;
  if (c->run_fine && net[1].W > c->Wmax) c->Wmax = net[1].W;
  if (pipe->kind == NRF_KIND_SMPL && warp && warp->width > c->Wmax) c->Wmax = warp->width;
  if (pipe->kind != NRF_KIND_NERF) {
    c->n_sel = pipe->pose_all ? pipe->pose_stride : 2;
    c->A_pose = pipe->pose_encoded ? c->n_sel * (2 * pipe->pose_freqs + (pipe->pose_identity ? 1 : 0)) : c->n_sel;
  }
  if (c->smpl) {
    if (!warp) { set_error("train: the smpl pipeline needs the warp net"); return NRF_E_INVALID; }
    if (!pipe->pose_encoded) { set_error("train: the smpl pipeline is trainable with human_pose_encoding=1 only (the reference's fine pass always feeds the warp net encoded inputs, smpl_nerf_pipeline.py:71-77)"); return NRF_E_INVALID; }
    if ((warp->width != 128 && warp->width != 256 && warp->width != 512) || warp->positions_dim != enc_dim(warp->in_freqs, warp->in_identity) || warp->positions_dim > 64) { set_error("train: unsupported warp net shape (width 128 / 256 / 512, <= 64 position features)"); return NRF_E_INVALID; }
    c->Ww = warp->width;
    if (warp->pose_dim != c->A_pose) { set_error("train: warp net pose_dim %d != pipeline pose features %d", warp->pose_dim, c->A_pose); return NRF_E_INVALID; }
    if (!coarse->per_sample_dirs || (c->run_fine && !fine->per_sample_dirs)) { set_error("train: smpl pipeline needs per_sample_dirs=1 nets"); return NRF_E_INVALID; }
  } else {
    if (coarse->per_sample_dirs) { set_error("train: per_sample_dirs=1 is only valid for the smpl pipeline"); return NRF_E_INVALID; }
    if (coarse->additional_input_dim != c->A_pose) { set_error("train: pose feature count %d does not match the net's additional_input_dim %d", c->A_pose, coarse->additional_input_dim); return NRF_E_INVALID; }
  }
  return NRF_OK;
}

static int sweep_alternates() {      // developer A/B: NRF_TRAIN_SWEEP=0 walks every GEMM front to back
  static int v = -1;
  if (v < 0) v = getenv("NRF_TRAIN_SWEEP") ? (atoi(getenv("NRF_TRAIN_SWEEP")) ? 1 : 0) : 1;
  return v;
}
static int grid1(int64_t total, int block) {
  int64_t g = (total + block - 1) / block;
  return static_cast<int>(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}
#define TRY(x) do { int rc__ = (x); if (rc__ != NRF_OK) return rc__; } while (0)
#define LAUNCH_CHECK(what) do { ++g_train_launches; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return cuda_fail(e__, what); } while (0)

struct TrainCtx {
  TrainCfg c; TNet net[2]; TrainWs ws;
  const NrfPipelineDesc* pipe; const NrfRayNetDesc* nd[2]; const NrfWarpNetDesc* wd;
  const float* const* par[3];      // coarse, fine, warp parameter tables (host arrays of device pointers)
  NrfRenderIO io;
  int n_sms; cudaStream_t st;
  int rev;                         // sweep direction of the next large GEMM: alternates, so that each starts where its predecessor ended (L2 reuse)
};

static int setup(TrainCtx& t, const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const float* const* pc, int n_pc,
                 const NrfRayNetDesc* fine, const float* const* pf, int n_pf, const NrfWarpNetDesc* warp, const float* const* pw, int n_pw,
                 const NrfRenderIO* io, int64_t B, void* workspace, size_t ws_bytes, int n_sms, void* stream) {
  TRY(make_cfg(pipe, coarse, fine, warp, B, &t.c, t.net));
  if (!io) { set_error("train: io is NULL"); return NRF_E_INVALID; }
  if (!pc || n_pc != 2 * (coarse->n_layers + 5)) { set_error("train: coarse RenderRayNet expects %d parameter tensors, got %d", 2 * (coarse->n_layers + 5), n_pc); return NRF_E_INVALID; }
  if (t.c.run_fine && (!pf || n_pf != 2 * (fine->n_layers + 5))) { set_error("train: fine RenderRayNet expects %d parameter tensors, got %d", 2 * (fine->n_layers + 5), n_pf); return NRF_E_INVALID; }
  if (t.c.smpl && (!pw || n_pw != 4)) { set_error("train: WarpFieldNet expects 4 parameter tensors, got %d", n_pw); return NRF_E_INVALID; }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) { set_error("train: workspace must be a 256-byte aligned device buffer"); return NRF_E_INVALID; }
  Bump m{static_cast<uint8_t*>(workspace), 0};
  layout_ws(m, t.c, t.net, &t.ws);
  if (t.ws.bytes > ws_bytes) { set_error("train: workspace too small (%zu bytes given, %zu needed)", ws_bytes, t.ws.bytes); return NRF_E_INVALID; }
  t.pipe = pipe; t.nd[0] = coarse; t.nd[1] = fine; t.wd = warp; t.par[0] = pc; t.par[1] = pf; t.par[2] = pw;
  t.io = *io; t.n_sms = n_sms; t.st = static_cast<cudaStream_t>(stream);
  if (!io->ray_samples || !io->ray_origin || !io->ray_dir || !io->z_vals) { set_error("train: ray inputs are NULL"); return NRF_E_INVALID; }
  if (t.c.kind != NRF_KIND_NERF && !io->goal_pose) { set_error("train: goal_pose is NULL"); return NRF_E_INVALID; }
  if (t.c.run_fine && !io->u_fine) { set_error("train: u_fine is NULL"); return NRF_E_INVALID; }
  return NRF_OK;
}

// ------------------------------------------------------------------------------ forward
static int split_weights(TrainCtx& t) {
  static thread_local SplitTable tab;
  tab.n = 0; tab.bf16 = t.c.passes == 6 ? 1 : 0;
  auto flush = [&]() -> int {
    if (tab.n == 0) return NRF_OK;
    split_planes_kernel<<<dim3(8, tab.n), 256, 0, t.st>>>(tab);
    tab.n = 0;
    LAUNCH_CHECK("split_planes_kernel");
    return NRF_OK;
  };
  cudaError_t e0 = cudaMemsetAsync(t.ws.wmax, 0, kWmaxSlots * sizeof(unsigned int), t.st);
  if (e0 != cudaSuccess) return cuda_fail(e0, "cudaMemsetAsync(wmax)");
  auto add = [&](const float* src, int rows, int cols, int ld, int col0, const Planes& dst, unsigned int* wmax) -> int {
    SplitJob& j = tab.j[tab.n++];
    j.src = src; j.rows = rows; j.cols = cols; j.ld = ld; j.col0 = col0; j.hi = dst.hi; j.lo = dst.lo; j.ll = dst.ll; j.ld_dst = dst.ld; j.cols_pad = dst.cols;
    j.wmax = wmax;
    return tab.n == 36 ? flush() : NRF_OK;
  };
  const int n_pass = t.c.run_fine ? 2 : 1;
  for (int p = 0; p < n_pass; ++p)
    for (int l = 0; l < t.net[p].n; ++l) {
      const TLayer& L = t.net[p].L[l];
      const float* W = t.par[p][L.pidx];
      if (!W || !t.par[p][L.pidx + 1]) { set_error("train: parameter %d of net %d is NULL", L.pidx, p); return NRF_E_INVALID; }
      if (L.in_act) TRY(add(W, L.n_out, L.in_act, L.ld, L.col_act, t.ws.w[p].act[l], t.ws.wmax + wslot(p, l)));
      if (L.aux) TRY(add(W, L.n_out, L.aux_cols, L.ld, L.col_aux, t.ws.w[p].aux[l], nullptr));
    }
  if (t.c.smpl) TRY(add(t.par[2][0], t.c.Ww, t.wd->positions_dim, t.wd->positions_dim + t.wd->pose_dim, 0, t.ws.warp_w, nullptr));
  TRY(flush());
  for (int p = 0; p < n_pass; ++p) {
    const int nl = t.net[p].nl;
    absmax_kernel<<<1, 128, 0, t.st>>>(t.par[p][2 * nl + 8], 3 * (t.net[p].W / 2), t.ws.wmax + wslot(p, 16));
    absmax_kernel<<<1, 128, 0, t.st>>>(t.par[p][2 * nl + 2], t.net[p].W, t.ws.wmax + wslot(p, 17));
  }
  if (t.c.smpl) absmax_kernel<<<1, 128, 0, t.st>>>(t.par[2][2], 3 * t.c.Ww, t.ws.wmax + 40);
  LAUNCH_CHECK("absmax_kernel");
  return NRF_OK;
}

static int ray_bias(TrainCtx& t, const float* W, int ld, int col0, int K, const float* bias, const float* feat, int n_out, float* out) {
  const size_t smem = static_cast<size_t>(8) * K * sizeof(float);
  if (smem > 200 * 1024) { set_error("train: %d per-ray features do not fit the bias kernel", K); return NRF_E_INVALID; }
  cudaError_t e = cudaFuncSetAttribute(ray_bias2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem > 48 * 1024 ? smem : 48 * 1024));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(ray_bias2)");
  ray_bias2_kernel<<<static_cast<unsigned>((t.c.B + 7) / 8), 256, smem, t.st>>>(W, ld, col0, K, bias, feat, t.c.B, n_out, out);
  LAUNCH_CHECK("ray_bias2_kernel");
  return NRF_OK;
}

static int heads(TrainCtx& t, const float* x, int64_t S, int K, const float* W, const float* b, int nh, float* out, int out_ld, int c0) {
  heads_kernel<<<grid1(S, 8), 256, static_cast<size_t>(nh) * K * sizeof(float), t.st>>>(x, S, K, W, b, nh, out, out_ld, c0);
  LAUNCH_CHECK("heads_kernel");
  return NRF_OK;
}

static int encode(TrainCtx& t, const float* x, int64_t S, int freqs, int identity, const Planes& dst) {
  encode_planes_kernel<<<static_cast<unsigned>((S + kEncRows - 1) / kEncRows), 256, 0, t.st>>>(x, S, freqs, identity, dst.hi, dst.lo, dst.ll);
  LAUNCH_CHECK("encode_planes_kernel");
  return NRF_OK;
}

static int forward_pass(TrainCtx& t, int p) {
  const TrainCfg& c = t.c;
  t.rev = 1;                       // the encodings were written front to back: the first GEMM starts at the end
  PassWs& w = t.ws.pass[p];
  const TNet& net = t.net[p];
  const NrfRayNetDesc* d = t.nd[p];
  const bool last = p == (c.run_fine ? 1 : 0);
  const float* pts = w.pts;
  if (c.smpl) {
    // ---- warp field (models/smpl_nerf_pipeline.py:30-49, 71-79): x -> x + W2 relu(W1 [enc(x), enc(pose)] + b1) + b2
    const int Pw = t.wd->positions_dim, Aw = t.wd->pose_dim;
    TRY(encode(t, pts, w.S, t.wd->in_freqs, t.wd->in_identity, w.wpe));
    const float* bias = t.par[2][1];
    if (Aw > 0) { TRY(ray_bias(t, t.par[2][0], Pw + Aw, Pw, Aw, t.par[2][1], t.ws.pose_feat, c.Ww, w.rbw)); bias = w.rbw; }
    TileGemmArgs g{};
    g.a[0] = w.wpe; g.b[0] = t.ws.warp_w; g.n_src = 1; g.N = c.Ww; g.passes = c.passes; g.epi = GEPI_PLANES; g.relu = 1;
    g.bias = bias; g.bias_ld = Aw > 0 ? c.Ww : 0; g.rows_per_ray = w.n; g.out = w.warph; g.out_f32 = w.warph_f32; g.out_f32_ld = c.Ww;
    g.status = t.io.status;
    g.reverse = t.rev & sweep_alternates(); t.rev ^= 1;
    TRY(launch_tile_gemm(g, t.n_sms, t.st));
    TRY(heads(t, w.warph_f32, w.S, c.Ww, t.par[2][2], t.par[2][3], 3, w.warp_raw, 3, 0));
    smpl_points_kernel<<<grid1(w.S, 256), 256, 0, t.st>>>(pts, w.warp_raw, t.io.ray_origin, w.S, w.n, w.warped, w.u, w.dnorm,
                                                           last ? t.io.warp_out : nullptr, last ? t.io.warped_out : nullptr);
    LAUNCH_CHECK("smpl_points_kernel");
    TRY(encode(t, w.warped, w.S, d->pos_freqs, d->pos_identity, w.encx));
    TRY(encode(t, w.u, w.S, d->dir_freqs, d->dir_identity, w.encd));
  } else {
    TRY(encode(t, pts, w.S, d->pos_freqs, d->pos_identity, w.encx));
  }
  for (int l = 0; l < net.n; ++l) {
    const TLayer& L = net.L[l];
    const float* W = t.par[p][L.pidx];
    const float* b = t.par[p][L.pidx + 1];
    const float* bias = b;
    if (L.ray_src) {
      TRY(ray_bias(t, W, L.ld, L.col_ray, L.ray_k, b, L.ray_src == 1 ? t.ws.pose_feat : t.ws.dir_feat, L.n_out, w.rb[l]));
      bias = w.rb[l];
    }
    TileGemmArgs g{};
    int ns = 0;
    if (L.in_act) { g.a[ns] = w.act[l - 1]; g.b[ns] = t.ws.w[p].act[l]; ++ns; }
    if (L.aux) { g.a[ns] = L.aux == 1 ? w.encx : w.encd; g.b[ns] = t.ws.w[p].aux[l]; ++ns; }
    g.n_src = ns; g.N = L.n_out; g.passes = c.passes; g.epi = GEPI_PLANES; g.relu = L.relu;
    g.bias = bias; g.bias_ld = L.ray_src ? L.n_out : 0; g.rows_per_ray = w.n; g.out = w.act[l];
    if (L.role == ROLE_LINEAR) { g.out_f32 = w.a_f32; g.out_f32_ld = net.W; }
    if (L.role == ROLE_RGB) { g.out_f32 = w.h2_f32; g.out_f32_ld = net.W / 2; }
    if (L.relu) { g.bits_out = w.bits[l]; g.bits_ld = L.n_out / 32; }
    g.status = t.io.status;
    g.reverse = t.rev & sweep_alternates(); t.rev ^= 1;
    TRY(launch_tile_gemm(g, t.n_sms, t.st));
  }
  const int nl = net.nl;
  TRY(heads(t, w.h2_f32, w.S, net.W / 2, t.par[p][2 * nl + 8], t.par[p][2 * nl + 9], 3, w.raw, 4, 0));
  TRY(heads(t, w.a_f32, w.S, net.W, t.par[p][2 * nl + 2], t.par[p][2 * nl + 3], 1, w.raw, 4, 3));
  if (float* tap = p == 0 ? t.io.raw_coarse : t.io.raw_fine) {      // debug tap (stage-wise parity tests)
    cudaError_t et = cudaMemcpyAsync(tap, w.raw, static_cast<size_t>(w.S) * 4 * sizeof(float), cudaMemcpyDeviceToDevice, t.st);
    if (et != cudaSuccess) return cuda_fail(et, "cudaMemcpyAsync(raw tap)");
  }
  // ---- compositing: the SMPL coarse pass scales its deltas by |warped - o| per sample, every other pass by |ray_direction|
  const bool per_sample = c.smpl && p == 0;
  const float* dn = per_sample ? w.dnorm : t.ws.ray_norm;
  const int wpb = 4;
  const size_t smem = static_cast<size_t>(wpb) * (4 * w.n + 4 * ((w.n + 3) & ~3) + kTeamScratch) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(composite_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem > 48 * 1024 ? smem : 48 * 1024));
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(composite_fwd)");
  float* rgb = p == 0 ? t.io.rgb : t.io.rgb_fine;
  float* alpha = last ? t.io.alpha_out : t.ws.alpha_c;
  if (!rgb) { set_error("train: rgb output of pass %d is NULL", p); return NRF_E_INVALID; }
  composite_fwd_kernel<<<grid1(c.B, wpb), wpb * 32, smem, t.st>>>(w.raw, w.z, dn, per_sample ? 1 : 0, p == 0 ? t.io.noise_coarse : t.io.noise_fine, c.B, w.n,
                                                                   c.white, rgb, p == 0 ? t.ws.weights_c : nullptr, alpha);
  LAUNCH_CHECK("composite_fwd_kernel");
  if (p == 0 && t.io.weights_coarse) {
    cudaError_t et = cudaMemcpyAsync(t.io.weights_coarse, t.ws.weights_c, static_cast<size_t>(c.B) * c.nc * sizeof(float), cudaMemcpyDeviceToDevice, t.st);
    if (et != cudaSuccess) return cuda_fail(et, "cudaMemcpyAsync(weights tap)");
  }
  return NRF_OK;
}

static int forward_all(TrainCtx& t) {
  const TrainCfg& c = t.c;
  if (c.B == 0) return NRF_OK;
  TRY(split_weights(t));
  {
    const NrfRayNetDesc* d = t.nd[0];
    const int D = (d->use_directional_input && !c.smpl) ? d->directions_dim : 0;
    ray_feats_kernel<<<grid1(c.B * (c.A_pose + D + 1), 256), 256, 0, t.st>>>(
        t.io.goal_pose, t.pipe->pose_stride, c.n_sel, t.pipe->pose_col0, t.pipe->pose_col1, t.pipe->pose_all ? 1 : 0, t.pipe->pose_freqs,
        t.pipe->pose_identity, t.pipe->pose_encoded ? 1 : 0, t.io.ray_dir, d->dir_freqs, d->dir_identity, c.B, c.A_pose, D, t.ws.pose_feat,
        t.ws.dir_feat, t.ws.ray_norm);
    LAUNCH_CHECK("ray_feats_kernel");
  }
  t.ws.pass[0].pts = t.io.ray_samples;
  t.ws.pass[0].z = const_cast<float*>(t.io.z_vals);
  TRY(forward_pass(t, 0));
  if (c.run_fine) {
    // the sampler is detached in the reference (utils.py:260): plain forward kernels, nothing saved but z_all / the points
    float* pts = t.io.samples_out ? t.io.samples_out : t.ws.pts_fine;
    if (t.io.z_all_in) {      // teacher-forced depths (stage-wise tests)
      points_from_z_kernel<<<grid1(c.B * c.na, 256), 256, 0, t.st>>>(t.io.ray_origin, t.io.ray_dir, t.io.z_all_in, c.B, c.na, t.ws.z_all, pts);
      LAUNCH_CHECK("points_from_z_kernel");
    } else {
      TRY(nrf_fine_sampling(t.io.ray_origin, t.io.ray_dir, t.io.z_vals, t.ws.weights_c, t.io.u_fine, c.B, c.nc, c.nf, t.ws.z_all, pts, t.st));
      ++g_train_launches;
    }
    if (t.io.z_all) {
      cudaError_t et = cudaMemcpyAsync(t.io.z_all, t.ws.z_all, static_cast<size_t>(c.B) * c.na * sizeof(float), cudaMemcpyDeviceToDevice, t.st);
      if (et != cudaSuccess) return cuda_fail(et, "cudaMemcpyAsync(z_all tap)");
    }
    t.ws.pass[1].pts = pts;
    t.ws.pass[1].z = t.ws.z_all;
    TRY(forward_pass(t, 1));
  }
  return NRF_OK;
}

// ------------------------------------------------------------------------------ backward
struct Grads { float* const* g[3]; };

static int dw_into(TrainCtx& t, const Planes& dy, const float* sc, int M, const Planes& x, int n0, int N, int cols, float* dst, int ld, int col0,
                   float* db = nullptr) {
  // dW[M, cols] += dY^T X[:, n0 : n0 + N], in blocks of at most 256 X columns (one TMEM accumulator); M < 128 (the 64-wide layers of
  // a width-128 net) runs as one 128-row tile whose upper half multiplies zero-filled (out-of-bounds) dY columns and is dropped
  const int Mp = (M + 127) & ~127;
  for (int nb = 0; nb < N; nb += 256) {
    const int Nb = N - nb < 256 ? N - nb : 256;
    const int cb = cols - nb < Nb ? cols - nb : Nb;
    if (cb <= 0) break;
    int split = 0;
    const bool with_db = db != nullptr && nb == 0;       // the bias gradient rides on the first block: (dY_hi + dY_lo)^T . ones
    TRY(launch_dw_gemm(dy, 0, Mp, x, n0 + nb, Nb, t.c.passes, t.ws.partial, 148, &split, t.n_sms, t.st, with_db ? t.ws.cs_partial : nullptr, t.rev & sweep_alternates()));
    t.rev ^= 1;
    const int main_blocks = (M * cb + 63) / 64;      // + the bias gradient's blocks (same launch)
    dw_reduce_kernel<<<main_blocks + (with_db ? (M + 63) / 64 : 0), kDwReduceThreads, 0, t.st>>>(t.ws.partial, split, Mp, M, Nb, cb, sc, dst, ld, col0 + nb,
                                                                                                main_blocks, with_db ? t.ws.cs_partial : nullptr, db);
    LAUNCH_CHECK("dw_reduce_kernel");
  }
  return NRF_OK;
}

static int colsum_and_bias(TrainCtx& t, const Planes& dy, const float* sc, int F, int n, float* db) {
  ray_colsum_kernel<<<static_cast<unsigned>(t.c.B), F, 0, t.st>>>(dy.hi, dy.lo, dy.ld, F, n, t.ws.dysum);
  LAUNCH_CHECK("ray_colsum_kernel");
  bias_grad_kernel<<<(F + 31) / 32, 1024, 0, t.st>>>(t.ws.dysum, t.c.B, F, sc, db);
  LAUNCH_CHECK("bias_grad_kernel");
  return NRF_OK;
}

static int rayfeat_dw(TrainCtx& t, const float* sc, const float* feat, int n_out, int K, float* dW, int ld, int col0) {
  // partial sums per ray slice into the dW scratch (n_out * K <= 512 * 1,408 floats per slice fits its 148 x 128 x 256), then the fixed-order reduce
  const size_t cap = static_cast<size_t>(148) * 128 * 256, per = static_cast<size_t>(n_out) * K;
  if (per > cap) { set_error("train: per-ray input block %d x %d too large", n_out, K); return NRF_E_INVALID; }
  const int slices = static_cast<int>(cap / per < static_cast<size_t>(kRayFeatSlices) ? cap / per : kRayFeatSlices);
  rayfeat_dw_kernel<<<dim3((K + 15) / 16, (n_out + 15) / 16, slices), 256, 0, t.st>>>(t.ws.dysum, feat, t.c.B, n_out, K, t.ws.partial);
  LAUNCH_CHECK("rayfeat_dw_kernel");
  const int main_blocks = static_cast<int>((per + 63) / 64);
  dw_reduce_kernel<<<main_blocks, kDwReduceThreads, 0, t.st>>>(t.ws.partial, slices, n_out, n_out, K, K, sc, dW, ld, col0, main_blocks, nullptr, nullptr);
  LAUNCH_CHECK("dw_reduce_kernel");
  return NRF_OK;
}

static int head_bwd(TrainCtx& t, const float* x, int64_t S, int K, const float* g, int g_ld, int c0, int nh, const float* W, const float* sc_out,
                    int relu_mask, float* dW, float* db, const Planes* dy, unsigned int* l1max) {
  const int rows = 256;           // K in {64, 128, 256, 512}: 256 threads = 1024 / K rows in flight
  head_bwd_kernel<<<static_cast<unsigned>((S + rows - 1) / rows), 256, 0, t.st>>>(x, S, K, g, g_ld, c0, nh, W, sc_out, relu_mask, dW, db,
                                                                                  dy ? dy->hi : nullptr, dy ? dy->lo : nullptr, dy ? dy->ld : 0, rows, l1max);
  LAUNCH_CHECK("head_bwd_kernel");
  return NRF_OK;
}

static int next_scale(TrainCtx& t, const unsigned int* lmax, float mul, const unsigned int* wmax, const unsigned int* ea, const unsigned int* eb, float* sc) {
  scale_from_bound_kernel<<<1, 1, 0, t.st>>>(lmax, mul, wmax, ea, eb, sc);
  LAUNCH_CHECK("scale_from_bound_kernel");
  return NRF_OK;
}

static Planes view(const Planes& p, int cols, int64_t rows) { Planes v = p; v.cols = cols; v.rows = rows; return v; }     // same storage, narrower logical extent (ld kept)

static int backward_pass(TrainCtx& t, int p, const Grads& G) {
  const TrainCfg& c = t.c;
  t.rev = 1;                       // head_bwd_kernel wrote the first dY front to back
  PassWs& w = t.ws.pass[p];
  const TNet& net = t.net[p];
  const int nl = net.nl;
  float* const* g = G.g[p];
  int cur = 0;
  // scale / bound slots of this pass's chain: slot 0 is global (max |d raw|), this pass uses 1 + 24 p ...
  int slot = 1 + 24 * p;
  auto SC = [&](int k) { return t.ws.sc + 2 * k; };
  auto MX = [&](int k) { return t.ws.mx + k; };
  // ---- rgb head -> dY of the last (ReLU) layer: |dY| <= 3 max|d raw| max|W_rgb|
  Planes dy = view(t.ws.dy[cur], net.W / 2, w.S);
  TRY(next_scale(t, MX(0), 3.f, t.ws.wmax + wslot(p, 16), nullptr, nullptr, SC(slot)));
  TRY(head_bwd(t, w.h2_f32, w.S, net.W / 2, w.g_raw, 4, 0, 3, t.par[p][2 * nl + 8], SC(slot), 1, g[2 * nl + 8], g[2 * nl + 9], &dy, MX(slot)));
  // ---- sigma head (its contribution to d(additional_linear_layer output) is the rank-1 term of the dir layer's dX epilogue)
  TRY(head_bwd(t, w.a_f32, w.S, net.W, w.g_raw, 4, 3, 1, t.par[p][2 * nl + 2], nullptr, 0, g[2 * nl + 2], g[2 * nl + 3], nullptr, nullptr));
  bool encx_written = false;
  for (int l = net.n - 1; l >= 0; --l) {
    const TLayer& L = net.L[l];
    float* dW = g[L.pidx];
    float* db = g[L.pidx + 1];
    const float* sc = SC(slot);
    dy = view(t.ws.dy[cur], L.n_out, w.S);
    // parameter gradients
    // (the bias gradient comes out of the first dW GEMM of the layer; layers with per-ray inputs need the per-ray sums of dY anyway)
    float* db_gemm = L.ray_src ? nullptr : db;
    if (L.in_act) { TRY(dw_into(t, dy, sc, L.n_out, w.act[l - 1], 0, L.in_act, L.in_act, dW, L.ld, L.col_act, db_gemm)); db_gemm = nullptr; }
    if (L.aux) TRY(dw_into(t, dy, sc, L.n_out, L.aux == 1 ? w.encx : w.encd, 0, 64, L.aux_cols, dW, L.ld, L.col_aux, db_gemm));
    if (L.ray_src) {
      TRY(colsum_and_bias(t, dy, sc, L.n_out, w.n, db));
      TRY(rayfeat_dw(t, sc, L.ray_src == 1 ? t.ws.pose_feat : t.ws.dir_feat, L.n_out, L.ray_k, dW, L.ld, L.col_ray));
    }
    // gradient of the encodings (only the SMPL pipeline differentiates them: they are functions of the warp net); REAL units
    if (c.smpl && L.aux) {
      TileGemmArgs a{};
      a.a[0] = dy; a.b[0] = t.ws.w[p].aux[l]; a.n_src = 1; a.b_mn = 1; a.N = 64; a.passes = c.passes; a.epi = GEPI_F32;
      a.out_f32 = L.aux == 1 ? t.ws.g_encx : t.ws.g_encd; a.out_f32_ld = 64; a.accumulate = (L.aux == 1 && encx_written) ? 1 : 0;
      a.sc_in = sc;
      a.reverse = t.rev & sweep_alternates(); t.rev ^= 1;
    TRY(launch_tile_gemm(a, t.n_sms, t.st));
      if (L.aux == 1) encx_written = true;
    }
    // dX -> the previous layer's dY (ReLU' of the previous layer fused; + the sigma head's rank-1 term below the dir layer)
    if (L.in_act) {
      const TLayer& Pv = net.L[l - 1];
      const bool dir = L.role == ROLE_DIR;
      // |dX| <= (largest row L1 norm of dY) max|W| (+ max|d sigma| max|w_sigma| for the rank-1 term)
      TRY(next_scale(t, MX(slot), 1.f, t.ws.wmax + wslot(p, l), dir ? MX(0) : nullptr, dir ? t.ws.wmax + wslot(p, 17) : nullptr, SC(slot + 1)));
      TileGemmArgs a{};
      a.a[0] = dy; a.b[0] = t.ws.w[p].act[l]; a.n_src = 1; a.b_mn = 1; a.N = L.in_act; a.passes = c.passes; a.epi = GEPI_PLANES;
      a.out = view(t.ws.dy[cur ^ 1], L.in_act, w.S);
      if (Pv.relu) { a.mask_bits = w.bits[l - 1]; a.bits_ld = Pv.n_out / 32; }
      if (dir) { a.row_scale = w.g_raw + 3; a.row_scale_ld = 4; a.col_vec = t.par[p][2 * nl + 2]; }
      a.sc_in = sc; a.sc_out = SC(slot + 1); a.l1max = MX(slot + 1); a.status = t.io.status;
      a.reverse = t.rev & sweep_alternates(); t.rev ^= 1;
    TRY(launch_tile_gemm(a, t.n_sms, t.st));
      cur ^= 1;
      ++slot;
    }
  }
  if (c.smpl) {
    const NrfRayNetDesc* d = t.nd[p];
    ++slot;
    smpl_points_bwd_kernel<<<grid1((w.S + 7) / 8 * 8 * 4, 128), 128, 0, t.st>>>(t.ws.g_encx, d->pos_freqs, d->pos_identity, t.ws.g_encd, d->dir_freqs, d->dir_identity,
                                                               w.warped, w.u, w.dnorm, p == 0 ? w.g_dnorm : nullptr, w.S, t.ws.g_warp, MX(slot));
    LAUNCH_CHECK("smpl_points_bwd_kernel");
    float* const* gw = G.g[2];
    Planes dyw = view(t.ws.dy[0], c.Ww, w.S);
    TRY(next_scale(t, MX(slot), 3.f, t.ws.wmax + 40, nullptr, nullptr, SC(slot)));
    TRY(head_bwd(t, w.warph_f32, w.S, c.Ww, t.ws.g_warp, 3, 0, 3, t.par[2][2], SC(slot), 1, gw[2], gw[3], &dyw, nullptr));
    const int Pw = t.wd->positions_dim, Aw = t.wd->pose_dim;
    TRY(dw_into(t, dyw, SC(slot), c.Ww, w.wpe, 0, 64, Pw, gw[0], Pw + Aw, 0));
    TRY(colsum_and_bias(t, dyw, SC(slot), c.Ww, w.n, gw[1]));
    if (Aw > 0) TRY(rayfeat_dw(t, SC(slot), t.ws.pose_feat, c.Ww, Aw, gw[0], Pw + Aw, Pw));
  }
  return NRF_OK;
}

}  // namespace nrf

using namespace nrf;

extern "C" size_t nrf_train_workspace_bytes(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const NrfRayNetDesc* fine,
                                            const NrfWarpNetDesc* warp, int64_t B) {
  TrainCfg c; TNet net[2]; TrainWs ws;
  if (make_cfg(pipe, coarse, fine, warp, B, &c, net) != NRF_OK) return 0;
  Bump m{nullptr, 0};
  layout_ws(m, c, net, &ws);
  return ws.bytes;
}

extern "C" int nrf_train_forward(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const float* const* params_coarse, int n_coarse,
                                 const NrfRayNetDesc* fine, const float* const* params_fine, int n_fine, const NrfWarpNetDesc* warp,
                                 const float* const* params_warp, int n_warp, const NrfRenderIO* io, int64_t B, void* workspace,
                                 size_t workspace_bytes, int n_sms, void* stream) {
  static thread_local TrainCtx t;
  TRY(setup(t, pipe, coarse, params_coarse, n_coarse, fine, params_fine, n_fine, warp, params_warp, n_warp, io, B, workspace, workspace_bytes, n_sms, stream));
  if (!io->rgb || (t.c.run_fine && !io->rgb_fine)) { set_error("train: rgb / rgb_fine outputs are NULL"); return NRF_E_INVALID; }
  return forward_all(t);
}

extern "C" int nrf_train_backward(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const float* const* params_coarse, int n_coarse,
                                  const NrfRayNetDesc* fine, const float* const* params_fine, int n_fine, const NrfWarpNetDesc* warp,
                                  const float* const* params_warp, int n_warp, const NrfRenderIO* io, int64_t B, void* workspace,
                                  size_t workspace_bytes, const float* grad_rgb, const float* grad_rgb_fine, float* const* grads_coarse,
                                  float* const* grads_fine, float* const* grads_warp, int n_sms, void* stream) {
  static thread_local TrainCtx t;
  TRY(setup(t, pipe, coarse, params_coarse, n_coarse, fine, params_fine, n_fine, warp, params_warp, n_warp, io, B, workspace, workspace_bytes, n_sms, stream));
  const TrainCfg& c = t.c;
  if (c.B == 0) return NRF_OK;
  if (c.passes == 6) { set_error("train: precision 2 (exact, bf16 x 3) is an inference mode; train with precision 0 or 1"); return NRF_E_INVALID; }
  if (!grad_rgb || !grads_coarse || (c.run_fine && (!grad_rgb_fine || !grads_fine)) || (c.smpl && !grads_warp)) { set_error("train: gradient pointers are NULL"); return NRF_E_INVALID; }
  for (int i = 0; i < n_coarse; ++i) if (!grads_coarse[i]) { set_error("train: coarse gradient %d is NULL", i); return NRF_E_INVALID; }
  if (c.run_fine) for (int i = 0; i < n_fine; ++i) if (!grads_fine[i]) { set_error("train: fine gradient %d is NULL", i); return NRF_E_INVALID; }
  if (c.smpl) for (int i = 0; i < 4; ++i) if (!grads_warp[i]) { set_error("train: warp gradient %d is NULL", i); return NRF_E_INVALID; }
  // the workspace holds what nrf_train_forward left there for this batch; re-derive the pointers the forward set at run time
  t.ws.pass[0].pts = t.io.ray_samples; t.ws.pass[0].z = const_cast<float*>(t.io.z_vals);
  if (c.run_fine) { t.ws.pass[1].pts = t.io.samples_out ? t.io.samples_out : t.ws.pts_fine; t.ws.pass[1].z = t.ws.z_all; }
  cudaError_t e = cudaMemsetAsync(t.ws.mx, 0, kScaleSlots * sizeof(unsigned int), t.st);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(mx)");
  const int n_pass = c.run_fine ? 2 : 1;
  for (int p = 0; p < n_pass; ++p) {
    PassWs& w = t.ws.pass[p];
    const bool per_sample = c.smpl && p == 0;
    const int wpb = 4;
    const size_t smem = static_cast<size_t>(wpb) * (4 * w.n + 6 * ((w.n + 3) & ~3)) * sizeof(float);
    e = cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem > 48 * 1024 ? smem : 48 * 1024));
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(composite_bwd)");
    composite_bwd_kernel<<<grid1(c.B, wpb), wpb * 32, smem, t.st>>>(w.raw, w.z, per_sample ? w.dnorm : t.ws.ray_norm, per_sample ? 1 : 0,
                                                                     p == 0 ? t.io.noise_coarse : t.io.noise_fine, c.B, w.n, c.white,
                                                                     p == 0 ? grad_rgb : grad_rgb_fine, w.g_raw, per_sample ? w.g_dnorm : nullptr, t.ws.mx);
    LAUNCH_CHECK("composite_bwd_kernel");
  }
  Grads G; G.g[0] = grads_coarse; G.g[1] = grads_fine; G.g[2] = grads_warp;
  for (int p = 0; p < n_pass; ++p) TRY(backward_pass(t, p, G));
  return NRF_OK;
}

extern "C" long long nrf_train_launch_count(int reset) { return reset ? g_train_launches.exchange(0) : g_train_launches.load(); }

// ------------------------------------------------------------------------------ per-ray pose bias of AppendSmplParamsPipeline (tcgen05)
// out[b, e, n] = bias_e[n] + sum_k W_e[n, col0_e + k] * feats[b, k]     b < B, n < 256, k < A  (A = 69 or 1380)
// for the n_ext layers of a RenderRayNet that read the A additional inputs (models/render_ray_net.py:43-49 with the input built by
// models/append_smpl_params_pipeline.py:46-48: pose features FIRST).  The pose is constant along a ray, so this replaces A of
// the K columns of two layers for every SAMPLE by one GEMM per RAY -- three fp16 hi/lo passes on the tensor cores
// (tile_gemm, weight chunks streamed with the feature chunks: K = 1408 does not fit a resident slice).
static size_t ray_bias_ws(const NetPlan& plan, int64_t B, int A, Planes* feats, Planes w[NRF_MAX_SKIPS + 1], uint8_t* base) {
  const int Kp = (A + 63) & ~63;
  Bump m{base, 0};
  *feats = m.planes(B, Kp, 2);
  for (int e = 0; e < plan.n_ext_slots; ++e) w[e] = m.planes(kWidth, Kp, 2);
  return (m.off + 255) & ~static_cast<size_t>(255);
}

extern "C" size_t nrf_ray_bias_workspace_bytes(const NrfRayNetDesc* d, int64_t B) {
  NetPlan plan;
  if (plan_raynet(d, &plan) != NRF_OK || plan.n_ext_slots < 1 || B < 0) return 0;
  Planes f, w[NRF_MAX_SKIPS + 1];
  return ray_bias_ws(plan, B, d->additional_input_dim, &f, w, nullptr);
}

extern "C" int nrf_ray_bias(const NrfRayNetDesc* d, const float* const* params, int n_params, const float* feats, int64_t B,
                            float* out, int32_t* nonuniform, void* workspace, size_t workspace_bytes, void* stream) {
  static thread_local NetPlan plan;
  TRY(plan_raynet(d, &plan));
  if (plan.n_ext_slots < 1) { set_error("ray_bias: the net has no external pose-bias layers (ext_pose_bias = 0 or additional_input_dim = 0)"); return NRF_E_INVALID; }
  const int nl = d->n_layers, A = d->additional_input_dim, P = d->positions_dim;
  if (!params || n_params != 2 * (nl + 5)) { set_error("ray_bias: RenderRayNet expects %d parameter tensors, got %d", 2 * (nl + 5), n_params); return NRF_E_INVALID; }
  if (!feats || !out) { set_error("ray_bias: NULL argument"); return NRF_E_INVALID; }
  if (B < 0) { set_error("ray_bias: B < 0"); return NRF_E_INVALID; }
  if (B == 0) return NRF_OK;
  if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) { set_error("ray_bias: out must be 16-byte aligned"); return NRF_E_INVALID; }
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) { set_error("ray_bias: workspace must be a 256-byte aligned device buffer"); return NRF_E_INVALID; }
  Planes fp, wp[NRF_MAX_SKIPS + 1];
  const size_t need = ray_bias_ws(plan, B, A, &fp, wp, static_cast<uint8_t*>(workspace));
  if (need > workspace_bytes) { set_error("ray_bias: workspace too small (%zu bytes given, %zu needed)", workspace_bytes, need); return NRF_E_INVALID; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static thread_local SplitTable tab;
  tab.n = 0; tab.bf16 = 0;
  { SplitJob& j = tab.j[tab.n++]; j.ll = nullptr; j.src = feats; j.rows = static_cast<int32_t>(B); j.cols = A; j.ld = A; j.col0 = 0; j.hi = fp.hi; j.lo = fp.lo; j.ld_dst = fp.ld; j.cols_pad = fp.cols; j.wmax = nullptr; }
  const float* bias[NRF_MAX_SKIPS + 1];
  for (int li = 0; li < plan.n_layers; ++li) {
    const Layer& L = plan.layers[li];
    if (L.ray_src != RAY_POSE_EXT) continue;
    // layer 0 = positions_pose_input: columns [pose(A) | xyz(P)]; a skip layer: [activations(256) | pose(A) | xyz(P)]   (render_ray_net.py:22-31,43-49)
    if (!params[L.pidx] || !params[L.pidx + 1]) { set_error("ray_bias: parameter %d is NULL", L.pidx); return NRF_E_INVALID; }
    if ((reinterpret_cast<uintptr_t>(params[L.pidx + 1]) & 15u) != 0) { set_error("ray_bias: bias tensors must be 16-byte aligned"); return NRF_E_INVALID; }
    SplitJob& j = tab.j[tab.n++];
    j.ll = nullptr;
    j.src = params[L.pidx]; j.rows = kWidth; j.cols = A; j.ld = li == 0 ? A + P : kWidth + A + P; j.col0 = li == 0 ? 0 : kWidth;
    j.hi = wp[L.ext_idx].hi; j.lo = wp[L.ext_idx].lo; j.ld_dst = wp[L.ext_idx].ld; j.cols_pad = wp[L.ext_idx].cols; j.wmax = nullptr;
    bias[L.ext_idx] = params[L.pidx + 1];
  }
  split_planes_kernel<<<dim3(148, tab.n), 256, 0, st>>>(tab);
  LAUNCH_CHECK("split_planes_kernel");
  if (nonuniform) TRY(launch_rows_differ(feats, B, A, nonuniform, st));
  for (int e = 0; e < plan.n_ext_slots; ++e) {
    TileGemmArgs g{};
    g.a[0] = fp; g.b[0] = wp[e]; g.n_src = 1; g.N = kWidth; g.passes = 3; g.epi = GEPI_F32; g.bias = bias[e];
    g.out_f32 = out + static_cast<size_t>(e) * kWidth; g.out_f32_ld = plan.n_ext_slots * kWidth;
    TRY(launch_tile_gemm(g, 0, st));
  }
  return NRF_OK;
}

// ---- building blocks exported for stage-wise tests (tests/test_gpu_train.py) ----
extern "C" int nrf_split_planes(const float* src, int64_t rows, int32_t cols, int32_t ld, void* hi, void* lo, void* ll, int32_t ld_dst, int32_t cols_pad,
                                void* stream) {
  if (!src || !hi || rows < 0 || cols < 1 || cols_pad < cols || (ll && !lo)) { set_error("split_planes: bad arguments"); return NRF_E_INVALID; }
  if (rows == 0) return NRF_OK;
  static thread_local SplitTable tab;
  tab.n = 1; tab.bf16 = ll ? 1 : 0;
  SplitJob& j = tab.j[0];
  j.wmax = nullptr;
  j.src = src; j.rows = static_cast<int32_t>(rows); j.cols = cols; j.ld = ld; j.col0 = 0; j.hi = static_cast<__half*>(hi); j.lo = static_cast<__half*>(lo);
  j.ll = static_cast<__half*>(ll); j.ld_dst = ld_dst; j.cols_pad = cols_pad;
  split_planes_kernel<<<dim3(64, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(tab);
  LAUNCH_CHECK("split_planes_kernel");
  return NRF_OK;
}

extern "C" int nrf_gemm_planes(int32_t b_mn, const void* a_hi, const void* a_lo, const void* a_ll, int64_t S, int32_t K, const void* b_hi, const void* b_lo,
                               const void* b_ll, int32_t N, int32_t passes, const float* bias, int32_t relu, float* out_f32, void* out_hi, void* out_lo,
                               void* out_ll, void* stream) {
  auto H = [](const void* p) { return static_cast<__half*>(const_cast<void*>(p)); };
  TileGemmArgs g{};
  g.a[0] = Planes{H(a_hi), H(a_lo), S, K, K, H(a_ll)};
  g.b[0] = b_mn ? Planes{H(b_hi), H(b_lo), K, N, N, H(b_ll)} : Planes{H(b_hi), H(b_lo), N, K, K, H(b_ll)};
  g.n_src = 1; g.b_mn = b_mn; g.N = N; g.passes = passes; g.relu = relu; g.bias = bias;
  if (out_hi) { g.epi = GEPI_PLANES; g.out = Planes{H(out_hi), H(out_lo), S, N, N, H(out_ll)}; g.out_f32 = out_f32; g.out_f32_ld = N; }
  else { g.epi = GEPI_F32; g.out_f32 = out_f32; g.out_f32_ld = N; }
  return launch_tile_gemm(g, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int nrf_gemm_dw(const void* a_hi, const void* a_lo, int32_t M, const void* b_hi, const void* b_lo, int32_t N, int64_t S, int32_t passes,
                           float* partial, int32_t max_split, float* out /* [M, N], accumulated into */, void* stream) {
  if (!partial || !out || max_split < 1) { set_error("gemm_dw: bad arguments"); return NRF_E_INVALID; }
  Planes a{static_cast<__half*>(const_cast<void*>(a_hi)), static_cast<__half*>(const_cast<void*>(a_lo)), S, M, M, nullptr};
  Planes b{static_cast<__half*>(const_cast<void*>(b_hi)), static_cast<__half*>(const_cast<void*>(b_lo)), S, N, N, nullptr};
  int split = 0;
  TRY(launch_dw_gemm(a, 0, M, b, 0, N, passes, partial, max_split, &split, 0, static_cast<cudaStream_t>(stream)));
  // unit scale: scale2 = {1, 1} lives behind the partial sums
  float* scale2 = partial + static_cast<size_t>(max_split) * M * N;
  const float one[2] = {1.f, 1.f};
  cudaError_t e = cudaMemcpyAsync(scale2, one, sizeof(one), cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(scale)");
  dw_reduce_kernel<<<(M * N + 63) / 64, kDwReduceThreads, 0, static_cast<cudaStream_t>(stream)>>>(partial, split, M, M, N, N, scale2, out, N, 0, (M * N + 63) / 64,
                                                                                                        nullptr, nullptr);
  LAUNCH_CHECK("dw_reduce_kernel");
  return NRF_OK;
}
