/* nrf_b200.h -- C ABI of the B200-native NeRF volume-rendering engine (libnrf_b200.so).
 *
 * Drop-in boundary for the hot path of HannesStark/SMPL-NeRF (citations are paths in that repo):
 *
 *   models/nerf_pipeline.py:14-67, models/smpl_nerf_pipeline.py:16-100,
 *   models/append_to_nerf_pipeline.py:14-90          -> nrf_render()           (one fused kernel)
 *   models/append_smpl_params_pipeline.py:14-91        -> nrf_ray_bias() + nrf_render()   (69-parameter pose hoisted per ray)
 *   solver/nerf_solver.py:81-87 (pipeline(data); loss.backward()) -> nrf_train_forward() + nrf_train_backward()
 *   models/render_ray_net.py:6-61 (weights, [out,in])  -> nrf_pack_raynet()      (fp32 -> packed fp16 hi/lo)
 *   models/warp_field_net.py:6-21                      -> nrf_pack_warpnet()
 *   utils.py:114-131  PositionalEncoder.encode         -> nrf_positional_encoding()
 *   utils.py:134-191  raw2outputs                      -> nrf_raw2outputs()
 *   utils.py:194-228  sample_pdf                       -> nrf_sample_pdf()
 *   utils.py:231-264  fine_sampling                    -> nrf_fine_sampling()
 *   utils.py:26-54 get_rays + datasets/transforms.py:82-89 CoarseSampling -> nrf_generate_rays()
 *   torchsearchsorted/src/cuda/searchsorted_cuda_wrapper.cpp:18-20
 *       searchsorted_cuda_wrapper(a, v, res, side_left) -> nrf_searchsorted()
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name says host
 *   - all tensors are contiguous row-major fp32 (int64 for searchsorted results), as the reference's
 *     collate + .to(device) produces them
 *   - nothing here allocates device memory or synchronises: work is enqueued on `stream`
 *     (a cudaStream_t passed as void*); outputs and the workspace are caller-owned
 *   - return value 0 = ok, negative = error (NRF_E_*); nrf_last_error() gives the per-thread message
 *   - re-entrant per device; no global mutable state besides the per-thread error string
 */
#ifndef NRF_B200_H_
#define NRF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRF_ABI_VERSION 4

enum {
  NRF_OK = 0,
  NRF_E_INVALID = -1,     /* bad argument / unsupported shape (message says which) */
  NRF_E_CUDA = -2,        /* a CUDA runtime call failed (message carries cudaGetErrorString) */
  NRF_E_UNSUPPORTED = -3, /* device is not sm_100 */
};

enum { NRF_KIND_NERF = 0, NRF_KIND_SMPL = 1, NRF_KIND_APPEND = 2 };

#define NRF_MAX_SKIPS 4

/* RenderRayNet hyper-parameters (models/render_ray_net.py:8-17) + the encoders that feed it
 * (utils.py:115-125).  positions_dim must equal 3*(pos_identity + 2*pos_freqs), directions_dim
 * 3*(dir_identity + 2*dir_freqs): the engine computes the encodings itself. */
typedef struct NrfRayNetDesc {
  int32_t n_layers;              /* netdepth (>= 2)                                   */
  int32_t width;                 /* netwidth; this build supports 256                 */
  int32_t positions_dim;         /* 60                                                */
  int32_t directions_dim;        /* 24                                                */
  int32_t additional_input_dim;  /* A: pose features prepended to the xyz encoding     */
  int32_t use_directional_input; /* 0/1                                               */
  int32_t n_skips;
  int32_t skips[NRF_MAX_SKIPS];  /* indices into positional_net                       */
  int32_t pos_freqs, pos_identity;
  int32_t dir_freqs, dir_identity;
  int32_t per_sample_dirs;       /* 1: directions differ per sample (SmplNerfPipeline) */
  int32_t ext_pose_bias;         /* 1: the contribution of the A additional inputs to the first and the skip layers
                                    arrives precomputed per ray (nrf_ray_bias -> NrfRenderIO.ray_bias_*); lifts the
                                    A <= 64 limit (AppendSmplParamsPipeline: A = 69 or 1380)                       */
  int32_t fold_linear;           /* 1: fold additional_linear_layer (models/render_ray_net.py:51, no activation behind it) into its
                                    two consumers at pack time -- sigma_out_layer and directional_input (:52-57) -- as fp64
                                    products rounded once to fp32: W_dir' = W_dir[:, :W] W_add, w_sigma' = w_sigma W_add (biases
                                    alike).  Removes one 256x256 layer per sample (-10.8% MACs); fp32-level reassociation. */
} NrfRayNetDesc;

/* WarpFieldNet (models/warp_field_net.py:8-15): Linear(positions_dim + pose_dim -> width), ReLU,
 * Linear(width -> 3).  in_freqs/in_identity describe how xyz is encoded for it (10/0 when
 * human_pose_encoding=1, 0/1 = raw xyz otherwise). */
typedef struct NrfWarpNetDesc {
  int32_t width;         /* 256 */
  int32_t positions_dim; /* 3*(in_identity + 2*in_freqs) */
  int32_t pose_dim;      /* per-ray pose features appended after the xyz features */
  int32_t in_freqs, in_identity;
} NrfWarpNetDesc;

typedef struct NrfPipelineDesc {
  int32_t kind;             /* NRF_KIND_*                                              */
  int32_t n_coarse;         /* samples per ray in ray_samples / z_vals                 */
  int32_t n_fine;           /* args.number_fine_samples (ignored when run_fine == 0)   */
  int32_t run_fine;         /* args.run_fine                                           */
  int32_t white_background; /* args.white_background                                   */
  int32_t pose_freqs, pose_identity; /* human_pose_encoder                             */
  int32_t pose_encoded;     /* args.human_pose_encoding                                */
  int32_t pose_stride;      /* floats per row of goal_pose (69)                        */
  int32_t pose_col0, pose_col1; /* the two columns the pipelines read (38, 41)         */
  int32_t precision;        /* 0 = parity (fp16 hi/lo split, 3 MMA passes); 1 = fast (1 pass) */
  int32_t pose_all;         /* training entry points only: 1 = the pipeline feeds ALL pose_stride pose parameters
                               (AppendSmplParamsPipeline, models/append_smpl_params_pipeline.py:30-37), 0 = the two columns above */
} NrfPipelineDesc;

/* Inputs/outputs of one render call.  NULL is allowed where marked optional. */
typedef struct NrfRenderIO {
  /* inputs */
  const float* ray_samples;  /* [B, n_coarse, 3]                                             */
  const float* ray_origin;   /* [B, 3]  ("ray_translation")                                 */
  const float* ray_dir;      /* [B, 3]                                                       */
  const float* z_vals;       /* [B, n_coarse]                                                */
  const float* goal_pose;    /* [B, pose_stride]  (smpl / append kinds)                      */
  const float* u_fine;       /* [n_fine] = torch.linspace(0,1,n_fine) (utils.py:206)         */
  const float* noise_coarse; /* optional [B, n_coarse]: N(0,sigma_noise_std) draw (utils.py:174) */
  const float* noise_fine;   /* optional [B, n_coarse+n_fine]                                */
  const float* z_all_in;     /* optional [B, n_coarse+n_fine]: use these fine depths instead of
                                sampling (stage-wise parity tests, "teacher forcing")        */
  const float* ray_bias_coarse; /* [B, nrf_raynet_ext_slots, 256] from nrf_ray_bias (ext_pose_bias nets)   */
  const float* ray_bias_fine;   /* same for the fine net                                                    */
  const int32_t* ray_bias_nonuniform; /* optional [1] written by nrf_ray_bias: 0 = every ray has the pose of ray 0
                                   (a rendered frame), so only row 0 of ray_bias_* was computed and is read      */
  /* outputs (fp32) */
  float* rgb;         /* [B,3] coarse colour                                                 */
  float* rgb_fine;    /* [B,3] (run_fine=0: may be NULL; the pipelines return rgb twice)     */
  float* samples_out; /* [B,n,3] fine sample points (run_fine=1; n = n_coarse+n_fine)        */
  float* alpha_out;   /* [B,n]   "densities" = alpha of the last pass                        */
  float* warp_out;    /* [B,n,3] smpl kind: warp of the last pass                            */
  float* warped_out;  /* [B,n,3] smpl kind: warped samples of the last pass                  */
  /* optional debug taps */
  float* raw_coarse;     /* [B,n_coarse,4] (rgb_raw, sigma_raw) of the coarse net            */
  float* raw_fine;       /* [B,n,4]                                                          */
  float* weights_coarse; /* [B,n_coarse]                                                     */
  float* z_new;          /* [B,n_fine] inverse-CDF samples before the merge                  */
  float* z_all;          /* [B,n] merged depths                                              */
  int32_t* status;       /* optional [1]: bit0 set if an activation left the fp16 range      */
  long long* trace;      /* optional [1 + 3*cap], trace[0] = cap on entry: CTA 0 appends (event, layer
                            counter, SM clock) triples of its MMA issuer / epilogue timeline (developer tap) */
} NrfRenderIO;

const char* nrf_last_error(void);
int nrf_abi_version(void);
/* 0 if device `dev` can run the engine (compute capability 10.x), NRF_E_UNSUPPORTED otherwise. */
int nrf_device_supported(int dev);

/* Size in bytes of the packed form of a net (device buffer the caller allocates, 1024-aligned). */
size_t nrf_raynet_packed_bytes(const NrfRayNetDesc* d);
size_t nrf_warpnet_packed_bytes(const NrfWarpNetDesc* d);

/* Pack fp32 nn.Linear parameters ([out,in] row-major weights, biases) into the engine's layout.
 * `params` is a HOST array of DEVICE pointers in state_dict order:
 *   positions_pose_input.{weight,bias}, positional_net.{0..n_layers-2}.{weight,bias},
 *   additional_linear_layer.{weight,bias}, sigma_out_layer.{weight,bias},
 *   directional_input.{weight,bias}, directional_net.0.{weight,bias}, rgb_out_layer.{weight,bias}
 * Must be re-run whenever the parameters change (optimizer step, load_state_dict). */
int nrf_pack_raynet(const NrfRayNetDesc* d, const float* const* params, int n_params, void* packed, void* stream);
/* params: linear1.{weight,bias}, linear2.{weight,bias} */
int nrf_pack_warpnet(const NrfWarpNetDesc* d, const float* const* params, int n_params, void* packed, void* stream);

/* Per-ray bias vectors of the layers that read the A additional inputs (first layer and every skip layer):
 *   out[b, e, :] = bias_e + W_e[:, pose columns] * feats[b, :]        (tcgen05, three fp16 hi/lo passes, fp32 accumulate)
 * feats: [B, A] (pose parameters, positionally encoded by the caller when the pipeline encodes them);
 * params: as for nrf_pack_raynet (the ORIGINAL fp32 nn.Linear tensors are read in place); out: [B, n_ext, 256]
 * with n_ext = nrf_raynet_ext_slots(d).  Must run on the same stream before nrf_render.
 * workspace: caller-owned device buffer of nrf_ray_bias_workspace_bytes(d, B) bytes, 256-byte aligned (fp16 hi/lo planes of
 * the features and of the pose columns of the weights).
 * nonuniform (optional, DEVICE int32[1]): set to 0 when all B feature rows are bit-identical (every ray of a rendered
 * frame carries the same pose) -- nrf_render then reads row 0 only -- else to 1; hand it to nrf_render through
 * NrfRenderIO.ray_bias_nonuniform.  Decided on the device, no host synchronisation. */
int nrf_raynet_ext_slots(const NrfRayNetDesc* d);
size_t nrf_ray_bias_workspace_bytes(const NrfRayNetDesc* d, int64_t B);
int nrf_ray_bias(const NrfRayNetDesc* d, const float* const* params, int n_params, const float* feats, int64_t B,
                 float* out, int32_t* nonuniform, void* workspace, size_t workspace_bytes, void* stream);

/* The fused forward of the three pipelines for B rays.  packed_warp/warp may be NULL unless
 * kind == NRF_KIND_SMPL.  n_sms <= 0 means "all SMs of the current device". */
int nrf_render(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const void* packed_coarse,
               const NrfRayNetDesc* fine, const void* packed_fine, const NrfWarpNetDesc* warp,
               const void* packed_warp, const NrfRenderIO* io, int64_t n_rays, int n_sms, void* stream);

/* Number of kernels nrf_render launches per call (for bench.py's gpu_launches accounting). */
int nrf_render_launches(void);

/* ---- stand-alone ops (same arithmetic as the fused kernel's stages) ---- */
/* utils.py:127-131: x[n, c] -> out[n, c*(identity + 2*freqs)] */
int nrf_positional_encoding(const float* x, int64_t n, int32_t c, int32_t freqs, int32_t identity, float* out,
                            void* stream);
/* Backward of the encoding: grad_out[n, c*(identity + 2*freqs)] -> grad_x[n, c] (needed where the encoded quantity is itself
 * a network output: the SMPL warp field's warped samples, models/smpl_nerf_pipeline.py:49-50). */
int nrf_positional_encoding_backward(const float* x, const float* grad_out, int64_t n, int32_t c, int32_t freqs, int32_t identity,
                                     float* grad_x, void* stream);
/* utils.py:134-191: raw[B,n,4], z[B,n], dirs[B,n,3], optional noise[B,n] -> rgb[B,3], weights[B,n], alpha[B,n] */
int nrf_raw2outputs(const float* raw, const float* z, const float* dirs, const float* noise, int64_t B, int32_t n,
                    int32_t white_background, float* rgb, float* weights, float* alpha, void* stream);
/* Backward of raw2outputs (what loss.backward() in solver/nerf_solver.py:85 needs from this stage): d(loss)/d(raw)[B,n,4]
 * from d(loss)/d(rgb)[B,3] and the optional d(loss)/d(weights)[B,n], d(loss)/d(alpha)[B,n] (NULL = zero).  The forward
 * quantities are recomputed; `noise` must be the draw the forward used. */
int nrf_raw2outputs_backward(const float* raw, const float* z, const float* dirs, const float* noise, int64_t B, int32_t n,
                             int32_t white_background, const float* grad_rgb, const float* grad_weights,
                             const float* grad_alpha, float* grad_raw, void* stream);
/* utils.py:194-228: bins[B,m], weights[B,m-1], u[n_fine] -> samples[B,n_fine] */
int nrf_sample_pdf(const float* bins, const float* weights, const float* u, int64_t B, int32_t m, int32_t n_fine,
                   float* samples, void* stream);
/* utils.py:231-264: origin[B,3], dir[B,3], z[B,nc], weights[B,nc], u[n_fine] -> z_all[B,nc+nf], pts[B,nc+nf,3] */
int nrf_fine_sampling(const float* origin, const float* dir, const float* z, const float* weights, const float* u,
                      int64_t B, int32_t n_coarse, int32_t n_fine, float* z_all, float* pts, void* stream);
/* utils.py:26-54 get_rays + datasets/transforms.py:82-89 CoarseSampling + :13-19 ToTensor for one H x W view, on the
 * device: float64 arithmetic in numpy's operation order, rounded once to fp32.  camera_transform_host: HOST 4x4
 * row-major camera-to-world matrix; lower/span: DEVICE [n_coarse] float64 bin tables (lower edge, upper - lower;
 * built once on the host from near/far exactly like transforms.py:82-86); jitter: DEVICE [H*W] float64, the one
 * np.random.rand() scalar per ray (drawn by the caller so the host RNG stream is the reference's).
 * Outputs (fp32): ray_samples[H*W, n_coarse, 3], ray_origin[H*W, 3], ray_dir[H*W, 3], z_vals[H*W, n_coarse]. */
int nrf_generate_rays(int32_t H, int32_t W, double focal, const double* camera_transform_host, const double* lower,
                      const double* span, const double* jitter, int32_t n_coarse, float* ray_samples, float* ray_origin,
                      float* ray_dir, float* z_vals, void* stream);
/* The same for the rays [ray0, ray0 + n_rays) of the view only (row-major pixel index): a rank of a sharded render generates
 * just its own rays.  jitter: [n_rays] (the window's scalars); outputs are [n_rays, ...]. */
int nrf_generate_rays_range(int32_t H, int32_t W, double focal, const double* camera_transform_host, const double* lower,
                            const double* span, const double* jitter, int32_t n_coarse, int64_t ray0, int64_t n_rays,
                            float* ray_samples, float* ray_origin, float* ray_dir, float* z_vals, void* stream);
/* util/scores.py:88-173 ssim / _ssim_per_channel: x, y [n_planes, H, W] (n_planes = N * C channel planes), kernel2d: DEVICE
 * [ks, ks] window (the caller builds it like gaussian_filter, util/scores.py:68-86); c1 = (k1 data_range)^2, c2 = (k2 data_range)^2.
 * ssim_out / cs_out: [n_planes] means over the valid positions (cs_out optional).  partial: nrf_ssim_partial_floats() floats. */
int64_t nrf_ssim_partial_floats(int64_t n_planes, int32_t H, int32_t W, int32_t ks);
int nrf_ssim(const float* x, const float* y, int64_t n_planes, int32_t H, int32_t W, const float* kernel2d, int32_t ks, float c1,
             float c2, float* partial, float* ssim_out, float* cs_out, void* stream);
/* inference.py:260-262: rgb [n_pixels, 3] float -> uint8 (clip to [0,1], * 255, truncate), channels flipped to BGR when to_bgr. */
int nrf_quantize_image(const float* rgb, int64_t n_pixels, uint8_t* out, int32_t to_bgr, void* stream);
/* torchsearchsorted: a[rows_a, na], v[rows_v, nv] (rows broadcast when one side has 1 row) -> res int64 */
int nrf_searchsorted(const float* a, int64_t rows_a, int64_t na, const float* v, int64_t rows_v, int64_t nv,
                     int64_t* res, int32_t side_left, void* stream);

/* ---- training (the callers are solver/nerf_solver.py:81-87 and solver/smpl_nerf_solver.py:74-83:
 *      out = pipeline(data); loss = MSE(out[0], gt) + MSE(out[1], gt); loss.backward()) ----
 * nrf_train_forward computes the same outputs as nrf_render, layer by layer (one tcgen05 GEMM per nn.Linear), and leaves in
 * `workspace` (caller-owned device buffer of nrf_train_workspace_bytes(), 256-byte aligned) what nrf_train_backward needs.
 * params_*: HOST arrays of DEVICE pointers to the ORIGINAL fp32 nn.Linear tensors in state_dict order (as nrf_pack_*).
 * nrf_train_backward takes d(loss)/d(rgb) [B,3] and d(loss)/d(rgb_fine) [B,3] and ACCUMULATES d(loss)/d(parameter) into
 * grads_* (same order and shapes as params_*; the caller zeroes them).  The other outputs (sample points, alpha, warp) are
 * not differentiated, like the hierarchical sampler (utils.py:260 detaches it).  `io` must be the same in both calls.
 * pipe->precision: 0 = three fp16 MMA passes in forward and backward (fp32-equivalent), 1 = one pass (mixed precision). */
size_t nrf_train_workspace_bytes(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const NrfRayNetDesc* fine,
                                 const NrfWarpNetDesc* warp, int64_t B);
int nrf_train_forward(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const float* const* params_coarse, int n_coarse,
                      const NrfRayNetDesc* fine, const float* const* params_fine, int n_fine, const NrfWarpNetDesc* warp,
                      const float* const* params_warp, int n_warp, const NrfRenderIO* io, int64_t B, void* workspace,
                      size_t workspace_bytes, int n_sms, void* stream);
int nrf_train_backward(const NrfPipelineDesc* pipe, const NrfRayNetDesc* coarse, const float* const* params_coarse, int n_coarse,
                       const NrfRayNetDesc* fine, const float* const* params_fine, int n_fine, const NrfWarpNetDesc* warp,
                       const float* const* params_warp, int n_warp, const NrfRenderIO* io, int64_t B, void* workspace,
                       size_t workspace_bytes, const float* grad_rgb, const float* grad_rgb_fine, float* const* grads_coarse,
                       float* const* grads_fine, float* const* grads_warp, int n_sms, void* stream);
/* Kernels launched by the training entry points on this thread since the last reset (bench.py's gpu_launches accounting). */
long long nrf_train_launch_count(int reset);
/* Building blocks of the training path, exported for stage-wise tests.  "planes" = a matrix as two fp16 tensors hi, lo
 * (x ~= hi + lo), row-major [rows, ld]; lo may be NULL with passes = 1.  With a third plane ll (non-NULL) the three planes hold
 * BFLOAT16 bit patterns and x = hi + lo + ll exactly (the exact mode, passes = 6).
 * nrf_split_planes: fp32 [rows, cols] (row pitch ld) -> planes [rows, cols_pad] (row pitch ld_dst), zero padded.
 * nrf_gemm_planes:  b_mn = 0: C[S,N] = A[S,K] B[N,K]^T (the forward of nn.Linear);  b_mn = 1: C[S,N] = A[S,K] B[K,N] (dX = dY W).
 *                   out_hi != NULL: C = [relu](acc + bias) as planes (+ fp32 copy in out_f32 if given); else fp32 C in out_f32.
 * nrf_gemm_dw:      out[M,N] += A[S,M]^T B[S,N] (dW = dY^T X), split over the SMs; partial: >= (max_split * M * N + 2) floats. */
int nrf_split_planes(const float* src, int64_t rows, int32_t cols, int32_t ld, void* hi, void* lo, void* ll, int32_t ld_dst, int32_t cols_pad,
                     void* stream);
int nrf_gemm_planes(int32_t b_mn, const void* a_hi, const void* a_lo, const void* a_ll, int64_t S, int32_t K, const void* b_hi, const void* b_lo,
                    const void* b_ll, int32_t N, int32_t passes, const float* bias, int32_t relu, float* out_f32, void* out_hi, void* out_lo,
                    void* out_ll, void* stream);
int nrf_gemm_dw(const void* a_hi, const void* a_lo, int32_t M, const void* b_hi, const void* b_lo, int32_t N, int64_t S, int32_t passes,
                float* partial, int32_t max_split, float* out, void* stream);

/* tcgen05 self-test: D[128,256] = A[128,64] * B[256,64]^T with fp16 operands staged through the swizzled
 * shared-memory layout / descriptors the renderer uses (SWIZZLE_128B K-major tiles, N = 256 per
 * instruction).  a, b: fp32 (rounded to fp16 inside). */
int nrf_selftest_umma(const float* a, const float* b, float* d, void* stream);
/* The same through a CTA pair (tcgen05.mma.cta_group::2, M = 256): D[256,256] = A[256,64] * B[256,64]^T,
 * each CTA staging its 128 rows of A and its half of B, as the renderer does. */
int nrf_selftest_umma2(const float* a, const float* b, float* d, void* stream);

/* Developer diagnostic: tcgen05.mma issue-rate probe on n_ctas SMs (one CTA each).  mode bit0: N=256 per
 * instruction (else 128); bit1: stream `wsrc` (device, >= 64 KiB) through a TMA ring concurrently;
 * bit2: two A passes per B stage.  cycles[n_ctas] receives the SM-clock cycles of `iters` K=64 steps. */
int nrf_bench_umma(int mode, int iters, const void* wsrc, size_t wsrc_bytes, long long* cycles, int n_ctas, void* stream);
/* CTA-pair variant (cta_group::2, M=256, N=256, 16 KB half-stages per CTA through an n_slots-deep (<= 4) ring
 * with the renderer's relay protocol).  mode bit1: TMA stream on; bit2: two A passes per stage; bit4: relay
 * with a release.cluster arrive.  cycles[n_pairs]. */
int nrf_bench_umma2(int mode, int iters, int n_slots, const void* wsrc, size_t wsrc_bytes, long long* cycles, int n_pairs,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NRF_B200_H_ */
