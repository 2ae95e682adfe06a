// Internal plan structures shared by the host planner, the packer and the fused kernel.
//
// A "net" (RenderRayNet or WarpFieldNet) is lowered to a short list of MMA layers.  Each MMA layer
//   D[128 rows, N] = sum over K-chunks  A_chunk[128, 64] * W_chunk[N, 64]^T
// reads its K-chunks either from the activation buffer (chunks 0..3 = the previous layer's output,
// 64 features each) or from the "aux" buffer (the positional encoding the epilogue warps computed).
// Inputs that are constant along a ray (pose features, per-ray view direction) never enter the MMA:
// their contribution is a per-ray bias vector computed on CUDA cores (ray_src / ray_slot).
//
// Packed blob of a net (device memory, 1024-byte aligned):
//   [ weight stream | fp32 section ]
// weight stream = for each layer, for each 64-feature K-chunk, four stages
//   hi/half0, hi/half1, lo/half0, lo/half1
// where half h holds output features [h*N/2, (h+1)*N/2) and hi/lo are the fp16 split of the fp32
// weight (w ~= hi + lo).  A stage is a [N/2 x 64] fp16 K-major tile in the UMMA SWIZZLE_128B layout
// (16 KB for the 256-wide layers), i.e. exactly the bytes the tensor core wants in shared memory, so
// staging is one 1-D bulk copy.  The renderer runs as CTA pairs (tcgen05 cta_group::2): CTA r of a pair
// stages only half r of every stage and each M=256 x N instruction reads both halves.
#pragma once
#include <stdint.h>
#include <atomic>
#include <cuda_runtime.h>

#include "../../include/nrf_b200.h"

namespace nrf {

constexpr int kTileRows = 128;  // samples per MMA tile (UMMA M)
constexpr int kChunkK = 64;     // features per K-chunk (one 128-byte swizzle row of fp16)
constexpr int kWidth = 256;     // hidden width this build supports
constexpr int kMaxLayers = 16;
constexpr int kMaxK = 6;
constexpr int kSrcAux = 4;      // ksrc value meaning "aux buffer"
constexpr int kMaxRayFeat = 64; // max per-ray (pose) features folded into a bias
constexpr int kMaxRaySlots = 3;
constexpr int kMaxFineRows = 512;

enum EpiKind : uint8_t {
  EPI_RELU = 0,    // out = relu(acc + bias)            -> next layer's activations
  EPI_LINEAR = 1,  // out = acc + bias                  -> next layer's activations
  EPI_RGB = 2,     // h = relu(acc + bias); rgb_raw = Wrgb h + brgb   (last layer of RenderRayNet)
  EPI_WARP = 3,    // h = relu(acc + bias); warp = W2 h + b2          (WarpFieldNet)
};
enum LayerFlags : uint8_t {
  LF_SIGMA_HEAD = 1,   // also compute sigma_raw = wsigma . out + bsigma
  LF_WRITE_DIRPE = 2,  // epilogue also writes the per-sample direction encoding into aux
  LF_AUX_WAIT = 4,     // the MMA issuer must wait for fresh aux contents before this layer
};
enum RaySrc : uint8_t { RAY_NONE = 0, RAY_POSE = 1, RAY_DIR = 2, RAY_POSE_EXT = 3 };   // EXT: bias vector precomputed by nrf_ray_bias

struct Layer {
  uint32_t stream_ofs;  // byte offset of the layer's first stage inside the weight stream
  uint32_t bias_ofs;    // float offset (fp32 section): bias[n_out]
  uint32_t rayw_ofs;    // float offset: per-ray weight columns, transposed [ray_k][n_out]
  uint16_t n_out;       // 256 or 128
  uint16_t ray_k;       // number of per-ray input features (0 = none)
  uint8_t ray_src;      // RaySrc
  int8_t ray_slot;      // slot of the shared-memory per-ray bias table (-1: plain bias)
  uint8_t nk;           // number of K-chunks
  uint8_t ksrc[kMaxK];  // 0..3 activation chunk, kSrcAux
  uint8_t epi;          // EpiKind
  uint8_t flags;        // LayerFlags
  uint8_t ext_idx;      // RAY_POSE_EXT: index of this layer's vector inside a ray's [n_ext][256] block
  uint8_t role;         // LayerRole: which nn.Linear of the reference net feeds this MMA layer
  uint8_t pidx;         // index of that nn.Linear's weight in the state_dict-ordered parameter list (bias = pidx + 1)
};
enum LayerRole : uint8_t { ROLE_FIRST = 0, ROLE_TRUNK = 1, ROLE_LINEAR = 2, ROLE_DIR = 3, ROLE_RGB = 4, ROLE_WARP = 5 };

struct NetPlan {
  int32_t n_layers;
  int32_t n_ray_slots;
  int32_t n_ext_slots;    // number of RAY_POSE_EXT layers
  uint32_t stream_bytes;  // weight stream size (all hi+lo stages)
  uint32_t f32_ofs;       // byte offset of the fp32 section in the blob
  uint32_t total_bytes;
  uint32_t head_ofs;      // float offset: RenderRayNet: wrgb[3][128], brgb[3];  WarpNet: w2[3][256], b2[3]
  uint32_t sigma_ofs;     // float offset: wsigma[256], bsigma[1]
  int32_t in_freqs, in_identity;    // encoding of xyz that feeds the aux K-chunk
  int32_t dir_freqs, dir_identity;  // encoding of the view direction
  uint32_t flag_ofs;      // float offset (fp32 section) of one int32: set to 1 by the packer when a weight left the fp16 range (|w| > 65504)
  int32_t folded;         // 1: additional_linear_layer is folded into the sigma head and directional_input at pack time
  uint32_t fold_ofs;      // float offset (fp32 section) of the folded tensors: Wdir'[128][256 + D], bdir'[128], wsigma'[256], bsigma'[1]
  Layer layers[kMaxLayers];
};

// Host-side planners (nrf_pack.cu).  Return 0 or an NRF_E_* code and set the error string.
int plan_raynet(const NrfRayNetDesc* d, NetPlan* plan);
int plan_warpnet(const NrfWarpNetDesc* d, NetPlan* plan);

int launch_rows_differ(const float* feats, int64_t B, int A, int32_t* flag, cudaStream_t stream);   // nrf_ops.cu
void set_error(const char* fmt, ...);
// per-thread count of kernels launched by the training path (nrf_train_launch_count(): bench.py's gpu_launches accounting)
extern std::atomic<long long> g_train_launches;      // process-wide: the backward runs on autograd's worker thread
int cuda_fail(int err, const char* what);

// Number of aux features (<= 64) an encoder produces for a 3-vector, and the reference column of
// aux feature f (or -1 for padding).  Engine order: [sin,cos] pairs, COMPONENT-major (pair p = comp * L + k, so
// that a thread's 8 consecutive pairs are runs of consecutive frequencies of one component and can use the
// double-angle recurrence), then the identity components; reference order (utils.py:119-131): identity first,
// then for each frequency sin(x,y,z) followed by cos(x,y,z).
__host__ __device__ inline int enc_dim(int freqs, int identity) { return 3 * (2 * freqs + (identity ? 1 : 0)); }
__host__ __device__ inline int enc_ref_col(int f, int freqs, int identity) {
  if (f < 6 * freqs) {
    int p = f >> 1, s = f & 1, comp = p / freqs, k = p % freqs;
    return (identity ? 3 : 0) + k * 6 + s * 3 + comp;
  }
  int c = f - 6 * freqs;
  return (identity && c < 3) ? c : -1;
}

}  // namespace nrf
