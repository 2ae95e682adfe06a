// tcgen05 GEMM building blocks of the TRAINING path (the backward of
// solver/nerf_solver.py:81-87 `loss.backward()` through models/render_ray_net.py:42-61 and
// models/warp_field_net.py:17-21, plus the layer-by-layer forward that saves what the backward needs).
//
// Every matrix lives in HBM as two fp16 "planes" (hi, lo: x ~= hi + lo, ~22 significant bits) in plain row-major
// [rows, features] form.  Because tcgen05.mma takes either operand K-major or MN-major, the SAME plane tensors feed all
// three products of a linear layer without any transposition pass:
//
//   forward   Y[S, out]  = X[S, in]   . W[out, in]^T     A = X  K-major,  B = W  K-major     (tile_gemm, b_mn = 0)
//   dX        dX[S, in]  = dY[S, out] . W[out, in]       A = dY K-major,  B = W  N-major     (tile_gemm, b_mn = 1)
//   dW        dW[out,in] = dY[S, out]^T . X[S, in]       A = dY M-major,  B = X  N-major     (dw_gemm, K = samples)
//
// Operands are staged by 2-D tensor TMA (SWIZZLE_128B boxes of 64 features) straight into the UMMA canonical layouts.
// passes = 3 runs hi*hi + lo*hi + hi*lo (fp32-equivalent, like the renderer's parity mode); passes = 1 runs hi*hi;
// passes = 6 (tile_gemm only, bf16 planes hi/lo/ll = 24 significant bits with fp32's exponent range) adds ll*hi + lo*lo + hi*ll:
// the operands are then EXACT fp32 values and no activation magnitude can leave the representable range.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nrf {

enum GemmEpi : int32_t {
  GEPI_PLANES = 0,   // out = [relu](acc + bias [+ row_scale * col_vec]) [masked] -> fp16 hi/lo planes (+ optional fp32 copy)
  GEPI_F32 = 1,      // out_f32 (+)= acc
};

struct TileGemmParams {
  CUtensorMap a_map[2][3];        // A sources x planes (hi, lo, ll): [S, 64 * kc[j]], box {64, 128}
  CUtensorMap b_map[2][3];        // B sources x planes: b_mn = 0: [N, 64 * kc[j]] box {64, n_tile}; b_mn = 1: [64 * kc[j], N] box {64, 64}
  int32_t kc[2];
  int32_t n_src, b_mn, n_tile, passes, n_stages;
  int32_t bf16;                   // 1: the planes hold bfloat16 (exact mode: 3 planes = 24 significant bits, fp32 range), else fp16
  CUtensorMap o_map[2];           // staged epilogue: output planes hi / lo as [S, N], box {32, 32} (one warp's share of a unit, SWIZZLE_64B), TMA stores
  int32_t staged;                 // 1: the epilogue stages the fp16 output planes through shared memory (coalesced TMA stores)
  CUtensorMap f_map;              // staged fp32 copy (out_f32 beside the planes): [S, N] floats seen as [S, 2N] 16-bit elements, box {64, 32}
  int32_t f32_staged;             // 1: the fp32 copy of a unit also leaves through the staging buffer (one [32 x 32] float box per warp)
  int32_t b_stream;               // 1: K too large for a resident weight slice -- the B chunk travels with every A chunk through the ring
  int64_t S;
  int32_t epi, relu;
  const float* bias; int32_t bias_ld; int32_t rows_per_ray;      // bias[(row / rows_per_ray) * bias_ld + col]  (bias_ld = 0: one vector)
  __half* out_hi; __half* out_lo; __half* out_ll; int32_t out_ld;
  float* out_f32; int32_t out_f32_ld; int32_t accumulate;
  const uint32_t* mask_bits; uint32_t* bits_out; int32_t bits_ld;  // ReLU' bit masks, bits_ld words per row: keep out[row, col] only where bit col of
                                                                   // mask_bits[row] is set (dX); bits_out[row] receives bit col <=> hi(out[row, col]) != 0 (forward)
  const float* row_scale; int32_t row_scale_ld; const float* col_vec;   // + row_scale[row * ld] * sc_out[0] * col_vec[col]
  // gradient scaling (all NULL in the forward): the A planes hold real * sc_in[0]; the output planes are written as
  // real * sc_out[0] (sc_x = {s, 1/s}, powers of two); GEPI_F32 writes REAL values.  l1max (float bits, atomicMax) receives
  // parts * max over (row, column segment) of the segment's L1 norm of the REAL output: an upper bound of the largest row L1
  // norm, from which the host derives the next layer's scale (|dX| <= |dY row|_1 max|W|).
  const float* sc_in; const float* sc_out; unsigned int* l1max;
  int32_t* status;                                                 // bit 1: an output saturated the fp16 range
  int32_t reverse;                                                 // 1: walk the sample tiles back to front (L2 reuse across consecutive kernels)
  int32_t pf_dist;                                                 // L2 prefetch distance in tiles (developer A/B: NRF_GEMM_PF; default 1)
  long long* trace;                                                // developer tap (NRF_GEMM_TRACE): CTA 0's epilogue timeline, SM clocks
};

struct DwGemmParams {
  CUtensorMap a_hi, a_lo;         // dY planes [S, Fa], box {64, 64}
  CUtensorMap b_hi, b_lo;         // X planes [S, Fb],  box {64, 64}
  int32_t m0, n0, N, passes, n_stages;
  int64_t S;
  float* partial;                 // [n_split][M_total][N]
  int32_t M_total;
  int32_t reverse;                // 1: walk the 64-sample chunks back to front (see TileGemmParams::reverse)
  float* colsum_partial;          // optional [n_split][M_total]: sum over this CTA's samples of (hi + lo)[s, m] -- the bias gradient --
                                  // from two extra N = 16 MMAs per K-step against a constant tile of ones (no extra pass over dY)
};

// host helpers (nrf_gemm.cu).  All return NRF_OK or an error code with the message set.
int encode_planes_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_cols, uint32_t box_rows,
                      int swizzle_bytes = 128);

// cols: logical feature count (multiple of 64).  ll: third plane of the exact mode (bf16 x 3), NULL otherwise; the element type of the
// storage is 16-bit either way (fp16 or, in the exact mode, bfloat16 bit patterns).
struct Planes { __half* hi; __half* lo; int64_t rows; int32_t cols; int32_t ld; __half* ll; };

struct TileGemmArgs {
  Planes a[2]; Planes b[2]; int n_src; int b_mn; int N;      // N: output columns (multiple of 64)
  int passes;                  // 1, 3 (fp16 planes) or 6 (bf16 x 3 planes)
  int epi, relu;
  const float* bias; int bias_ld; int rows_per_ray;
  Planes out; float* out_f32; int out_f32_ld; int accumulate;
  const uint32_t* mask_bits; uint32_t* bits_out; int bits_ld;
  const float* row_scale; int row_scale_ld; const float* col_vec;
  const float* sc_in; const float* sc_out; unsigned int* l1max;
  int32_t* status;
  int reverse;
};
int launch_tile_gemm(const TileGemmArgs& a, int n_sms, cudaStream_t stream);

// dW[M, N] = sum_s A[s, m0 + m] * B[s, n0 + n]: partial sums per CTA into `partial` ([n_split][M][N] floats, n_split returned)
int launch_dw_gemm(const Planes& a, int m0, int M, const Planes& b, int n0, int N, int passes, float* partial, int max_split,
                   int* n_split_out, int n_sms, cudaStream_t stream, float* colsum_partial = nullptr, int reverse = 0);
int dw_gemm_max_split(int n_sms, int M);

}  // namespace nrf
