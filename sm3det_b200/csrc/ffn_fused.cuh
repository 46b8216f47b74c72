// Fused dense-FFN kernels for the narrow ConvNeXt stages (C <= 256): the [T, 4C] hidden tensor never leaves the SM.
//
// Reference math (mmrotate/models/backbones/convnext_moe.py): FFN.forward :397-405 inside
// ConvNeXtBlock._inner_forward :343-372 --   out = x + drop_path(gamma * (W2 gelu(W1 v + b1) + b2))   -- and what autograd
// derives for it.  Round 1 ran this as GEMM1 -> act_pack -> GEMM2 with the hidden
// tensor crossing HBM 4x in the forward and ~10x in the backward (805 MB per crossing at stage 0, bs 8); at C = 96 / 192
// those GEMMs are HBM streams (K or N <= 192), not tensor-bound.  Here one persistent CTA per SM walks 128-token tiles:
//
//   ffn_chain_kernel<MODE, HC>      per tile, per hidden chunk j of HC columns:
//        GEMM-a   acc_h[128 x HC]  = A1 . Wa1_j^T                      (tcgen05, 3-pass split-bf16, accumulator in TMEM)
//        (MODE 1) acc_d[128 x HC]  = A2 . Wa2_j^T
//        middle   MODE 0: y = gelu(acc_h + b1_j)         MODE 1: y = acc_d * gelu'(acc_h + b1_j)
//                 -> split hi/lo -> shared memory, directly in the K-major SWIZZLE_64B operand layout
//        GEMM-b   acc_o[128 x C]  += y . Wb_j^T
//      final epilogue: MODE 0: out = resid + row_scale * gamma * (acc_o + b2), aux = acc_o + b2;   MODE 1: out = acc_o
//      MODE 0 = forward (A1 = v, Wa1 = W1, Wb = W2);  MODE 1 = backward into dv (A1 = v: the hidden pre-activation is
//      RECOMPUTED, A2 = dz, Wa1 = W1, Wa2 = (gamma W2)^T, Wb = W1^T) -- the tensor pipe has >2x slack on these shapes,
//      HBM does not, so nothing hidden-sized is saved by the forward at all.
//
//   ffn_wgrad_kernel<HC>            weight gradients, hidden slice [hb, hb+HW) per CTA, token tiles streamed:
//        recompute acc_h / acc_d as above;  middle: dh = acc_d * gelu'(.), a = gelu(.) -> two smem operand blocks;
//        accW1[c, hid] += v^T dh,  accW2[c, hid] += dz^T a   (the SAME smem images serve as K-major and MN-major operands:
//        a [128 x 32] SWIZZLE_64B block is both), accumulated in TMEM over all the CTA's token tiles, flushed once with
//        fp32 reductions:  dW1[hid, c] += accW1[c, hid],  dW2[c, hid] += gamma_c accW2[c, hid],  db1[hid] += accW1[C, hid]
//        (the bias gradient rides on the tensor cores: a constant "ones" channel appended to v, when C < 128).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sm3 {
namespace ffn {

struct ChainParams {
  // operands (all pre-split bf16 hi|lo tile images, see gemm_tc.cuh pack_act / pack_b)
  const uint16_t* a1;     // K-major image of A1 [M, C], 128-row tiles
  const uint16_t* a2;     // MODE 1: K-major image of A2 [M, C]
  const uint16_t* wa1;    // image of Wa1 [H4, C] packed with tile width HC
  const uint16_t* wa2;    // MODE 1: image of Wa2 [H4, C], tile width HC
  const uint16_t* wb;     // image of Wb [C, H4] packed with tile width C
  const float* bias1;     // [H4]
  const float* bias2;     // [C] or null
  const float* col_scale; // [C] or null (gamma)
  const float* row_scale; // [M] or null (drop-path)
  const float* resid;     // [M, C] or null (shortcut)
  float* out;             // [M, C]
  float* aux_out;         // [M, C] or null: acc_o + bias2 before scaling (y2, needed for dgamma)
  float* h_out;           // MODE 0: [M, H4] or null: hidden pre-activation A1 Wa1^T + b1 (saved for a GEMM-based backward)
  int M, C, H4, HC, passes, mode;
  int debug;              // perf experiments only: bit0 = middle stage skips the GELU math, bit1 = skip the operand split/stores
};
int chain(const ChainParams& p, cudaStream_t stream);

struct WgradParams {
  const uint16_t* a1;     // K-major image of v  [M, C]
  const uint16_t* a2;     // K-major image of dz [M, C]
  const uint16_t* wa1;    // image of W1 [H4, C], tile width HC
  const uint16_t* wa2;    // image of (gamma W2)^T [H4, C], tile width HC
  const float* bias1;     // [H4]
  const float* gamma;     // [C]: dW2 rows are scaled by it (dz already carries drop-path)
  float* dw1;             // [H4, C]  accumulated (pre-zeroed)
  float* dw2;             // [C, H4]  accumulated
  float* db1;             // [H4]     accumulated
  int M, C, H4, HC, passes;
  int debug;
};
int wgrad(const WgradParams& p, cudaStream_t stream);

// largest hidden chunk the shared-memory budget allows for (mode, C); 0 if the shape is unsupported
int chain_chunk(int mode, int C);
int wgrad_chunk(int C);

}  // namespace ffn
}  // namespace sm3
