// LayerNorm-over-channels kernels (NHWC tokens) and the patchify stem, forward and backward.
//
// Replaces (reference convnext_moe.py): LayerNorm2d.forward :34-47 in all three places it is used --
// the block norm after the depthwise conv (:351, channel_last), the downsample norms (:549-551) and
// the per-stage output norms (:811-817, channel_first incl. the NHWC->NCHW permute+contiguous) --
// and the 4x4/s4 stem Conv2d (:532-536 / :787-791).  F.layer_norm semantics: biased variance,
// y = (x-mean)*rsqrt(var+eps)*w + b.
#include "common.cuh"
#include "kernels.h"

namespace sm3 {

// ------------------------------------------------------------------------------------------------
// One warp per token.  lane holds channels lane, lane+32, ...  (C % 32 == 0, C <= 32*MAXV)
constexpr int LN_MAXV = 32;  // C <= 1024

enum LnOut : int { LN_OUT_NHWC = 0, LN_OUT_PATCH2 = 1, LN_OUT_NCHW = 2 };

template <int V>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                    const float* __restrict__ b, float* __restrict__ y,
                                                    float* __restrict__ stats, int T, int C, float eps,
                                                    int out_mode, int H, int W) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= T) return;
  const float* xr = x + (long long)warp * C;
  float v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) { v[i] = __ldg(xr + lane + 32 * i); s += v[i]; }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  if (stats && lane == 0) { stats[2 * (long long)warp] = mean; stats[2 * (long long)warp + 1] = rstd; }
  float* yr;
  if (out_mode == LN_OUT_PATCH2) {
    // token (n,h,w) -> row (n, h/2, w/2), column block (h%2)*2 + (w%2)
    const int wq = warp % W, hq = (warp / W) % H, n = warp / (W * H);
    const long long row = ((long long)n * (H / 2) + hq / 2) * (W / 2) + wq / 2;
    yr = y + row * (4LL * C) + (long long)((hq & 1) * 2 + (wq & 1)) * C;
  } else {
    yr = y + (long long)warp * C;
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = lane + 32 * i;
    yr[c] = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
  }
}

// LN + NHWC->NCHW: block = 32 consecutive tokens (same image row segment), 8 warps (4 tokens each).
template <int V>
__global__ void __launch_bounds__(256) ln_fwd_nchw_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, float* __restrict__ y,
                                                         float* __restrict__ stats, int T, int C, float eps, int HW) {
  extern __shared__ float tile[];  // [C][33]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = (long long)blockIdx.x * 32;
  for (int tt = warp; tt < 32; tt += 8) {
    const long long t = t0 + tt;
    if (t >= T) break;
    const float* xr = x + t * C;
    float v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { v[i] = __ldg(xr + lane + 32 * i); s += v[i]; }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    if (stats && lane == 0) { stats[2 * t] = mean; stats[2 * t + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = lane + 32 * i;
      tile[c * 33 + tt] = (v[i] - mean) * rstd * __ldg(w + c) + __ldg(b + c);
    }
  }
  __syncthreads();
  // HW % 32 == 0 is required (H, W multiples of 32 upstream => every level has HW % 32 == 0 unless
  // the level is smaller than 32 tokens; handled by the generic bound check below)
  for (int c = warp; c < C; c += 8) {
    const long long t = t0 + lane;
    if (t < T) {
      const long long n = t / HW, hw = t % HW;
      y[(n * C + c) * HW + hw] = tile[c * 33 + lane];
    }
  }
}

// LayerNorm whose output goes straight into the K-major bf16 hi|lo operand image of the following GEMM (the fused FFN's A
// operand): fp32 `v` is never written and the separate pack_act pass (read v, write image) disappears.  G lanes per token,
// each lane owns one 16-byte operand chunk = 8 consecutive channels (G = 16 for C <= 128, 32 for C <= 256).
template <int G>
__global__ void __launch_bounds__(256) ln_fwd_img_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, uint8_t* __restrict__ img,
                                                        float* __restrict__ y, float* __restrict__ stats, long long T,
                                                        long long T_pad, int C, float eps) {
  constexpr int TPW = 32 / G;                       // tokens per warp
  const int lane = threadIdx.x & 31, g = lane % G;
  const long long t = ((long long)(blockIdx.x * blockDim.x + threadIdx.x) >> 5) * TPW + lane / G;
  const int chunks = C / 8, kblocks = C / 32;
  const bool act = g < chunks, row_ok = t < T;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (act && row_ok) {
    const float4 p0 = ldg_f4(x + t * C + g * 8), p1 = ldg_f4(x + t * C + g * 8 + 4);
    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += v[e];
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
  if (act) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { const float d = v[e] - mean; q += d * d; }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
  if (!act || t >= T_pad) return;
  if (stats && g == 0 && row_ok) { stats[2 * t] = mean; stats[2 * t + 1] = rstd; }
  float o8[8];
  if (row_ok) {
    const float4 w0 = ldg_f4(w + g * 8), w1 = ldg_f4(w + g * 8 + 4), b0 = ldg_f4(b + g * 8), b1 = ldg_f4(b + g * 8 + 4);
    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) o8[e] = (v[e] - mean) * rstd * ww[e] + bb[e];
    if (y) {
      *reinterpret_cast<float4*>(y + t * C + g * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
      *reinterpret_cast<float4*>(y + t * C + g * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) o8[e] = 0.f;        // rows of the last 128-row tile beyond T: zero operand rows
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 8; e += 2) {
    const uint32_t u0 = __float_as_uint(o8[e]), u1 = __float_as_uint(o8[e + 1]);
    hi[e / 2] = __byte_perm(u0, u1, 0x7632);
    const uint32_t r0 = __float_as_uint(o8[e] - __uint_as_float(u0 & 0xFFFF0000u)) + 0x8000u;
    const uint32_t r1 = __float_as_uint(o8[e + 1] - __uint_as_float(u1 & 0xFFFF0000u)) + 0x8000u;
    lo[e / 2] = __byte_perm(r0, r1, 0x7632);
  }
  const long long rt = t >> 7;
  const uint32_t rr = (uint32_t)(t & 127), kb = (uint32_t)g >> 2, c = (uint32_t)g & 3u;
  uint8_t* dst = img + (rt * kblocks + kb) * 16384LL + ((rr >> 3) * 512u + (rr & 7u) * 64u + ((c ^ ((rr >> 1) & 3u)) << 4));
  *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(dst + 8192) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

int layernorm_fwd_img(const float* x, const float* w, const float* b, unsigned short* img, float* y, float* stats, long long T,
                      int C, float eps, cudaStream_t stream) {
  SM3_REQUIRE(x && w && b && img && T > 0, SM3_ERR_INVALID_ARG, "layernorm_fwd_img: null/empty argument");
  SM3_REQUIRE(C % 32 == 0 && C <= 256, SM3_ERR_UNSUPPORTED_SHAPE, "layernorm_fwd_img: C=%d must be a multiple of 32 <= 256", C);
  const long long T_pad = (T + 127) / 128 * 128;
  if (C <= 128) {
    const long long warps = (T_pad + 1) / 2;
    ln_fwd_img_kernel<16><<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(x, w, b, reinterpret_cast<uint8_t*>(img), y, stats, T, T_pad, C, eps);
  } else {
    ln_fwd_img_kernel<32><<<(unsigned)((T_pad + 7) / 8), 256, 0, stream>>>(x, w, b, reinterpret_cast<uint8_t*>(img), y, stats, T, T_pad, C, eps);
  }
  return check_launch("layernorm_fwd_img");
}

int layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* stats, long long T, int C,
                  float eps, int out_mode, int H, int W, cudaStream_t stream) {
  SM3_REQUIRE(x && w && b && y && T > 0, SM3_ERR_INVALID_ARG, "layernorm_fwd: null/empty argument");
  SM3_REQUIRE(C % 32 == 0 && C <= 32 * LN_MAXV, SM3_ERR_UNSUPPORTED_SHAPE, "layernorm_fwd: C=%d must be a multiple of 32 <= 1024", C);
  SM3_REQUIRE(T < (1LL << 31), SM3_ERR_UNSUPPORTED_SHAPE, "layernorm_fwd: too many tokens");
  const int V_ = C / 32;
  if (out_mode == LN_OUT_NCHW) {
    SM3_REQUIRE(H > 0 && W > 0, SM3_ERR_INVALID_ARG, "layernorm_fwd: NCHW output needs H, W");
    const int blocks = (int)((T + 31) / 32);
    const size_t smem = (size_t)C * 33 * sizeof(float);
    SM3_V_DISPATCH(V_, {
      cudaFuncSetAttribute(ln_fwd_nchw_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      ln_fwd_nchw_kernel<V><<<blocks, 256, smem, stream>>>(x, w, b, y, stats, (int)T, C, eps, H * W);
    });
  } else {
    if (out_mode == LN_OUT_PATCH2)
      SM3_REQUIRE(H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, SM3_ERR_INVALID_ARG, "layernorm_fwd: patch output needs even H, W");
    const int blocks = (int)((T + 7) / 8);
    SM3_V_DISPATCH(V_, (ln_fwd_kernel<V><<<blocks, 256, 0, stream>>>(x, w, b, y, stats, (int)T, C, eps, out_mode, H, W)));
  }
  return check_launch("layernorm_fwd");
}

// ------------------------------------------------------------------------------------------------
// Backward.  dy is read through the same three layouts the forward wrote (in_mode); dx is NHWC.
//   xhat = (x-mean)*rstd ; g = dy*w ; dx = rstd*(g - mean_c(g) - xhat*mean_c(g*xhat))
//   dw += sum_t dy*xhat ; db += sum_t dy     (per-block partials in registers -> atomics per block)
// If dx_accum != 0 the result is added to dx (residual branches meeting at one tensor).
// One token of the backward: d[] = dy for this lane's channels.  Accumulates the parameter partials, writes dx.
template <int V>
__device__ __forceinline__ void ln_bwd_token(const float (&d)[V], const float* __restrict__ xr, float mean, float rstd,
                                             const float (&wv)[V], float (&adw)[V], float (&adb)[V], float* __restrict__ dxr,
                                             int C, int lane, int dx_accum) {
  float g[V], xh[V];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    xh[i] = (__ldg(xr + lane + 32 * i) - mean) * rstd;
    adw[i] = fmaf(d[i], xh[i], adw[i]);
    adb[i] += d[i];
    g[i] = d[i] * wv[i];
    s1 += g[i];
    s2 = fmaf(g[i], xh[i], s2);
  }
  s1 = warp_sum(s1) / (float)C;
  s2 = warp_sum(s2) / (float)C;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = lane + 32 * i;
    const float r = rstd * (g[i] - s1 - xh[i] * s2);
    dxr[c] = dx_accum ? dxr[c] + r : r;
  }
}

// parameter partials of the 8 warps of a block -> shared memory -> one atomic per channel per block
template <int V>
__device__ __forceinline__ void ln_bwd_flush(const float (&adw)[V], const float (&adb)[V], float* red /*[2][8][C]*/,
                                             float* __restrict__ dw, float* __restrict__ db, int C, int warp, int lane) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
    red[warp * C + lane + 32 * i] = adw[i];
    red[(8 + warp) * C + lane + 32 * i] = adb[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    const int which = c / C, cc = c % C;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[(which * 8 + w) * C + cc];
    atomicAdd((which ? db : dw) + cc, t);
  }
}

template <int V>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                    const float* __restrict__ stats, const float* __restrict__ w,
                                                    float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                                    int T, int C, int in_mode, int H, int W, int dx_accum,
                                                    int tokens_per_warp) {
  extern __shared__ float red[];                   // [2][8][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * 8 + warp;
  float adw[V], adb[V], wv[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { adw[i] = 0.f; adb[i] = 0.f; wv[i] = __ldg(w + lane + 32 * i); }
  const long long tb = (long long)gwarp * tokens_per_warp;
  const int HW = H * W;
  for (int k = 0; k < tokens_per_warp; ++k) {
    const long long t = tb + k;
    if (t >= T) break;
    const float mean = __ldg(stats + 2 * t), rstd = __ldg(stats + 2 * t + 1);
    const float* dr;
    if (in_mode == LN_OUT_PATCH2) {
      const int wq = (int)(t % W), hq = (int)((t / W) % H); const long long n = t / HW;
      const long long row = (n * (H / 2) + hq / 2) * (W / 2) + wq / 2;
      dr = dy + row * (4LL * C) + (long long)((hq & 1) * 2 + (wq & 1)) * C;
    } else {
      dr = dy + t * C;
    }
    float d[V];
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = __ldg(dr + lane + 32 * i);
    ln_bwd_token<V>(d, x + t * C, mean, rstd, wv, adw, adb, dx + t * C, C, lane, dx_accum);
  }
  ln_bwd_flush<V>(adw, adb, red, dw, db, C, warp, lane);
}

// dy in NCHW (the output norms): a block stages 32 consecutive tokens x C channels of dy through shared memory with
// coalesced reads along hw (lane = token) and then works token-major (lane = channel) like the kernel above.
template <int V>
__global__ void __launch_bounds__(256) ln_bwd_nchw_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ stats, const float* __restrict__ w,
                                                         float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                                         int T, int C, int HW, int dx_accum, int chunks_per_block) {
  extern __shared__ float smem[];
  float* tile = smem;                              // [C][33]
  float* red = smem + C * 33;                      // [2][8][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float adw[V], adb[V], wv[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { adw[i] = 0.f; adb[i] = 0.f; wv[i] = __ldg(w + lane + 32 * i); }
  for (int q = 0; q < chunks_per_block; ++q) {
    const long long t0 = ((long long)blockIdx.x * chunks_per_block + q) * 32;
    if (t0 >= T) break;
    __syncthreads();                               // previous chunk's tile fully consumed
    {
      const long long t = t0 + lane;
      if (t < T) {
        const long long n = t / HW, hw = t % HW;
        for (int c = warp; c < C; c += 8) tile[c * 33 + lane] = __ldg(dy + (n * C + c) * HW + hw);
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int tt = warp * 4; tt < warp * 4 + 4; ++tt) {
      const long long t = t0 + tt;
      if (t >= T) break;
      const float mean = __ldg(stats + 2 * t), rstd = __ldg(stats + 2 * t + 1);
      float d[V];
#pragma unroll
      for (int i = 0; i < V; ++i) d[i] = tile[(lane + 32 * i) * 33 + tt];
      ln_bwd_token<V>(d, x + t * C, mean, rstd, wv, adw, adb, dx + t * C, C, lane, dx_accum);
    }
  }
  ln_bwd_flush<V>(adw, adb, red, dw, db, C, warp, lane);
}

int layernorm_bwd(const float* dy, const float* x, const float* stats, const float* w, float* dx, float* dw,
                  float* db, long long T, int C, int in_mode, int H, int W, int dx_accum, cudaStream_t stream) {
  SM3_REQUIRE(dy && x && stats && w && dx && dw && db && T > 0, SM3_ERR_INVALID_ARG, "layernorm_bwd: null/empty argument");
  SM3_REQUIRE(C % 32 == 0 && C <= 32 * LN_MAXV, SM3_ERR_UNSUPPORTED_SHAPE, "layernorm_bwd: C=%d", C);
  const int V_ = C / 32;
  // Latency-bound per token (two dependent load rounds + two warp reductions): keep >= 2 full waves of resident warps
  // busy (register use grows with V, so fewer warps fit for wide rows) and reduce the parameter partials per BLOCK.
  const int warps_per_sm = V_ <= 6 ? 64 : V_ <= 12 ? 32 : 16;
  const long long target = (long long)num_sms() * warps_per_sm * 2;
  if (in_mode == LN_OUT_NCHW) {
    SM3_REQUIRE(H > 0 && W > 0, SM3_ERR_INVALID_ARG, "layernorm_bwd: NCHW input needs H, W");
    const long long chunks = (T + 31) / 32;
    long long cpb = (chunks * 8 + target - 1) / target;          // a chunk keeps 8 warps busy
    if (cpb < 1) cpb = 1;
    const int blocks = (int)((chunks + cpb - 1) / cpb);
    const size_t smem = (size_t)C * (33 + 16) * sizeof(float);
    SM3_V_DISPATCH(V_, {
      cudaFuncSetAttribute(ln_bwd_nchw_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      ln_bwd_nchw_kernel<V><<<blocks, 256, smem, stream>>>(dy, x, stats, w, dx, dw, db, (int)T, C, H * W, dx_accum, (int)cpb);
    });
    return check_launch("layernorm_bwd_nchw");
  }
  int tpw = (int)((T + target - 1) / target);
  if (tpw < 2) tpw = 2;
  const long long warps = (T + tpw - 1) / tpw;
  const int blocks = (int)((warps + 7) / 8);
  const size_t smem = (size_t)16 * C * sizeof(float);
  SM3_V_DISPATCH(V_, {
    cudaFuncSetAttribute(ln_bwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ln_bwd_kernel<V><<<blocks, 256, smem, stream>>>(dy, x, stats, w, dx, dw, db, (int)T, C, in_mode, H, W, dx_accum, tpw);
  });
  return check_launch("layernorm_bwd");
}

// ------------------------------------------------------------------------------------------------
// Stem: y[n,ho,wo,:] = LN( conv4x4s4(x)[n,:,ho,wo] )   x NCHW [N,Cin,H,W] -> NHWC [N,H/ps,W/ps,C0]
// One warp computes 4 output pixels at a time; weights transposed in smem [K][C0], K = Cin*ps*ps.
template <int V>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wt /*[K][C0]*/,
                                                      const float* __restrict__ bias, const float* __restrict__ lnw,
                                                      const float* __restrict__ lnb, float* __restrict__ y,
                                                      float* __restrict__ conv_out, float* __restrict__ stats,
                                                      int N, int Cin, int H, int W, int ps, int C0, float eps) {
  extern __shared__ float smem[];
  const int K = Cin * ps * ps;
  float* s_w = smem;                 // [K][C0]
  float* s_p = smem + K * C0;        // [32 pixels][K+1]
  const int Ho = H / ps, Wo = W / ps;
  const long long P = (long long)N * Ho * Wo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < K * C0; i += blockDim.x) s_w[i] = __ldg(wt + i);
  const long long p0 = (long long)blockIdx.x * 32;
  // gather 32 patches: element (pix, k=(c,i,j)) = x[n, c, ho*ps+i, wo*ps+j]
  for (int idx = threadIdx.x; idx < 32 * K; idx += blockDim.x) {
    const int k = idx / 32, pix = idx % 32;   // consecutive threads -> consecutive pixels (strided by ps in w)
    const long long p = p0 + pix;
    float val = 0.f;
    if (p < P) {
      const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho); const long long n = p / ((long long)Wo * Ho);
      const int j = k % ps, i = (k / ps) % ps, c = k / (ps * ps);
      val = __ldg(x + ((n * Cin + c) * H + ho * ps + i) * W + wo * ps + j);
    }
    s_p[pix * (K + 1) + k] = val;
  }
  __syncthreads();
  // warp handles pixels warp*4 .. warp*4+3
  float acc[4][V];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[q][i] = __ldg(bias + lane + 32 * i);
  const float* pp = s_p + (warp * 4) * (K + 1);
  for (int k = 0; k < K; ++k) {
    float wv[V];
#pragma unroll
    for (int i = 0; i < V; ++i) wv[i] = s_w[k * C0 + lane + 32 * i];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float pv = pp[q * (K + 1) + k];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[q][i] = fmaf(pv, wv[i], acc[q][i]);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long long p = p0 + warp * 4 + q;
    if (p >= P) break;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += acc[q][i];
    const float mean = warp_sum(s) / (float)C0;
    float qq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { const float d = acc[q][i] - mean; qq += d * d; }
    const float rstd = rsqrtf(warp_sum(qq) / (float)C0 + eps);
    if (stats && lane == 0) { stats[2 * p] = mean; stats[2 * p + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = lane + 32 * i;
      if (conv_out) conv_out[p * C0 + c] = acc[q][i];
      y[p * C0 + c] = (acc[q][i] - mean) * rstd * __ldg(lnw + c) + __ldg(lnb + c);
    }
  }
}

int stem_fwd(const float* x, const float* wt, const float* bias, const float* lnw, const float* lnb, float* y,
             float* conv_out, float* stats, int N, int Cin, int H, int W, int ps, int C0, float eps,
             cudaStream_t stream) {
  SM3_REQUIRE(x && wt && bias && lnw && lnb && y, SM3_ERR_INVALID_ARG, "stem_fwd: null argument");
  SM3_REQUIRE(H % ps == 0 && W % ps == 0, SM3_ERR_UNSUPPORTED_SHAPE, "stem_fwd: H,W must be multiples of the patch size");
  SM3_REQUIRE(C0 % 32 == 0 && C0 <= 512, SM3_ERR_UNSUPPORTED_SHAPE, "stem_fwd: C0=%d must be a multiple of 32 <= 512", C0);
  const int K = Cin * ps * ps;
  const size_t smem = ((size_t)K * C0 + 32 * (K + 1)) * sizeof(float);
  SM3_REQUIRE(smem <= 200 * 1024, SM3_ERR_UNSUPPORTED_SHAPE, "stem_fwd: patch weights do not fit shared memory");
  const long long P = (long long)N * (H / ps) * (W / ps);
  const int blocks = (int)((P + 31) / 32);
  const int V_ = C0 / 32;
  SM3_V_DISPATCH(V_, {
    cudaFuncSetAttribute(stem_fwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stem_fwd_kernel<V><<<blocks, 256, smem, stream>>>(x, wt, bias, lnw, lnb, y, conv_out, stats, N, Cin, H, W, ps, C0, eps);
  });
  return check_launch("stem_fwd");
}

// Stem weight gradient: dWt[k][c] += sum_p patch[p][k] * du[p][c] ; dbias[c] += sum_p du[p][c]
// (du = gradient w.r.t. the conv output, NHWC).  Block = 256 pixels chunk, thread (k-slice, c).
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ du,
                                                        float* __restrict__ dwt, float* __restrict__ dbias, int N,
                                                        int Cin, int H, int W, int ps, int C0, int pix_per_block) {
  extern __shared__ float smem[];
  const int K = Cin * ps * ps;
  float* s_p = smem;             // [32][K+1]
  float* s_d = smem + 32 * (K + 1);  // [32][C0]
  const int Ho = H / ps, Wo = W / ps;
  const long long P = (long long)N * Ho * Wo;
  const long long pb = (long long)blockIdx.x * pix_per_block;
  // each thread owns outputs (k, c) for idx = tid, tid+256, ... < K*C0 (+ bias row k = K)
  const int total = (K + 1) * C0;
  constexpr int MAXO = 32;
  float acc[MAXO];
#pragma unroll
  for (int o = 0; o < MAXO; ++o) acc[o] = 0.f;
  for (long long p0 = pb; p0 < pb + pix_per_block && p0 < P; p0 += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * K; idx += blockDim.x) {
      const int k = idx / 32, pix = idx % 32;
      const long long p = p0 + pix;
      float val = 0.f;
      if (p < P && p < pb + pix_per_block) {
        const int wo = (int)(p % Wo), ho = (int)((p / Wo) % Ho); const long long n = p / ((long long)Wo * Ho);
        const int j = k % ps, i = (k / ps) % ps, c = k / (ps * ps);
        val = __ldg(x + ((n * Cin + c) * H + ho * ps + i) * W + wo * ps + j);
      }
      s_p[pix * (K + 1) + k] = val;
    }
    for (int idx = threadIdx.x; idx < 32 * C0; idx += blockDim.x) {
      const int pix = idx / C0, c = idx % C0;
      const long long p = p0 + pix;
      s_d[idx] = (p < P && p < pb + pix_per_block) ? __ldg(du + p * C0 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      const int idx = threadIdx.x + o * 256;
      if (idx >= total) break;
      const int k = idx / C0, c = idx % C0;
      float a = 0.f;
      if (k < K) {
#pragma unroll 8
        for (int pix = 0; pix < 32; ++pix) a = fmaf(s_p[pix * (K + 1) + k], s_d[pix * C0 + c], a);
      } else {
#pragma unroll 8
        for (int pix = 0; pix < 32; ++pix) a += s_d[pix * C0 + c];
      }
      acc[o] += a;
    }
  }
#pragma unroll
  for (int o = 0; o < MAXO; ++o) {
    const int idx = threadIdx.x + o * 256;
    if (idx >= total) break;
    const int k = idx / C0, c = idx % C0;
    if (k < K) atomicAdd(dwt + k * C0 + c, acc[o]);
    else atomicAdd(dbias + c, acc[o]);
  }
}

int stem_wgrad(const float* x, const float* du, float* dwt, float* dbias, int N, int Cin, int H, int W, int ps,
               int C0, cudaStream_t stream) {
  SM3_REQUIRE(x && du && dwt && dbias, SM3_ERR_INVALID_ARG, "stem_wgrad: null argument");
  const int K = Cin * ps * ps;
  SM3_REQUIRE((K + 1) * C0 <= 32 * 256, SM3_ERR_UNSUPPORTED_SHAPE, "stem_wgrad: (K+1)*C0=%d exceeds 8192", (K + 1) * C0);
  const long long P = (long long)N * (H / ps) * (W / ps);
  int blocks = num_sms() * 2;
  long long ppb = (P + blocks - 1) / blocks;
  ppb = (ppb + 31) / 32 * 32;
  blocks = (int)((P + ppb - 1) / ppb);
  const size_t smem = ((size_t)32 * (K + 1) + 32 * C0) * sizeof(float);
  stem_wgrad_kernel<<<blocks, 256, smem, stream>>>(x, du, dwt, dbias, N, Cin, H, W, ps, C0, (int)ppb);
  return check_launch("stem_wgrad");
}

}  // namespace sm3
