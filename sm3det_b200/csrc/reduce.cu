// Small HBM-bound helpers of the backward pass: column reductions (bias / layer-scale gradients),
// the token-major gather-sum that mirrors the dispatch gather (autograd's index-backward of
// x[_batch_index], reference convnext_moe.py:265, done without atomics), and a row/col scaling pass.
#include "common.cuh"
#include "kernels.h"

namespace sm3 {

// out[g][c] += sum_{r in [seg_begin[g], seg_end[g])} a[r,c] * (b ? b[r,c] : 1) * (rs ? rs[r] : 1)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                    const float* __restrict__ rs, const int* __restrict__ seg_begin,
                                                    const int* __restrict__ seg_end, float* __restrict__ out,
                                                    long long rows, int C, int chunks) {
  const int g = blockIdx.z;
  long long r0 = 0, r1 = rows;
  if (seg_begin) { r0 = __ldg(seg_begin + g); r1 = __ldg(seg_end + g); }
  const long long len = r1 - r0;
  if (len <= 0) return;
  const long long per = (len + chunks - 1) / chunks;
  const long long cb = r0 + (long long)blockIdx.x * per;
  const long long ce = min(r1, cb + per);
  const int c = (blockIdx.y * 64 + (threadIdx.x & 63));
  const int rl = threadIdx.x >> 6;  // 4 row lanes
  const bool col_ok = c < C;
  float acc = 0.f;
  for (long long r = cb + rl; col_ok && r < ce; r += 4) {
    float v = __ldg(a + r * C + c);
    if (b) v *= __ldg(b + r * C + c);
    if (rs) v *= __ldg(rs + r);
    acc += v;
  }
  __shared__ float s[4][64];
  s[rl][threadIdx.x & 63] = acc;
  __syncthreads();
  if (rl == 0) {
    const float t = s[0][threadIdx.x] + s[1][threadIdx.x] + s[2][threadIdx.x] + s[3][threadIdx.x];
    if (cb < ce && col_ok) atomicAdd(out + (long long)g * C + c, t);
  }
}

int colsum(const float* a, const float* b, const float* rs, const int* seg_begin, const int* seg_end, int G,
           float* out, long long rows, int C, cudaStream_t stream) {
  SM3_REQUIRE(a && out && G >= 1 && C >= 1, SM3_ERR_INVALID_ARG, "colsum: bad argument");
  SM3_REQUIRE((seg_begin == nullptr) == (seg_end == nullptr), SM3_ERR_INVALID_ARG, "colsum: seg_begin/seg_end");
  const int gy = (C + 63) / 64;
  long long chunks = (long long)num_sms() * 8 / ((long long)gy * G);
  if (chunks < 1) chunks = 1;
  if (chunks > (rows + 63) / 64) chunks = (rows + 63) / 64;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)gy, (unsigned)G);
  colsum_kernel<<<grid, 256, 0, stream>>>(a, b, rs, seg_begin, seg_end, out, rows, C, (int)chunks);
  return check_launch("colsum");
}

// out[t,:] = (add ? add[t,:] : 0) + sum_j src[slot_of[t,j], :]   (slot < 0 skipped)
__global__ void __launch_bounds__(256) gather_sum_kernel(const float* __restrict__ src, const int* __restrict__ slot_of,
                                                        const float* __restrict__ add, float* __restrict__ out,
                                                        long long total, int C, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Q = C >> 2;
  const long long t = i / Q;
  const int c = (int)(i % Q) * 4;
  float4 y = add ? ldg_f4(add + t * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = 0; j < k; ++j) {
    const int s = __ldg(slot_of + t * k + j);
    if (s < 0) continue;
    const float4 v = ldg_f4(src + (long long)s * C + c);
    y.x += v.x; y.y += v.y; y.z += v.z; y.w += v.w;
  }
  *reinterpret_cast<float4*>(out + t * C + c) = y;
}

int gather_sum(const float* src, const int* slot_of, const float* add, float* out, int T, int C, int k,
               cudaStream_t stream) {
  SM3_REQUIRE(src && slot_of && out && C % 4 == 0, SM3_ERR_INVALID_ARG, "gather_sum: bad argument");
  const long long total = (long long)T * (C / 4);
  gather_sum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, slot_of, add, out, total, C, k);
  return check_launch("gather_sum");
}

// out[r,c] = x[r,c] * (rs ? rs[r] : 1) * (cs ? cs[c] : 1)
__global__ void __launch_bounds__(256) scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ rs,
                                                        const float* __restrict__ cs, float* __restrict__ out,
                                                        long long total, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Q = C >> 2;
  const long long r = i / Q;
  const int c = (int)(i % Q) * 4;
  float4 v = ldg_f4(x + r * C + c);
  const float s = rs ? __ldg(rs + r) : 1.0f;
  float4 m = cs ? ldg_f4(cs + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  v.x *= s * m.x; v.y *= s * m.y; v.z *= s * m.z; v.w *= s * m.w;
  *reinterpret_cast<float4*>(out + r * C + c) = v;
}

int scale_rows(const float* x, const float* rs, const float* cs, float* out, long long rows, int C,
               cudaStream_t stream) {
  SM3_REQUIRE(x && out && C % 4 == 0, SM3_ERR_INVALID_ARG, "scale_rows: bad argument");
  const long long total = rows * (C / 4);
  scale_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, rs, cs, out, total, C);
  return check_launch("scale_rows");
}

// Peer row gather over NVLink (expert-parallel dispatch / combine): out[r,:] = scale[r] * base[src_rank[r]][row*C ..]
// where row = src_row[r], or token_list[src_rank[r]][src_row[r]] when token lists are given (the dispatch reads the source
// rank's expert-sorted pair list and then the token row, both through peer pointers).  src_rank < 0 -> zero row.
// bases / token_lists are device arrays of `world` device pointers into symmetric (P2P-mapped) buffers.
__global__ void __launch_bounds__(256) gather_rows_peer_kernel(const float* const* __restrict__ bases,
                                                              const int* const* __restrict__ token_lists,
                                                              const int* __restrict__ src_rank, const int* __restrict__ src_row,
                                                              const float* __restrict__ scale, float* __restrict__ out,
                                                              long long total, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Q = C >> 2;
  const long long r = i / Q;
  const int c = (int)(i % Q) * 4;
  const int s = __ldg(src_rank + r);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s >= 0) {
    long long row = __ldg(src_row + r);
    if (token_lists) row = token_lists[s][row];
    if (row >= 0) {
      v = *reinterpret_cast<const float4*>(bases[s] + row * C + c);      // plain (coherent) load: peer memory
      if (scale) { const float g = __ldg(scale + r); v.x *= g; v.y *= g; v.z *= g; v.w *= g; }
    }
  }
  *reinterpret_cast<float4*>(out + r * C + c) = v;
}

int gather_rows_peer(const float* const* bases, const int* const* token_lists, const int* src_rank, const int* src_row,
                     const float* scale, float* out, long long rows, int C, cudaStream_t stream) {
  SM3_REQUIRE(bases && src_rank && src_row && out && C % 4 == 0 && rows >= 0, SM3_ERR_INVALID_ARG, "gather_rows_peer: bad argument");
  if (rows == 0) return SM3_OK;
  const long long total = rows * (C / 4);
  gather_rows_peer_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(bases, token_lists, src_rank, src_row, scale, out, total, C);
  return check_launch("gather_rows_peer");
}

// ------------------------------------------------------------------------------------------------
// Expert-parallel exchange plan, built ON THE DEVICE from the all-gathered per-rank (count, segment start) tables -- no host
// round trip per layer (round 1 did `.cpu()` + a Python loop here).  One block.  For the calling rank `me` of `W`:
//   expert side: row r of my padded expert-major space <- (source rank, slot in that rank's expert-sorted pair list);
//                experts are laid out one after another, each padded to 128 rows; inside an expert, sources in rank order;
//   source side: my slot l (expert g) lives on rank g / E_loc at row  l - seg_s[g] + seg_owner[g] + off[g][me].
// Rows beyond `cap` are dropped and *overflow = rows needed (the host checks it off the critical path).
__global__ void __launch_bounds__(1024) ep_plan_kernel(const EpPlanArgs a) {
  __shared__ int s_cnt[8][32], s_seg[8][32];   // [rank][expert]
  __shared__ int s_start[32];                   // first row of global expert g on its owner
  __shared__ int s_off[32][8];                  // offset of source s inside expert g's segment
  const int W = a.W, E = a.E, El = E / W, me = a.me, tid = threadIdx.x;
  for (int i = tid; i < W * E; i += blockDim.x) {
    const int r = i / E, e = i % E;
    s_cnt[r][e] = a.allm[(r * 2 + 0) * E + e];
    s_seg[r][e] = a.allm[(r * 2 + 1) * E + e];
  }
  __syncthreads();
  if (tid == 0) {
    int my_rows = 0, my_tiles = 0;
    for (int d = 0; d < W; ++d) {
      int pos = 0;
      for (int el = 0; el < El; ++el) {
        const int g = d * El + el;
        s_start[g] = pos;
        int acc = 0;
        for (int s = 0; s < W; ++s) { s_off[g][s] = acc; acc += s_cnt[s][g]; }
        const int nt = (acc + 127) / 128;
        if (d == me) {
          a.seg_begin[el] = pos; a.seg_end[el] = pos + acc;
          for (int t = 0; t < nt; ++t) if (my_tiles + t < a.cap / 128) a.tile_group[my_tiles + t] = el;
          my_tiles += nt;
        }
        pos += nt * 128;
      }
      if (d == me) my_rows = pos;
    }
    *a.num_tiles = min(my_tiles, a.cap / 128);
    if (my_rows > a.cap) atomicMax(a.overflow, my_rows);
  }
  for (int r = tid; r < a.cap; r += blockDim.x) { a.src_rank[r] = -1; a.src_slot[r] = 0; }
  __syncthreads();
  for (int el = 0; el < El; ++el) {
    const int g = me * El + el;
    for (int s = 0; s < W; ++s) {
      const int L = s_cnt[s][g], D = s_start[g] + s_off[g][s], S = s_seg[s][g];
      for (int i = tid; i < L; i += blockDim.x)
        if (D + i < a.cap) { a.src_rank[D + i] = s; a.src_slot[D + i] = S + i; }
    }
  }
  const int live_rows = __ldg(a.num_tiles_s) * 128;
  for (int l = tid; l < a.R_s; l += blockDim.x) {
    int g = __ldg(a.tile_group_s + (l >> 7));
    g = g < 0 ? 0 : (g >= E ? E - 1 : g);
    const bool live = l < live_rows && __ldg(a.pair_token + l) >= 0;
    const int row = l - s_seg[me][g] + s_start[g] + s_off[g][me];
    a.comb_rank[l] = live ? g / El : -1;
    a.comb_row[l] = live ? row : 0;
  }
}

int ep_plan(const EpPlanArgs& a, cudaStream_t stream) {
  SM3_REQUIRE(a.allm && a.tile_group_s && a.num_tiles_s && a.pair_token && a.src_rank && a.src_slot && a.tile_group &&
              a.num_tiles && a.seg_begin && a.seg_end && a.comb_rank && a.comb_row && a.overflow, SM3_ERR_INVALID_ARG, "ep_plan: null argument");
  SM3_REQUIRE(a.W >= 1 && a.W <= 8 && a.E >= a.W && a.E <= 32 && a.E % a.W == 0 && a.me >= 0 && a.me < a.W && a.cap % 128 == 0 && a.cap > 0,
              SM3_ERR_UNSUPPORTED_SHAPE, "ep_plan: W=%d E=%d cap=%d", a.W, a.E, a.cap);
  ep_plan_kernel<<<1, 1024, 0, stream>>>(a);
  return check_launch("ep_plan_kernel");
}

}  // namespace sm3
