// Fused dense-FFN kernels (see ffn_fused.cuh for the what and why).  sm_100a only: tcgen05.mma with TMEM accumulators,
// cp.async.bulk (TMA bulk copies) completing on mbarriers, one persistent CTA per SM.
#define SM3_GEMM_KERNEL_IMPL
#include "gemm_tc.cuh"
#include "ffn_fused.cuh"
#include "kernels.h"

namespace sm3 {
namespace ffn {
using namespace gemm;

constexpr int THREADS = 384;                 // 12 warps: 0 A/Wa producer, 1 Wb producer, 2 MMA issuer, 3 idle, 4-11 middle/epilogue
constexpr int W_PROD_A = 0, W_PROD_B = 1, W_MMA = 2, EPI0 = 4, NE = 8;
constexpr uint32_t SMEM_LIMIT = 232448u - 1024u;      // 227 KB opt-in maximum minus the 1 KB alignment slack
constexpr uint32_t STAGING = NE * EPI_STAGE_BYTES;    // per-warp transpose staging of the final epilogue
constexpr uint32_t BARS = 256;

// shared-memory carve-up, computed on the host and passed by value
struct Layout {
  uint32_t a, ones, a2, wa, wb, act, act2, stage, bar, total;
  uint32_t a_tile, wa_chunk, wa_stage, wb_stage, act_bytes;
  int sa, sb;
};

struct ChainK { ChainParams p; Layout L; int m_tiles, nch; };

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// split 8 fp32 (one 16-byte operand chunk) into the hi / lo bf16 planes
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 8; e += 2) {
    const uint32_t u0 = __float_as_uint(v[e]), u1 = __float_as_uint(v[e + 1]);
    h[e / 2] = __byte_perm(u0, u1, 0x7632);
    const uint32_t r0 = __float_as_uint(v[e] - __uint_as_float(u0 & 0xFFFF0000u)) + 0x8000u;
    const uint32_t r1 = __float_as_uint(v[e + 1] - __uint_as_float(u1 & 0xFFFF0000u)) + 0x8000u;
    l[e / 2] = __byte_perm(r0, r1, 0x7632);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// 3-pass (or 1-pass) split-bf16 product of one 16-k step:  D (+)= (Ahi + Alo)(Bhi + Blo) minus the lo*lo term
__device__ __forceinline__ void mma3(uint32_t d, uint64_t ahi, uint64_t alo, uint64_t bhi, uint64_t blo, uint32_t idesc,
                                     uint32_t accum, int passes) {
  if (passes == 1) {
    tc_mma(d, ahi, bhi, idesc, accum);
  } else {
    tc_mma(d, alo, bhi, idesc, accum);
    tc_mma(d, ahi, blo, idesc, 1u);
    tc_mma(d, ahi, bhi, idesc, 1u);
  }
}

// ================================================================================================================
template <int MODE, int HC>
__global__ void __launch_bounds__(THREADS, 1) ffn_chain_kernel(const __grid_constant__ ChainK k) {
  const ChainParams& p = k.p;
  const Layout& L = k.L;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = sb0 + L.bar;
  const uint32_t a_full = bar, a_empty = bar + 8;
  auto wa_full = [&](int s) { return bar + 16u + 8u * s; };
  auto wa_empty = [&](int s) { return bar + 32u + 8u * s; };
  auto wb_full = [&](int s) { return bar + 48u + 8u * s; };
  auto wb_empty = [&](int s) { return bar + 64u + 8u * s; };
  auto h_full = [&](int b) { return bar + 80u + 8u * b; };
  auto h_empty = [&](int b) { return bar + 96u + 8u * b; };
  const uint32_t act_full = bar + 112, act_empty = bar + 120, o_full = bar + 128, o_empty = bar + 136, tmem_slot = bar + 144;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NA = (MODE == 1) ? 2 : 1;

  if (threadIdx.x == 0) {
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(wa_full(s), 1); mbar_init(wa_empty(s), 1); mbar_init(wb_full(s), 1); mbar_init(wb_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(h_full(b), 1); mbar_init(h_empty(b), NE); }
    mbar_init(act_full, NE); mbar_init(act_empty, 1); mbar_init(o_full, 1); mbar_init(o_empty, NE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  // TMEM columns: [b * NA * HC, ...) = acc_h[b] (, acc_d[b]);  acc_o behind them
  const uint32_t col_o = 2u * NA * HC;
  const int nch = k.nch, C = p.C;
  const int kbc = C / 32;

  if (warp == W_PROD_A) {
    if (lane == 0) {
      uint32_t n_t = 0, n_w = 0;
      for (int t = blockIdx.x; t < k.m_tiles; t += gridDim.x, ++n_t) {
        mbar_wait(a_empty, (n_t & 1u) ^ 1u);
        if ((p.debug & 8) && n_t > 0) { mbar_arrive(a_full); goto weights; }
        expect_tx(a_full, NA * L.a_tile);
        bulk_g2s(sb0 + L.a, reinterpret_cast<const uint8_t*>(p.a1) + (long long)t * L.a_tile, L.a_tile, a_full);
        if (MODE == 1) bulk_g2s(sb0 + L.a + L.a_tile, reinterpret_cast<const uint8_t*>(p.a2) + (long long)t * L.a_tile, L.a_tile, a_full);
      weights:
        for (int j = 0; j < nch; ++j, ++n_w) {
          const int s = n_w % L.sa;
          mbar_wait(wa_empty(s), ((n_w / L.sa) & 1u) ^ 1u);
          if ((p.debug & 4) && n_w >= (uint32_t)L.sa) { mbar_arrive(wa_full(s)); continue; }   // timing experiment: stale weights
          expect_tx(wa_full(s), L.wa_stage);
          const uint32_t dst = sb0 + L.wa + s * L.wa_stage;
          bulk_g2s(dst, reinterpret_cast<const uint8_t*>(p.wa1) + (long long)j * L.wa_chunk, L.wa_chunk, wa_full(s));
          if (MODE == 1) bulk_g2s(dst + L.wa_chunk, reinterpret_cast<const uint8_t*>(p.wa2) + (long long)j * L.wa_chunk, L.wa_chunk, wa_full(s));
        }
      }
    }
  } else if (warp == W_PROD_B) {
    if (lane == 0) {
      uint32_t n = 0;
      for (int t = blockIdx.x; t < k.m_tiles; t += gridDim.x) {
        for (int j = 0; j < nch; ++j, ++n) {
          const int s = n % L.sb;
          mbar_wait(wb_empty(s), ((n / L.sb) & 1u) ^ 1u);
          if ((p.debug & 4) && n >= (uint32_t)L.sb) { mbar_arrive(wb_full(s)); continue; }
          expect_tx(wb_full(s), L.wb_stage);
          bulk_g2s(sb0 + L.wb + s * L.wb_stage, reinterpret_cast<const uint8_t*>(p.wb) + (long long)j * L.wb_stage, L.wb_stage, wb_full(s));
        }
      }
    }
  } else if (warp == W_MMA) {
    if (lane == 0) {
      const uint32_t idesc_a = make_instr_desc(HC, false, false), idesc_b = make_instr_desc(C, false, false);
      uint32_t n_t = 0, n_c = 0, n_b = 0;
      for (int t = blockIdx.x; t < k.m_tiles; t += gridDim.x, ++n_t) {
        mbar_wait(a_full, n_t & 1u);
        tc_fence_after();
        for (int step = 0; step <= nch; ++step) {
          if (step < nch) {
            const int b = n_c & 1, s = n_c % L.sa;
            mbar_wait(h_empty(b), ((n_c >> 1) & 1u) ^ 1u);
            mbar_wait(wa_full(s), (n_c / L.sa) & 1u);
            tc_fence_after();
            const uint32_t wst = sb0 + L.wa + s * L.wa_stage;
            // the single issuing thread is the bottleneck of these narrow MMAs (N = 32..96 is 16..48 tensor cycles each):
            // descriptors are built once per operand block and only ADVANCED inside the loops (address field, 16-byte units)
#pragma unroll
            for (int g = 0; g < NA; ++g) {
              const uint32_t acc = tmem_base + (uint32_t)((b * NA + g) * HC);
              uint64_t da = make_smem_desc(sb0 + L.a + g * L.a_tile, false), db = make_smem_desc(wst + g * L.wa_chunk, false);
              for (int kb = 0; kb < kbc; ++kb) {
                mma3(acc, da, da + 512u, db, db + HC * 4u, idesc_a, kb > 0 ? 1u : 0u, p.passes);
                mma3(acc, da + 2u, da + 514u, db + 2u, db + HC * 4u + 2u, idesc_a, 1u, p.passes);
                da += 1024u; db += HC * 8u;
              }
            }
            tc_commit(wa_empty(s));
            tc_commit(h_full(b));
            if (step == nch - 1) tc_commit(a_empty);
            ++n_c;
          }
          if (step >= 1) {
            const int jb = step - 1, s = n_b % L.sb;
            if (jb == 0) mbar_wait(o_empty, (n_t & 1u) ^ 1u);
            mbar_wait(act_full, n_b & 1u);
            mbar_wait(wb_full(s), (n_b / L.sb) & 1u);
            tc_fence_after();
            const uint32_t acc = tmem_base + col_o;
            uint64_t da = make_smem_desc(sb0 + L.act, false), db = make_smem_desc(sb0 + L.wb + s * L.wb_stage, false);
            const uint64_t blo = (uint64_t)(C * 4);              // lo plane of a Wb k-block: C * 64 bytes behind the hi plane
#pragma unroll
            for (int kb = 0; kb < HC / 32; ++kb) {
              mma3(acc, da, da + 512u, db, db + blo, idesc_b, (jb > 0 || kb > 0) ? 1u : 0u, p.passes);
              mma3(acc, da + 2u, da + 514u, db + 2u, db + blo + 2u, idesc_b, 1u, p.passes);
              da += 1024u; db += 2u * blo;
            }
            tc_commit(wb_empty(s));
            tc_commit(act_empty);
            if (jb == nch - 1) tc_commit(o_full);
            ++n_b;
          }
        }
      }
    }
  } else if (warp >= EPI0) {
    // ------------------------------ middle stage + final epilogue -------------------------------------------
    const int e = warp - EPI0, q = warp & 3, half = e >> 2;
    constexpr int HCW = HC / 2;                         // hidden columns of a chunk handled by this warp
    const int c0 = half * HCW;
    const uint32_t lane_t = (uint32_t)(q * 32) << 16;
    const uint32_t r = (uint32_t)(q * 32 + lane);       // token row inside the tile = TMEM lane
    const uint32_t stage_base = sb0 + L.stage + (uint32_t)e * EPI_STAGE_BYTES;
    const int rl = lane >> 2, c4 = (lane & 3) * 4;
    const int nchunks = C / 16;
    uint32_t n_c = 0, n_t = 0;
    for (int t = blockIdx.x; t < k.m_tiles; t += gridDim.x, ++n_t) {
      for (int j = 0; j < nch; ++j, ++n_c) {
        const int b = n_c & 1;
        mbar_wait(h_full(b), (n_c >> 1) & 1u);
        tc_fence_after();
        uint32_t rh[HCW], rd[HCW];
        const uint32_t th = tmem_base + lane_t + (uint32_t)((b * NA) * HC + c0);
#pragma unroll
        for (int g = 0; g < HCW / 16; ++g) {
          tc_ld16_issue(th + g * 16, *reinterpret_cast<uint32_t(*)[16]>(&rh[g * 16]));
          if constexpr (MODE == 1) tc_ld16_issue(th + HC + g * 16, *reinterpret_cast<uint32_t(*)[16]>(&rd[g * 16]));
        }
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_empty(b));
        float y[HCW];
        const float* b1 = p.bias1 + (long long)j * HC + c0;
        // training forward: the pre-activation h = A1 Wa1^T + b1 is stored once (fp32, row-major) for the GEMM-based
        // backward; each lane owns one token row and writes HCW * 4 contiguous bytes of it (whole 128-byte lines)
        float* hrow = nullptr;
        if (MODE == 0 && p.h_out) {
          const long long row = (long long)t * 128 + r;
          if (row < p.M) hrow = p.h_out + row * p.H4 + (long long)j * HC + c0;
        }
#pragma unroll
        for (int i = 0; i < HCW; i += 4) {
          const float4 bv = ldg_f4(b1 + i);
          const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
          if (MODE == 0 && hrow)
            *reinterpret_cast<float4*>(hrow + i) = make_float4(__uint_as_float(rh[i]) + bb[0], __uint_as_float(rh[i + 1]) + bb[1],
                                                               __uint_as_float(rh[i + 2]) + bb[2], __uint_as_float(rh[i + 3]) + bb[3]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float x = __uint_as_float(rh[i + u]) + bb[u];
            if (p.debug & 1) y[i + u] = x;
            else if constexpr (MODE == 0) y[i + u] = gelu_fast(x);
            else y[i + u] = __uint_as_float(rd[i + u]) * gelu_grad_fast(x);
          }
        }
        mbar_wait(act_empty, (n_c & 1u) ^ 1u);          // GEMM-b of the previous chunk has finished reading the block
#pragma unroll
        for (int i = 0; i < HCW; i += 8) {
          if (p.debug & 2) break;
          uint4 hi, lo;
          split8(&y[i], hi, lo);
          const int kk = c0 + i;
          const uint32_t o = sb0 + L.act + (uint32_t)(kk >> 5) * 16384u + kmajor_sw64_offset(r, (uint32_t)((kk & 31) >> 3));
          sts128(o, hi.x, hi.y, hi.z, hi.w);
          sts128(o + 8192u, lo.x, lo.y, lo.z, lo.w);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(act_full);
      }
      // ---- final epilogue of the tile: TMEM -> registers -> per-warp smem transpose -> coalesced 64-byte row segments
      mbar_wait(o_full, n_t & 1u);
      tc_fence_after();
      const long long row0 = (long long)t * 128 + q * 32;
      const uint32_t to = tmem_base + lane_t + col_o;
      for (int c = half; c < nchunks; c += 2) {
        uint32_t vr[16];
        tc_ld16_issue(to + (uint32_t)(c * 16), vr);
        const int n = c * 16 + c4;
        float4 rr[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const long long row = row0 + it * 8 + rl;
          rr[it] = (p.resid && row < p.M) ? ldg_f4(p.resid + row * C + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tc_wait_ld();
        if (c + 2 >= nchunks) {                          // this warp's last read of acc_o
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
        }
        __syncwarp();                                    // the previous chunk's reads of the staging rows are done
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          sts128(stage_base + (uint32_t)(lane * EPI_STAGE_ROW_FLOATS + 4 * jj) * 4u, vr[4 * jj], vr[4 * jj + 1], vr[4 * jj + 2], vr[4 * jj + 3]);
        __syncwarp();
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), sv = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.bias2) bv = ldg_f4(p.bias2 + n);
        if (p.col_scale) sv = ldg_f4(p.col_scale + n);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr_ = it * 8 + rl;
          const long long row = row0 + rr_;
          if (row >= p.M) continue;
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                       : "r"(stage_base + (uint32_t)(rr_ * EPI_STAGE_ROW_FLOATS + c4) * 4u) : "memory");
          x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
          if (p.aux_out) *reinterpret_cast<float4*>(p.aux_out + row * C + n) = x;
          x.x *= sv.x; x.y *= sv.y; x.z *= sv.z; x.w *= sv.w;
          if (p.row_scale) { const float rs = __ldg(p.row_scale + row); x.x *= rs; x.y *= rs; x.z *= rs; x.w *= rs; }
          x.x += rr[it].x; x.y += rr[it].y; x.z += rr[it].z; x.w += rr[it].w;
          *reinterpret_cast<float4*>(p.out + row * C + n) = x;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ================================================================================================================
static bool chain_layout(int mode, int C, int HC, int sa, int sb, Layout& L) {
  const uint32_t nA = mode == 1 ? 2 : 1;
  L.a_tile = (uint32_t)(C / 32) * 16384u;
  L.wa_chunk = (uint32_t)(C / 32) * HC * 128u;
  L.wa_stage = nA * L.wa_chunk;
  L.wb_stage = (uint32_t)(HC / 32) * C * 128u;
  L.act_bytes = (uint32_t)(HC / 32) * 16384u;
  L.sa = sa; L.sb = sb;
  uint32_t o = 0;
  L.a = o; o += nA * L.a_tile;
  L.ones = 0; L.a2 = L.a + L.a_tile;
  L.wa = o; o += sa * L.wa_stage;
  L.wb = o; o += sb * L.wb_stage;
  L.act = o; o += L.act_bytes;
  L.act2 = 0;
  L.stage = o; o += STAGING;
  L.bar = o; o += BARS;
  L.total = o + 1024u;
  return L.total <= SMEM_LIMIT + 1024u;
}

// HC_req > 0: only layouts with that chunk width (the caller packed its weights for it)
static bool pick_chain(int mode, int C, int HC_req, int& HC, Layout& L) {
  static const int cand[][3] = {{64, 2, 2}, {32, 2, 2}, {64, 2, 1}, {32, 2, 1}, {32, 1, 1}};
  if (C % 32 != 0 || C < 32 || C > 256) return false;
  for (auto& c : cand)
    if ((HC_req == 0 || c[0] == HC_req) && chain_layout(mode, C, c[0], c[1], c[2], L)) { HC = c[0]; return true; }
  return false;
}

int chain_chunk(int mode, int C) {
  int HC = 0; Layout L;
  return pick_chain(mode, C, 0, HC, L) ? HC : 0;
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

int chain(const ChainParams& p, cudaStream_t stream) {
  SM3_REQUIRE(p.mode == 0 || p.mode == 1, SM3_ERR_INVALID_ARG, "ffn chain: mode");
  SM3_REQUIRE(p.a1 && p.wa1 && p.wb && p.bias1 && p.out && (p.mode == 0 || (p.a2 && p.wa2)), SM3_ERR_INVALID_ARG, "ffn chain: null operand");
  SM3_REQUIRE(p.M > 0 && p.H4 > 0 && p.C > 0, SM3_ERR_INVALID_ARG, "ffn chain: bad shape");
  ChainK k{};
  k.p = p;
  int HC = 0;
  SM3_REQUIRE(p.HC == 32 || p.HC == 64, SM3_ERR_INVALID_ARG, "ffn chain: chunk must be 32 or 64 (sm3_ffn_fused_chunk), got %d", p.HC);
  SM3_REQUIRE(pick_chain(p.mode, p.C, p.HC, HC, k.L), SM3_ERR_UNSUPPORTED_SHAPE, "ffn chain: C=%d with chunk %d does not fit shared memory (mode %d)", p.C, p.HC, p.mode);
  SM3_REQUIRE(p.H4 % HC == 0 && p.C % 16 == 0, SM3_ERR_UNSUPPORTED_SHAPE, "ffn chain: H4=%d not a multiple of the chunk %d", p.H4, HC);
  SM3_REQUIRE(2 * (p.mode == 1 ? 2 : 1) * HC + p.C <= 512, SM3_ERR_UNSUPPORTED_SHAPE, "ffn chain: TMEM budget");
  SM3_REQUIRE(aligned16(p.a1) && aligned16(p.wa1) && aligned16(p.wb) && aligned16(p.out) && aligned16(p.bias1) &&
              (!p.a2 || aligned16(p.a2)) && (!p.wa2 || aligned16(p.wa2)) && (!p.resid || aligned16(p.resid)) &&
              (!p.aux_out || aligned16(p.aux_out)) && (!p.bias2 || aligned16(p.bias2)) && (!p.col_scale || aligned16(p.col_scale)),
              SM3_ERR_INVALID_ARG, "ffn chain: pointers must be 16B aligned");
  k.p.passes = (p.passes == 1) ? 1 : 3;
  k.m_tiles = (p.M + 127) / 128;
  k.nch = p.H4 / HC;
  int grid = persistent_grid_sms();
  if (grid > k.m_tiles) grid = k.m_tiles;
#define SM3_CHAIN_LAUNCH(MODE_, HC_)                                                                                        \
  do {                                                                                                                      \
    cudaFuncSetAttribute(ffn_chain_kernel<MODE_, HC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.L.total);        \
    ffn_chain_kernel<MODE_, HC_><<<grid, THREADS, k.L.total, stream>>>(k);                                                  \
  } while (0)
  if (p.mode == 0) { if (HC == 64) SM3_CHAIN_LAUNCH(0, 64); else SM3_CHAIN_LAUNCH(0, 32); }
  else { if (HC == 64) SM3_CHAIN_LAUNCH(1, 64); else SM3_CHAIN_LAUNCH(1, 32); }
#undef SM3_CHAIN_LAUNCH
  return check_launch("ffn_chain_kernel");
}

}  // namespace ffn
}  // namespace sm3
