// Host launcher for the split-bf16 tcgen05 GEMM (see gemm_tc.cuh).
#define SM3_GEMM_KERNEL_IMPL
#include "gemm_tc.cuh"
#include <mutex>

namespace sm3 {
const char* last_error();
namespace gemm {

int pick_bn(int N) {
  static const int cand[] = {256, 224, 192, 160, 128, 96, 64, 32};
  for (int bn : cand)
    if (N % bn == 0) return bn;
  return 0;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch(Params p, cudaStream_t stream) {
  SM3_REQUIRE(p.A && p.B && p.D, SM3_ERR_INVALID_ARG, "gemm: null operand");
  SM3_REQUIRE(p.M > 0 && p.N > 0 && p.K >= 0, SM3_ERR_INVALID_ARG, "gemm: bad shape %d %d %d", p.M, p.N, p.K);
  SM3_REQUIRE((p.a_smn == 1) != (p.a_sk == 1) || (p.a_smn == 1 && p.M == 1), SM3_ERR_INVALID_ARG,
              "gemm: A needs exactly one unit stride");
  SM3_REQUIRE((p.b_smn == 1) != (p.b_sk == 1), SM3_ERR_INVALID_ARG, "gemm: B needs exactly one unit stride");
  const bool a_mn = (p.a_smn == 1 && p.a_sk != 1), b_mn = (p.b_smn == 1 && p.b_sk != 1);
  if (p.BN == 0) p.BN = pick_bn(p.N);
  SM3_REQUIRE(p.BN >= 32 && p.BN <= MAX_BN && p.BN % 32 == 0 && p.N % p.BN == 0, SM3_ERR_UNSUPPORTED_SHAPE,
              "gemm: N=%d has no tile width (multiple of 32 <= 256 dividing N)", p.N);
  SM3_REQUIRE(aligned16(p.A) && aligned16(p.B) && aligned16(p.D), SM3_ERR_INVALID_ARG, "gemm: pointers must be 16B aligned");
  if (!a_mn) SM3_REQUIRE(p.a_smn % 4 == 0 && p.K % 4 == 0, SM3_ERR_UNSUPPORTED_SHAPE, "gemm: K-major A needs K%%4==0, lda%%4==0");
  else       SM3_REQUIRE(p.a_sk % 4 == 0 && p.M % 4 == 0, SM3_ERR_UNSUPPORTED_SHAPE, "gemm: MN-major A needs M%%4==0, lda%%4==0");
  if (!b_mn) SM3_REQUIRE(p.b_smn % 4 == 0 && p.K % 4 == 0, SM3_ERR_UNSUPPORTED_SHAPE, "gemm: K-major B needs K%%4==0, ldb%%4==0");
  else       SM3_REQUIRE(p.b_sk % 4 == 0, SM3_ERR_UNSUPPORTED_SHAPE, "gemm: MN-major B needs ldb%%4==0");
  SM3_REQUIRE(!(p.a_row_index && a_mn), SM3_ERR_INVALID_ARG, "gemm: row gather needs K-major A");
  SM3_REQUIRE(!(p.b_k_index && !b_mn), SM3_ERR_INVALID_ARG, "gemm: k gather needs MN-major B");
  SM3_REQUIRE(p.ldd % 4 == 0, SM3_ERR_UNSUPPORTED_SHAPE, "gemm: ldd%%4");
  if (p.epi & (EPI_GELU | EPI_AUXSTORE)) SM3_REQUIRE(!p.aux_out || (aligned16(p.aux_out) && p.ld_aux % 4 == 0), SM3_ERR_INVALID_ARG, "gemm: aux_out");
  if (p.epi & EPI_DGELU) SM3_REQUIRE(p.aux_in && aligned16(p.aux_in) && p.ld_aux % 4 == 0, SM3_ERR_INVALID_ARG, "gemm: aux_in");
  if (p.epi & EPI_BIAS) SM3_REQUIRE(p.bias && aligned16(p.bias) && p.bias_group_stride % 4 == 0, SM3_ERR_INVALID_ARG, "gemm: bias");
  if (p.epi & EPI_COLSCALE) SM3_REQUIRE(p.col_scale && aligned16(p.col_scale), SM3_ERR_INVALID_ARG, "gemm: col_scale");
  if (p.epi & EPI_ROWSCALE) SM3_REQUIRE(p.row_scale, SM3_ERR_INVALID_ARG, "gemm: row_scale");
  if (p.epi & EPI_RESID) SM3_REQUIRE(p.resid && aligned16(p.resid) && p.ld_resid % 4 == 0, SM3_ERR_INVALID_ARG, "gemm: resid");

  p.n_tiles = p.N / p.BN;
  p.m_tiles = (p.M + BM - 1) / BM;
  if (p.sched == SCHED_DENSE) {
    p.k_splits = 1; p.num_groups = 1;
    p.num_tiles = p.m_tiles * p.n_tiles;
  } else if (p.sched == SCHED_GROUPED) {
    SM3_REQUIRE(p.tile_group && p.num_m_tiles_dev, SM3_ERR_INVALID_ARG, "gemm: grouped schedule needs tile map");
    p.k_splits = 1;
    p.num_tiles = p.m_tiles * p.n_tiles;  // upper bound; the device scalar decides
  } else if (p.sched == SCHED_SPLITK) {
    SM3_REQUIRE(p.k_splits >= 1 && p.num_groups >= 1, SM3_ERR_INVALID_ARG, "gemm: split-K needs k_splits/num_groups");
    SM3_REQUIRE((p.epi & EPI_ATOMIC) || p.k_splits == 1, SM3_ERR_INVALID_ARG, "gemm: split-K>1 needs EPI_ATOMIC");
    p.num_tiles = p.num_groups * p.m_tiles * p.n_tiles * p.k_splits;
  } else {
    SM3_REQUIRE(false, SM3_ERR_INVALID_ARG, "gemm: bad schedule %d", p.sched);
  }
  // 32-bit element offsets inside the kernel: every operand must span < 2^32 floats (16 GiB)
  {
    const long long a_ext = a_mn ? (long long)p.K * p.a_sk : (long long)(p.a_row_index ? (1LL << 31) / (p.a_smn ? p.a_smn : 1) : p.M) * p.a_smn;
    const long long b_ext = b_mn ? (long long)(p.b_k_index ? 1 : p.K) * p.b_sk + p.N : (long long)p.N * p.b_smn;
    SM3_REQUIRE(a_ext < (1LL << 32) && b_ext < (1LL << 32), SM3_ERR_UNSUPPORTED_SHAPE, "gemm: operand larger than 2^32 elements");
  }
  int grid = num_sms();
  if (grid > p.num_tiles) grid = p.num_tiles;
  if (grid < 1) grid = 1;
#define SM3_GEMM_LAUNCH(AMN, BMN)                                                                                   \
  do {                                                                                                              \
    cudaFuncSetAttribute(gemm_bf16x3_kernel<AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES); \
    gemm_bf16x3_kernel<AMN, BMN><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);                                     \
  } while (0)
  if (!a_mn && !b_mn) SM3_GEMM_LAUNCH(false, false);
  else if (!a_mn && b_mn) SM3_GEMM_LAUNCH(false, true);
  else if (a_mn && b_mn) SM3_GEMM_LAUNCH(true, true);
  else SM3_REQUIRE(false, SM3_ERR_UNSUPPORTED_SHAPE, "gemm: MN-major A with K-major B is not instantiated");
#undef SM3_GEMM_LAUNCH
  return check_launch("gemm_bf16x3_kernel");
}

}  // namespace gemm
}  // namespace sm3
