// Common helpers for the sm3det_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "sm3det_b200 kernels are written for sm_100a only"
#endif

namespace sm3 {

// ---- error plumbing (thread-local message, C ABI returns negative codes) -------------------
enum : int {
  SM3_OK = 0,
  SM3_ERR_INVALID_ARG = -1,
  SM3_ERR_UNSUPPORTED_SHAPE = -2,
  SM3_ERR_CUDA = -3,
  SM3_ERR_WORKSPACE = -4,
};

void set_last_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError() -> SM3_ERR_CUDA (+ message)

#define SM3_REQUIRE(cond, code, ...)                      \
  do {                                                    \
    if (!(cond)) {                                        \
      ::sm3::set_last_error(__VA_ARGS__);                 \
      return (code);                                      \
    }                                                     \
  } while (0)

int num_sms();  // cached per device

// ---- small device helpers ------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU and its derivative (matches torch.nn.GELU(approximate='none'))
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

}  // namespace sm3
