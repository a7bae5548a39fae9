// common.cuh -- shared host/device helpers of libicdrag (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/icdrag.h"

#ifndef __CUDA_ARCH__
#define ICD_HOST_ONLY 1
#endif

namespace icd {

// ------------------------------------------------------------------ errors / accounting
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define ICD_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      icd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (_e == cudaErrorMemoryAllocation) ? ICD_E_NOMEM : ICD_E_CUDA;                  \
    }                                                                                       \
  } while (0)

#define ICD_CHECK_ARG(cond, msg)                       \
  do {                                                 \
    if (!(cond)) {                                     \
      icd::set_error("%s: %s", __func__, msg);         \
      return ICD_E_ARG;                                \
    }                                                  \
  } while (0)

#define ICD_TRY(expr)              \
  do {                             \
    int _s = (expr);               \
    if (_s != ICD_OK) return _s;   \
  } while (0)

// memory space of a caller pointer: true = device (or managed), false = host
inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// ------------------------------------------------------------------ candidate ordering
// Total order of candidates everywhere in the library: score descending, then id ascending
// (what a stable sort over insertion order gives; oracle/search.py::_cut).
__host__ __device__ __forceinline__ bool cand_before(float sa, int64_t ia, float sb, int64_t ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

__device__ __forceinline__ float level_weight_f(uint8_t level) {
  // services/milvus_service.py:550-558
  return level == 1 ? 1.2f : (level == 3 ? 0.8f : 1.0f);
}
__device__ __forceinline__ double level_weight_d(uint8_t level) {
  return level == 1 ? 1.2 : (level == 3 ? 0.8 : 1.0);
}

// order-preserving float <-> int key (signed int compare == float compare, -0 < +0 harmless)
__device__ __forceinline__ int float_key(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// ------------------------------------------------------------------ bf16 <-> f32 bit tricks
__device__ __forceinline__ float bf16lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// 128-bit streaming load that does not allocate in L1 (rows are read once per scan)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ------------------------------------------------------------------ GELU of the encoder's FFN (gemm_tc.cu, skinny_linear.cu)
// HF "gelu" = x * Phi(x) with Phi(x) = 0.5 * (1 + erf(x / sqrt(2))).  The epilogue of the FFN-up GEMM
// (gemm_tc.cu) evaluates 32 K of these per 128 x 256 tile while the next tile's MMAs run (K = 768: the MMAs of a tile
// take about as long as its epilogue), so every instruction counts.  Phi is evaluated as a logistic of an odd quintic
// fitted to the erf form, written through tanh so that it costs ONE MUFU op:
//   Phi(x) ~= 1 / (1 + exp(-2 u)) = 0.5 (1 + tanh(u)),  u = c x + a x^3 + b x^5
//   gelu(x) ~= hx + hx tanh(u),  hx = x / 2
// (c, a, b from a minimax fit: max |x Phi(x) - gelu_erf(x)| = 2.6e-5 with exact tanh; tanh.approx.f32 adds at most
// 2^-11 |hx|, below the bf16 rounding of the stored value; tests/test_encoder_gpu.py checks the formula against
// erf).  x^2 is clamped so the quintic stays monotone.  8 FP32 instructions, no branches.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) {
  constexpr float kC = 7.97507884e-01f, kA = 3.70056460e-02f, kB = -3.51516788e-04f;
  const float t = fminf(x * x, 36.0f);
  float p = fmaf(kB, t, kA);
  p = fmaf(p, t, kC);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(p * x), hx);
}

}  // namespace icd
