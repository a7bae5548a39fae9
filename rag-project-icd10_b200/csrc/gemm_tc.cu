// gemm_tc.cu -- bf16 tcgen05 GEMM with fused epilogues for the BERT encoder (sm_100a).
//
//   C[M, N] = epilogue( A[M, K] * W[N, K]^T + bias[N] )          A, W, C bf16; fp32 accumulate
//
// Replaces the nn.Linear calls inside SentenceTransformer.encode (reference call sites
// services/embedding_service.py:81,97-102,120): QKV projection, attention output projection,
// FFN up (+ erf-GELU) and FFN down (+ residual).  One persistent CTA per SM walks 128 x 256
// output tiles; A and W tiles arrive through TMA (128-byte swizzle) into a 4-stage mbarrier
// ring; tcgen05.mma (M=128, N=256, K=16) accumulates in TMEM, double-buffered (2 x 256
// columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer,
// warps 2..9 epilogue (lane quarter = warp % 4, column half = (warp - 2) / 4).
//
// Roofline: bf16 tensor pipe; 2*M*N*K flops per launch.
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "encoder_kernels.h"
#include "ptx.cuh"

namespace icd {
namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int kStages = 4;
constexpr int kABytes = BM * BK * 2;  // 16 KiB
constexpr int kBBytes = BN * BK * 2;  // 32 KiB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 320;
constexpr int kEpiWarps = 8;
constexpr int kTmemCols = 512;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
  const float* bias;               // [N]
  const __nv_bfloat16* residual;   // [M, N] or null
  __nv_bfloat16* out;              // [M, N]
  int M, N, K;
  int epi;                         // GemmEpilogue
};

// HF "gelu" = x * Phi(x) with Phi(x) = 0.5 * (1 + erf(x / sqrt(2))).  The epilogue of the FFN-up GEMM
// evaluates 32 K of these per 128 x 256 tile while the next tile's MMAs run, so it has to fit ~20 issue
// slots per element: Phi is evaluated as a logistic of an odd quintic fitted to the erf form,
//   Phi(x) ~= 1 / (1 + exp(-2 (c x + a x^3 + b x^5))),   max |x Phi(x) - gelu_erf(x)| = 2.6e-5 over all x
// (c, a, b from a minimax fit, tests/test_encoder_gpu.py checks the formula against erf): 7 FP32
// instructions + ex2.approx + rcp.approx, no branches.  x^2 is clamped so the quintic stays monotone.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) {
  constexpr float kS = -2.0f * 1.4426950408889634f;  // exp(-2 u) = exp2(kS * u)
  constexpr float kC = kS * 7.97507884e-01f, kA = kS * 3.70056460e-02f, kB = kS * -3.51516788e-04f;
  const float t = fminf(x * x, 36.0f);
  float p = fmaf(kB, t, kA);
  p = fmaf(p, t, kC);
  const float e = ex2_approx(p * x);
  return x * rcp_approx(1.0f + e);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = p.N / BN;
  const int tiles = m_tiles * n_tiles;
  const int nkb = p.K / BK;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(ptx::smem_u32(&tfull_bar[b]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty_bar[b]), kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_holder), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
          ptx::mbar_expect_tx(fb, kStageBytes);
          unsigned char* sa = smem + (size_t)stage * kStageBytes;
          ptx::tma_load_2d(ptx::smem_u32(sa), &tmap_a, fb, kb * BK, m0);
          ptx::tma_load_2d(ptx::smem_u32(sa + kABytes), &tmap_b, fb, kb * BK, n0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(ptx::smem_u32(&tempty_bar[buf]), ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4) {
            ptx::mma_ss(d_tmem, ptx::make_desc_k128(sa + k4 * 32), ptx::make_desc_k128(sb + k4 * 32), idesc,
                        (kb | k4) ? 1u : 0u);
          }
          ptx::tc_commit(ptx::smem_u32(&empty_bar[stage]));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::tc_commit(ptx::smem_u32(&tfull_bar[buf]));
      }
    }
  } else {
    // epilogue: thread = row of the tile (TMEM lane), this warp covers 128 of the 256 columns
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = 32 * quarter + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * quarter) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
      const int row = m0 + row_in_tile;
      ptx::mbar_wait(ptx::smem_u32(&tfull_bar[buf]), (it >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c32 = 0; c32 < (BN / 2) / 32; ++c32) {
        const int col0 = half * (BN / 2) + c32 * 32;  // column within the tile
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(buf * BN + col0), r);
        ptx::tmem_ld_wait();
        if (c32 == (BN / 2) / 32 - 1) {
          // last read of this accumulator buffer by this warp: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty_bar[buf]));
        }
        if (row < p.M) {
          const int gcol = n0 + col0;
          const float4* bias4 = reinterpret_cast<const float4*>(p.bias + gcol);
          __nv_bfloat16* orow = p.out + (size_t)row * p.N + gcol;
          const uint4* res4 = p.residual ? reinterpret_cast<const uint4*>(p.residual + (size_t)row * p.N + gcol) : nullptr;
#pragma unroll
          for (int v = 0; v < 4; ++v) {  // 8 columns per 16-byte store
            float x[8];
            const float4 b0 = bias4[2 * v], b1 = bias4[2 * v + 1];
            x[0] = __uint_as_float(r[8 * v + 0]) + b0.x;
            x[1] = __uint_as_float(r[8 * v + 1]) + b0.y;
            x[2] = __uint_as_float(r[8 * v + 2]) + b0.z;
            x[3] = __uint_as_float(r[8 * v + 3]) + b0.w;
            x[4] = __uint_as_float(r[8 * v + 4]) + b1.x;
            x[5] = __uint_as_float(r[8 * v + 5]) + b1.y;
            x[6] = __uint_as_float(r[8 * v + 6]) + b1.z;
            x[7] = __uint_as_float(r[8 * v + 7]) + b1.w;
            if (p.epi == EPI_BIAS_GELU) {
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = gelu_erf(x[e]);
            } else if (p.epi == EPI_BIAS_RESIDUAL) {
              const uint4 rr = res4[v];
              x[0] += bf16lo_to_f32(rr.x);
              x[1] += bf16hi_to_f32(rr.x);
              x[2] += bf16lo_to_f32(rr.y);
              x[3] += bf16hi_to_f32(rr.y);
              x[4] += bf16lo_to_f32(rr.z);
              x[5] += bf16hi_to_f32(rr.z);
              x[6] += bf16lo_to_f32(rr.w);
              x[7] += bf16hi_to_f32(rr.w);
            }
            uint4 o;
            o.x = pack_bf16(x[0], x[1]);
            o.y = pack_bf16(x[2], x[3]);
            o.z = pack_bf16(x[4], x[5]);
            o.w = pack_bf16(x[6], x[7]);
            *reinterpret_cast<uint4*>(orow + 8 * v) = o;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int gemm_tile_n() { return BN; }

int launch_gemm_tc(const GemmArgs& a, cudaStream_t st) {
  if (a.N % BN != 0 || a.K % BK != 0 || a.M <= 0) {
    set_error("gemm_tc: needs N %% %d == 0 and K %% %d == 0 (M=%d N=%d K=%d)", BN, BK, a.M, a.N, a.K);
    return ICD_E_ARG;
  }
  ICD_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CUtensorMap ta, tb;
  memcpy(&ta, a.tmap_a, sizeof(ta));
  memcpy(&tb, a.tmap_b, sizeof(tb));
  GemmParams p{};
  p.bias = a.bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a.residual);
  p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.epi = a.epi;
  const int tiles = ((a.M + BM - 1) / BM) * (a.N / BN);
  const int grid = std::min(tiles, kSMs);
  gemm_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(ta, tb, p);
  count_launch();
  ICD_CUDA(cudaGetLastError());
  return ICD_OK;
}

int gemm_make_map_a(void* map128, const void* base, int64_t rows, int K) {
  return make_tmap_bf16_2d(map128, base, (uint64_t)rows, (uint64_t)K, BM, BK, true);
}
int gemm_make_map_b(void* map128, const void* base, int64_t rows, int K) {
  return make_tmap_bf16_2d(map128, base, (uint64_t)rows, (uint64_t)K, BN, BK, true);
}

}  // namespace icd
