// attention_tc.cu -- tensor-core self-attention for short sequences (S <= 128, head dim 64), sm_100a.
//
// Replaces BertSelfAttention inside SentenceTransformer.encode (reference call sites
// services/embedding_service.py:81,97-102,120):  ctx = softmax(Q K^T / 8 + mask) V  per (sequence, head).
//
// A work item is (token tile, head).  A token tile packs G = floor(128 / S) whole sequences into
// the 128 MMA rows (G * S <= 128 rows used); the block-diagonal + length mask keeps sequences
// apart.  Per item:
//   TMA      Q, K, V tiles [128 x 64] bf16 (one 128-byte-swizzled K block each) -> shared memory
//   MMA 1    S = Q K^T           tcgen05 SS, M=128 N=128 K=64, fp32 in TMEM (128 columns)
//   softmax  thread == TMEM lane == query row: tcgen05.ld, mask, max, exp2, row sum; the
//            un-normalised P goes back to TMEM as packed bf16 (64 columns) with tcgen05.st
//   MMA 2    O = P V             tcgen05 TS: A = P from TMEM, B = V read MN-major from the same
//            swizzled tile TMA wrote (no transpose pass), M=128 N=64 K=128, fp32 in TMEM
//   epilogue tcgen05.ld O, scale by 1 / row sum, bf16, 128 contiguous bytes per row to ctx
// Two softmax warpgroups ping-pong over two TMEM sets (2 x 256 columns) and a 3-slot shared
// memory ring, so the MUFU-bound softmax of one item overlaps the loads and MMAs of the next.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer,
// warps 2..5 / 6..9 softmax warpgroups 0 / 1 (TMEM lane quarter = warp % 4).
//
// Roofline: HBM (reads the 2304-wide qkv rows once, writes ctx once); the math is ~1 % of a layer.
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "encoder_kernels.h"
#include "ptx.cuh"

namespace icd {
namespace {

constexpr int H = 768;
constexpr int HD = 64;
constexpr int kHeads = 12;
constexpr int kTile = 128;
constexpr int kSlots = 3;
constexpr int kTileBytes = kTile * HD * 2;  // 16 KiB
constexpr int kSlotBytes = 3 * kTileBytes;  // Q, K, V
constexpr int kThreads = 320;
constexpr int kTmemCols = 512;
constexpr int kSetCols = 256;  // S 128 | P 64 | O 64
constexpr int kSmemBytes = kSlots * kSlotBytes + 1024 + 256;

struct AttParams {
  const int32_t* lens;
  __nv_bfloat16* ctx;
  int B, S, G, tiles;
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSlots * kSlotBytes);
  uint64_t* full_bar = bars;             // [3] TMA landed
  uint64_t* empty_bar = bars + 3;        // [3] P V of the item retired: slot reusable
  uint64_t* sfull_bar = bars + 6;        // [2] S ready in TMEM
  uint64_t* pready_bar = bars + 8;       // [2] P written by all 128 rows
  uint64_t* ofull_bar = bars + 10;       // [2] O ready in TMEM
  uint64_t* setfree_bar = bars + 12;     // [2] TMEM set drained by all 128 rows
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 14);

  const int n_items = p.tiles * kHeads;
  const int n_local = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmap_qkv);
    for (int s = 0; s < kSlots; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&sfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&pready_bar[s]), 128);
      ptx::mbar_init(ptx::smem_u32(&ofull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&setfree_bar[s]), 128);
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
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      for (int j = 0; j < n_local; ++j) {
        const int item = (int)blockIdx.x + j * (int)gridDim.x;
        const int tile = item / kHeads, head = item % kHeads;
        const int row0 = tile * p.G * p.S;
        const int slot = j % kSlots;
        ptx::mbar_wait(ptx::smem_u32(&empty_bar[slot]), ((j / kSlots) & 1) ^ 1);
        const uint32_t fb = ptx::smem_u32(&full_bar[slot]);
        ptx::mbar_expect_tx(fb, kSlotBytes);
        const uint32_t dst = ptx::smem_u32(smem + (size_t)slot * kSlotBytes);
        ptx::tma_load_2d(dst, &tmap_qkv, fb, head * HD, row0);
        ptx::tma_load_2d(dst + kTileBytes, &tmap_qkv, fb, H + head * HD, row0);
        ptx::tma_load_2d(dst + 2 * kTileBytes, &tmap_qkv, fb, 2 * H + head * HD, row0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(kTile, kTile);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(kTile, HD) | (1u << 16);  // B (= V) is MN-major
      auto issue_s = [&](int j) {
        const int slot = j % kSlots, set = j & 1;
        ptx::mbar_wait(ptx::smem_u32(&full_bar[slot]), (j / kSlots) & 1);
        ptx::mbar_wait(ptx::smem_u32(&setfree_bar[set]), ((j >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t sq = ptx::smem_u32(smem + (size_t)slot * kSlotBytes);
        const uint32_t sk = sq + kTileBytes;
        const uint32_t d = tmem_base + (uint32_t)(set * kSetCols);
#pragma unroll
        for (int k4 = 0; k4 < HD / 16; ++k4)
          ptx::mma_ss(d, ptx::make_desc_k128(sq + k4 * 32), ptx::make_desc_k128(sk + k4 * 32), idesc_s, k4 ? 1u : 0u);
        ptx::tc_commit(ptx::smem_u32(&sfull_bar[set]));
      };
      auto issue_pv = [&](int j) {
        const int slot = j % kSlots, set = j & 1;
        ptx::mbar_wait(ptx::smem_u32(&pready_bar[set]), (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t sv = ptx::smem_u32(smem + (size_t)slot * kSlotBytes) + 2 * kTileBytes;
        const uint32_t d = tmem_base + (uint32_t)(set * kSetCols + 192);
        const uint32_t a = tmem_base + (uint32_t)(set * kSetCols + 128);
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks)  // 16 keys per MMA: two 8-key groups, 1024 bytes apart
          ptx::mma_ts(d, a + (uint32_t)(ks * 8), ptx::make_desc_k128(sv + ks * 2048), idesc_o, ks ? 1u : 0u);
        ptx::tc_commit(ptx::smem_u32(&ofull_bar[set]));
        ptx::tc_commit(ptx::smem_u32(&empty_bar[slot]));
      };
      for (int j = 0; j < n_local; ++j) {
        issue_s(j);
        if (j > 0) issue_pv(j - 1);
      }
      if (n_local > 0) issue_pv(n_local - 1);
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int wg = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = 32 * quarter + lane;  // row of the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * quarter) << 16);
    const int g = r / p.S;
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e): softmax in base 2
    for (int j = wg; j < n_local; j += 2) {
      const int item = (int)blockIdx.x + j * (int)gridDim.x;
      const int tile = item / kHeads, head = item % kHeads;
      const int set = j & 1;
      const uint32_t par = (uint32_t)(j >> 1) & 1;
      const int seq = tile * p.G + g;
      const bool row_used = g < p.G && seq < p.B;
      const int len = row_used ? min(p.lens[seq], p.S) : 0;
      const int lo = g * p.S, hi = lo + len;  // key columns this row attends to
      const bool q_live = row_used && (r - lo) < len;
      const int wlo = __reduce_min_sync(0xffffffffu, lo);
      const int whi = __reduce_max_sync(0xffffffffu, hi);

      ptx::mbar_wait(ptx::smem_u32(&sfull_bar[set]), par);
      ptx::tc_fence_after();
      const uint32_t s_addr = lane_addr + (uint32_t)(set * kSetCols);
      uint32_t sv[kTile];
#pragma unroll
      for (int c32 = 0; c32 < 4; ++c32) {
        if (c32 * 32 < whi && c32 * 32 + 32 > wlo) {  // warp-uniform: some row of this warp needs the chunk
          ptx::tmem_ld_32x32b_x32(s_addr + c32 * 32, sv + c32 * 32);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) sv[c32 * 32 + c] = 0xff800000u;  // -inf
        }
      }
      ptx::tmem_ld_wait();
      float m = -INFINITY;
#pragma unroll
      for (int c = 0; c < kTile; ++c) {
        const float v = (c >= lo && c < hi) ? __uint_as_float(sv[c]) : -INFINITY;
        sv[c] = __float_as_uint(v);
        m = fmaxf(m, v);
      }
      const float msc = (m == -INFINITY) ? 0.f : m * sc;
      float sum = 0.f;
      const uint32_t p_addr = s_addr + 128;
#pragma unroll
      for (int c16 = 0; c16 < 4; ++c16) {  // 32 keys -> 16 packed columns
        uint32_t pk[16];
        if (c16 * 32 < whi && c16 * 32 + 32 > wlo) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float e0 = exp2f(fmaf(__uint_as_float(sv[c16 * 32 + 2 * c]), sc, -msc));
            const float e1 = exp2f(fmaf(__uint_as_float(sv[c16 * 32 + 2 * c + 1]), sc, -msc));
            sum += e0 + e1;
            pk[c] = pack_bf16(e0, e1);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) pk[c] = 0u;
        }
        ptx::tmem_st_32x32b_x16(p_addr + c16 * 16, pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(ptx::smem_u32(&pready_bar[set]));

      ptx::mbar_wait(ptx::smem_u32(&ofull_bar[set]), par);
      ptx::tc_fence_after();
      uint32_t ov[HD];
      ptx::tmem_ld_32x32b_x32(s_addr + 192, ov);
      ptx::tmem_ld_32x32b_x32(s_addr + 224, ov + 32);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(ptx::smem_u32(&setfree_bar[set]));

      if (row_used) {
        // padded query positions are never read downstream (masked keys, masked pooling): zeros
        const float inv = (q_live && sum > 0.f) ? 1.0f / sum : 0.f;
        __nv_bfloat16* orow = p.ctx + ((size_t)tile * p.G * p.S + r) * H + head * HD;
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(ov[c * 8 + 0]) * inv, __uint_as_float(ov[c * 8 + 1]) * inv);
          o.y = pack_bf16(__uint_as_float(ov[c * 8 + 2]) * inv, __uint_as_float(ov[c * 8 + 3]) * inv);
          o.z = pack_bf16(__uint_as_float(ov[c * 8 + 4]) * inv, __uint_as_float(ov[c * 8 + 5]) * inv);
          o.w = pack_bf16(__uint_as_float(ov[c * 8 + 6]) * inv, __uint_as_float(ov[c * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 8) = o;
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

int attention_make_map(void* map128, const void* qkv, int64_t rows) {
  return make_tmap_bf16_2d(map128, qkv, (uint64_t)rows, (uint64_t)(3 * H), kTile, HD, true);
}

int launch_attention_tc(const void* tmap_qkv, const int32_t* lens, int B, int S, void* ctx, cudaStream_t st) {
  if (S < 1 || S > kTile) {
    set_error("attention: S=%d outside [1, 128]", S);
    return ICD_E_UNSUPPORTED;
  }
  AttParams p{};
  p.lens = lens;
  p.ctx = reinterpret_cast<__nv_bfloat16*>(ctx);
  p.B = B;
  p.S = S;
  p.G = kTile / S;
  p.tiles = (B + p.G - 1) / p.G;
  CUtensorMap tm;
  memcpy(&tm, tmap_qkv, sizeof(tm));
  ICD_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const int grid = std::min(p.tiles * kHeads, kSMs);
  attention_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(tm, p);
  count_launch();
  ICD_CUDA(cudaGetLastError());
  return ICD_OK;
}

}  // namespace icd
