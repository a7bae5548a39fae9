// encoder_kernels.cu -- the memory-bound pieces of the BERT encoder around the tcgen05 GEMMs:
// embedding gather + LayerNorm, LayerNorm, small-sequence attention, masked mean pool + L2
// normalise.  Arithmetic follows transformers.BertModel as driven by
// SentenceTransformer.encode (reference call sites services/embedding_service.py:81,97,120).
// Hidden size is 768 (24 elements per lane with a warp per token), heads of 64.
#include "common.cuh"
#include "encoder_kernels.h"
#include "ptx.cuh"

namespace icd {
namespace {

constexpr int H = 768;
constexpr int kPerLane = H / 32;  // 24 = 3 chunks of 8

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// lane owns elements c*256 + lane*8 + e  (c = 0..2, e = 0..7): 16-byte bf16 accesses, coalesced
__device__ __forceinline__ void ln_store(const float* x, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float eps, __nv_bfloat16* out_row, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) s += x[i];
  const float mean = warp_sum(s) * (1.0f / H);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const float d = x[i] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(warp_sum(v) * (1.0f / H) + eps);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = c * 256 + lane * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + col), g1 = *reinterpret_cast<const float4*>(gamma + col + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + col), b1 = *reinterpret_cast<const float4*>(beta + col + 4);
    const float* xx = x + c * 8;
    uint4 o;
    o.x = pack2((xx[0] - mean) * rstd * g0.x + b0.x, (xx[1] - mean) * rstd * g0.y + b0.y);
    o.y = pack2((xx[2] - mean) * rstd * g0.z + b0.z, (xx[3] - mean) * rstd * g0.w + b0.w);
    o.z = pack2((xx[4] - mean) * rstd * g1.x + b1.x, (xx[5] - mean) * rstd * g1.y + b1.y);
    o.w = pack2((xx[6] - mean) * rstd * g1.z + b1.z, (xx[7] - mean) * rstd * g1.w + b1.w);
    *reinterpret_cast<uint4*>(out_row + col) = o;
  }
}

__global__ void __launch_bounds__(256)
embed_ln_kernel(const int32_t* __restrict__ ids, int M, int S, int vocab, int unk, const float* __restrict__ word,
                const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out) {
  ptx::griddep_launch();
  ptx::griddep_wait();   // `out` is read by kernels of the previous forward that may still be running
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 8 + warp;
  if (tok >= M) return;
  int id = ids[tok];
  if ((unsigned)id >= (unsigned)vocab) id = unk;  // an id outside the table (mismatched vocab.txt) reads [UNK], not foreign memory
  const int t = tok % S;
  float x[kPerLane];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int col = c * 256 + lane * 8;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 w = *reinterpret_cast<const float4*>(word + (size_t)id * H + col + 4 * h);
      const float4 p = *reinterpret_cast<const float4*>(pos + (size_t)t * H + col + 4 * h);
      const float4 ty = *reinterpret_cast<const float4*>(type0 + col + 4 * h);
      x[c * 8 + 4 * h + 0] = w.x + ty.x + p.x;
      x[c * 8 + 4 * h + 1] = w.y + ty.y + p.y;
      x[c * 8 + 4 * h + 2] = w.z + ty.z + p.z;
      x[c * 8 + 4 * h + 3] = w.w + ty.w + p.w;
    }
  }
  ln_store(x, gamma, beta, eps, out + (size_t)tok * H, lane);
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16* __restrict__ in, int M, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 8 + warp;
  if (tok >= M) return;
  float x[kPerLane];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + (size_t)tok * H + c * 256 + lane * 8);
    x[c * 8 + 0] = bf16lo_to_f32(v.x);
    x[c * 8 + 1] = bf16hi_to_f32(v.x);
    x[c * 8 + 2] = bf16lo_to_f32(v.y);
    x[c * 8 + 3] = bf16hi_to_f32(v.y);
    x[c * 8 + 4] = bf16lo_to_f32(v.z);
    x[c * 8 + 5] = bf16hi_to_f32(v.z);
    x[c * 8 + 6] = bf16lo_to_f32(v.w);
    x[c * 8 + 7] = bf16hi_to_f32(v.w);
  }
  ln_store(x, gamma, beta, eps, out + (size_t)tok * H, lane);
}

// ---------------------------------------------------------------- attention for 128 < S <= 512
// The tensor-core kernel (attention_tc.cu) holds one whole sequence in a 128-row tile; longer sequences -- only the
// token-classification path sends them (the reference's NER pipeline reads up to 512 tokens in one pass,
// services/medical_ner_service.py:177-229) -- take this kernel: one CTA per (sequence, head, 128-query tile), the
// K and V rows of the head staged once in shared memory as bf16 (<= 128 KiB), thread == query row, online softmax
// over the keys j < len in chunks of 4, fp32 accumulation.  CUDA cores: 2 * len * 64 FMAs per query row; at the
// handful of long texts a request carries this is microseconds, and the layer's GEMMs stay on the tensor cores.
constexpr int HD = 64;
constexpr int kLongMaxS = 512;
__global__ void __launch_bounds__(128)
attention_long_kernel(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ lens, int S,
                      __nv_bfloat16* __restrict__ ctx) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  extern __shared__ __align__(16) unsigned char sm_raw[];
  uint4* Ks = reinterpret_cast<uint4*>(sm_raw);   // [len][8] uint4 = [len][64] bf16
  const int b = blockIdx.x, h = blockIdx.y;
  const int len = min(lens[b], S);
  uint4* Vs = Ks + (size_t)len * (HD / 8);
  const size_t row0 = (size_t)b * S;
  for (int i = threadIdx.x; i < len * (HD / 8); i += blockDim.x) {
    const int j = i / (HD / 8), c = i % (HD / 8);
    const __nv_bfloat16* base = qkv + (row0 + j) * (3 * H) + h * HD + c * 8;
    Ks[i] = *reinterpret_cast<const uint4*>(base + H);
    Vs[i] = *reinterpret_cast<const uint4*>(base + 2 * H);
  }
  __syncthreads();
  const int t = blockIdx.z * 128 + threadIdx.x;
  if (t >= S) return;
  __nv_bfloat16* orow = ctx + (row0 + t) * H + h * HD;
  if (t >= len) {  // padded query rows are never read downstream (masked keys, masked pooling)
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(0, 0, 0, 0);
    return;
  }
  float q[HD];
  {
    const __nv_bfloat16* qb = qkv + (row0 + t) * (3 * H) + h * HD;
    const float sc = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) and log2(e): the softmax runs in base 2
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 v = *reinterpret_cast<const uint4*>(qb + c * 8);
      q[c * 8 + 0] = bf16lo_to_f32(v.x) * sc; q[c * 8 + 1] = bf16hi_to_f32(v.x) * sc;
      q[c * 8 + 2] = bf16lo_to_f32(v.y) * sc; q[c * 8 + 3] = bf16hi_to_f32(v.y) * sc;
      q[c * 8 + 4] = bf16lo_to_f32(v.z) * sc; q[c * 8 + 5] = bf16hi_to_f32(v.z) * sc;
      q[c * 8 + 6] = bf16lo_to_f32(v.w) * sc; q[c * 8 + 7] = bf16hi_to_f32(v.w) * sc;
    }
  }
  float acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) acc[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < len; j0 += 4) {
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = min(j0 + u, len - 1);  // the tail repeats the last key; its weight is zeroed below
      const uint4* kr = Ks + (size_t)j * (HD / 8);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 k = kr[c];
        a0 = fmaf(q[c * 8 + 0], bf16lo_to_f32(k.x), a0); a1 = fmaf(q[c * 8 + 1], bf16hi_to_f32(k.x), a1);
        a0 = fmaf(q[c * 8 + 2], bf16lo_to_f32(k.y), a0); a1 = fmaf(q[c * 8 + 3], bf16hi_to_f32(k.y), a1);
        a0 = fmaf(q[c * 8 + 4], bf16lo_to_f32(k.z), a0); a1 = fmaf(q[c * 8 + 5], bf16hi_to_f32(k.z), a1);
        a0 = fmaf(q[c * 8 + 6], bf16lo_to_f32(k.w), a0); a1 = fmaf(q[c * 8 + 7], bf16hi_to_f32(k.w), a1);
      }
      s[u] = (j0 + u < len) ? a0 + a1 : -INFINITY;
    }
    const float mn = fmaxf(fmaxf(m, fmaxf(s[0], s[1])), fmaxf(s[2], s[3]));
    const float corr = exp2f(m - mn);  // 0 for the first chunk (m = -inf)
    float pw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) pw[u] = exp2f(s[u] - mn);  // exp2(-inf) = 0 for the masked tail
    l = fmaf(l, corr, (pw[0] + pw[1]) + (pw[2] + pw[3]));
    m = mn;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint4* vr = Vs + (size_t)min(j0 + u, len - 1) * (HD / 8);
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 v = vr[c];
        acc[c * 8 + 0] = fmaf(pw[u], bf16lo_to_f32(v.x), acc[c * 8 + 0]); acc[c * 8 + 1] = fmaf(pw[u], bf16hi_to_f32(v.x), acc[c * 8 + 1]);
        acc[c * 8 + 2] = fmaf(pw[u], bf16lo_to_f32(v.y), acc[c * 8 + 2]); acc[c * 8 + 3] = fmaf(pw[u], bf16hi_to_f32(v.y), acc[c * 8 + 3]);
        acc[c * 8 + 4] = fmaf(pw[u], bf16lo_to_f32(v.z), acc[c * 8 + 4]); acc[c * 8 + 5] = fmaf(pw[u], bf16hi_to_f32(v.z), acc[c * 8 + 5]);
        acc[c * 8 + 6] = fmaf(pw[u], bf16lo_to_f32(v.w), acc[c * 8 + 6]); acc[c * 8 + 7] = fmaf(pw[u], bf16hi_to_f32(v.w), acc[c * 8 + 7]);
      }
    }
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    uint4 o;
    o.x = pack2(acc[c * 8 + 0] * inv, acc[c * 8 + 1] * inv);
    o.y = pack2(acc[c * 8 + 2] * inv, acc[c * 8 + 3] * inv);
    o.z = pack2(acc[c * 8 + 4] * inv, acc[c * 8 + 5] * inv);
    o.w = pack2(acc[c * 8 + 6] * inv, acc[c * 8 + 7] * inv);
    *reinterpret_cast<uint4*>(orow + c * 8) = o;
  }
}

// ---------------------------------------------------------------- masked mean + L2 normalise
// sentence-transformers Pooling(mean) + F.normalize: sum(h*mask)/clamp(sum(mask),1e-9), then
// x / max(||x||, 1e-12).  One CTA (256 threads, 3 columns each) per sequence.
__global__ void __launch_bounds__(256)
pool_normalise_kernel(const __nv_bfloat16* __restrict__ h, const int32_t* __restrict__ lens, int S,
                      void* __restrict__ out, int out_dtype, int normalise) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  __shared__ float red[8];
  const int b = blockIdx.x;
  const int len = min(lens[b], S);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int t = 0; t < len; ++t) {
    const __nv_bfloat16* row = h + ((size_t)b * S + t) * H;
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += __bfloat162float(row[c * 256 + threadIdx.x]);
  }
  const float denom = fmaxf((float)len, 1e-9f);
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    acc[c] /= denom;
    ss = fmaf(acc[c], acc[c], ss);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float inv = normalise ? 1.0f / fmaxf(sqrtf(tot), 1e-12f) : 1.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = acc[c] * inv;
    const size_t o = (size_t)b * H + c * 256 + threadIdx.x;
    if (out_dtype == ICD_F32)
      reinterpret_cast<float*>(out)[o] = v;
    else
      reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(v);
  }
}

// Token-classification head (BertForTokenClassification.classifier): one warp per token row.  The row (768 bf16)
// sits in registers, 24 elements per lane; the L weight rows stream from L1/L2 (the same 3 KiB rows for every
// warp) as coalesced fp32 and each label costs one warp reduction.  HBM-bound on the hidden states (1.5 KB per
// token in, 4 L bytes out).
__global__ void __launch_bounds__(256)
token_head_kernel(const __nv_bfloat16* __restrict__ h, int M, const float* __restrict__ w,
                  const float* __restrict__ b, int L, float* __restrict__ out) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= M) return;
  float x[24];
  const __nv_bfloat16* row = h + (size_t)m * H;
#pragma unroll
  for (int j = 0; j < 24; ++j) x[j] = __bfloat162float(row[j * 32 + lane]);
  for (int l = 0; l < L; ++l) {
    const float* wr = w + (size_t)l * H;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 24; ++j) acc = fmaf(x[j], __ldg(wr + j * 32 + lane), acc);
    acc = warp_sum(acc);
    if (lane == 0) out[(size_t)m * L + l] = acc + __ldg(b + l);
  }
}

}  // namespace

static int g_encoder_pdl = 1;  // validated in r02o: parity suite green, batch-1 forward 0.85 -> 0.72 ms
int encoder_pdl() { return g_encoder_pdl; }
void encoder_set_pdl(int on) { g_encoder_pdl = on ? 1 : 0; }

int launch_token_head(const void* h, int M, const float* w, const float* b, int L, float* out, cudaStream_t st) {
  if (M <= 0) return ICD_OK;
  ICD_CUDA(launch_chained(token_head_kernel, dim3((M + 7) / 8), dim3(256), 0, st, 1, reinterpret_cast<const __nv_bfloat16*>(h), M, w,
                          b, L, out));
  count_launch();
  return ICD_OK;
}

int launch_embed_ln(const int32_t* ids, int M, int S, int vocab, int unk, const float* word, const float* pos,
                    const float* type0, const float* gamma, const float* beta, float eps, void* out, cudaStream_t st) {
  ICD_CUDA(launch_chained(embed_ln_kernel, dim3((M + 7) / 8), dim3(256), 0, st, 1, ids, M, S, vocab, unk, word, pos, type0, gamma, beta,
                          eps, reinterpret_cast<__nv_bfloat16*>(out)));
  count_launch();
  return ICD_OK;
}

int launch_layernorm(const void* x, int M, const float* gamma, const float* beta, float eps, void* out,
                     cudaStream_t st) {
  ICD_CUDA(launch_chained(layernorm_kernel, dim3((M + 7) / 8), dim3(256), 0, st, 1, reinterpret_cast<const __nv_bfloat16*>(x), M, gamma,
                          beta, eps, reinterpret_cast<__nv_bfloat16*>(out)));
  count_launch();
  return ICD_OK;
}

int launch_attention_long(const void* qkv, const int32_t* lens, int B, int S, void* ctx, cudaStream_t st) {
  if (S < 1 || S > kLongMaxS) {
    set_error("attention: S=%d outside [1, %d]", S, kLongMaxS);
    return ICD_E_UNSUPPORTED;
  }
  const size_t smem = (size_t)2 * S * HD * 2;
  static bool attr_set = false;
  if (!attr_set) {
    ICD_CUDA(cudaFuncSetAttribute(attention_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kLongMaxS * HD * 2));
    attr_set = true;
  }
  ICD_CUDA(launch_chained(attention_long_kernel, dim3(B, 12, (S + 127) / 128), dim3(128), smem, st, 1,
                          reinterpret_cast<const __nv_bfloat16*>(qkv), lens, S, reinterpret_cast<__nv_bfloat16*>(ctx)));
  count_launch();
  return ICD_OK;
}

int launch_pool_normalise(const void* h, const int32_t* lens, int B, int S, void* out, int out_dtype,
                          cudaStream_t st) {
  ICD_CUDA(launch_chained(pool_normalise_kernel, dim3(B), dim3(256), 0, st, 1, reinterpret_cast<const __nv_bfloat16*>(h), lens, S, out,
                          out_dtype & 0xff, (out_dtype & ICD_OUT_NO_NORMALISE) ? 0 : 1));
  count_launch();
  return ICD_OK;
}

}  // namespace icd
