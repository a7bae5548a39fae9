// api_encoder.cu -- C ABI of the BERT sentence encoder (include/icdrag.h, icd_encoder_*).
//
// Stands where SentenceTransformer(...).encode(normalize_embeddings=True) stands in the
// reference (services/embedding_service.py:61,81,97-102,120).  Weight blob order (fp32, HF
// BertModel names) -- mirrored by rag-project-icd10_b200/engine/weights.py:
//   embeddings.word_embeddings.weight        [V, H]
//   embeddings.position_embeddings.weight    [P, H]
//   embeddings.token_type_embeddings.weight  [T, H]
//   embeddings.LayerNorm.weight / .bias      [H] [H]
//   per layer l:
//     attention.self.query.weight, key.weight, value.weight      3 x [H, H]
//     attention.self.query.bias, key.bias, value.bias            3 x [H]
//     attention.output.dense.weight [H, H], .bias [H]
//     attention.output.LayerNorm.weight / .bias                  [H] [H]
//     intermediate.dense.weight [I, H], .bias [I]
//     output.dense.weight [H, I], .bias [H]
//     output.LayerNorm.weight / .bias                            [H] [H]
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "encoder_kernels.h"
#include "kernels.h"

namespace icd {

struct LayerW {
  // bf16 GEMM weights ([out, in] row-major == K-major B operand) and their TMA maps
  void *wqkv, *wo, *w1, *w2;
  alignas(128) unsigned char m_qkv[128];
  alignas(128) unsigned char m_o[128];
  alignas(128) unsigned char m_1[128];
  alignas(128) unsigned char m_2[128];
  // fp32 small parameters
  float *bqkv, *bo, *ln1g, *ln1b, *b1, *b2, *ln2g, *ln2b;
};

}  // namespace icd

struct icd_encoder {
  icd_bert_cfg cfg;
  int device = 0;
  // fp32 embedding tables and LN
  float *word = nullptr, *pos = nullptr, *type = nullptr, *eg = nullptr, *eb = nullptr;
  std::vector<icd::LayerW*> layers;
  std::vector<void*> allocs;
  // activations (bf16) for max_tokens rows
  int max_tokens = 0;
  void *h = nullptr, *h1 = nullptr, *t = nullptr, *ctx = nullptr, *qkv = nullptr, *f = nullptr;
  alignas(128) unsigned char m_h[128];
  alignas(128) unsigned char m_h1[128];
  alignas(128) unsigned char m_ctx[128];
  alignas(128) unsigned char m_f[128];
  alignas(128) unsigned char m_qkv[128];
  alignas(128) unsigned char m_t[128];        // GEMM output maps (box 128 x 64)
  alignas(128) unsigned char m_qkv_out[128];
  int32_t *ids = nullptr, *lens = nullptr;
  int ids_cap = 0, lens_cap = 0;
  void* out_stage = nullptr;
  size_t out_cap = 0;
  int last_M = 0;
};

namespace icd {

static int64_t weight_count(const icd_bert_cfg& c) {
  const int64_t H = c.hidden, I = c.intermediate;
  int64_t n = (int64_t)c.vocab_size * H + (int64_t)c.max_position * H + (int64_t)c.type_vocab * H + 2 * H;
  n += (int64_t)c.layers * (3 * H * H + 3 * H + H * H + H + 2 * H + I * H + I + H * I + H + 2 * H);
  return n;
}

static int dev_alloc(icd_encoder* e, void** p, size_t bytes) {
  ICD_CUDA(cudaMalloc(p, bytes));
  e->allocs.push_back(*p);
  return ICD_OK;
}

// upload `n` fp32 values from host `src` as fp32 (keep) or bf16 (GEMM weights)
static int upload(icd_encoder* e, const float* src, int64_t n, bool as_bf16, void** out, float* stage) {
  if (!as_bf16) {
    ICD_TRY(dev_alloc(e, out, (size_t)n * 4));
    ICD_CUDA(cudaMemcpy(*out, src, (size_t)n * 4, cudaMemcpyDefault));
    return ICD_OK;
  }
  ICD_TRY(dev_alloc(e, out, (size_t)n * 2));
  ICD_CUDA(cudaMemcpy(stage, src, (size_t)n * 4, cudaMemcpyDefault));
  ICD_TRY(launch_f32_to_bf16(stage, *out, n, 0));
  ICD_CUDA(cudaStreamSynchronize(0));
  return ICD_OK;
}

static int reserve_tokens(icd_encoder* e, int max_tokens) {
  if (max_tokens <= e->max_tokens) return ICD_OK;
  const int H = e->cfg.hidden, I = e->cfg.intermediate;
  const int M = ((max_tokens + 127) / 128) * 128;
  void** bufs[6] = {&e->h, &e->h1, &e->t, &e->ctx, &e->qkv, &e->f};
  const size_t widths[6] = {(size_t)H, (size_t)H, (size_t)H, (size_t)H, (size_t)3 * H, (size_t)I};
  for (int i = 0; i < 6; ++i) {
    if (*bufs[i]) cudaFree(*bufs[i]);
    *bufs[i] = nullptr;
  }
  e->max_tokens = 0;
  for (int i = 0; i < 6; ++i) {
    ICD_CUDA(cudaMalloc(bufs[i], (size_t)M * widths[i] * 2));
    ICD_CUDA(cudaMemset(*bufs[i], 0, (size_t)M * widths[i] * 2));
  }
  ICD_TRY(gemm_make_map_a(e->m_h, e->h, M, H));
  ICD_TRY(gemm_make_map_a(e->m_h1, e->h1, M, H));
  ICD_TRY(gemm_make_map_a(e->m_ctx, e->ctx, M, H));
  ICD_TRY(gemm_make_map_a(e->m_f, e->f, M, I));
  ICD_TRY(attention_make_map(e->m_qkv, e->qkv, M));
  ICD_TRY(gemm_make_map_out(e->m_t, e->t, M, H));
  ICD_TRY(gemm_make_map_out(e->m_qkv_out, e->qkv, M, 3 * H));
  e->max_tokens = M;
  return ICD_OK;
}

}  // namespace icd

using namespace icd;

extern "C" {

int64_t icd_encoder_weight_count(const icd_bert_cfg* cfg) { return cfg ? weight_count(*cfg) : -1; }

int icd_encoder_create(const float* weights, int64_t count, const icd_bert_cfg* cfg, int device, icd_encoder** out) {
  ICD_CHECK_ARG(weights && cfg && out, "null argument");
  ICD_CHECK_ARG(cfg->hidden == 768 && cfg->heads == 12, "this build supports hidden=768, heads=12 (head dim 64)");
  ICD_CHECK_ARG(cfg->intermediate > 0 && cfg->intermediate % gemm_tile_n() == 0, "intermediate must be a multiple of 256");
  ICD_CHECK_ARG(cfg->layers >= 1 && cfg->vocab_size > 0 && cfg->max_position > 0 && cfg->type_vocab > 0, "bad config");
  ICD_CHECK_ARG(count == weight_count(*cfg), "weight blob size does not match the config");
  int ndev = 0;
  ICD_CUDA(cudaGetDeviceCount(&ndev));
  ICD_CHECK_ARG(device >= 0 && device < ndev, "no such CUDA device");
  ICD_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ICD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libicdrag is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return ICD_E_UNSUPPORTED;
  }
  icd_encoder* e = new icd_encoder();
  e->cfg = *cfg;
  e->device = device;
  const int64_t H = cfg->hidden, I = cfg->intermediate;
  float* stage = nullptr;
  const int64_t stage_n = std::max<int64_t>(3 * H * H, I * H);
  int st = ICD_OK;
  auto fail = [&](int s) {
    if (stage) cudaFree(stage);
    icd_encoder_destroy(e);
    return s;
  };
  if (cudaMalloc((void**)&stage, (size_t)stage_n * 4) != cudaSuccess) {
    set_error("staging allocation failed");
    return fail(ICD_E_NOMEM);
  }
  const float* w = weights;
  auto take = [&](int64_t n, bool bf16, void** dst) {
    if (st != ICD_OK) return;
    st = upload(e, w, n, bf16, dst, stage);
    w += n;
  };
  take((int64_t)cfg->vocab_size * H, false, (void**)&e->word);
  take((int64_t)cfg->max_position * H, false, (void**)&e->pos);
  take((int64_t)cfg->type_vocab * H, false, (void**)&e->type);
  take(H, false, (void**)&e->eg);
  take(H, false, (void**)&e->eb);
  for (int l = 0; l < cfg->layers && st == ICD_OK; ++l) {
    LayerW* L = new LayerW();
    e->layers.push_back(L);
    take(3 * H * H, true, &L->wqkv);
    take(3 * H, false, (void**)&L->bqkv);
    take(H * H, true, &L->wo);
    take(H, false, (void**)&L->bo);
    take(H, false, (void**)&L->ln1g);
    take(H, false, (void**)&L->ln1b);
    take(I * H, true, &L->w1);
    take(I, false, (void**)&L->b1);
    take(H * I, true, &L->w2);
    take(H, false, (void**)&L->b2);
    take(H, false, (void**)&L->ln2g);
    take(H, false, (void**)&L->ln2b);
    if (st != ICD_OK) break;
    st = gemm_make_map_b(L->m_qkv, L->wqkv, 3 * H, (int)H);
    if (st == ICD_OK) st = gemm_make_map_b(L->m_o, L->wo, H, (int)H);
    if (st == ICD_OK) st = gemm_make_map_b(L->m_1, L->w1, I, (int)H);
    if (st == ICD_OK) st = gemm_make_map_b(L->m_2, L->w2, H, (int)I);
  }
  if (st != ICD_OK) return fail(st);
  cudaFree(stage);
  stage = nullptr;
  *out = e;
  return ICD_OK;
}

int icd_encoder_destroy(icd_encoder* e) {
  if (!e) return ICD_OK;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (void* p : e->allocs) cudaFree(p);
  for (auto* L : e->layers) delete L;
  void* bufs[] = {e->h, e->h1, e->t, e->ctx, e->qkv, e->f, e->ids, e->lens, e->out_stage};
  for (void* p : bufs)
    if (p) cudaFree(p);
  delete e;
  return ICD_OK;
}

int icd_encoder_reserve(icd_encoder* e, int max_tokens) {
  ICD_CHECK_ARG(e != nullptr && max_tokens > 0, "bad argument");
  ICD_CUDA(cudaSetDevice(e->device));
  return reserve_tokens(e, max_tokens);
}

int icd_encoder_forward(icd_encoder* e, const int32_t* ids, const int32_t* lens, int B, int S, void* out,
                        int out_dtype, void* stream, int sync) {
  ICD_CHECK_ARG(e != nullptr, "encoder is null");
  ICD_CHECK_ARG(B >= 0 && S >= 1 && S <= 128, "need 1 <= S <= 128");
  ICD_CHECK_ARG(S <= e->cfg.max_position, "S exceeds max_position");
  const int out_flags = out_dtype;
  out_dtype &= 0xff;
  ICD_CHECK_ARG(out_dtype == ICD_F32 || out_dtype == ICD_BF16, "unknown out dtype");
  if (B == 0) return ICD_OK;
  ICD_CHECK_ARG(ids && lens && out, "null buffer");
  ICD_CHECK_ARG((int64_t)B * S <= 0x7fffffffLL / 4, "batch too large");
  ICD_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int M = B * S;
  ICD_TRY(reserve_tokens(e, M));
  const int H = e->cfg.hidden, I = e->cfg.intermediate;
  const float eps = e->cfg.ln_eps;

  // token ids / lengths -> device
  const int32_t* d_ids = ids;
  const int32_t* d_lens = lens;
  if (!is_device_ptr(ids)) {
    if (M > e->ids_cap) {
      if (e->ids) cudaFree(e->ids);
      e->ids = nullptr;
      ICD_CUDA(cudaMalloc((void**)&e->ids, (size_t)M * 4));
      e->ids_cap = M;
    }
    ICD_CUDA(cudaMemcpyAsync(e->ids, ids, (size_t)M * 4, cudaMemcpyHostToDevice, st));
    d_ids = e->ids;
  }
  if (!is_device_ptr(lens)) {
    if (B > e->lens_cap) {
      if (e->lens) cudaFree(e->lens);
      e->lens = nullptr;
      ICD_CUDA(cudaMalloc((void**)&e->lens, (size_t)B * 4));
      e->lens_cap = B;
    }
    ICD_CUDA(cudaMemcpyAsync(e->lens, lens, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    d_lens = e->lens;
  }
  const bool host_out = !is_device_ptr(out);
  void* d_out = out;
  if (host_out) {
    const size_t need = (size_t)B * H * 4;
    if (need > e->out_cap) {
      if (e->out_stage) cudaFree(e->out_stage);
      e->out_stage = nullptr;
      ICD_CUDA(cudaMalloc(&e->out_stage, need));
      e->out_cap = need;
    }
    d_out = e->out_stage;
  }

  static const bool cuda_core_attention = getenv("ICD_ATTN_CUDA_CORE") != nullptr;  // A/B timing only
  ICD_TRY(launch_embed_ln(d_ids, M, S, e->word, e->pos, e->type, e->eg, e->eb, eps, e->h, st));
  for (size_t l = 0; l < e->layers.size(); ++l) {
    LayerW* L = e->layers[l];
    GemmArgs g{};
    // qkv = h Wqkv^T + b
    g = GemmArgs{e->m_h, L->m_qkv, L->bqkv, nullptr, e->m_qkv_out, M, 3 * H, H, EPI_BIAS};
    ICD_TRY(launch_gemm_tc(g, st));
    if (cuda_core_attention)
      ICD_TRY(launch_attention(e->qkv, d_lens, B, S, e->ctx, st));
    else
      ICD_TRY(launch_attention_tc(e->m_qkv, d_lens, B, S, e->ctx, st));
    // t = ctx Wo^T + bo + h ; h1 = LN(t)
    g = GemmArgs{e->m_ctx, L->m_o, L->bo, e->m_h, e->m_t, M, H, H, EPI_BIAS_RESIDUAL};
    ICD_TRY(launch_gemm_tc(g, st));
    ICD_TRY(launch_layernorm(e->t, M, L->ln1g, L->ln1b, eps, e->h1, st));
    // f = gelu(h1 W1^T + b1)
    g = GemmArgs{e->m_h1, L->m_1, L->b1, nullptr, e->m_f, M, I, H, EPI_BIAS_GELU};
    ICD_TRY(launch_gemm_tc(g, st));
    // t = f W2^T + b2 + h1 ; h = LN(t)
    g = GemmArgs{e->m_f, L->m_2, L->b2, e->m_h1, e->m_t, M, H, I, EPI_BIAS_RESIDUAL};
    ICD_TRY(launch_gemm_tc(g, st));
    ICD_TRY(launch_layernorm(e->t, M, L->ln2g, L->ln2b, eps, e->h, st));
  }
  ICD_TRY(launch_pool_normalise(e->h, d_lens, B, S, d_out, out_flags, st));
  e->last_M = M;
  if (host_out) {
    ICD_CUDA(cudaMemcpyAsync(out, d_out, (size_t)B * H * (out_dtype == ICD_F32 ? 4 : 2), cudaMemcpyDeviceToHost, st));
  }
  if (sync || host_out || !is_device_ptr(ids) || !is_device_ptr(lens)) ICD_CUDA(cudaStreamSynchronize(st));
  return ICD_OK;
}

int icd_encoder_read_hidden(icd_encoder* e, int layer_unused, float* out, int64_t count) {
  (void)layer_unused;
  ICD_CHECK_ARG(e && out, "null argument");
  ICD_CHECK_ARG(count == (int64_t)e->last_M * e->cfg.hidden, "count must be tokens*hidden of the last forward");
  ICD_CUDA(cudaSetDevice(e->device));
  float* tmp = nullptr;
  ICD_CUDA(cudaMalloc((void**)&tmp, (size_t)count * 4));
  int st = launch_bf16_to_f32(e->h, tmp, count, 0);
  if (st == ICD_OK) {
    cudaError_t err = cudaMemcpy(out, tmp, (size_t)count * 4, cudaMemcpyDefault);
    if (err != cudaSuccess) {
      set_error("read_hidden: %s", cudaGetErrorString(err));
      st = ICD_E_CUDA;
    }
  }
  cudaFree(tmp);
  return st;
}

}  // extern "C"
