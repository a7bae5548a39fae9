"""EncoderEngine -- the object EmbeddingService holds where the reference holds a
``SentenceTransformer`` (/root/reference/services/embedding_service.py:61): same members the
reference uses -- ``encode(str | list, batch_size=, show_progress_bar=, normalize_embeddings=)``,
``get_sentence_embedding_dimension()``, ``max_seq_length`` -- with the BERT forward, pooling and
normalisation running in libicdrag.so on the GPU.  Only tokenisation stays on the host.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Union

import numpy as np

from .. import _native as N
from . import weights as W

MAX_S = 128  # kernel limit == sentence-transformers max_seq_length of text2vec-base-chinese


class EncoderEngine:
    def __init__(self, model_name_or_path: Optional[str] = None, device: Union[str, int, None] = None, *,
                 cfg: Optional[N.BertCfg] = None, blob: Optional[np.ndarray] = None, tokenizer=None,
                 max_seq_length: int = 128, max_tokens: int = 4096 * 64):
        N.require_gpu()
        self.device_index = _device_index(device)
        if blob is None:
            path = W.resolve_model_dir(model_name_or_path)
            cfg, blob, meta = W.load_model_dir(path)
            max_seq_length = meta["max_seq_length"]
            if tokenizer is None:
                tokenizer = load_tokenizer(path)
        if tokenizer is None:
            raise ValueError("a tokenizer is required")
        self.cfg = cfg
        self.tokenizer = tokenizer
        self.max_seq_length = min(int(max_seq_length), MAX_S, cfg.max_position)
        self.max_tokens = int(max_tokens)
        self._h = C.c_void_p()
        blob = np.ascontiguousarray(blob, np.float32)
        N.check(N.lib().icd_encoder_create(N.buf_ptr(blob), blob.size, C.byref(cfg), self.device_index,
                                           C.byref(self._h)), "icd_encoder_create")

    # ---------------------------------------------------------------- SentenceTransformer surface
    def get_sentence_embedding_dimension(self) -> int:
        return int(self.cfg.hidden)

    def encode(self, sentences, batch_size: int = 32, show_progress_bar=None, normalize_embeddings: bool = False,
               convert_to_numpy: bool = True, **_ignored):
        single = isinstance(sentences, str)
        items: List[str] = [sentences] if single else list(sentences)
        out = np.zeros((len(items), self.cfg.hidden), np.float32)
        if items:
            ids = self.tokenize(items)
            self._encode_ids(ids, out, normalize_embeddings)
        return out[0] if single else out

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().icd_encoder_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- host side
    def tokenize(self, texts: Sequence[str]) -> List[List[int]]:
        enc = self.tokenizer(list(texts), padding=False, truncation=True, max_length=self.max_seq_length,
                             add_special_tokens=True, return_attention_mask=False, return_token_type_ids=False)
        return enc["input_ids"]

    def _encode_ids(self, ids: List[List[int]], out: np.ndarray, normalise: bool) -> None:
        """Length-bucketed batches: sort by token count, cut where B*S would exceed max_tokens."""
        order = sorted(range(len(ids)), key=lambda j: -len(ids[j]))
        lo = 0
        while lo < len(order):
            S = max(1, len(ids[order[lo]]))
            B = max(1, min(len(order) - lo, self.max_tokens // S))
            idx = order[lo:lo + B]
            mat = np.zeros((B, S), np.int32)
            lens = np.zeros((B,), np.int32)
            for r, j in enumerate(idx):
                row = ids[j]
                mat[r, :len(row)] = row
                lens[r] = len(row)
            res = self.forward_ids(mat, lens, normalise=normalise)
            out[idx] = res
            lo += B

    def forward_ids(self, ids, lens, normalise: bool = True, out=None, stream: int = 0, sync: bool = True):
        """ids [B,S] int32, lens [B] int32 (numpy host or torch host/device) -> [B,hidden] float32."""
        B, S = int(ids.shape[0]), int(ids.shape[1])
        if out is None:
            if N._is_torch(ids) and ids.is_cuda:
                import torch
                out = torch.empty((B, self.cfg.hidden), dtype=torch.float32, device=ids.device)
            else:
                out = np.empty((B, self.cfg.hidden), np.float32)
        dt = N.vec_dtype(out) | (0 if normalise else 0x100)
        N.check(N.lib().icd_encoder_forward(self._h, N.buf_ptr(ids), N.buf_ptr(lens), B, S, N.buf_ptr(out), dt,
                                            C.c_void_p(stream), 1 if sync else 0), "icd_encoder_forward")
        return out

    def set_token_head(self, weight: np.ndarray, bias: np.ndarray) -> None:
        """classifier.weight [labels, hidden] / classifier.bias [labels] of a BertForTokenClassification."""
        weight = np.ascontiguousarray(weight, np.float32)
        bias = np.ascontiguousarray(bias, np.float32)
        if weight.ndim != 2 or weight.shape[1] != self.cfg.hidden or bias.shape != (weight.shape[0],):
            raise ValueError("token head must be weight [labels, hidden] and bias [labels]")
        N.check(N.lib().icd_encoder_set_token_head(self._h, N.buf_ptr(weight), N.buf_ptr(bias), int(weight.shape[0])),
                "icd_encoder_set_token_head")
        self.num_labels = int(weight.shape[0])

    def token_logits(self, ids, lens, out=None, stream: int = 0, sync: bool = True):
        """ids [B,S] int32, lens [B] int32 -> per-token logits [B,S,labels] float32 (set_token_head first)."""
        B, S = int(ids.shape[0]), int(ids.shape[1])
        L = getattr(self, "num_labels", 0)
        if L <= 0:
            raise N.NativeError("no token head set on this encoder")
        if out is None:
            if N._is_torch(ids) and ids.is_cuda:
                import torch
                out = torch.empty((B, S, L), dtype=torch.float32, device=ids.device)
            else:
                out = np.empty((B, S, L), np.float32)
        N.check(N.lib().icd_encoder_token_logits(self._h, N.buf_ptr(ids), N.buf_ptr(lens), B, S, N.buf_ptr(out),
                                                 C.c_void_p(stream), 1 if sync else 0), "icd_encoder_token_logits")
        return out

    def read_hidden(self, tokens: int) -> np.ndarray:
        out = np.empty((tokens, self.cfg.hidden), np.float32)
        N.check(N.lib().icd_encoder_read_hidden(self._h, 0, N.buf_ptr(out), out.size), "icd_encoder_read_hidden")
        return out


def _device_index(device) -> int:
    if device is None:
        return 0
    if isinstance(device, int):
        return device
    s = str(device)
    if s in ("cuda", "auto"):
        return 0
    if s.startswith("cuda:"):
        return int(s.split(":")[1])
    raise N.NativeError(f"EncoderEngine runs on CUDA devices only (got device={device!r}); there is no CPU fallback")


def bert_tokenizer_from_vocab(vocab_path: str, do_lower_case: bool = True):
    """BertTokenizerFast over a vocab.txt.  transformers >= 5 takes the vocabulary as `vocab=` (a token -> id dict)
    and silently ignores `vocab_file=` (leaving a 5-token vocabulary that maps every character to [UNK]);
    transformers 4 takes `vocab_file=`.  Either way the result is checked against the file."""
    from transformers import BertTokenizerFast
    with open(vocab_path, encoding="utf-8") as fh:
        tokens = [line.rstrip("\n") for line in fh]
    while tokens and tokens[-1] == "":
        tokens.pop()
    vocab = {t: i for i, t in enumerate(tokens)}
    tok = None
    for kwargs in ({"vocab": vocab}, {"vocab_file": vocab_path}):
        try:
            cand = BertTokenizerFast(do_lower_case=do_lower_case, **kwargs)
        except Exception:
            continue
        if cand.vocab_size == len(vocab):
            tok = cand
            break
    if tok is None:
        raise RuntimeError(f"could not build a BERT tokenizer over {vocab_path} ({len(vocab)} entries)")
    return tok


def load_tokenizer(path: str):
    """The model directory's own tokenizer; falls back to BertTokenizerFast over vocab.txt."""
    from transformers import AutoTokenizer
    vocab_path = os.path.join(path, "vocab.txt")
    try:
        if os.path.exists(os.path.join(path, "tokenizer_config.json")) or os.path.exists(os.path.join(path, "tokenizer.json")):
            tok = AutoTokenizer.from_pretrained(path, local_files_only=True)
            if tok.vocab_size > 5 or not os.path.exists(vocab_path):   # a 5-token vocabulary = the file was ignored
                return tok
    except Exception:
        pass
    return bert_tokenizer_from_vocab(vocab_path)


# ------------------------------------------------------------------------------------------
def synthetic_engine(num_layers: int = 12, seed: int = 0, device: int = 0, vocab_size: int = 21128,
                     max_tokens: int = 4096 * 64):
    """Random-init encoder of the text2vec-base-chinese architecture (no checkpoint exists
    offline): HF-style N(0, 0.02) init from a numpy generator.  Used by bench.py and smoke()."""
    cfg = N.BertCfg(vocab_size=vocab_size, hidden=768, layers=num_layers, heads=12, intermediate=3072,
                    max_position=512, type_vocab=2, ln_eps=1e-12)
    n = int(N.lib().icd_encoder_weight_count(cfg))
    rng = np.random.default_rng(seed)
    blob = (rng.standard_normal(n, dtype=np.float32) * 0.02)
    # LayerNorm gains must sit near 1: walk the canonical order and patch them
    off = 0
    H, I = 768, 3072
    shapes = [vocab_size * H, 512 * H, 2 * H]
    off = sum(shapes)
    blob[off:off + H] = 1.0 + 0.1 * rng.standard_normal(H, dtype=np.float32); off += 2 * H
    for _ in range(num_layers):
        off += 3 * H * H + 3 * H + H * H + H
        blob[off:off + H] = 1.0 + 0.1 * rng.standard_normal(H, dtype=np.float32); off += 2 * H
        off += I * H + I + H * I + H
        blob[off:off + H] = 1.0 + 0.1 * rng.standard_normal(H, dtype=np.float32); off += 2 * H
    assert off == n
    class _NoTok:
        def __call__(self, *a, **k):
            raise RuntimeError("synthetic engine has no tokenizer; use forward_ids")
    return EncoderEngine(cfg=cfg, blob=blob, tokenizer=_NoTok(), device=device, max_tokens=max_tokens)


def smoke() -> None:
    """One tiny forward on cuda:0 against the CPU oracle (HF BertModel fp32)."""
    import torch
    from oracle import encoder as oenc
    state = oenc.synthetic_state_dict(seed=3, num_layers=2, vocab_size=1000, max_position=512)
    cfg = N.BertCfg(vocab_size=1000, hidden=768, layers=2, heads=12, intermediate=3072, max_position=512,
                    type_vocab=2, ln_eps=1e-12)
    blob = W.pack_state_dict(state, cfg)
    eng = EncoderEngine(cfg=cfg, blob=blob, tokenizer=object(), device=0, max_tokens=4096)
    rng = np.random.default_rng(0)
    B, S = 6, 24
    lens = np.array([24, 20, 13, 7, 2, 24], np.int32)
    ids = np.zeros((B, S), np.int32)
    for b in range(B):
        ids[b, :lens[b]] = rng.integers(1, 1000, size=lens[b])
    got = eng.forward_ids(ids, lens)
    from transformers import BertModel
    model = BertModel(oenc.bert_config(2, 1000, 512), add_pooling_layer=False)
    model.load_state_dict(state, strict=False)
    model.eval()
    mask = torch.from_numpy((np.arange(S)[None, :] < lens[:, None]).astype(np.int64))
    with torch.no_grad():
        h = model(input_ids=torch.from_numpy(ids.astype(np.int64)), attention_mask=mask).last_hidden_state
        m = mask.unsqueeze(-1).float()
        ref = torch.nn.functional.normalize((h * m).sum(1) / m.sum(1).clamp(min=1e-9), dim=1).numpy()
    cos = (got * ref).sum(1)
    assert cos.min() >= 0.999, cos
    eng.close()


def bench_encoder(dev, peaks, batch: int = 4096, seq: int = 64, steps: int = 10, warmup: int = 3, barrier=None):
    """BASELINE configs[2]: encoder throughput at S=64, B=4096 on synthetic ids, random-init
    weights of the text2vec-base-chinese architecture.  Returns the `encoder` object of bench.py:
    `value` with ids and outputs resident in HBM, `e2e` through icd_encoder_forward with pinned HOST ids /
    lens / output (copies inside the timed region)."""
    import torch
    eng = synthetic_engine(device=dev.index or 0, max_tokens=batch * seq)
    g = torch.Generator(device=dev).manual_seed(7)
    ids = torch.randint(1000, 21128, (batch, seq), generator=g, device=dev, dtype=torch.int32)
    ids[:, 0] = 101
    ids[:, -1] = 102
    lens = torch.full((batch,), seq, dtype=torch.int32, device=dev)
    out = torch.empty((batch, 768), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    launches0 = N.lib().icd_launch_count()
    eng.forward_ids(ids, lens, out=out, stream=stream.cuda_stream, sync=False)
    launches = int(N.lib().icd_launch_count() - launches0)
    for _ in range(warmup):
        eng.forward_ids(ids, lens, out=out, stream=stream.cuda_stream, sync=False)
    torch.cuda.synchronize(dev)
    if barrier is not None:
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        eng.forward_ids(ids, lens, out=out, stream=stream.cuda_stream, sync=False)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if barrier is not None:
        barrier()
    ms = e0.elapsed_time(e1) / steps
    # end to end: host token ids in, host embeddings out, every step
    h_ids, h_lens = ids.cpu().pin_memory(), lens.cpu().pin_memory()
    h_out = torch.empty((batch, 768), dtype=torch.float32).pin_memory()
    eng.forward_ids(h_ids, h_lens, out=h_out, stream=stream.cuda_stream, sync=True)
    e0.record(stream)
    for _ in range(steps):
        eng.forward_ids(h_ids, h_lens, out=h_out, stream=stream.cuda_stream, sync=True)
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms_e2e = e0.elapsed_time(e1) / steps
    same = bool(torch.allclose(h_out, out.cpu(), atol=1e-6))
    flops = batch * seq * 12 * (2 * (4 * 768 * 768 + 2 * 768 * 3072) + 4 * seq * 768)
    tf = flops / (ms * 1e-3) / 1e12
    norm = float(out.norm(dim=1).mean())
    eng.close()
    return {"metric": "text2vec sentences/sec", "value": batch / (ms * 1e-3), "unit": "sentences/s",
            "ms_per_batch": ms, "steps": steps, "warmup": warmup, "flops_per_batch": flops, "batch": batch,
            "seq_len": seq, "layers": 12, "dtype": "bf16",
            "tflops": tf, "frac_of_bf16_sustained": tf / peaks["bf16_tflops_sustained"], "mean_norm": norm,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": tf / peaks["bf16_tflops_sustained"], "traffic": None,
                         "kernel": "gemm_tc_kernel (4 of the 5 launches per layer)", "peak_source": peaks["source"] + " (sustained)"},
            "e2e": {"value": batch / (ms_e2e * 1e-3), "unit": "sentences/s", "h2d_bytes_per_step": batch * seq * 4 + batch * 4,
                    "d2h_bytes_per_step": batch * 768 * 4, "host_equals_device": same},
            "gpu_launches_per_batch": launches,
            "data": "synthetic ids, random-init weights"}
