"""Oracle: the arithmetic inside SentenceTransformer.encode (test infrastructure only).

Third-party engine restated: sentence-transformers>=4.1.0 (/root/reference/requirements.txt:7)
loading shibing624/text2vec-base-chinese (/root/reference/env.example:16): a BertModel
(12 layers, 12 heads x 64, FFN 3072, erf-GELU, LayerNorm eps 1e-12, learned absolute
positions, post-LN) -> masked mean over tokens -> L2 normalise; `encode` sorts inputs by
descending text length, runs slices of `batch_size` padded to the longest member and
truncated at max_seq_length = 128, and returns float32.  Call sites:
/root/reference/services/embedding_service.py:81 (encode_single), :97-102 (encode_batch,
batch_size 32), :120 (encode_query).  Neither the package nor the weights are available
offline, so this oracle runs transformers.BertModel in fp32 on seeded synthetic weights and
a synthetic vocab: parity unpinned for the engine arithmetic.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Sequence

import numpy as np
import torch

MAX_SEQ_LENGTH = 128
SPECIALS = {"[PAD]": 0, "[UNK]": 100, "[CLS]": 101, "[SEP]": 102, "[MASK]": 103}


def bert_config(num_layers: int = 12, vocab_size: int = 21128, max_position: int = 512):
    from transformers import BertConfig
    return BertConfig(vocab_size=vocab_size, hidden_size=768, num_hidden_layers=num_layers,
                      num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                      max_position_embeddings=max_position, type_vocab_size=2,
                      layer_norm_eps=1e-12, hidden_dropout_prob=0.0,
                      attention_probs_dropout_prob=0.0)


def synthetic_state_dict(seed: int = 0, num_layers: int = 12, vocab_size: int = 21128,
                         max_position: int = 512) -> Dict[str, torch.Tensor]:
    """Seeded HF-named fp32 weights: N(0, 0.02) matrices (HF init), with LayerNorm gains,
    LayerNorm offsets and biases perturbed so every fused epilogue is exercised."""
    from transformers import BertModel
    torch.manual_seed(seed)
    model = BertModel(bert_config(num_layers, vocab_size, max_position), add_pooling_layer=False)
    g = torch.Generator().manual_seed(seed + 1)
    sd = {}
    for name, p in model.state_dict().items():
        t = p.detach().clone().float()
        if name.endswith("LayerNorm.weight"):
            t = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif name.endswith("LayerNorm.bias") or name.endswith(".bias"):
            t = 0.05 * torch.randn(t.shape, generator=g)
        sd[name] = t.contiguous()
    return sd


def save_hf_dir(path: str, state: Dict[str, torch.Tensor], vocab: Sequence[str], num_layers: int,
                max_position: int = 512) -> str:
    """Write config.json + model.safetensors + vocab.txt, the layout of an HF model dir."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    cfg = bert_config(num_layers, len(vocab), max_position).to_dict()
    cfg["model_type"] = "bert"
    with open(os.path.join(path, "config.json"), "w") as fh:
        json.dump(cfg, fh)
    save_file({k: v.contiguous() for k, v in state.items()}, os.path.join(path, "model.safetensors"))
    with open(os.path.join(path, "vocab.txt"), "w", encoding="utf-8") as fh:
        fh.write("\n".join(vocab) + "\n")
    return path


def make_vocab(texts: Sequence[str], size: int = 21128) -> List[str]:
    """Synthetic vocab.txt: bert-base-chinese special ids, every character of `texts`,
    printable ASCII, a few '##' pieces, padded with [unusedN] to `size`."""
    vocab = ["[unused%d]" % i for i in range(size)]
    for tok, idx in SPECIALS.items():
        vocab[idx] = tok
    vocab[0] = "[PAD]"
    chars = set()
    for t in texts:
        chars.update(t.lower())
    chars.update(chr(c) for c in range(33, 127))
    chars = sorted(c for c in chars if not c.isspace())
    pieces = list(chars) + ["##" + c for c in "abcdefghijklmnopqrstuvwxyz0123456789"] + \
             ["icd", "query", "passage", "##cd", "##ery", "##ssage"]
    seen = set(vocab)
    slot = 104
    for p in pieces:
        if p in seen:
            continue
        while vocab[slot] in SPECIALS:
            slot += 1
        vocab[slot] = p
        seen.add(p)
        slot += 1
    assert slot <= size
    return vocab


def make_tokenizer(vocab_path: str):
    """BertTokenizerFast(do_lower_case=True) over vocab.txt.  transformers >= 5 wants `vocab=` (a dict) and ignores
    `vocab_file=`, which used to leave every character mapped to [UNK]; the size check catches that."""
    from transformers import BertTokenizerFast
    with open(vocab_path, encoding="utf-8") as fh:
        tokens = [line.rstrip("\n") for line in fh]
    while tokens and tokens[-1] == "":
        tokens.pop()
    vocab = {t: i for i, t in enumerate(tokens)}
    for kwargs in ({"vocab": vocab}, {"vocab_file": vocab_path}):
        try:
            tok = BertTokenizerFast(do_lower_case=True, **kwargs)
        except Exception:
            continue
        if tok.vocab_size == len(vocab):
            return tok
    raise RuntimeError("could not build the oracle tokenizer over " + vocab_path)


class OracleEncoder:
    """fp32 CPU restatement of SentenceTransformer(...).encode(normalize_embeddings=True)."""

    def __init__(self, state: Dict[str, torch.Tensor], vocab_path: str, num_layers: int,
                 max_position: int = 512):
        from transformers import BertModel
        vocab_size = state["embeddings.word_embeddings.weight"].shape[0]
        self.model = BertModel(bert_config(num_layers, vocab_size, max_position), add_pooling_layer=False)
        missing = self.model.load_state_dict(state, strict=False)
        assert not [m for m in missing.missing_keys if "position_ids" not in m], missing
        self.model.eval()
        self.tok = make_tokenizer(vocab_path)
        self.max_seq_length = MAX_SEQ_LENGTH

    def tokenize(self, texts: Sequence[str]):
        enc = self.tok(list(texts), padding=True, truncation=True, max_length=self.max_seq_length,
                       return_tensors="pt")
        return enc["input_ids"], enc["attention_mask"]

    @torch.no_grad()
    def forward_ids(self, ids: torch.Tensor, mask: torch.Tensor) -> np.ndarray:
        h = self.model(input_ids=ids, attention_mask=mask).last_hidden_state.float()
        m = mask.unsqueeze(-1).float()
        pooled = (h * m).sum(1) / m.sum(1).clamp(min=1e-9)
        return torch.nn.functional.normalize(pooled, p=2, dim=1).numpy().astype(np.float32)

    def encode(self, texts, batch_size: int = 32) -> np.ndarray:
        single = isinstance(texts, str)
        items = [texts] if single else list(texts)
        order = sorted(range(len(items)), key=lambda j: -len(items[j]))
        out = np.zeros((len(items), 768), np.float32)
        for lo in range(0, len(items), batch_size):
            idx = order[lo:lo + batch_size]
            ids, mask = self.tokenize([items[j] for j in idx])
            out[idx] = self.forward_ids(ids, mask)
        return out[0] if single else out
