"""TokenClassifierEngine -- the object MedicalNERService holds where the reference holds
``pipeline("ner", model=AutoModelForTokenClassification..., aggregation_strategy="simple")``
(/root/reference/services/medical_ner_service.py:76-90, called as ``self.ner_pipeline(text)`` at :182).

Same call shape and result layout -- a list of ``{"entity_group", "score", "word", "start", "end"}`` per
text -- with the BERT forward and the classifier head running in libicdrag.so on the GPU
(icd_encoder_set_token_head / icd_encoder_token_logits).  Tokenisation, the softmax over <= 64 labels and
the "simple" grouping of adjacent tokens stay on the host, as they do inside the transformers pipeline
(third party, transformers TokenClassificationPipeline: postprocess / gather_pre_entities / aggregate /
group_entities; restated here, checked against the pipeline itself in tests/test_ner_*.py).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from .. import _native as N
from . import weights as W
from .encoder import EncoderEngine, load_tokenizer


def _get_tag(entity_name: str):
    if entity_name.startswith("B-"):
        return "B", entity_name[2:]
    if entity_name.startswith("I-"):
        return "I", entity_name[2:]
    return "I", entity_name  # not in B-/I- format: continuation


def softmax_rows(logits: np.ndarray) -> np.ndarray:
    maxes = np.max(logits, axis=-1, keepdims=True)
    e = np.exp(logits - maxes)
    return e / e.sum(axis=-1, keepdims=True)


def aggregate_simple(tokenizer, id2label: Dict[int, str], sentence: str, input_ids: Sequence[int],
                     scores: np.ndarray, offsets: Sequence[Sequence[int]], special_mask: Sequence[int],
                     ignore_labels=("O",)) -> List[dict]:
    """aggregation_strategy="simple": per token argmax label, adjacent tokens with the same tag (and no new
    "B-") merge into one group whose score is the mean of its tokens' scores."""
    entities = []
    for idx in range(len(input_ids)):
        if special_mask[idx]:
            continue
        word = tokenizer.convert_ids_to_tokens(int(input_ids[idx]))
        start, end = int(offsets[idx][0]), int(offsets[idx][1])
        if int(input_ids[idx]) == tokenizer.unk_token_id:
            word = sentence[start:end]
        k = int(scores[idx].argmax())
        entities.append({"entity": id2label[k], "score": scores[idx][k], "index": idx, "word": word,
                         "start": start, "end": end})
    groups, cur = [], []

    def flush():
        tag = cur[0]["entity"].split("-", 1)[-1]
        groups.append({"entity_group": tag, "score": np.mean(np.nanmean([e["score"] for e in cur])),
                       "word": tokenizer.convert_tokens_to_string([e["word"] for e in cur]),
                       "start": cur[0]["start"], "end": cur[-1]["end"]})

    for ent in entities:
        if not cur:
            cur.append(ent)
            continue
        bi, tag = _get_tag(ent["entity"])
        _, last_tag = _get_tag(cur[-1]["entity"])
        if tag == last_tag and bi != "B":
            cur.append(ent)
        else:
            flush()
            cur = [ent]
    if cur:
        flush()
    return [g for g in groups if g["entity_group"] not in ignore_labels]


class TokenClassifierEngine:
    """Callable like the reference's ``ner_pipeline``: ``engine(text)`` -> list of entity groups,
    ``engine([t1, t2, ...])`` -> list of lists (one GPU batch per length bucket)."""

    def __init__(self, model_name_or_path: Optional[str] = None, device: Union[str, int, None] = None, *,
                 encoder: Optional[EncoderEngine] = None, head_weight: Optional[np.ndarray] = None,
                 head_bias: Optional[np.ndarray] = None, id2label: Optional[Dict[int, str]] = None,
                 tokenizer=None, max_tokens: int = 1024 * 128):
        if encoder is None:
            path = W.resolve_model_dir(model_name_or_path)
            cfg, blob, _meta = W.load_model_dir(path)
            state = W.load_state(path)
            head_weight = np.asarray(state["classifier.weight"], np.float32)
            head_bias = np.asarray(state["classifier.bias"], np.float32)
            with open(os.path.join(path, "config.json"), encoding="utf-8") as fh:
                id2label = {int(k): v for k, v in json.load(fh).get("id2label", {}).items()}
            tokenizer = tokenizer or load_tokenizer(path)
            encoder = EncoderEngine(cfg=cfg, blob=blob, tokenizer=tokenizer, device=device,
                                    max_seq_length=min(512, cfg.max_position), max_tokens=max_tokens)
        if head_weight is None or head_bias is None:
            raise ValueError("a classifier head (weight, bias) is required")
        self.encoder = encoder
        self.tokenizer = tokenizer or encoder.tokenizer
        self.encoder.set_token_head(head_weight, head_bias)
        n = int(np.asarray(head_weight).shape[0])
        self.id2label = id2label or {i: f"LABEL_{i}" for i in range(n)}
        if len(self.id2label) != n:
            raise ValueError("id2label does not match the classifier head")
        self.max_seq_length = self.encoder.max_seq_length   # kernel limit 128 tokens per sequence

    def __call__(self, inputs, **_ignored):
        single = isinstance(inputs, str)
        texts = [inputs] if single else list(inputs)
        out = self.extract(texts)
        return out[0] if single else out

    def extract(self, texts: Sequence[str]) -> List[List[dict]]:
        if not texts:
            return []
        enc = self.tokenizer(list(texts), padding=False, truncation=True, max_length=self.max_seq_length,
                             return_special_tokens_mask=True, return_offsets_mapping=True,
                             return_attention_mask=False, return_token_type_ids=False)
        ids = enc["input_ids"]
        results: List[Optional[List[dict]]] = [None] * len(texts)
        order = sorted(range(len(ids)), key=lambda j: -len(ids[j]))
        lo = 0
        while lo < len(order):
            S = max(1, len(ids[order[lo]]))
            B = max(1, min(len(order) - lo, self.encoder.max_tokens // S))
            idx = order[lo:lo + B]
            mat = np.zeros((B, S), np.int32)
            lens = np.zeros((B,), np.int32)
            for r, j in enumerate(idx):
                mat[r, :len(ids[j])] = ids[j]
                lens[r] = len(ids[j])
            logits = self.encoder.token_logits(mat, lens)
            for r, j in enumerate(idx):
                n = int(lens[r])
                scores = softmax_rows(logits[r, :n].astype(np.float32))
                results[j] = aggregate_simple(self.tokenizer, self.id2label, texts[j], ids[j], scores,
                                              enc["offset_mapping"][j], enc["special_tokens_mask"][j])
            lo += B
        return results  # type: ignore[return-value]

    def close(self) -> None:
        self.encoder.close()
