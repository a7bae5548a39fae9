"""Shared fixtures of the NER (token-classification) tests: a seeded synthetic BertForTokenClassification and
the transformers pipeline the reference builds (medical_ner_service.py:76-90) as the oracle."""
import os

import numpy as np
import torch

from oracle import encoder as oenc

LABELS = ["O", "B-DiseaseNameOrComprehensiveCertificate", "I-DiseaseNameOrComprehensiveCertificate", "B-Symptom",
          "I-Symptom", "B-BodyParts", "I-BodyParts", "B-Drug", "I-Drug"]
TEXTS = ["急性胃肠炎伴发热三天", "患者右下腹疼痛,考虑急性阑尾炎", "2型糖尿病 高血压病3级", "COVID-19 感染后咳嗽", "头痛",
         "慢性阻塞性肺疾病急性加重期,给予沙丁胺醇雾化吸入治疗", "左侧股骨颈骨折术后"]


def build(tmpdir, hidden=768, layers=2, heads=12, inter=3072, seed=5, head_scale=0.6):
    """Writes an HF model dir (config.json, model.safetensors, vocab.txt) and returns (dir, hf_model, tokenizer)."""
    from transformers import BertConfig, BertForTokenClassification
    from safetensors.torch import save_file
    vocab = oenc.make_vocab(TEXTS, size=3000)
    torch.manual_seed(seed)
    cfg = BertConfig(vocab_size=len(vocab), hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                     intermediate_size=inter, hidden_act="gelu", max_position_embeddings=512, type_vocab_size=2,
                     layer_norm_eps=1e-12, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                     num_labels=len(LABELS), id2label={i: l for i, l in enumerate(LABELS)},
                     label2id={l: i for i, l in enumerate(LABELS)})
    model = BertForTokenClassification(cfg).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("LayerNorm.weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("LayerNorm.bias") or (name.endswith(".bias") and "classifier" not in name):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif name.startswith("classifier.weight"):
                p.copy_(head_scale * torch.randn(p.shape, generator=g))   # decisive logits: few argmax near-ties
            elif name.startswith("classifier.bias"):
                p.copy_(0.3 * torch.randn(p.shape, generator=g))
    os.makedirs(tmpdir, exist_ok=True)
    cfg.save_pretrained(tmpdir)
    save_file({k: v.contiguous() for k, v in model.state_dict().items()}, os.path.join(tmpdir, "model.safetensors"))
    with open(os.path.join(tmpdir, "vocab.txt"), "w", encoding="utf-8") as fh:
        fh.write("\n".join(vocab) + "\n")
    tok = oenc.make_tokenizer(os.path.join(tmpdir, "vocab.txt"))
    return tmpdir, model, tok


def hf_pipeline(model, tok):
    from transformers import pipeline
    return pipeline("ner", model=model, tokenizer=tok, aggregation_strategy="simple", device=-1)


@torch.no_grad()
def hf_logits(model, tok, text, max_length=128):
    enc = tok(text, return_tensors="pt", truncation=True, max_length=max_length)
    return model(**enc).logits[0].float().numpy()


def same_groups(a, b, score_tol):
    assert [(g["entity_group"], g["word"], g["start"], g["end"]) for g in a] == \
           [(g["entity_group"], g["word"], g["start"], g["end"]) for g in b], (a, b)
    for x, y in zip(a, b):
        assert abs(float(x["score"]) - float(y["score"])) <= score_tol, (x, y)
