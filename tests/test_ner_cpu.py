"""Host logic of the NER drop-in (engine/token_classifier.py) against the transformers pipeline the reference
constructs (medical_ner_service.py:76-90): the "simple" aggregation restated on the host must reproduce the
pipeline's entity groups when fed the pipeline model's own logits.  No GPU: the engine's GPU encoder is
replaced by a stub that returns the HF model's CPU logits."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ner_common as nc  # noqa: E402


class _StubEncoder:
    """token_logits computed by the HF model on the CPU (what the GPU path must match)."""
    max_tokens = 4096
    max_seq_length = 128

    def __init__(self, model, tok):
        self.model, self.tokenizer = model, tok

    def set_token_head(self, w, b):
        self.num_labels = int(np.asarray(w).shape[0])

    def token_logits(self, ids, lens):
        import torch
        mask = (np.arange(ids.shape[1])[None, :] < lens[:, None]).astype(np.int64)
        with torch.no_grad():
            out = self.model(input_ids=torch.from_numpy(ids.astype(np.int64)), attention_mask=torch.from_numpy(mask))
        return out.logits.float().numpy()

    def close(self):
        pass


@pytest.fixture(scope="module")
def tiny(tmp_path_factory):
    d, model, tok = nc.build(str(tmp_path_factory.mktemp("ner_tiny")), hidden=64, layers=1, heads=4, inter=128)
    return model, tok


def test_simple_aggregation_matches_transformers_pipeline(tiny):
    model, tok = tiny
    TC = importlib.import_module("rag-project-icd10_b200.engine.token_classifier")
    eng = TC.TokenClassifierEngine(encoder=_StubEncoder(model, tok), head_weight=np.zeros((len(nc.LABELS), 64), np.float32),
                                   head_bias=np.zeros(len(nc.LABELS), np.float32),
                                   id2label=dict(enumerate(nc.LABELS)), tokenizer=tok)
    pipe = nc.hf_pipeline(model, tok)
    got_all = eng(nc.TEXTS)                       # one batch, length-bucketed
    n_groups = 0
    for text, got in zip(nc.TEXTS, got_all):
        ref = pipe(text)
        nc.same_groups(got, ref, score_tol=1e-5)
        assert eng(text) == got or nc.same_groups(eng(text), got, 1e-6) is None   # single-string call shape
        n_groups += len(ref)
    assert n_groups > 10      # the synthetic head must actually produce entities
    assert eng([]) == []


def test_softmax_and_tag_rules():
    TC = importlib.import_module("rag-project-icd10_b200.engine.token_classifier")
    x = np.array([[1.0, 2.0, 3.0], [1000.0, 1000.0, 999.0]], np.float32)
    p = TC.softmax_rows(x)
    np.testing.assert_allclose(p.sum(1), 1.0, rtol=1e-6)
    assert np.isfinite(p).all() and p[0].argmax() == 2
    assert TC._get_tag("B-Symptom") == ("B", "Symptom") and TC._get_tag("I-Drug") == ("I", "Drug")
    assert TC._get_tag("O") == ("I", "O")


def test_grouping_rules_match_pipeline_on_crafted_label_sequences(tiny):
    """B-/I- boundary rules of the "simple" strategy on hand-made label sequences (two B- in a row stay apart, an I-
    of another type starts a new group, O runs are dropped, [UNK] takes its word from the text), fed through the
    transformers pipeline's own aggregate() and through the restatement."""
    from transformers.pipelines.token_classification import AggregationStrategy
    model, tok = tiny
    TC = importlib.import_module("rag-project-icd10_b200.engine.token_classifier")
    pipe = nc.hf_pipeline(model, tok)
    id2label = dict(enumerate(nc.LABELS))
    text = "急性胃肠炎伴发热☃三天"            # the snowman is not in the vocabulary -> [UNK]
    enc = tok(text, return_special_tokens_mask=True, return_offsets_mapping=True)
    ids, offsets, special = enc["input_ids"], enc["offset_mapping"], enc["special_tokens_mask"]
    assert tok.unk_token_id in ids
    n = len(ids)
    rng = np.random.default_rng(0)
    sequences = [
        [1, 2, 2, 1, 2, 0, 3, 4, 0, 0, 7],      # B I I B I O B I O O B
        [2, 2, 4, 4, 6, 0, 0, 1, 1, 1, 8],      # I-runs of different types, B B B
        [0] * 11, [3] * 11, [5, 6, 5, 6, 5, 6, 5, 6, 5, 6, 5],
    ] + [rng.integers(0, len(nc.LABELS), size=11).tolist() for _ in range(20)]
    for labels in sequences:
        labels = (labels * 3)[:n - 2]
        scores = np.full((n, len(nc.LABELS)), 0.01, np.float32)
        conf = rng.uniform(0.4, 0.9, size=n).astype(np.float32)
        for t, lab in enumerate(labels):
            scores[t + 1, lab] = conf[t + 1]                      # position 0 / n-1 are [CLS] / [SEP]
        pre = pipe.gather_pre_entities(text, np.asarray(ids), scores, offsets, np.asarray(special), AggregationStrategy.SIMPLE)
        ref = [g for g in pipe.aggregate(pre, AggregationStrategy.SIMPLE) if g["entity_group"] != "O"]
        got = TC.aggregate_simple(tok, id2label, text, ids, scores, offsets, special)
        nc.same_groups(got, ref, score_tol=1e-7)


def test_long_texts_go_through_overlapping_windows_like_pipeline_stride(tiny):
    """A text longer than one 128-token window: the engine cuts it into overlapping windows and merges their entities
    with the pipeline's own overlap rule -- the algorithm of pipeline(..., stride=n) with a 128-token model limit."""
    from transformers import pipeline
    model, tok = tiny
    TC = importlib.import_module("rag-project-icd10_b200.engine.token_classifier")
    eng = TC.TokenClassifierEngine(encoder=_StubEncoder(model, tok), head_weight=np.zeros((len(nc.LABELS), 64), np.float32),
                                   head_bias=np.zeros(len(nc.LABELS), np.float32), id2label=dict(enumerate(nc.LABELS)),
                                   tokenizer=tok, stride=16)
    long_text = ",".join(nc.TEXTS * 5)                       # ~ 400 tokens: four windows
    assert len(tok(long_text)["input_ids"]) > 300
    old = tok.model_max_length
    tok.model_max_length = 128
    try:
        ref_pipe = pipeline("ner", model=model, tokenizer=tok, aggregation_strategy="simple", device=-1, stride=16)
        ref = ref_pipe(long_text)
    finally:
        tok.model_max_length = old
    got = eng(long_text)
    assert len(ref) > 20
    nc.same_groups(got, ref, score_tol=1e-5)
    # a batch mixing short and long texts keeps every text's own result
    mixed = eng([nc.TEXTS[0], long_text, nc.TEXTS[1]])
    nc.same_groups(mixed[1], ref, score_tol=1e-5)
    nc.same_groups(mixed[0], eng(nc.TEXTS[0]), 1e-6)
    nc.same_groups(mixed[2], eng(nc.TEXTS[1]), 1e-6)
    # beyond the reference's 512-token horizon nothing is read
    huge = ",".join(nc.TEXTS * 12)
    assert len(tok(huge)["input_ids"]) > 700
    ends = [g["end"] for g in eng(huge)]
    cut = tok(huge, truncation=True, max_length=640, return_offsets_mapping=True)["offset_mapping"][-2][1]
    assert ends and max(ends) <= cut
