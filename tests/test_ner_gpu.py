"""GPU token-classification path (icd_encoder_set_token_head / icd_encoder_token_logits + TokenClassifierEngine)
against the transformers model and pipeline the reference constructs (medical_ner_service.py:76-90, :182) on the
same seeded synthetic weights: per-token logits close to the fp32 HF logits (bf16 activations), and the entity
groups of the "simple" aggregation identical wherever no argmax sits on a near-tie."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ner(tmp_path_factory):
    import ner_common as nc
    d, model, tok = nc.build(str(tmp_path_factory.mktemp("ner_model")))
    TC = importlib.import_module("rag-project-icd10_b200.engine.token_classifier")
    eng = TC.TokenClassifierEngine(d, device=0)
    yield nc, eng, model, tok
    eng.close()


def test_token_logits_match_hf_model(ner):
    nc, eng, model, tok = ner
    worst = 0.0
    for text in nc.TEXTS:
        ref = nc.hf_logits(model, tok, text)                       # [n, labels] fp32
        ids = np.asarray(tok(text)["input_ids"], np.int32)[None, :]
        got = eng.encoder.token_logits(ids, np.array([ids.shape[1]], np.int32))[0]
        assert got.shape == ref.shape
        scale = np.abs(ref).max()
        worst = max(worst, float(np.abs(got - ref).max() / scale))
        cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
        assert cos.min() >= 0.999, (text, cos.min())
    assert worst <= 0.03, worst          # bf16 activations through 2 layers vs fp32


def test_padding_and_batching_do_not_change_logits(ner):
    nc, eng, model, tok = ner
    enc = tok(nc.TEXTS)["input_ids"]
    S = max(len(r) for r in enc)
    mat = np.zeros((len(enc), S), np.int32)
    lens = np.array([len(r) for r in enc], np.int32)
    for i, r in enumerate(enc):
        mat[i, :len(r)] = r
    import importlib
    N = importlib.import_module("rag-project-icd10_b200._native")
    batch = eng.encoder.token_logits(mat, lens)
    scale = float(np.abs(batch).max())
    for i, r in enumerate(enc):
        ids = np.asarray(r, np.int32)[None, :]
        # a short text alone takes the few-token weight-streaming path (csrc/skinny_linear.cu), the batch the tile
        # kernels: same arithmetic in another summation order (bf16 streams: within 1 % of the logit scale) ...
        one = eng.encoder.token_logits(ids, lens[i:i + 1])[0]
        np.testing.assert_allclose(batch[i, :len(r)], one, atol=0.01 * scale, rtol=0.02)
        # ... and through the same kernels the padding and the batch around a text change next to nothing
        try:
            N.tune(enc_skinny=0)
            one = eng.encoder.token_logits(ids, lens[i:i + 1])[0]
        finally:
            N.tune(enc_skinny=1)
        np.testing.assert_allclose(batch[i, :len(r)], one, atol=0.05, rtol=0.02)


def test_entity_groups_match_transformers_pipeline(ner):
    nc, eng, model, tok = ner
    pipe = nc.hf_pipeline(model, tok)
    got_all = eng(nc.TEXTS)
    compared = groups = 0
    for text, got in zip(nc.TEXTS, got_all):
        ref_logits = nc.hf_logits(model, tok, text)
        top2 = np.sort(ref_logits, axis=1)[:, -2:]
        if (top2[:, 1] - top2[:, 0]).min() < 0.25:      # an argmax near-tie: bf16 may legitimately flip it
            continue
        nc.same_groups(got, pipe(text), score_tol=0.03)
        compared += 1
        groups += len(got)
    assert compared >= 4 and groups >= 8, (compared, groups)
    single = eng(nc.TEXTS[0])
    assert isinstance(single, list) and (not single or set(single[0]) == {"entity_group", "score", "word", "start", "end"})
    assert eng([]) == []


def test_token_head_errors(ner):
    nc, eng, model, tok = ner
    N = importlib.import_module("rag-project-icd10_b200._native")
    with pytest.raises(ValueError):
        eng.encoder.set_token_head(np.zeros((3, 5), np.float32), np.zeros(3, np.float32))
    with pytest.raises(N.NativeError):
        eng.encoder.set_token_head(np.zeros((N.MAX_LABELS + 1, 768), np.float32), np.zeros(N.MAX_LABELS + 1, np.float32))
    eng.encoder.set_token_head(*[np.asarray(p.detach().numpy(), np.float32) for p in (model.classifier.weight, model.classifier.bias)])


def test_long_text_is_read_in_one_pass_like_the_reference_pipeline(ner):
    """A ~400-token text: the reference pipeline reads it in one 512-token pass (medical_ner_service.py:177-229).  So
    does the engine (sequences beyond 128 tokens take the long-sequence attention kernel): logits against the fp32 HF
    model, entity groups against the pipeline without stride."""
    nc, eng, model, tok = ner
    assert eng.max_seq_length == 512
    long_text = ",".join(nc.TEXTS * 5)
    ids = np.asarray(tok(long_text, truncation=True, max_length=512)["input_ids"], np.int32)[None, :]
    assert 128 < ids.shape[1] <= 512
    ref = nc.hf_logits(model, tok, long_text, max_length=512)
    got = eng.encoder.token_logits(ids, np.array([ids.shape[1]], np.int32))[0]
    assert got.shape == ref.shape
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    assert cos.min() >= 0.999 and float(np.abs(got - ref).max() / np.abs(ref).max()) <= 0.03
    top2 = np.sort(ref, axis=1)[:, -2:]
    if (top2[:, 1] - top2[:, 0]).min() >= 0.25:
        nc.same_groups(eng(long_text), nc.hf_pipeline(model, tok)(long_text), score_tol=0.03)
    # a batch mixing short and long texts: every text equals its own single call
    both = eng([nc.TEXTS[0], long_text, nc.TEXTS[1]])
    assert [len(x) for x in both] == [len(eng(nc.TEXTS[0])), len(eng(long_text)), len(eng(nc.TEXTS[1]))]
