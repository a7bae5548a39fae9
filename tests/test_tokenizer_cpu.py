"""The host feeder's tokeniser (csrc/tokenizer.cc through engine/tokenizer.py) must return exactly what the model's own
BertTokenizerFast returns -- the tokenizer the reference reaches through SentenceTransformer.encode
(/root/reference/services/embedding_service.py:81,97-102,120).  No GPU involved."""
import importlib
import os
import random

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def toks(tmp_path_factory):
    import bench
    E = importlib.import_module("rag-project-icd10_b200.engine.encoder")
    T = importlib.import_module("rag-project-icd10_b200.engine.tokenizer")
    texts = ["query: " + t for t in bench.icd_texts()]
    d = tmp_path_factory.mktemp("vocab")
    extra = "the of and ing ##ing ##ed ##s un ##able é ü ##é αβγ ##βγ привет ##ет ① ℃ ㎎ ω ı icd10 ##10 ab ##ab abc".split()
    vp = bench.write_synthetic_vocab(texts + [" ".join(extra)], str(d / "vocab.txt"))
    lines = open(vp, encoding="utf-8").read().split("\n")
    free = [i for i, l in enumerate(lines) if l.startswith("[unused")][300:]
    for j, e in enumerate(extra):
        lines[free[j]] = e
    open(vp, "w", encoding="utf-8").write("\n".join(lines))
    hf = E.bert_tokenizer_from_vocab(vp)
    nt = T.NativeTokenizer(hf)
    assert nt.native, "the BertTokenizerFast pipeline must be recognised"
    return T, nt, texts


def _same(nt, texts, max_len):
    ids, lens = nt.encode(texts, max_len)
    ref = nt._hf_ids(texts, max_len)
    bad = [i for i, r in enumerate(ref) if lens[i] != len(r) or ids[i, :lens[i]].tolist() != r]
    assert not bad, (len(bad), [(texts[i], ids[i, :lens[i]].tolist(), ref[i]) for i in bad[:3]])


def test_all_icd_texts_tokenise_identically_without_fallback(toks):
    T, nt, texts = toks
    before = nt.fallbacks
    _same(nt, texts, 128)
    assert nt.fallbacks == before, "every real ICD string must take the native path"
    _same(nt, texts[:5000], 16)          # truncation


def test_every_native_code_point_agrees_with_the_wrapped_tokenizer(toks):
    """All BMP code points the tables claim to know, in three contexts (inside a word, alone, next to CJK / digits)."""
    T, nt, _ = toks
    cls, _, _ = T.build_char_tables(True, True, True)
    cps = [cp for cp in range(1, 65536) if (cls[cp] & 3) != 3]
    assert len(cps) > 50000
    texts = []
    for cp in cps:
        c = chr(cp)
        texts += ["a" + c + "b", c, "中" + c + "1 " + c]
    _same(nt, texts, 32)


def test_edge_cases_and_fallback_sentences(toks):
    T, nt, _ = toks
    cases = ["", " ", "a" * 101, "a" * 100, "ab" * 60, "[MASK] test", "x[CLS]y", "[cls]", "é", "中́文", "İstanbul", "ΑΣ", "ß", "≠",
             "Å", "Å", "ǅ", "ﬁ", "…", "½", "²", "a\x0bb", "a\x85b", "a b", "a᠎b", "﻿abc", "a​b", "😀 ok",
             "\U00020000", "abc\x00def", "\x00", "a\ud800b".encode("utf-16", "surrogatepass").decode("utf-16", "replace"),
             "急性胃肠炎 发热", "ICD-10: A00.001", "query: 2型糖尿病伴有多个并发症", "Ⅱ型呼吸衰竭（COPD）", "ＡＢＣ１２３"]
    _same(nt, cases, 128)
    _same(nt, cases, 8)
    before = nt.fallbacks
    nt.encode(["plain ascii", "中文"], 16)
    assert nt.fallbacks == before
    nt.encode(["😀"], 16)
    assert nt.fallbacks == before + 1


def test_fuzz_against_the_wrapped_tokenizer(toks):
    T, nt, _ = toks
    rng = random.Random(7)
    ranges = [(0x20, 0x7e)] * 6 + [(0x4e00, 0x9fa5)] * 4 + [(0xa0, 0xff), (0x100, 0x24f), (0x370, 0x3ff), (0x400, 0x4ff),
              (0x300, 0x36f), (0x2000, 0x206f), (0x2100, 0x218f), (0x2190, 0x22ff), (0x2460, 0x24ff), (0x3000, 0x303f),
              (0x3040, 0x30ff), (0xff00, 0xffef), (0x1, 0x1f), (0x7f, 0x9f), (0xac00, 0xd7a3), (0x3200, 0x33ff), (0xf900, 0xfaff),
              (0xfe30, 0xfe6f), (0x1e00, 0x1eff), (0x1f00, 0x1fff), (0x590, 0x6ff), (0x900, 0x97f), (0xe00, 0xe7f)]

    def rnd():
        s = []
        for _ in range(rng.randint(0, 40)):
            a, b = rng.choice(ranges)
            s.append(chr(rng.randint(a, b)))
            if rng.random() < 0.15:
                s.append(" ")
        return "".join(s)
    texts = [rnd() for _ in range(20000)]
    _same(nt, texts, 128)
    _same(nt, texts, 12)


def test_pack_batch_gathers_and_pads(toks):
    T, nt, texts = toks
    ids, lens = nt.encode(texts[:50], 128)
    rows = np.array([7, 3, 49, 0], np.int64)
    S = int(lens[rows].max())
    out = np.full((4, S), -1, np.int32)
    ol = np.zeros(4, np.int32)
    nt.pack(ids, lens, rows, S, out, ol)
    for b, r in enumerate(rows):
        assert ol[b] == lens[r] and out[b, :lens[r]].tolist() == ids[r, :lens[r]].tolist() and not out[b, lens[r]:].any()


def test_non_bert_tokenizers_are_left_to_the_wrapped_tokenizer():
    T = importlib.import_module("rag-project-icd10_b200.engine.tokenizer")

    class Odd:
        class backend_tokenizer:
            @staticmethod
            def to_str():
                return '{"normalizer": {"type": "NFKC"}, "model": {"type": "BPE"}}'

        def __call__(self, texts, **kw):
            return {"input_ids": [[101, 5, 102] for _ in texts]}
    nt = T.NativeTokenizer(Odd())
    assert not nt.native
    ids, lens = nt.encode(["x", "y"], 16)
    assert lens.tolist() == [3, 3] and ids[1, :3].tolist() == [101, 5, 102]
