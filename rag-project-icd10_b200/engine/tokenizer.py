"""NativeTokenizer -- the host feeder in front of the GPU encoder (csrc/tokenizer.cc, icd_tokenizer_*).

The reference tokenises inside ``SentenceTransformer.encode`` with the model directory's BertTokenizerFast
(/root/reference/services/embedding_service.py:81,97-102,120).  That stays the definition of correct: this class
wraps such a tokenizer, reproduces its normaliser / pre-tokeniser / WordPiece in multi-threaded C++ for the
characters whose Unicode properties are beyond doubt, and sends every other sentence (flagged by the C++ side)
through the wrapped tokenizer itself.  When the wrapped tokenizer is not a plain BERT WordPiece pipeline the native
path is off and everything goes through it -- results never depend on which path ran.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import unicodedata as ud
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .. import _native as N

_CLS_MAP, _CLS_SPACE, _CLS_REMOVE, _CLS_FALLBACK = 0, 1, 2, 3
_FLAG_CJK, _FLAG_PUNCT = 4, 8
_SPECIALS = {"[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"}

# Unicode White_Space minus the control characters the normaliser removes first (tokenizers' BertNormalizer:
# is_control wins over is_whitespace, except for \t \n \r)
_WHITESPACE = {0x09, 0x0A, 0x0D, 0x20, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000} | set(range(0x2000, 0x200B))

_TABLE_CACHE = {}


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0xF900 <= cp <= 0xFAFF or 0x20000 <= cp <= 0x2A6DF or
            0x2A700 <= cp <= 0x2B73F or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or 0x2F800 <= cp <= 0x2FA1F)


def _is_punct(cp: int, cat: str) -> bool:
    return 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126 or cat.startswith("P")


def build_char_tables(lowercase: bool, strip_accents: bool, chinese: bool = True):
    """Per BMP code point: class byte (bits 0-1 kind, bit 2 CJK input, bit 3 punctuation as an OUTPUT character) and the
    normalised replacement sequence.  A code point is FALLBACK (sentence goes to the wrapped tokenizer) unless its
    category and canonical decomposition are the same in Unicode 3.2 and in this Python's Unicode: whatever Unicode
    version the reference's tokenizer was built against, it agrees on those."""
    key = (lowercase, strip_accents, chinese)
    if key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    old = ud.ucd_3_2_0
    cls = np.zeros(65536, np.uint8)
    offs = np.zeros(65537, np.uint32)
    pool: List[int] = []
    stable_cat = [False] * 65536
    cats = [""] * 65536
    for cp in range(65536):
        ch = chr(cp)
        cats[cp] = ud.category(ch)
        stable_cat[cp] = not (0xD800 <= cp <= 0xDFFF) and old.category(ch) == cats[cp]
    for cp in range(65536):
        ch = chr(cp)
        cat = cats[cp]
        offs[cp] = len(pool)
        if _is_punct(cp, cat) and stable_cat[cp]:
            cls[cp] |= _FLAG_PUNCT
        if not stable_cat[cp]:
            cls[cp] |= _CLS_FALLBACK
            continue
        if cp == 0 or cp == 0xFFFD:
            cls[cp] |= _CLS_REMOVE
            continue
        if cp in (0x09, 0x0A, 0x0D):
            cls[cp] |= _CLS_SPACE
            continue
        if cat == "Cn":                  # unassigned: tokenizers' tables only know assigned characters -> wrapped tokenizer
            cls[cp] |= _CLS_FALLBACK
            continue
        if cat in ("Cc", "Cf", "Co"):
            cls[cp] |= _CLS_REMOVE
            continue
        if cp in _WHITESPACE:
            cls[cp] |= _CLS_SPACE
            continue
        if cat.startswith("Z"):          # a separator outside the White_Space list: not sure -> wrapped tokenizer
            cls[cp] |= _CLS_FALLBACK
            continue
        if chinese and _is_cjk(cp):
            cls[cp] |= _FLAG_CJK
        seq = ch
        if strip_accents:
            nfd = ud.normalize("NFD", ch)
            if nfd != old.normalize("NFD", ch):
                cls[cp] = (int(cls[cp]) & 0xFC) | _CLS_FALLBACK
                continue
            # canonical reordering only permutes combining marks; all Mn are dropped, so it is invisible unless a
            # combining character of another category is involved
            if any(ud.combining(c) != 0 and ud.category(c) != "Mn" for c in nfd):
                cls[cp] = (int(cls[cp]) & 0xFC) | _CLS_FALLBACK
                continue
            seq = "".join(c for c in nfd if ud.category(c) != "Mn")
        elif ud.combining(ch) != 0 and cat != "Mn":
            cls[cp] = (int(cls[cp]) & 0xFC) | _CLS_FALLBACK
            continue
        if lowercase:
            seq = "".join(c.lower() for c in seq)
        out = [ord(c) for c in seq]
        if any(o >= 65536 or not stable_cat[o] for o in out):
            cls[cp] = (int(cls[cp]) & 0xFC) | _CLS_FALLBACK
            continue
        pool.extend(out)
    offs[65536] = len(pool)
    res = (cls, offs, np.asarray(pool, np.uint32))
    _TABLE_CACHE[key] = res
    return res


def _plain_bert_pipeline(tok) -> Optional[dict]:
    """The wrapped tokenizer's settings when it is exactly BertNormalizer -> BertPreTokenizer -> WordPiece('##') ->
    [CLS] A [SEP] with no added tokens beyond the five specials; None otherwise."""
    try:
        j = json.loads(tok.backend_tokenizer.to_str())
    except Exception:
        return None
    nz, pt, md, pp = j.get("normalizer") or {}, j.get("pre_tokenizer") or {}, j.get("model") or {}, j.get("post_processor") or {}
    if nz.get("type") != "BertNormalizer" or not nz.get("clean_text", True):
        return None
    if pt.get("type") != "BertPreTokenizer" or md.get("type") != "WordPiece":
        return None
    if md.get("continuing_subword_prefix", "##") != "##" or md.get("unk_token", "[UNK]") != "[UNK]":
        return None
    if any(a.get("content") not in _SPECIALS or not a.get("special", False) for a in j.get("added_tokens", [])):
        return None
    single = pp.get("single") or []
    names = [next(iter(x.values())).get("id") for x in single]
    if pp.get("type") != "TemplateProcessing" or names != ["[CLS]", "A", "[SEP]"]:
        return None
    lower = bool(nz.get("lowercase", True))
    strip = nz.get("strip_accents")
    return {"lowercase": lower, "strip_accents": lower if strip is None else bool(strip),
            "chinese": bool(nz.get("handle_chinese_chars", True)),
            "max_word": int(md.get("max_input_chars_per_word", 100))}


class NativeTokenizer:
    """encode(texts, max_len) -> (ids [n, max_len] int32, rows valid up to lens; lens [n] int32)."""

    def __init__(self, tokenizer, threads: Optional[int] = None):
        self.hf = tokenizer
        self.threads = int(threads or os.cpu_count() or 1)
        self._h = C.c_void_p()
        self.native = False
        self.fallbacks = 0
        cfg = _plain_bert_pipeline(tokenizer)
        if cfg is None or cfg["max_word"] != 100:
            return
        vocab = tokenizer.get_vocab()
        if any("\n" in t for t in vocab):
            return
        toks = list(vocab.keys())
        blob = ("\n".join(toks) + "\n").encode("utf-8")
        ids = np.asarray([vocab[t] for t in toks], np.int32)
        cls, offs, pool = build_char_tables(cfg["lowercase"], cfg["strip_accents"], cfg["chinese"])
        self._keep = (blob, ids, cls, offs, pool)
        N.check(N.lib().icd_tokenizer_create(blob, len(blob), N.buf_ptr(ids), len(toks), N.buf_ptr(cls), N.buf_ptr(offs),
                                             N.buf_ptr(pool) if pool.size else None, int(pool.size), C.byref(self._h)),
                "icd_tokenizer_create")
        self.native = True

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            N.lib().icd_tokenizer_destroy(self._h)
            self._h = C.c_void_p()
            self.native = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _hf_ids(self, texts: Sequence[str], max_len: int) -> List[List[int]]:
        enc = self.hf(list(texts), padding=False, truncation=True, max_length=max_len, add_special_tokens=True,
                      return_attention_mask=False, return_token_type_ids=False)
        return enc["input_ids"]

    def encode(self, texts: Sequence[str], max_len: int, ids_out: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        n = len(texts)
        ids = ids_out if ids_out is not None else np.empty((n, max_len), np.int32)
        if ids.shape[0] < n or ids.shape[1] < max_len or ids.dtype != np.int32 or not ids.flags["C_CONTIGUOUS"]:
            raise ValueError("ids_out must be a C-contiguous int32 array of at least [n, max_len]")
        lens = np.zeros((n,), np.int32)
        if n == 0:
            return ids, lens
        if not self.native:
            for i, row in enumerate(self._hf_ids(texts, max_len)):
                lens[i] = len(row)
                ids[i, :len(row)] = row
            return ids, lens
        joined = "\x00".join(texts)
        slow = np.zeros((n,), np.uint8)
        if joined.count("\x00") != n - 1:
            # a text with an embedded NUL (the normaliser would drop it): keep boundaries exact, one call per text
            for i, t in enumerate(texts):
                i2, l2 = self.encode([t.replace("\x00", "")], max_len)
                ids[i, :max_len], lens[i] = i2[0], l2[0]
            return ids, lens
        try:
            blob = joined.encode("utf-8")
        except UnicodeEncodeError:        # lone surrogates: not valid UTF-8, the wrapped tokenizer decides
            blob = None
        if blob is None:
            slow[:] = 1
        else:
            N.check(N.lib().icd_tokenizer_encode(self._h, blob, len(blob), n, int(max_len), N.buf_ptr(ids),
                                                 int(ids.shape[1]), N.buf_ptr(lens), N.buf_ptr(slow), self.threads),
                    "icd_tokenizer_encode")
        bad = np.flatnonzero(slow)
        if bad.size:
            self.fallbacks += int(bad.size)
            for i, row in zip(bad, self._hf_ids([texts[int(i)] for i in bad], max_len)):
                lens[i] = len(row)
                ids[i, :len(row)] = row
        return ids, lens

    def pack(self, ids: np.ndarray, lens: np.ndarray, rows: np.ndarray, S: int, out_ids: np.ndarray, out_lens: np.ndarray) -> None:
        """Gather `rows` of the ragged id table into a zero-padded [B, S] batch (host buffers, typically pinned)."""
        rows = np.ascontiguousarray(rows, np.int64)
        N.check(N.lib().icd_pack_batch(N.buf_ptr(ids), int(ids.shape[1]), N.buf_ptr(lens), N.buf_ptr(rows), int(rows.size),
                                       int(S), N.buf_ptr(out_ids), N.buf_ptr(out_lens)), "icd_pack_batch")
