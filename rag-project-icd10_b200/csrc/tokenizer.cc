// tokenizer.cc -- host-side BERT WordPiece tokeniser feeding the GPU encoder (include/icdrag.h, icd_tokenizer_*).
//
// Stands where SentenceTransformer.encode tokenises its inputs in the reference (the model directory's
// BertTokenizerFast, reached from services/embedding_service.py:81,97-102,120).  "Only tokenisation stays on the
// host" (BASELINE north_star) -- but at 16 k sentences/s on 8 cores the Python-driven tokeniser, not the GPU
// (90 k+ sentences/s), bounded every text-in API.  This is the same algorithm, multi-threaded, writing straight into
// the id matrix the encoder reads:
//
//   normalise   (tokenizers' BertNormalizer: drop control characters, whitespace -> ' ', isolate CJK ideographs,
//                NFD + strip Mn + lower-case when the model lower-cases)
//   pre-tokenise (BertPreTokenizer: split on whitespace, isolate punctuation)
//   WordPiece   (greedy longest match, "##" continuation pieces, words longer than 100 characters -> [UNK])
//   template    ([CLS] ... [SEP], truncated to max_len)
//
// Unicode knowledge does not live here: the caller passes, per BMP code point, a class byte and the normalised
// replacement sequence (engine/tokenizer.py builds both from Python's unicodedata and marks every code point whose
// properties differ between Unicode versions as FALLBACK).  A sentence that contains a FALLBACK code point, a
// supplementary-plane character, malformed UTF-8 or a literal special token ("[CLS]", "[MASK]", ...) is not tokenised
// here: its needs_fallback byte is set and the caller runs the model's own tokenizer on it -- so results are identical
// to BertTokenizerFast by construction on exotic input and by test on everything else (tests/test_tokenizer_cpu.py).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/icdrag.h"

namespace icd {
void set_error(const char* fmt, ...);
}

namespace {

// class byte of an input code point (engine/tokenizer.py::build_char_tables)
enum : uint8_t { kClsMap = 0, kClsSpace = 1, kClsRemove = 2, kClsFallback = 3, kClsMask = 3, kFlagCjk = 4, kFlagPunct = 8 };

inline uint64_t fnv1a(const char* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) {
    h ^= (unsigned char)p[i];
    h *= 1099511628211ull;
  }
  return h;
}

// open-addressing map: token bytes -> id
struct PieceMap {
  struct Slot {
    uint32_t off, len;
    int32_t id;
  };
  std::vector<Slot> slots;
  std::string pool;
  uint64_t mask = 0;
  int max_len = 0;  // longest key in bytes

  void build(const std::vector<std::pair<std::string, int32_t>>& items) {
    size_t cap = 64;
    while (cap < items.size() * 2 + 8) cap <<= 1;
    slots.assign(cap, Slot{0, 0, -1});
    mask = cap - 1;
    for (const auto& it : items) {
      max_len = std::max(max_len, (int)it.first.size());
      uint64_t h = fnv1a(it.first.data(), it.first.size()) & mask;
      bool dup = false;
      while (slots[h].id >= 0) {
        if (slots[h].len == it.first.size() && !memcmp(pool.data() + slots[h].off, it.first.data(), it.first.size())) {
          slots[h].id = it.second;  // the same token twice in a vocab file: the later line wins, as in a Python dict
          dup = true;
          break;
        }
        h = (h + 1) & mask;
      }
      if (dup) continue;
      slots[h] = Slot{(uint32_t)pool.size(), (uint32_t)it.first.size(), it.second};
      pool.append(it.first);
    }
  }
  inline int32_t find(const char* p, size_t n) const {
    if ((int)n > max_len) return -1;
    uint64_t h = fnv1a(p, n) & mask;
    while (slots[h].id >= 0) {
      if (slots[h].len == n && !memcmp(pool.data() + slots[h].off, p, n)) return slots[h].id;
      h = (h + 1) & mask;
    }
    return -1;
  }
};

inline int put_utf8(uint32_t cp, char* out) {
  if (cp < 0x80) {
    out[0] = (char)cp;
    return 1;
  }
  if (cp < 0x800) {
    out[0] = (char)(0xC0 | (cp >> 6));
    out[1] = (char)(0x80 | (cp & 0x3F));
    return 2;
  }
  out[0] = (char)(0xE0 | (cp >> 12));
  out[1] = (char)(0x80 | ((cp >> 6) & 0x3F));
  out[2] = (char)(0x80 | (cp & 0x3F));
  return 3;
}

}  // namespace

struct icd_tokenizer {
  PieceMap first, cont;  // whole-word / word-initial pieces, and "##" continuation pieces (stored without the prefix)
  std::vector<uint8_t> cls;       // [65536]
  std::vector<uint32_t> map_off;  // [65537]
  std::vector<uint32_t> map_pool;
  int32_t unk = 100, cls_id = 101, sep_id = 102;
  int max_word_chars = 100;
  std::vector<std::string> specials;  // literal special-token strings: a sentence containing one takes the fallback
};

namespace {

struct Scratch {
  std::vector<char> bytes;       // UTF-8 of the current word
  std::vector<uint32_t> ends;    // byte offset after each character of the word
  std::vector<int32_t> pieces;   // ids of the current word
};

// WordPiece over the word held in s.bytes / s.ends; appends to out (bounded by cap).  Returns the new count.
inline int wordpiece(const icd_tokenizer& t, Scratch& s, int32_t* out, int n, int cap) {
  const int nch = (int)s.ends.size();
  if (nch == 0) return n;
  if (nch > t.max_word_chars) {
    if (n < cap) out[n++] = t.unk;
    return n;
  }
  s.pieces.clear();
  int start = 0;
  while (start < nch) {
    const uint32_t b0 = start ? s.ends[start - 1] : 0u;
    const PieceMap& m = start ? t.cont : t.first;
    int end = nch, id = -1;
    while (end > start) {
      const uint32_t b1 = s.ends[end - 1];
      if ((int)(b1 - b0) <= m.max_len) {
        id = m.find(s.bytes.data() + b0, b1 - b0);
        if (id >= 0) break;
      }
      --end;
    }
    if (id < 0) {  // one piece without a match: the whole word is unknown
      if (n < cap) out[n++] = t.unk;
      return n;
    }
    s.pieces.push_back(id);
    start = end;
  }
  for (int32_t id : s.pieces) {
    if (n >= cap) break;
    out[n++] = id;
  }
  return n;
}

// one sentence [p, e) -> ids (<= max_len, with [CLS]/[SEP]).  Returns the length, or -1 when the sentence needs the
// model's own tokenizer.
int encode_one(const icd_tokenizer& t, const unsigned char* p, const unsigned char* e, int max_len, int32_t* out, Scratch& s) {
  if (max_len < 2) return -1;
  const int cap = max_len - 1;  // room for [SEP]
  int n = 0;
  out[n++] = t.cls_id;
  // literal special tokens are matched by the reference tokenizer before normalisation
  if (memchr(p, '[', (size_t)(e - p))) {
    for (const std::string& sp : t.specials) {
      if (sp.empty() || (size_t)(e - p) < sp.size()) continue;
      const unsigned char* q = p;
      while (q + sp.size() <= e && (q = (const unsigned char*)memchr(q, sp[0], (size_t)(e - q) - sp.size() + 1))) {
        if (!memcmp(q, sp.data(), sp.size())) return -1;
        ++q;
      }
    }
  }
  s.bytes.clear();
  s.ends.clear();
  auto flush = [&]() {
    if (!s.ends.empty()) {
      n = wordpiece(t, s, out, n, cap);
      s.bytes.clear();
      s.ends.clear();
    }
  };
  while (p < e) {
    uint32_t cp;
    const unsigned char c = *p;
    if (c < 0x80) {
      cp = c;
      p += 1;
    } else if ((c & 0xE0) == 0xC0) {
      if (p + 1 >= e || (p[1] & 0xC0) != 0x80 || c < 0xC2) return -1;
      cp = ((uint32_t)(c & 0x1F) << 6) | (p[1] & 0x3F);
      p += 2;
    } else if ((c & 0xF0) == 0xE0) {
      if (p + 2 >= e || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80) return -1;
      cp = ((uint32_t)(c & 0x0F) << 12) | ((uint32_t)(p[1] & 0x3F) << 6) | (p[2] & 0x3F);
      if (cp < 0x800) return -1;
      p += 3;
    } else {
      return -1;  // supplementary planes (and malformed lead bytes): the model's own tokenizer decides
    }
    const uint8_t k = t.cls[cp];
    const uint8_t kind = k & kClsMask;
    if (kind == kClsFallback) return -1;
    if (kind == kClsRemove) continue;
    if (kind == kClsSpace) {
      flush();
      continue;
    }
    const bool cjk_in = (k & kFlagCjk) != 0;
    if (cjk_in) flush();
    const uint32_t m0 = t.map_off[cp], m1 = t.map_off[cp + 1];
    for (uint32_t mi = m0; mi < m1; ++mi) {
      const uint32_t oc = t.map_pool[mi];
      const bool isolate = cjk_in || (t.cls[oc] & kFlagPunct);
      if (isolate) flush();
      char buf[4];
      const int nb = put_utf8(oc, buf);
      s.bytes.insert(s.bytes.end(), buf, buf + nb);
      s.ends.push_back((uint32_t)s.bytes.size());
      if (isolate) flush();
    }
    if (n >= cap) break;  // already full: the rest is truncated away
  }
  flush();
  if (n > cap) n = cap;
  out[n++] = t.sep_id;
  return n;
}

}  // namespace

extern "C" {

int icd_tokenizer_create(const char* tokens, int64_t tokens_bytes, const int32_t* ids, int64_t n_tokens,
                         const uint8_t* char_class, const uint32_t* map_offsets, const uint32_t* map_pool,
                         int64_t pool_len, icd_tokenizer** out) {
  if (!tokens || !ids || !char_class || !map_offsets || !out || n_tokens <= 0 || (pool_len > 0 && !map_pool)) {
    icd::set_error("icd_tokenizer_create: null argument");
    return ICD_E_ARG;
  }
  if (map_offsets[65536] != (uint32_t)pool_len) {
    icd::set_error("icd_tokenizer_create: map_offsets[65536] must equal pool_len");
    return ICD_E_ARG;
  }
  icd_tokenizer* t = new icd_tokenizer();
  t->cls.assign(char_class, char_class + 65536);
  t->map_off.assign(map_offsets, map_offsets + 65537);
  t->map_pool.assign(map_pool, map_pool + pool_len);
  for (int64_t i = 0; i < pool_len; ++i) {
    if (t->map_pool[i] >= 65536) {
      delete t;
      icd::set_error("icd_tokenizer_create: replacement code points must lie in the BMP");
      return ICD_E_ARG;
    }
  }
  std::vector<std::pair<std::string, int32_t>> first, cont;
  const char* p = tokens;
  const char* e = tokens + tokens_bytes;
  int64_t i = 0;
  int unk = -1, cls = -1, sep = -1;
  while (p < e && i < n_tokens) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
    if (!nl) nl = e;
    std::string tok(p, nl);
    const int32_t id = ids[i++];
    p = nl + 1;
    if (tok == "[UNK]") unk = id;
    if (tok == "[CLS]") cls = id;
    if (tok == "[SEP]") sep = id;
    if (tok.size() > 2 && tok[0] == '[' && tok.back() == ']' &&
        (tok == "[UNK]" || tok == "[CLS]" || tok == "[SEP]" || tok == "[PAD]" || tok == "[MASK]"))
      t->specials.push_back(tok);
    if (tok.empty()) continue;
    if (tok.size() > 2 && tok[0] == '#' && tok[1] == '#') cont.emplace_back(tok.substr(2), id);
    else first.emplace_back(tok, id);
    // a bare "##x" piece can also open a word when the text itself contains "##x": '#' is punctuation, so it never does
  }
  if (i != n_tokens || unk < 0 || cls < 0 || sep < 0) {
    delete t;
    icd::set_error("icd_tokenizer_create: vocabulary needs %lld lines and [UNK]/[CLS]/[SEP] (parsed %lld)",
                   (long long)n_tokens, (long long)i);
    return ICD_E_ARG;
  }
  t->unk = unk;
  t->cls_id = cls;
  t->sep_id = sep;
  t->first.build(first);
  t->cont.build(cont);
  *out = t;
  return ICD_OK;
}

int icd_tokenizer_destroy(icd_tokenizer* t) {
  delete t;
  return ICD_OK;
}

int icd_tokenizer_encode(const icd_tokenizer* t, const char* texts, int64_t nbytes, int64_t n, int max_len,
                         int32_t* ids, int row_stride, int32_t* lens, uint8_t* needs_fallback, int threads) {
  if (!t || !ids || !lens || !needs_fallback || n < 0 || nbytes < 0 || (n > 0 && !texts)) {
    icd::set_error("icd_tokenizer_encode: bad argument");
    return ICD_E_ARG;
  }
  if (max_len < 2 || row_stride < max_len) {
    icd::set_error("icd_tokenizer_encode: need 2 <= max_len <= row_stride");
    return ICD_E_ARG;
  }
  if (n == 0) return ICD_OK;
  // sentence boundaries: the n texts are joined by single NUL bytes
  std::vector<int64_t> start((size_t)n + 1);
  {
    int64_t i = 0;
    const char* p = texts;
    const char* e = texts + nbytes;
    start[0] = 0;
    while (i + 1 < n) {
      const char* z = (const char*)memchr(p, 0, (size_t)(e - p));
      if (!z) break;
      start[++i] = (z - texts) + 1;
      p = z + 1;
    }
    if (i + 1 != n || memchr(p, 0, (size_t)(e - p))) {
      icd::set_error("icd_tokenizer_encode: expected %lld texts joined by NUL bytes", (long long)n);
      return ICD_E_ARG;
    }
    start[n] = nbytes + 1;
  }
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, (n + 255) / 256));
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    Scratch s;
    s.bytes.reserve(512);
    s.ends.reserve(128);
    s.pieces.reserve(128);
    for (;;) {
      const int64_t lo = next.fetch_add(256);
      if (lo >= n) break;
      const int64_t hi = std::min(n, lo + 256);
      for (int64_t i = lo; i < hi; ++i) {
        const unsigned char* p = (const unsigned char*)texts + start[i];
        const unsigned char* e = (const unsigned char*)texts + start[i + 1] - 1;
        const int len = encode_one(*t, p, e, max_len, ids + (size_t)i * row_stride, s);
        needs_fallback[i] = len < 0;
        lens[i] = len < 0 ? 0 : len;
      }
    }
  };
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int i = 0; i < nt; ++i) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  return ICD_OK;
}

int icd_pack_batch(const int32_t* ids, int row_stride, const int32_t* lens, const int64_t* rows, int B, int S,
                   int32_t* out_ids, int32_t* out_lens) {
  if (!ids || !lens || !rows || !out_ids || !out_lens || B < 0 || S < 1 || row_stride < 1) {
    icd::set_error("icd_pack_batch: bad argument");
    return ICD_E_ARG;
  }
  for (int b = 0; b < B; ++b) {
    const int64_t r = rows[b];
    const int len = std::min(lens[r], S);
    int32_t* dst = out_ids + (size_t)b * S;
    memcpy(dst, ids + (size_t)r * row_stride, (size_t)len * 4);
    memset(dst + len, 0, (size_t)(S - len) * 4);
    out_lens[b] = len;
  }
  return ICD_OK;
}

}  // extern "C"
