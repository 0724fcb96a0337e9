"""Synthetic corpora of the shapes BASELINE.json / SURVEY.md section 8(d) name (the reference's benchmark corpus,
400 Gutenberg books, is not shipped: benchmark/data/.gitkeep).  Measurement tooling, not part of the encode path.

Text is assembled on the torch device it is asked for (CPU here, CUDA on the GPU box) from tables of
pre-rendered items, so 1 GiB takes seconds:
  english      : words drawn Zipf(1.1) from the first 30 000 space-prefixed letters-only cl100k tokens, sentences
                 with , . ; : ! ? punctuation, ~2 % contractions, ~1 % digit groups, lines wrapped with \\n, blank
                 lines between paragraphs
  multilingual : per document one of english / latin-with-diacritics / cyrillic / cjk (no spaces) /
                 arabic+hebrew / indic, plus ~0.5 % emoji (ZWJ / VS16 sequences)
Only code points from old, stable Unicode blocks are used (same class in every JVM >= 8), and no
special-token strings occur.
"""
import math
import os
import unicodedata

import numpy as np
import torch

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

KINDS = ["word", "cap", "comma", "period", "wrap", "para", "contraction", "number", "emoji", "punct"]
KIND_P = [0.625, 0.05, 0.08, 0.06, 0.08, 0.01, 0.02, 0.01, 0.005, 0.06]
KIND_P_CHAT = [0.705, 0.05, 0.08, 0.06, 0.0, 0.0, 0.02, 0.01, 0.015, 0.06]  # no line wrapping / paragraphs (config 4)
ZIPF_K = 30000
ZIPF_S = 1.1

# script -> list of (lo, hi) letter ranges (all present since Unicode 1.1-3.0)
SCRIPTS = {
    "latin": [(0x41, 0x5A), (0x61, 0x7A), (0xC0, 0xD6), (0xD8, 0xF6), (0xF8, 0xFF), (0x100, 0x17E)],
    "cyrillic": [(0x410, 0x44F), (0x401, 0x401), (0x451, 0x451)],
    "cjk": [(0x4E00, 0x9FA5), (0x3041, 0x3093), (0x30A1, 0x30F6)],
    "semitic": [(0x621, 0x63A), (0x641, 0x64A), (0x5D0, 0x5EA)],
    "indic": [(0x905, 0x939), (0x93E, 0x94D), (0xB85, 0xBB9), (0xBBE, 0xBCD), (0x985, 0x9B9), (0x9BE, 0x9CD)],
}
EMOJI = ["\U0001F600", "\U0001F44D", "❤️", "\U0001F468‍\U0001F469‍\U0001F467", "\U0001F355", "☃️", "\U0001F680", "\U0001F389"]
MULTILINGUAL_MIX = [("english", 0.5), ("latin", 0.1), ("cyrillic", 0.1), ("cjk", 0.1), ("semitic", 0.1), ("indic", 0.1)]


def _load_ranks(name="cl100k_base.tiktoken"):
    import base64
    ranks = []
    with open(os.path.join(DATA_DIR, name), "rb") as f:
        for line in f.read().splitlines():
            if line:
                tok, rank = line.split(None, 1)
                ranks.append((int(rank), base64.b64decode(tok)))
    ranks.sort()
    return [t for _, t in ranks]


def _in_ranges(ch, ranges):
    o = ord(ch)
    return any(lo <= o <= hi for lo, hi in ranges)


def _script_words(tokens, script, rng, want):
    """Words of a script: vocabulary tokens that decode to letters of that script (rank order), topped up with
    deterministic random compositions from the script's ranges."""
    ranges = SCRIPTS[script]
    words, seen = [], set()
    for t in tokens:
        try:
            s = t.decode("utf-8")
        except UnicodeDecodeError:
            continue
        s = s.lstrip(" ")
        if not s or s in seen:
            continue
        if all(_in_ranges(c, ranges) and unicodedata.category(c) != "Cn" for c in s):
            if script == "latin" and all(ord(c) < 128 for c in s):
                continue
            if script == "indic" and unicodedata.category(s[0])[0] == "M":
                continue  # do not start a word with a combining mark
            seen.add(s)
            words.append(s)
            if len(words) >= want:
                return words
    pool = [chr(o) for lo, hi in ranges for o in range(lo, hi + 1) if unicodedata.category(chr(o)) != "Cn"]
    letters = [c for c in pool if unicodedata.category(c)[0] == "L"]
    lo_len, hi_len = (1, 3) if script == "cjk" else (2, 7)
    while len(words) < want:
        n = int(rng.integers(lo_len, hi_len + 1))
        w = letters[int(rng.integers(len(letters)))] + "".join(pool[int(rng.integers(len(pool)))] for _ in range(n - 1))
        if w not in seen:
            seen.add(w)
            words.append(w)
    return words


_SHORT_WORDS = {"a", "i", "of", "to", "in", "is", "it", "on", "at", "by", "an", "as", "be", "or", "we", "he", "so", "do", "if", "my", "no", "up",
                "me", "us", "am", "go"}


def _english_words(tokens, want=ZIPF_K):
    """Space-prefixed letters-only tokens in rank order.  Low ranks are dominated by merge fragments (" t", " th",
    " c"); those are dropped (one- and two-letter tokens unless real words, longer ones without a vowel) so that
    the word-length distribution is English-like."""
    words = []
    for t in tokens:
        if len(t) >= 2 and t[:1] == b" " and t[1:].isalpha() and t[1:].isascii():
            w = t[1:].decode("ascii")
            lw = w.lower()
            if len(w) <= 2 and lw not in _SHORT_WORDS:
                continue
            if len(w) >= 3 and not any(c in "aeiouy" for c in lw):
                continue
            words.append(w)
            if len(words) >= want:
                break
    return words


_DIACRITICS = {"a": "áàäâã", "e": "éèêë", "i": "íï", "o": "óöôõ", "u": "úü", "c": "ç", "n": "ñ", "s": "ß", "y": "ý"}


def _latin_words(tokens, rng, want):
    """Western-European-looking words: English word shapes with about every second word carrying a diacritic."""
    base = _english_words(tokens, want)
    out = []
    for i, w in enumerate(base):
        if i % 2 == 0:
            pos = [k for k, c in enumerate(w) if c in _DIACRITICS]
            if pos:
                k = pos[int(rng.integers(len(pos)))]
                alt = _DIACRITICS[w[k]]
                w = w[:k] + alt[int(rng.integers(len(alt)))] + w[k + 1:]
        out.append(w)
    return out


CONTRACTIONS = ["'s", "'t", "'re", "'ll", "'ve", "'m", "'d"]


def _render(lang, words, rng):
    """kind -> list of rendered items (bytes) for one language."""
    cjk = lang == "cjk"
    sep = "" if cjk else " "
    comma, period = ("，", "。") if cjk else (",", ".")
    others = ["！", "？", "；", "："] if cjk else ["!", "?", ";", ":"]
    cap = (lambda w: w[:1].upper() + w[1:]) if lang in ("english", "latin", "cyrillic") else (lambda w: w)
    numbers = ["".join(str(int(d)) for d in rng.integers(0, 10, size=int(rng.integers(1, 7)))) for _ in range(4096)]
    out = {
        "word": [sep + w for w in words],
        "cap": [sep + cap(w) for w in words],
        "comma": [sep + w + comma for w in words],
        "period": [sep + w + period for w in words],
        "wrap": ["\n" + w for w in words],
        "para": [period + "\n\n" + cap(w) for w in words],
        "number": [sep + n for n in numbers] if not cjk else numbers,
        "emoji": [sep + e for e in EMOJI],
        "punct": [sep + w + others[i % 4] for i, w in enumerate(words)],
    }
    if lang == "english":
        con = []
        for i, w in enumerate(words):
            c = CONTRACTIONS[i % len(CONTRACTIONS)]
            con.append(sep + w + (c.upper() if i % 10 == 9 else c))
        out["contraction"] = con
    else:
        out["contraction"] = out["word"]
    return {k: [s.encode("utf-8") for s in v] for k, v in out.items()}


class ItemTables:
    """Flat table of every rendered item of every language, as tensors on `device`."""

    def __init__(self, device, seed=1234):
        rng = np.random.Generator(np.random.PCG64(seed))
        tokens = _load_ranks()
        self.langs = [m[0] for m in MULTILINGUAL_MIX]
        items, base, size = [], np.zeros((len(self.langs), len(KINDS)), dtype=np.int64), np.zeros((len(self.langs), len(KINDS)), dtype=np.int64)
        for li, lang in enumerate(self.langs):
            if lang == "english":
                words = _english_words(tokens)
            elif lang == "latin":
                words = _latin_words(tokens, rng, 8000)
            else:
                words = _script_words(tokens, lang, rng, 4000)
            rendered = _render(lang, words, rng)
            for ki, kind in enumerate(KINDS):
                base[li, ki] = len(items)
                size[li, ki] = len(rendered[kind])
                items.extend(rendered[kind])
        self.maxlen = max(len(b) for b in items)
        lens = np.array([len(b) for b in items], dtype=np.int32)
        flat = np.zeros((len(items), self.maxlen), dtype=np.uint8)
        for i, b in enumerate(items):
            flat[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
        self.device = device
        self.table = torch.from_numpy(flat).to(device).reshape(-1)
        self.lens = torch.from_numpy(lens).to(device)
        self.base = torch.from_numpy(base).to(device)
        self.size = torch.from_numpy(size).to(device)
        zipf = 1.0 / np.arange(1, ZIPF_K + 1) ** ZIPF_S
        self.zipf_cdf = torch.from_numpy(np.cumsum(zipf / zipf.sum())).to(device)
        # expected item length per language (for sizing documents)
        zp = zipf / zipf.sum()
        self.avg_len = np.zeros((len(self.langs), 2))
        for variant, kp in enumerate((KIND_P, KIND_P_CHAT)):
            for li in range(len(self.langs)):
                e = 0.0
                for ki in range(len(KINDS)):
                    l = lens[base[li, ki]:base[li, ki] + size[li, ki]].astype(np.float64)
                    p = np.zeros(len(l))
                    np.add.at(p, np.arange(ZIPF_K) % len(l), zp)
                    e += kp[ki] * float((p * l).sum())
                self.avg_len[li, variant] = e


_TABLES = {}


def tables(device):
    key = str(device)
    if key not in _TABLES:
        _TABLES[key] = ItemTables(device)
    return _TABLES[key]


def generate(total_bytes, seed, device="cpu", doc_len=("loguniform", 1024, 65536), mix=None, chat=False):
    """Returns (uint8 tensor of ~total_bytes bytes, int64 document offsets [ndocs + 1]) on `device`.

    doc_len: ("loguniform", lo, hi) | ("fixed", n) | ("lognormal", mean, lo, hi)
    mix    : list of (language, probability); default english only
    """
    device = torch.device(device)
    T = tables(device)
    mix = mix or [("english", 1.0)]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    # ---- documents
    if doc_len[0] == "fixed":
        mean = float(doc_len[1])
    elif doc_len[0] == "loguniform":
        lo, hi = doc_len[1], doc_len[2]
        mean = (hi - lo) / math.log(hi / lo)
    else:
        mean = float(doc_len[1])
    ndocs = max(1, int(total_bytes / mean * 1.05) + 8)
    u = torch.rand(ndocs, generator=g, device=device, dtype=torch.float64)
    if doc_len[0] == "fixed":
        target = torch.full((ndocs,), float(doc_len[1]), device=device, dtype=torch.float64)
    elif doc_len[0] == "loguniform":
        target = torch.exp(math.log(doc_len[1]) + u * (math.log(doc_len[2]) - math.log(doc_len[1])))
    else:
        sigma = 0.9
        mu = math.log(doc_len[1]) - sigma * sigma / 2
        z = torch.randn(ndocs, generator=g, device=device, dtype=torch.float64)
        target = torch.exp(mu + sigma * z).clamp(doc_len[2], doc_len[3])
    csum = torch.cumsum(target, 0)
    ndocs = int(torch.searchsorted(csum, torch.tensor([float(total_bytes)], device=device, dtype=torch.float64)).item()) + 1
    ndocs = min(ndocs, target.numel())
    target = target[:ndocs]
    lang_ids = torch.tensor([T.langs.index(m[0]) for m in mix], device=device)
    lang_p = torch.tensor([m[1] for m in mix], device=device, dtype=torch.float64)
    lang = lang_ids[torch.multinomial(lang_p, ndocs, replacement=True, generator=g)]
    avg = torch.from_numpy(T.avg_len[:, 1 if chat else 0]).to(device)[lang]
    n_items = torch.clamp((target / avg).round().to(torch.int64), min=1)
    first_item = torch.zeros(ndocs + 1, dtype=torch.int64, device=device)
    first_item[1:] = torch.cumsum(n_items, 0)
    N = int(first_item[-1].item())
    kind_cdf = torch.tensor(np.cumsum(KIND_P_CHAT if chat else KIND_P), device=device, dtype=torch.float64)
    # ---- items, assembled in blocks to bound memory
    item_len_parts, out_parts = [], []
    BLOCK = 8 << 20
    doc_of_item_full = torch.repeat_interleave(torch.arange(ndocs, device=device, dtype=torch.int32), n_items)
    for s in range(0, N, BLOCK):
        e = min(N, s + BLOCK)
        li = lang[doc_of_item_full[s:e].long()]
        kind = torch.searchsorted(kind_cdf, torch.rand(e - s, generator=g, device=device, dtype=torch.float64)).clamp(max=len(KINDS) - 1)
        zidx = torch.searchsorted(T.zipf_cdf, torch.rand(e - s, generator=g, device=device, dtype=torch.float64)).clamp(max=ZIPF_K - 1)
        item = T.base[li, kind] + zidx % T.size[li, kind]
        lens = T.lens[item]
        item_len_parts.append(lens)
        rep = torch.repeat_interleave(item, lens.long())
        starts = torch.cumsum(lens.long(), 0) - lens.long()
        pos = torch.arange(rep.numel(), device=device, dtype=torch.int64) - torch.repeat_interleave(starts, lens.long())
        out_parts.append(T.table[rep * T.maxlen + pos])
        del rep, pos, starts
    data = torch.cat(out_parts) if out_parts else torch.zeros(0, dtype=torch.uint8, device=device)
    item_len = torch.cat(item_len_parts).long()
    item_off = torch.zeros(N + 1, dtype=torch.int64, device=device)
    item_off[1:] = torch.cumsum(item_len, 0)
    doc_off = item_off[first_item]
    return data, doc_off


# ---- the named configurations of BASELINE.json ----------------------------------------------------
def config2_english_64mib(device="cpu", total=64 << 20, seed=2002):
    """64 MiB synthetic English batch, 1 024 documents of ~64 KiB."""
    return generate(total, seed, device, doc_len=("fixed", 65536))


def config3_multilingual(device="cpu", total=1 << 30, seed=3003):
    """1 GiB synthetic multilingual corpus, documents log-uniform in [1 KiB, 64 KiB]."""
    return generate(total, seed, device, doc_len=("loguniform", 1024, 65536), mix=MULTILINGUAL_MIX)


def config4_chat(device="cpu", total=2560 << 20, seed=4004):
    """Short chat-length strings: log-normal, mean 256 B, clipped to [8, 2048]."""
    return generate(total, seed, device, doc_len=("lognormal", 256, 8, 2048), chat=True)


def config5_adversarial(n=1 << 20, seed=5005):
    """1 MiB single-piece documents (one per class); returns a list of bytes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    letters = bytes(rng.integers(97, 123, size=n, dtype=np.uint8))
    han = "".join(chr(int(c)) for c in rng.integers(0x4E00, 0x9FA5, size=n // 3))
    digits = bytes(rng.integers(48, 58, size=n, dtype=np.uint8))
    return [b"a" * n, letters, b" " * n, b"!" * n, b"ab" * (n // 2), han.encode("utf-8"), b"\n" * n, digits]
