"""ORACLE - TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/libjo_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; nothing under jtokkit_b200/ may.  See oracle/jo_oracle.h for what it restates and how it is pinned.
"""
import base64
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "..", "jtokkit_b200", "data")
LIB_PATH = os.path.join(HERE, "libjo_oracle.so")

# The reference's split regexes and special tokens (EncodingFactory.java:18-53,63,77,91,105), flags (:129).
X50K_PATTERN = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"
CL100K_PATTERN = r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"
UNICODE_CHARACTER_CLASS = 0x100
BUILTIN = {
    "r50k_base": (X50K_PATTERN, "r50k_base.tiktoken", {"<|endoftext|>": 50256}),
    "p50k_base": (X50K_PATTERN, "p50k_base.tiktoken", {"<|endoftext|>": 50256}),
    "p50k_edit": (X50K_PATTERN, "p50k_base.tiktoken",
                  {"<|endoftext|>": 50256, "<|fim_prefix|>": 50281, "<|fim_middle|>": 50282, "<|fim_suffix|>": 50283}),
    "cl100k_base": (CL100K_PATTERN, "cl100k_base.tiktoken",
                    {"<|endoftext|>": 100257, "<|fim_prefix|>": 100258, "<|fim_middle|>": 100259, "<|fim_suffix|>": 100260,
                     "<|endofprompt|>": 100276}),
}

E_SPECIAL, E_UNKNOWN_BYTES, E_UNKNOWN_ID, E_CAPACITY = -1, -2, -3, -4
MERGE_LITERAL, MERGE_HEAP, MERGE_AUTO = 0, 1, 2


def build():
    """Compile the C restatement (gcc); building the checker is not using it."""
    subprocess.check_call(["make", "-s", "-C", HERE])


def _lib():
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    vp, i64p, i32p, u8p = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    lib.jo_create.restype = vp
    lib.jo_create.argtypes = [C.c_char_p, C.c_int, vp, vp, vp, C.c_int64, vp, vp, vp, C.c_int64, C.c_char_p, C.c_int]
    lib.jo_destroy.argtypes = [vp]
    lib.jo_split.restype = C.c_int64
    lib.jo_split.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int64]
    lib.jo_encode.restype = C.c_int64
    lib.jo_encode.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, vp, C.c_int64]
    lib.jo_encode_with_special.restype = C.c_int64
    lib.jo_encode_with_special.argtypes = [vp, vp, C.c_int64, C.c_int, vp, C.c_int64]
    lib.jo_encode_max.restype = C.c_int64
    lib.jo_encode_max.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, vp, C.c_int64, C.POINTER(C.c_int)]
    lib.jo_contains_special.restype = C.c_int
    lib.jo_contains_special.argtypes = [vp, vp, C.c_int64]
    lib.jo_decode_bytes.restype = C.c_int64
    lib.jo_decode_bytes.argtypes = [vp, vp, C.c_int64, vp, C.c_int64, C.POINTER(C.c_int32)]
    lib.jo_merge_piece.restype = C.c_int64
    lib.jo_merge_piece.argtypes = [vp, vp, C.c_int64, C.c_int, vp, C.c_int64]
    lib.jo_encode_batch.restype = C.c_int64
    lib.jo_encode_batch.argtypes = [vp, vp, vp, C.c_int64, C.c_int, C.c_int, C.c_int, vp, vp]
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _lib()
    return _LIB


def load_tiktoken(path):
    """EncodingFactory.loadMergeableRanks (EncodingFactory.java:139-164)."""
    ranks = {}
    with open(path, "rb") as f:
        for line in f.read().splitlines():
            if not line:
                continue
            tok, rank = line.split(None, 1)
            ranks[base64.b64decode(tok)] = int(rank)
    return ranks


def flatten(d):
    """dict[bytes,int] -> (uint8 bytes, int64 offsets, int32 values) in insertion order."""
    keys = list(d.keys())
    off = np.zeros(len(keys) + 1, dtype=np.int64)
    if keys:
        off[1:] = np.cumsum([len(k) for k in keys])
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8).copy() if keys else np.zeros(0, dtype=np.uint8)
    vals = np.array([d[k] for k in keys], dtype=np.int32)
    return blob, off, vals


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class OracleEncoding:
    """CPU restatement of GptBytePairEncoding for one (pattern, encoder, special tokens) triple."""

    def __init__(self, name, pattern, flags, ranks, special_tokens):
        self.name = name
        self.ranks = ranks
        self.special_tokens = dict(special_tokens)
        kb, ko, kv = flatten(ranks)
        sb, so, sv = flatten({k.encode("utf-8"): v for k, v in special_tokens.items()})
        err = C.create_string_buffer(256)
        self._h = lib().jo_create(pattern.encode("utf-8"), flags, _p(kb), _p(ko), _p(kv), len(kv), _p(sb), _p(so), _p(sv), len(sv), err, 256)
        if not self._h:
            raise ValueError("oracle: cannot compile pattern: " + err.value.decode())

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.jo_destroy(self._h)
            self._h = None

    @staticmethod
    def builtin(name):
        pat, fname, special = BUILTIN[name]
        return OracleEncoding(name, pat, UNICODE_CHARACTER_CLASS, load_tiktoken(os.path.join(DATA, fname)), special)

    @staticmethod
    def _bytes(text):
        if isinstance(text, str):
            text = text.encode("utf-8", "replace")  # String.getBytes(UTF_8): a lone surrogate becomes '?'
        return np.frombuffer(text, dtype=np.uint8)

    def split(self, text):
        b = self._bytes(text)
        cap = b.size + 2
        st, en = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64)
        n = lib().jo_split(self._h, _p(b), b.size, _p(st), _p(en), cap)
        return list(zip(st[:n].tolist(), en[:n].tolist()))

    def encode_ordinary(self, text, merge=MERGE_AUTO):
        return self._encode(text, 0, merge)

    def encode(self, text, merge=MERGE_AUTO):
        return self._encode(text, 1, merge)

    def _encode(self, text, check_special, merge):
        b = self._bytes(text)
        out = np.empty(b.size + 1, dtype=np.int32)
        n = lib().jo_encode(self._h, _p(b), b.size, check_special, merge, _p(out), out.size)
        if n == E_SPECIAL:
            raise NotImplementedError("Encoding special tokens is not supported yet.")
        if n < 0:
            raise ValueError("oracle encode failed: %d" % n)
        return out[:n].tolist()

    def encode_with_special(self, text, merge=MERGE_AUTO):
        """tiktoken's encode(text, allowed_special="all"): special tokens in the text become their ids (not in the reference)."""
        b = self._bytes(text)
        out = np.empty(b.size + 1, dtype=np.int32)
        n = lib().jo_encode_with_special(self._h, _p(b), b.size, merge, _p(out), out.size)
        if n < 0:
            raise ValueError("oracle encode failed: %d" % n)
        return out[:n].tolist()

    def encode_max(self, text, max_tokens, ordinary=False):
        b = self._bytes(text)
        out = np.empty(b.size + 1, dtype=np.int32)
        trunc = C.c_int(0)
        n = lib().jo_encode_max(self._h, _p(b), b.size, 0 if ordinary else 1, max_tokens, _p(out), out.size, C.byref(trunc))
        if n == E_SPECIAL:
            raise NotImplementedError("Encoding special tokens is not supported yet.")
        if n < 0:
            raise ValueError("oracle encode failed: %d" % n)
        return out[:n].tolist(), bool(trunc.value)

    def decode_bytes(self, ids):
        a = np.asarray(ids, dtype=np.int32)
        cap = max(1, a.size) * 256
        out = np.empty(cap, dtype=np.uint8)
        bad = C.c_int32(0)
        n = lib().jo_decode_bytes(self._h, _p(a), a.size, _p(out), cap, C.byref(bad))
        if n == E_UNKNOWN_ID:
            raise ValueError("Unknown token for decoding: %d" % bad.value)
        if n < 0:
            raise ValueError("oracle decode failed: %d" % n)
        return out[:n].tobytes()

    def merge_piece(self, piece, merge):
        b = np.frombuffer(bytes(piece), dtype=np.uint8)
        out = np.empty(b.size + 1, dtype=np.int32)
        n = lib().jo_merge_piece(self._h, _p(b), b.size, merge, _p(out), out.size)
        if n < 0:
            raise ValueError("oracle merge failed: %d" % n)
        return out[:n].tolist()

    def encode_batch(self, utf8, doc_off, nthreads, check_special=False, merge=MERGE_AUTO):
        """Returns (ids laid out at byte offsets, per-document counts, total)."""
        utf8 = np.ascontiguousarray(utf8, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.int64)
        ids = np.empty(max(1, utf8.size), dtype=np.int32)
        counts = np.empty(max(1, doc_off.size - 1), dtype=np.int64)
        total = lib().jo_encode_batch(self._h, _p(utf8), _p(doc_off), doc_off.size - 1, nthreads, int(check_special), merge, _p(ids), _p(counts))
        return ids, counts[: doc_off.size - 1], total

    def encode_batch_compact(self, utf8, doc_off, nthreads, check_special=False, merge=MERGE_AUTO):
        """Returns (ids int32 concatenated, token offsets int64[ndocs+1], counts incl. negative statuses)."""
        ids, counts, _ = self.encode_batch(utf8, doc_off, nthreads, check_special, merge)
        pos = np.maximum(counts, 0)
        tok_off = np.zeros(counts.size + 1, dtype=np.int64)
        np.cumsum(pos, out=tok_off[1:])
        out = np.empty(int(tok_off[-1]), dtype=np.int32)
        for d in range(counts.size):
            c = int(pos[d])
            if c:
                out[tok_off[d]:tok_off[d] + c] = ids[doc_off[d]:doc_off[d] + c]
        return out, tok_off, counts
