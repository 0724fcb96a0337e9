"""TEST INFRASTRUCTURE ONLY: ctypes wrapper over tests/emu/libjtk_emu.so (host emulation of the tile logic)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class JtkParams(C.Structure):
    _fields_ = [("name", C.c_char_p), ("pattern", C.c_char_p), ("pattern_flags", C.c_int32),
                ("vocab_bytes", C.c_void_p), ("vocab_off", C.c_void_p), ("vocab_ranks", C.c_void_p), ("vocab_size", C.c_int64),
                ("special_bytes", C.c_void_p), ("special_off", C.c_void_p), ("special_ids", C.c_void_p), ("special_size", C.c_int64)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def flatten(d):
    keys = list(d.keys())
    off = np.zeros(len(keys) + 1, dtype=np.int64)
    if keys:
        off[1:] = np.cumsum([len(k) for k in keys])
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8).copy() if keys else np.zeros(0, dtype=np.uint8)
    vals = np.array([d[k] for k in keys], dtype=np.int32)
    return blob, off, vals


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", HERE])
        L = C.CDLL(os.path.join(HERE, "libjtk_emu.so"))
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.POINTER(JtkParams), C.c_char_p, C.c_int]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_run.restype = C.c_int64
        L.emu_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_tile.restype = C.c_int
        L.emu_general_split.restype = C.c_int
        L.emu_general_split.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.emu_pattern_kind.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class EmuEncoding:
    def __init__(self, name, pattern, flags, ranks, special):
        self._keep = flatten(ranks) + flatten({k.encode(): v for k, v in special.items()})
        kb, ko, kv, sb, so, sv = self._keep
        p = JtkParams(name.encode(), pattern.encode(), flags, _p(kb), _p(ko), _p(kv), len(kv), _p(sb), _p(so), _p(sv), len(sv))
        err = C.create_string_buffer(512)
        self._h = lib().emu_create(C.byref(p), err, 512)
        if not self._h:
            raise ValueError(err.value.decode())

    def stats(self):
        out = np.zeros(8, dtype=np.int64)
        lib().emu_stats(self._h, _p(out))
        return out

    def run(self, utf8, doc_off, flags=0, want_ids=True):
        utf8 = np.ascontiguousarray(utf8, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.int64)
        nd = doc_off.size - 1
        pf = np.zeros(max(1, utf8.size), dtype=np.uint8)
        ids = np.zeros(max(1, utf8.size), dtype=np.int32) if want_ids else None
        tok_off = np.zeros(nd + 1, dtype=np.int64)
        status = np.zeros(max(1, nd), dtype=np.int32)
        n = lib().emu_run(self._h, _p(utf8), utf8.size, _p(doc_off), nd, flags, _p(pf), _p(ids) if want_ids else None, _p(tok_off), _p(status))
        if n < 0:
            raise AssertionError("tile halo view disagrees with owner tile: %d" % n)
        return pf[:utf8.size], (ids[:n] if want_ids else None), tok_off, status[:nd]

    def pattern_kind(self):
        return lib().emu_pattern_kind(self._h)

    def dfa_info(self):
        """(states, symbols, reason) of the general pattern's DFA; states == 0 when the pattern keeps the backtracking program."""
        nsym = C.c_int(0)
        why = C.create_string_buffer(256)
        lib().emu_dfa_info.restype = C.c_int
        lib().emu_dfa_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        n = lib().emu_dfa_info(self._h, C.byref(nsym), why, 256)
        return n, nsym.value, why.value.decode()

    def general_split(self, utf8, doc_off, stack_cap=1024, no_dfa=False):
        """(start flags, skip flags) per byte from the general-pattern program; raises on backtrack-stack overflow."""
        utf8 = np.ascontiguousarray(utf8, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.int64)
        start = np.zeros(max(1, utf8.size), dtype=np.uint8)
        skip = np.zeros(max(1, utf8.size), dtype=np.uint8)
        rc = lib().emu_general_split(self._h, _p(utf8), _p(doc_off), doc_off.size - 1, _p(start), _p(skip), stack_cap, 1 if no_dfa else 0)
        if rc != 0:
            raise OverflowError("general split: rc %d" % rc)
        return start[:utf8.size], skip[:utf8.size]
