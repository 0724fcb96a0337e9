"""Host-side mirror of JTokkit's public API over the C ABI.

The reference's host language is Java and this image has no JVM, so the operator interface of the hot path is
mirrored here in Python with the reference's names, argument meaning and error behaviour (camelCase aliases
are provided so the parity tests read like the reference's JUnit tests):

  com.knuddels.jtokkit.api.Encoding            (api/Encoding.java:29,61,80,107,127,147,164,181,189)   -> Encoding
  com.knuddels.jtokkit.api.EncodingResult      (api/EncodingResult.java:8-38)                          -> EncodingResult
  com.knuddels.jtokkit.api.EncodingRegistry    (api/EncodingRegistry.java:20-67)                       -> EncodingRegistry
  com.knuddels.jtokkit.api.GptBytePairEncodingParams (api/GptBytePairEncodingParams.java:36-62)        -> GptBytePairEncodingParams
  com.knuddels.jtokkit.api.EncodingType / ModelType                                                     -> EncodingType / ModelType
  com.knuddels.jtokkit.Encodings               (Encodings.java:13-30)                                  -> Encodings
  com.knuddels.jtokkit.EncodingFactory         (EncodingFactory.java:60-137)                           -> EncodingFactory

Java exceptions map to Python ones: UnsupportedOperationException -> NotImplementedError,
IllegalArgumentException -> ValueError, IllegalStateException -> RuntimeError, NullPointerException -> KeyError.
Everything that touches text runs on the GPU through libjtokkit_b200.so; there is no CPU fallback.
"""
import base64
import ctypes as C
import enum
import os
import threading

import numpy as np

from . import _capi

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_LIVE_LOCK = threading.Lock()  # guards the per-encoding count of results that still own pinned buffers of the encoding


# ----------------------------------------------------------------------------- enums (static tables of the reference)
class EncodingType(enum.Enum):
    R50K_BASE = "r50k_base"
    P50K_BASE = "p50k_base"
    P50K_EDIT = "p50k_edit"
    CL100K_BASE = "cl100k_base"

    def get_name(self):
        return self.value

    getName = get_name

    @staticmethod
    def from_name(name):
        for t in EncodingType:
            if t.value == name:
                return t
        return None

    fromName = from_name


_MODELS = [
    # chat
    ("GPT_4", "gpt-4", "CL100K_BASE", 8192), ("GPT_4_32K", "gpt-4-32k", "CL100K_BASE", 32768),
    ("GPT_3_5_TURBO", "gpt-3.5-turbo", "CL100K_BASE", 4097), ("GPT_3_5_TURBO_16K", "gpt-3.5-turbo-16k", "CL100K_BASE", 16384),
    # text
    ("TEXT_DAVINCI_003", "text-davinci-003", "P50K_BASE", 4097), ("TEXT_DAVINCI_002", "text-davinci-002", "P50K_BASE", 4097),
    ("TEXT_DAVINCI_001", "text-davinci-001", "R50K_BASE", 2049), ("TEXT_CURIE_001", "text-curie-001", "R50K_BASE", 2049),
    ("TEXT_BABBAGE_001", "text-babbage-001", "R50K_BASE", 2049), ("TEXT_ADA_001", "text-ada-001", "R50K_BASE", 2049),
    ("DAVINCI", "davinci", "R50K_BASE", 2049), ("CURIE", "curie", "R50K_BASE", 2049), ("BABBAGE", "babbage", "R50K_BASE", 2049),
    ("ADA", "ada", "R50K_BASE", 2049),
    # code
    ("CODE_DAVINCI_002", "code-davinci-002", "P50K_BASE", 8001), ("CODE_DAVINCI_001", "code-davinci-001", "P50K_BASE", 8001),
    ("CODE_CUSHMAN_002", "code-cushman-002", "P50K_BASE", 2048), ("CODE_CUSHMAN_001", "code-cushman-001", "P50K_BASE", 2048),
    ("DAVINCI_CODEX", "davinci-codex", "P50K_BASE", 4096), ("CUSHMAN_CODEX", "cushman-codex", "P50K_BASE", 2048),
    # edit
    ("TEXT_DAVINCI_EDIT_001", "text-davinci-edit-001", "P50K_EDIT", 3000), ("CODE_DAVINCI_EDIT_001", "code-davinci-edit-001", "P50K_EDIT", 3000),
    # embeddings
    ("TEXT_EMBEDDING_ADA_002", "text-embedding-ada-002", "CL100K_BASE", 8191),
    # old embeddings
    ("TEXT_SIMILARITY_DAVINCI_001", "text-similarity-davinci-001", "R50K_BASE", 2046),
    ("TEXT_SIMILARITY_CURIE_001", "text-similarity-curie-001", "R50K_BASE", 2046),
    ("TEXT_SIMILARITY_BABBAGE_001", "text-similarity-babbage-001", "R50K_BASE", 2046),
    ("TEXT_SIMILARITY_ADA_001", "text-similarity-ada-001", "R50K_BASE", 2046),
    ("TEXT_SEARCH_DAVINCI_DOC_001", "text-search-davinci-doc-001", "R50K_BASE", 2046),
    ("TEXT_SEARCH_CURIE_DOC_001", "text-search-curie-doc-001", "R50K_BASE", 2046),
    ("TEXT_SEARCH_BABBAGE_DOC_001", "text-search-babbage-doc-001", "R50K_BASE", 2046),
    ("TEXT_SEARCH_ADA_DOC_001", "text-search-ada-doc-001", "R50K_BASE", 2046),
    ("CODE_SEARCH_BABBAGE_CODE_001", "code-search-babbage-code-001", "R50K_BASE", 2046),
    ("CODE_SEARCH_ADA_CODE_001", "code-search-ada-code-001", "R50K_BASE", 2046),
]


class _ModelTypeMixin:
    """api/ModelType.java:11-53: model name -> encoding type + maximum context length."""

    def get_name(self):
        return self.value[0]

    def get_encoding_type(self):
        return EncodingType[self.value[1]]

    def get_max_context_length(self):
        return self.value[2]

    getName, getEncodingType, getMaxContextLength = get_name, get_encoding_type, get_max_context_length


ModelType = enum.Enum("ModelType", [(k, (n, e, c)) for k, n, e, c in _MODELS], module=__name__, qualname="ModelType", type=_ModelTypeMixin)


def _model_from_name(name):
    """ModelType.fromName (api/ModelType.java:108-110); None stands for Optional.empty()."""
    for m in ModelType:
        if m.value[0] == name:
            return m
    return None


ModelType.from_name = staticmethod(_model_from_name)
ModelType.fromName = staticmethod(_model_from_name)


# ----------------------------------------------------------------------------- value types
class EncodingResult:
    """api/EncodingResult.java:8-38."""

    def __init__(self, tokens, truncated):
        self.tokens = list(tokens)
        self.truncated = bool(truncated)

    def get_tokens(self):
        return self.tokens

    def is_truncated(self):
        return self.truncated

    getTokens, isTruncated = get_tokens, is_truncated

    def __repr__(self):
        return "EncodingResult{tokens=%r, truncated=%s}" % (self.tokens, str(self.truncated).lower())


class Pattern:
    """The part of java.util.regex.Pattern that GptBytePairEncodingParams carries: source + flag bits."""
    CASE_INSENSITIVE, UNICODE_CASE, UNICODE_CHARACTER_CLASS = 0x02, 0x40, 0x100

    def __init__(self, pattern, flags=0):
        self._pattern, self._flags = pattern, flags

    @staticmethod
    def compile(pattern, flags=0):
        return Pattern(pattern, flags)

    def pattern(self):
        return self._pattern

    def flags(self):
        return self._flags


class GptBytePairEncodingParams:
    """api/GptBytePairEncodingParams.java:36-62: name, pattern, Map<byte[],Integer> encoder, Map<String,Integer> special tokens."""

    def __init__(self, name, pattern, encoder, special_tokens_encoder):
        self.name = name
        self.pattern = pattern if isinstance(pattern, Pattern) else Pattern(pattern)
        self.encoder = encoder
        self.special_tokens_encoder = special_tokens_encoder

    def get_name(self):
        return self.name

    def get_pattern(self):
        return self.pattern

    def get_encoder(self):
        return self.encoder

    def get_special_tokens_encoder(self):
        return self.special_tokens_encoder

    getName, getPattern, getEncoder, getSpecialTokensEncoder = get_name, get_pattern, get_encoder, get_special_tokens_encoder


# ----------------------------------------------------------------------------- batch result
class _ResultOwner:
    """Owns a C jtk_result: its pinned buffers go back to the library's pool when the last array viewing them is gone
    (or earlier, on BatchResult.close())."""

    def __init__(self, handle, encoding):
        self.handle = handle
        self.encoding = encoding  # the result's buffers go back to the encoding's pool: keep it alive until then

    def free(self):
        if self.handle is not None:
            h, self.handle = self.handle, None
            _capi.lib().jtk_result_free(h)
            enc, self.encoding = self.encoding, None
            if enc is not None:
                enc._result_released()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _PinnedView:
    """Array-interface window onto one buffer of a jtk_result; numpy keeps it (and through it the owner) as the array's base."""

    def __init__(self, owner, ptr, n, typestr):
        self.owner = owner
        self.__array_interface__ = {"data": (ptr or 0, False), "shape": (n,), "typestr": typestr, "version": 3}


def _view(owner, ptr, n, typestr, dtype):
    return np.asarray(_PinnedView(owner, ptr, n, typestr)) if n and ptr else np.zeros(0, dtype=dtype)


class BatchResult:
    """Result of Encoding.encode_batch: primitive arrays, no boxing (ids int32, token offsets int64, per-document status)."""

    def __init__(self, ids, token_offsets, doc_status, device_ms, gpu_launches):
        self.ids = ids
        self.token_offsets = token_offsets
        self.doc_status = doc_status
        self.device_ms = device_ms
        self.gpu_launches = gpu_launches
        self._owner = None  # zero-copy results: the C result the arrays are views of

    def close(self):
        """Returns the pinned result buffers to the library right away (zero-copy results only; the arrays must not be used
        afterwards).  Without it they go back when the last array viewing them is garbage collected."""
        if self._owner is not None:
            o, self._owner = self._owner, None
            self.ids = self.token_offsets = self.doc_status = None
            o.free()

    def __len__(self):
        return self.token_offsets.size - 1

    def tokens(self, d):
        return self.ids[self.token_offsets[d]:self.token_offsets[d + 1]].tolist()

    def counts(self):
        return np.diff(self.token_offsets)

    def to_lists(self):
        return [self.tokens(d) for d in range(len(self))]


def _flatten(keys, values):
    off = np.zeros(len(keys) + 1, dtype=np.int64)
    if keys:
        off[1:] = np.cumsum([len(k) for k in keys])
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8).copy() if keys else np.zeros(0, dtype=np.uint8)
    vals = np.array(values, dtype=np.int64)
    if vals.size and (vals.min() < -2 ** 31 or vals.max() > 2 ** 31 - 1):
        raise ValueError("token ids must fit a Java int")
    return blob, off, vals.astype(np.int32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _utf8(text):
    """String.getBytes(UTF_8) (ImmutableByteArray.java:16-19): a lone surrogate becomes '?'."""
    if isinstance(text, (bytes, bytearray, memoryview)):
        return bytes(text)
    return text.encode("utf-8", "replace")


def pack_documents(texts):
    """List of str/bytes -> (uint8 array, int64 offsets[n+1]); the flattening a Java shim does before the FFI call."""
    enc = [_utf8(t) for t in texts]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        off[1:] = np.cumsum([len(b) for b in enc])
    blob = np.frombuffer(b"".join(enc), dtype=np.uint8) if enc else np.zeros(0, dtype=np.uint8)
    return blob, off


class Encoding:
    """GPU-backed implementation of api/Encoding.java; replaces com.knuddels.jtokkit.GptBytePairEncoding."""

    def __init__(self, params, devices=None):
        if not isinstance(params, GptBytePairEncodingParams):
            raise TypeError("params must be GptBytePairEncodingParams")
        self._name = params.get_name()
        self._special = dict(params.get_special_tokens_encoder())
        keys = [bytes(k) for k in params.get_encoder().keys()]
        self._single = {k[0] for k in keys if len(k) == 1}
        kb, ko, kv = _flatten(keys, list(params.get_encoder().values()))
        sk = [k.encode("utf-8", "replace") for k in self._special.keys()]
        sb, so, sv = _flatten(sk, list(self._special.values()))
        p = _capi.JtkParams(self._name.encode("utf-8"), params.get_pattern().pattern().encode("utf-8"), params.get_pattern().flags(),
                            _ptr(kb), _ptr(ko), _ptr(kv), len(kv), _ptr(sb), _ptr(so), _ptr(sv), len(sv))
        devs = np.array(devices, dtype=np.int32) if devices else None
        h = C.c_void_p()
        rc = _capi.lib().jtk_encoding_create(C.byref(p), _ptr(devs), 0 if devs is None else devs.size, C.byref(h))
        if rc == _capi.JTK_E_PATTERN_UNSUPPORTED:
            raise ValueError("unsupported split pattern (no CPU fallback): " + _capi.last_error())
        if rc == _capi.JTK_E_ARG:
            raise ValueError("invalid encoding parameters: " + _capi.last_error())
        _capi.check(rc)
        self._h = h
        self.devices = list(devices) if devices else [0]

    @classmethod
    def from_tiktoken_file(cls, name, tiktoken_path, devices=None):
        """A predefined encoding created entirely by the C ABI (jtk_encoding_create_builtin): the library parses the
        .tiktoken file (EncodingFactory.loadMergeableRanks, EncodingFactory.java:139-164) and supplies the predefined pattern and
        special tokens for `name` - the path a non-JVM, non-Python caller uses."""
        self = cls.__new__(cls)
        self._name = name
        self._special = dict(_PREDEFINED[EncodingType.from_name(name)][2]) if EncodingType.from_name(name) else {}
        devs = np.array(devices, dtype=np.int32) if devices else None
        h = C.c_void_p()
        rc = _capi.lib().jtk_encoding_create_builtin(name.encode("utf-8"), os.fsencode(tiktoken_path), _ptr(devs), 0 if devs is None else devs.size, C.byref(h))
        if rc == _capi.JTK_E_ARG:
            raise ValueError("invalid encoding parameters: " + _capi.last_error())
        _capi.check(rc)
        self._h = h
        self.devices = list(devices) if devices else [0]
        return self

    def close(self):
        """Destroys the C handle.  Results that still view the encoding's pinned buffers keep it alive: the handle is destroyed
        when the last of them has been released (jtk_encoding_destroy requires every jtk_result to be freed first)."""
        with _LIVE_LOCK:
            self._closing = True
            if getattr(self, "_live", 0) > 0:
                return
            h, self._h = getattr(self, "_h", None), None
        if h:
            _capi.lib().jtk_encoding_destroy(h)

    def _result_released(self):
        with _LIVE_LOCK:
            self._live = getattr(self, "_live", 0) - 1
            destroy = getattr(self, "_closing", False) and self._live <= 0
            h = getattr(self, "_h", None) if destroy else None
            if destroy:
                self._h = None
        if h:
            _capi.lib().jtk_encoding_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ the new batch entry points
    def encode_batch(self, texts, ordinary=False, count_only=False):
        """encodeAll(Encoding, List<String>) of the JMH harness (AbstractBenchmark.java:37) as one device batch."""
        blob, off = pack_documents(texts)
        return self.encode_packed(blob, off, ordinary=ordinary, count_only=count_only)

    def encode_packed(self, utf8, doc_off, ordinary=False, count_only=False, copy=False, with_special_tokens=False):
        """HOST arrays in, HOST arrays out; host<->device copies happen inside the C call.
        The result arrays are views into the library's pinned result buffers (no copy: copying 400 MB of ids costs more than
        encoding them); the buffers return to the library's pool when the arrays are garbage collected or on
        BatchResult.close().  copy=True returns independent numpy arrays instead.  Pass pinned input (torch pin_memory,
        jtk_host_alloc) for full PCIe speed: pageable memory is copied in at ~8 GB/s by the driver.
        with_special_tokens: special tokens in the text become their ids (jtk_encode_batch_special; not in the reference)."""
        utf8 = np.ascontiguousarray(utf8, dtype=np.uint8)
        doc_off = np.ascontiguousarray(doc_off, dtype=np.int64)
        # the C ABI takes plain pointers: it cannot know the length of the byte buffer, so the bounds are checked here
        if doc_off.ndim != 1 or doc_off.size < 1 or utf8.ndim != 1:
            raise ValueError("doc_off must be a 1-D array of ndocs + 1 offsets and utf8 a 1-D byte array")
        if int(doc_off[0]) != 0 or int(doc_off[-1]) > utf8.size or (doc_off.size > 1 and bool((np.diff(doc_off) < 0).any())):
            raise ValueError("doc_off must start at 0, be non-decreasing and end at or before len(utf8)")
        flags = (0 if ordinary or with_special_tokens else _capi.CHECK_SPECIAL) | (_capi.COUNT_ONLY if count_only else 0)
        r = C.c_void_p()
        call = _capi.lib().jtk_encode_batch_special if with_special_tokens else _capi.lib().jtk_encode_batch
        _capi.check(call(self._h, _ptr(utf8), _ptr(doc_off), doc_off.size - 1, flags, C.byref(r)))
        L = _capi.lib()
        nd, nt = L.jtk_result_num_docs(r), L.jtk_result_num_tokens(r)
        with _LIVE_LOCK:
            self._live = getattr(self, "_live", 0) + 1
        owner = _ResultOwner(r, self)
        ids = None if count_only else _view(owner, L.jtk_result_ids(r), nt, "<i4", np.int32)
        tok_off = _view(owner, L.jtk_result_token_offsets(r), nd + 1, "<i8", np.int64)
        status = _view(owner, L.jtk_result_doc_status(r), nd, "<i4", np.int32)
        res = BatchResult(ids, tok_off, status, L.jtk_result_device_ms(r), L.jtk_result_gpu_launches(r))
        if copy:
            res.ids = None if ids is None else ids.copy()
            res.token_offsets, res.doc_status = tok_off.copy(), status.copy()
            del ids, tok_off, status
            owner.free()
        else:
            res._owner = owner
        return res

    def encode_device(self, d_utf8, d_doc_off, d_ids, d_tok_off, d_status, ordinary=False, count_only=False, time_kernel=False, device=None):
        """Device-resident batch: torch CUDA tensors in and out (uint8 bytes, int64 offsets, int32 ids, int64 token
        offsets, int32 status), enqueued on torch's current stream.  Returns (num_tokens, num_long_pieces, gpu_launches,
        tile_kernel_ms)."""
        import torch
        dev = d_utf8.device.index if device is None else device
        if d_doc_off.numel() < 1 or d_tok_off.numel() < d_doc_off.numel() or (d_status is not None and d_status.numel() < d_doc_off.numel() - 1):
            raise ValueError("d_doc_off needs ndocs + 1 entries, d_tok_off at least as many, d_status at least ndocs")
        flags = (0 if ordinary else _capi.CHECK_SPECIAL) | (_capi.COUNT_ONLY if count_only else 0) | (_capi.TIME_KERNEL if time_kernel else 0)
        info = _capi.JtkDeviceInfo()
        _capi.check(_capi.lib().jtk_encode_batch_device(
            self._h, dev, d_utf8.data_ptr(), d_utf8.numel(), d_doc_off.data_ptr(), d_doc_off.numel() - 1, flags,
            d_ids.data_ptr() if d_ids is not None else None, d_ids.numel() if d_ids is not None else 0, d_tok_off.data_ptr(),
            d_status.data_ptr() if d_status is not None else None, torch.cuda.current_stream(dev).cuda_stream, C.byref(info)))
        return info.num_tokens, info.num_long_pieces, info.gpu_launches, info.tile_kernel_ms

    def decode_device(self, d_ids, d_tok_off, d_out, d_byte_off, d_status, d_bad_ids, device=None):
        """Device-resident batch decodeBytes (jtk_decode_batch_device): torch CUDA tensors in and out (int32 ids, int64 token offsets
        -> uint8 bytes, int64 byte offsets, int32 status / first unknown id per document), on torch's current stream.
        d_out=None asks for the byte count only.  Returns (total_bytes, gpu_launches)."""
        import torch
        dev = d_ids.device.index if device is None else device
        total, launches = C.c_int64(0), C.c_int64(0)
        nd = d_tok_off.numel() - 1
        if d_out is not None and (d_byte_off.numel() < nd + 1 or d_status.numel() < nd or d_bad_ids.numel() < nd):
            raise ValueError("d_byte_off needs ndocs + 1 entries, d_status and d_bad_ids ndocs")
        rc = _capi.lib().jtk_decode_batch_device(
            self._h, dev, d_ids.data_ptr(), d_ids.numel(), d_tok_off.data_ptr(), nd, d_out.data_ptr() if d_out is not None else None,
            d_out.numel() if d_out is not None else 0, d_byte_off.data_ptr() if d_out is not None else None,
            d_status.data_ptr() if d_out is not None else None, d_bad_ids.data_ptr() if d_out is not None else None,
            torch.cuda.current_stream(dev).cuda_stream, C.byref(total), C.byref(launches))
        _capi.check(rc)
        return total.value, launches.value

    def encode_ordinary_batch(self, texts):
        return self.encode_batch(texts, ordinary=True)

    # Special-token ENCODING is not part of the reference (README.md:46 "not started"; Encoding.encode throws
    # UnsupportedOperationException instead), hence the methods of their own.  Semantics: tiktoken's
    # encode(text, allowed_special="all") - every occurrence of a registered special token becomes its id.
    def encode_with_special_tokens_batch(self, texts):
        blob, off = pack_documents(texts)
        return self.encode_packed(blob, off, with_special_tokens=True)

    def encode_with_special_tokens(self, text):
        if text is None:
            return []
        res = self.encode_with_special_tokens_batch([text])
        self._raise_for_status(res.doc_status)
        return res.tokens(0)

    encodeWithSpecialTokens = encode_with_special_tokens

    def count_tokens_batch(self, texts, ordinary=False):
        res = self.encode_batch(texts, ordinary=ordinary, count_only=True)
        self._raise_for_status(res.doc_status)
        return res.counts()

    def decode_bytes_batch(self, token_lists):
        """Batch decodeBytes; raises ValueError for an unknown id like the reference."""
        flat = np.array([t for toks in token_lists for t in toks], dtype=np.int64)
        if flat.size and (flat.min() < -2 ** 31 or flat.max() > 2 ** 31 - 1):
            bad = int(flat[(flat < -2 ** 31) | (flat > 2 ** 31 - 1)][0])
            raise ValueError("Unknown token for decoding: %d" % bad)
        flat = flat.astype(np.int32)
        off = np.zeros(len(token_lists) + 1, dtype=np.int64)
        if token_lists:
            off[1:] = np.cumsum([len(t) for t in token_lists])
        r = C.c_void_p()
        _capi.check(_capi.lib().jtk_decode_batch(self._h, _ptr(flat), _ptr(off), len(token_lists), C.byref(r)))
        L = _capi.lib()
        try:
            nd = len(token_lists)
            boff = np.ctypeslib.as_array(C.cast(L.jtk_result_byte_offsets(r), C.POINTER(C.c_int64)), shape=(nd + 1,)).copy()
            status = np.ctypeslib.as_array(C.cast(L.jtk_result_doc_status(r), C.POINTER(C.c_int32)), shape=(max(nd, 1),))[:nd].copy()
            bad = np.ctypeslib.as_array(C.cast(L.jtk_result_bad_ids(r), C.POINTER(C.c_int32)), shape=(max(nd, 1),))[:nd].copy()
            total = int(boff[-1])
            data = bytes(np.ctypeslib.as_array(C.cast(L.jtk_result_bytes(r), C.POINTER(C.c_uint8)), shape=(max(total, 1),))[:total])
        finally:
            L.jtk_result_free(r)
        for d in range(nd):
            if status[d] & _capi.DOC_UNKNOWN_ID:
                raise ValueError("Unknown token for decoding: %d" % int(bad[d]))
        return [data[boff[d]:boff[d + 1]] for d in range(nd)]

    def _raise_for_status(self, status, text=None):
        if status.size and (status & _capi.DOC_HAS_SPECIAL).any():
            raise NotImplementedError("Encoding special tokens is not supported yet.")  # UnsupportedOperationException
        if status.size and (status & _capi.DOC_UNKNOWN_BYTES).any():
            # IllegalArgumentException, TokenEncoder.java:67: "Unknown token for encoding: " + Arrays.toString(part), the part being ONE byte
            # (every part that is not a token is a single byte).  The device reports the document, not the part: the payload is given
            # when only one byte value without a single-byte token occurs in the text (then it is that byte), else left out, never guessed.
            absent = {b for b in _utf8(text)} - getattr(self, "_single", set(range(256))) if text is not None else set()
            if len(absent) == 1:
                b = absent.pop()
                raise ValueError("Unknown token for encoding: [%d]" % (b - 256 if b > 127 else b))
            raise ValueError("Unknown token for encoding")
        if status.size and (status & _capi.DOC_PATTERN_STACK).any():
            # general split patterns only: java.util.regex would die with StackOverflowError on such a text
            raise RecursionError("split pattern exhausted the device backtracking stack")

    # ------------------------------------------------------------------ api/Encoding.java
    def encode(self, text, max_tokens=None):
        """Encoding.encode(String) / encode(String, int maxTokens)."""
        return self._encode(text, max_tokens, ordinary=False)

    def encode_ordinary(self, text, max_tokens=None):
        """Encoding.encodeOrdinary(String) / encodeOrdinary(String, int maxTokens)."""
        return self._encode(text, max_tokens, ordinary=True)

    def _encode(self, text, max_tokens, ordinary):
        if text is None:  # GptBytePairEncoding.java:48-50,72-74
            return [] if max_tokens is None else EncodingResult([], False)
        if max_tokens is None:
            res = self.encode_batch([text], ordinary=ordinary)
            self._raise_for_status(res.doc_status, text)
            return res.tokens(0)
        b = np.frombuffer(_utf8(text), dtype=np.uint8)
        ids, n, trunc, st = C.c_void_p(), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        flags = 0 if ordinary else _capi.CHECK_SPECIAL
        _capi.check(_capi.lib().jtk_encode_max_tokens(self._h, _ptr(b), b.size, int(max_tokens), flags, C.byref(ids), C.byref(n), C.byref(trunc),
                                                      C.byref(st)))
        try:
            self._raise_for_status(np.array([st.value], dtype=np.int32), text)
            toks = np.ctypeslib.as_array(C.cast(ids, C.POINTER(C.c_int32)), shape=(max(n.value, 1),))[:n.value].tolist()
        finally:
            _capi.lib().jtk_free(ids)
        return EncodingResult(toks, bool(trunc.value))

    def count_tokens(self, text):
        """Encoding.countTokens: encode(text).size(), GptBytePairEncoding.java:121-124."""
        if text is None:
            return 0
        return int(self.count_tokens_batch([text])[0])

    def count_tokens_ordinary(self, text):
        if text is None:
            return 0
        return int(self.count_tokens_batch([text], ordinary=True)[0])

    def decode_bytes(self, tokens):
        """Encoding.decodeBytes, GptBytePairEncoding.java:136-151."""
        return self.decode_bytes_batch([list(tokens)])[0]

    def decode(self, tokens):
        """Encoding.decode: new String(decodeBytes(tokens), UTF_8) - malformed bytes become U+FFFD."""
        return self.decode_bytes(tokens).decode("utf-8", "replace")

    def get_name(self):
        return self._name

    # camelCase aliases of the Java interface
    encodeOrdinary, countTokens, countTokensOrdinary, decodeBytes, getName = encode_ordinary, count_tokens, count_tokens_ordinary, decode_bytes, get_name


# ----------------------------------------------------------------------------- factory + registries
X50K_PATTERN = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"
CL100K_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|"
                  r"\s+(?!\S)|\s+")
_X50K_SPECIAL = {"<|endoftext|>": 50256}
_P50K_EDIT_SPECIAL = {"<|endoftext|>": 50256, "<|fim_prefix|>": 50281, "<|fim_middle|>": 50282, "<|fim_suffix|>": 50283}
_CL100K_SPECIAL = {"<|endoftext|>": 100257, "<|fim_prefix|>": 100258, "<|fim_middle|>": 100259, "<|fim_suffix|>": 100260,
                   "<|endofprompt|>": 100276}
_PREDEFINED = {
    EncodingType.R50K_BASE: (X50K_PATTERN, "r50k_base.tiktoken", _X50K_SPECIAL),
    EncodingType.P50K_BASE: (X50K_PATTERN, "p50k_base.tiktoken", _X50K_SPECIAL),
    EncodingType.P50K_EDIT: (X50K_PATTERN, "p50k_base.tiktoken", _P50K_EDIT_SPECIAL),
    EncodingType.CL100K_BASE: (CL100K_PATTERN, "cl100k_base.tiktoken", _CL100K_SPECIAL),
}


class EncodingFactory:
    """EncodingFactory.java:60-164."""
    devices = None  # devices new encodings are replicated on (None = device 0)

    @staticmethod
    def load_mergeable_ranks(file_name):
        """loadMergeableRanks (:139-164): '<base64 token> <rank>' per line."""
        path = file_name if os.path.isabs(file_name) else os.path.join(DATA_DIR, os.path.basename(file_name))
        if not os.path.exists(path):
            raise RuntimeError("Could not find " + file_name + " in resources")
        ranks = {}
        with open(path, "rb") as f:
            for line in f.read().splitlines():
                if not line:
                    continue
                parts = line.split(None, 1)
                if len(parts) != 2:
                    raise RuntimeError("Invalid line in " + file_name + ": " + line.decode("utf-8", "replace"))
                ranks[base64.b64decode(parts[0])] = int(parts[1])
        return ranks

    @staticmethod
    def predefined_params(encoding_type):
        pat, fname, special = _PREDEFINED[encoding_type]
        return GptBytePairEncodingParams(encoding_type.get_name(), Pattern.compile(pat, Pattern.UNICODE_CHARACTER_CLASS),
                                         EncodingFactory.load_mergeable_ranks(fname), dict(special))

    @staticmethod
    def from_parameters(parameters, devices=None):
        """fromParameters (:117-119): the seam where the engine is chosen - here the CUDA engine."""
        return Encoding(parameters, devices=devices if devices is not None else EncodingFactory.devices)

    @staticmethod
    def r50k_base():
        return EncodingFactory.from_parameters(EncodingFactory.predefined_params(EncodingType.R50K_BASE))

    @staticmethod
    def p50k_base():
        return EncodingFactory.from_parameters(EncodingFactory.predefined_params(EncodingType.P50K_BASE))

    @staticmethod
    def p50k_edit():
        return EncodingFactory.from_parameters(EncodingFactory.predefined_params(EncodingType.P50K_EDIT))

    @staticmethod
    def cl100k_base():
        return EncodingFactory.from_parameters(EncodingFactory.predefined_params(EncodingType.CL100K_BASE))

    fromParameters, r50kBase, p50kBase, p50kEdit, cl100kBase = from_parameters, r50k_base, p50k_base, p50k_edit, cl100k_base


class EncodingRegistry:
    """AbstractEncodingRegistry.java:14-96 (ConcurrentHashMap<String, Encoding> + lookup rules)."""

    def __init__(self):
        self._encodings = {}
        self._lock = threading.Lock()

    def _lookup(self, name):
        return self._encodings.get(name)

    def get_encoding(self, key):
        if isinstance(key, EncodingType):
            self._before_lookup(key)
            enc = self._lookup(key.get_name())
            if enc is None:
                raise KeyError("No encoding registered for encoding type " + key.get_name())
            return enc
        return self._lookup(key)  # Optional.empty() -> None

    def get_encoding_for_model(self, key):
        if isinstance(key, ModelType):
            self._before_lookup(key.get_encoding_type())
            enc = self._lookup(key.get_encoding_type().get_name())
            if enc is None:
                raise KeyError("No encoding registered for model type " + key.get_name())
            return enc
        model = ModelType.from_name(key)
        if model is not None:
            return self.get_encoding_for_model(model)
        for prefix in (ModelType.GPT_4_32K, ModelType.GPT_4, ModelType.GPT_3_5_TURBO_16K, ModelType.GPT_3_5_TURBO):
            if key.startswith(prefix.get_name()):
                return self.get_encoding_for_model(prefix)
        return None

    def register_gpt_byte_pair_encoding(self, parameters):
        return self.register_custom_encoding(EncodingFactory.from_parameters(parameters))

    def register_custom_encoding(self, encoding):
        name = encoding.get_name() if hasattr(encoding, "get_name") else encoding.getName()
        with self._lock:
            if name in self._encodings:
                raise RuntimeError("Encoding " + name + " already registered")  # IllegalStateException
            self._encodings[name] = encoding
        return self

    def _add_encoding(self, encoding_type):
        with self._lock:
            if encoding_type.get_name() not in self._encodings:
                self._encodings[encoding_type.get_name()] = EncodingFactory.from_parameters(EncodingFactory.predefined_params(encoding_type))

    def _before_lookup(self, encoding_type):
        pass

    getEncoding, getEncodingForModel = get_encoding, get_encoding_for_model
    registerGptBytePairEncoding, registerCustomEncoding = register_gpt_byte_pair_encoding, register_custom_encoding


class DefaultEncodingRegistry(EncodingRegistry):
    """DefaultEncodingRegistry.java:16-20: all predefined encodings, eagerly."""

    def initialize_default_encodings(self):
        for t in EncodingType:
            self._add_encoding(t)


class LazyEncodingRegistry(EncodingRegistry):
    """LazyEncodingRegistry.java:17-34: predefined encodings are created on first use."""

    def _before_lookup(self, encoding_type):
        self._add_encoding(encoding_type)

    def get_encoding(self, key):
        if isinstance(key, str):
            t = EncodingType.from_name(key)
            if t is not None:
                self._add_encoding(t)
        return super().get_encoding(key)

    getEncoding = get_encoding


class Encodings:
    """Encodings.java:13-30."""

    @staticmethod
    def new_default_encoding_registry():
        r = DefaultEncodingRegistry()
        r.initialize_default_encodings()
        return r

    @staticmethod
    def new_lazy_encoding_registry():
        return LazyEncodingRegistry()

    newDefaultEncodingRegistry, newLazyEncodingRegistry = new_default_encoding_registry, new_lazy_encoding_registry
