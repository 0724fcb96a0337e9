"""ctypes binding of include/jtokkit_b200.h (the same entry points a Panama FFM / JNI shim binds).

The library must be present: there is no Python or CPU fallback.  Import fails loudly when
jtokkit_b200/libjtokkit_b200.so has not been built (python -c "import __graft_entry__ as g; g.build()").
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JTK_LIB", os.path.join(HERE, "libjtokkit_b200.so"))  # JTK_LIB: development override for A/B builds

JTK_OK, JTK_E_ARG, JTK_E_CUDA, JTK_E_PATTERN_UNSUPPORTED, JTK_E_NOMEM, JTK_E_CAPACITY = 0, -1, -2, -3, -4, -5
DOC_OK, DOC_HAS_SPECIAL, DOC_UNKNOWN_BYTES, DOC_UNKNOWN_ID, DOC_PATTERN_STACK = 0, 1, 2, 4, 8
ENCODE_ORDINARY, CHECK_SPECIAL, COUNT_ONLY, TIME_KERNEL = 0, 1, 2, 0x100
RE_CASE_INSENSITIVE, RE_UNICODE_CASE, RE_UNICODE_CHARACTER_CLASS = 0x02, 0x40, 0x100


class JtkParams(C.Structure):
    _fields_ = [("name", C.c_char_p), ("pattern", C.c_char_p), ("pattern_flags", C.c_int32),
                ("vocab_bytes", C.c_void_p), ("vocab_off", C.c_void_p), ("vocab_ranks", C.c_void_p), ("vocab_size", C.c_int64),
                ("special_bytes", C.c_void_p), ("special_off", C.c_void_p), ("special_ids", C.c_void_p), ("special_size", C.c_int64)]


class JtkDeviceInfo(C.Structure):
    _fields_ = [("num_tokens", C.c_int64), ("num_long_pieces", C.c_int64), ("gpu_launches", C.c_int64), ("reserved", C.c_int32),
                ("tile_kernel_ms", C.c_float)]


# every symbol include/jtokkit_b200.h declares: (restype, argtypes)
vp, i64, i32, u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32
SIGNATURES = {
    "jtk_encoding_create": (C.c_int, [C.POINTER(JtkParams), vp, C.c_int, C.POINTER(vp)]),
    "jtk_encoding_create_builtin": (C.c_int, [C.c_char_p, C.c_char_p, vp, C.c_int, C.POINTER(vp)]),
    "jtk_encoding_destroy": (None, [vp]),
    "jtk_encoding_name": (C.c_char_p, [vp]),
    "jtk_encoding_num_devices": (C.c_int, [vp]),
    "jtk_encode_batch": (C.c_int, [vp, vp, vp, i64, u32, C.POINTER(vp)]),
    "jtk_plan_chunks": (i64, [vp, i64, C.c_int, i64, vp, i64]),
    "jtk_encode_batch_special": (C.c_int, [vp, vp, vp, i64, u32, C.POINTER(vp)]),
    "jtk_result_num_docs": (i64, [vp]),
    "jtk_result_num_tokens": (i64, [vp]),
    "jtk_result_ids": (vp, [vp]),
    "jtk_result_token_offsets": (vp, [vp]),
    "jtk_result_doc_status": (vp, [vp]),
    "jtk_result_device_ms": (C.c_double, [vp]),
    "jtk_result_gpu_launches": (i64, [vp]),
    "jtk_result_free": (None, [vp]),
    "jtk_encode_batch_device": (C.c_int, [vp, C.c_int, vp, i64, vp, i64, u32, vp, i64, vp, vp, vp, C.POINTER(JtkDeviceInfo)]),
    "jtk_split_batch_device": (C.c_int, [vp, C.c_int, vp, i64, vp, i64, vp, vp]),
    "jtk_decode_batch": (C.c_int, [vp, vp, vp, i64, C.POINTER(vp)]),
    "jtk_decode_batch_device": (C.c_int, [vp, C.c_int, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(i64)]),
    "jtk_result_bytes": (vp, [vp]),
    "jtk_result_byte_offsets": (vp, [vp]),
    "jtk_result_bad_ids": (vp, [vp]),
    "jtk_encode_max_tokens": (C.c_int, [vp, vp, i64, i32, u32, C.POINTER(vp), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
    "jtk_free": (None, [vp]),
    "jtk_host_alloc": (vp, [i64]),
    "jtk_host_free": (None, [vp]),
    "jtk_last_error": (C.c_char_p, []),
    "jtk_version": (C.c_char_p, []),
}

_LIB = None


def lib():
    """Loads libjtokkit_b200.so; raises ImportError when it is missing (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("jtokkit_b200: %s is missing - build it with __graft_entry__.build() "
                              "(make -C jtokkit_b200/csrc); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def last_error():
    return lib().jtk_last_error().decode("utf-8", "replace")


class JtkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("jtokkit_b200 error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != JTK_OK:
        raise JtkError(rc, last_error())
