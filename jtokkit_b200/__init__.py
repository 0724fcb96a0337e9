"""jtokkit_b200 - B200-native implementation of JTokkit's encode hot path behind JTokkit's own API.

Host-side mirror of the reference interface (api.py) over the C ABI in include/jtokkit_b200.h
(libjtokkit_b200.so: hand-written sm_100a kernels).  No CPU fallback.
"""
from .api import (BatchResult, DefaultEncodingRegistry, Encoding, EncodingFactory, EncodingRegistry, EncodingResult, Encodings,  # noqa: F401
                  EncodingType, GptBytePairEncodingParams, LazyEncodingRegistry, ModelType, Pattern, pack_documents)

__all__ = ["BatchResult", "DefaultEncodingRegistry", "Encoding", "EncodingFactory", "EncodingRegistry", "EncodingResult", "Encodings",
           "EncodingType", "GptBytePairEncodingParams", "LazyEncodingRegistry", "ModelType", "Pattern", "pack_documents"]
