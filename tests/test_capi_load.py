"""The C-ABI library loads and exports every symbol include/jtokkit_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "jtokkit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jtk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from jtokkit_b200 import _capi
    lib = _capi.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(_capi.SIGNATURES.keys()) == names  # the ctypes binding covers the whole header
    assert b"sm_100a" in lib.jtk_version() and b"15.0.0" in lib.jtk_version()


def test_header_is_plain_c():
    """extern "C", plain pointers and sizes, no torch / CUDA types in the signatures."""
    text = open(os.path.join(ROOT, "include", "jtokkit_b200.h")).read()
    assert 'extern "C"' in text
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for banned in ["torch", "cudaStream_t", "at::", "std::", "#include <cuda"]:
        assert banned not in code
    import subprocess
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "jtokkit_b200.h")])


def test_registration_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device registration returns JTK_E_CUDA (on the GPU box this test is skipped)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import jtokkit_b200 as jt
    from jtokkit_b200 import _capi
    with pytest.raises(_capi.JtkError) as ei:
        jt.EncodingFactory.cl100k_base()
    assert ei.value.code == _capi.JTK_E_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_unsupported_pattern_is_rejected_at_registration():
    """Patterns outside the device pattern compiler's subset fail at registration (they never run on the CPU instead)."""
    import jtokkit_b200 as jt
    for pat in [r"(?<=ab)c", r"(?<!a|bc)d", r"(a)\1", r"(?<n>a)\k<n>", r"(?:a*)*", r"[a-z&&[^b]", r"\p{IsKlingon}+", r"\p{InGreek}", r"a{2,1}", r"(", r"x{17}y(?:ab){65}", r"\Ga", r"\X",
                r"(?>a+)b"]:
        params = jt.GptBytePairEncodingParams("custom", jt.Pattern.compile(pat), {b"a": 0}, {})
        with pytest.raises(ValueError):
            jt.EncodingFactory.from_parameters(params)


def test_unicode_properties_and_word_boundaries_compile():
    """What round 1 rejected and a JTokkit user can legally register (AbstractEncodingRegistry.java:63-66): general categories, \\w / \\d
    under UNICODE_CHARACTER_CLASS (the factory's own flag, EncodingFactory.java:129), \\b.  Registration gets as far as the device."""
    import torch
    import jtokkit_b200 as jt
    from jtokkit_b200 import _capi
    for pat, flags in [(r"\p{Lu}+|\p{Ll}+|.", 0), (r"\w+|\d+|\s+", 0x100), (r"\bword\b|.", 0), (r"\b\w+\b|\W", 0x100), (r"[\p{IsAlphabetic}\p{Mn}]+|\P{L}", 0x100),
                       # round 2: named groups, \A \Z \z, \Q..\E, \h \v, scripts, one-character look-behind
                       (r"(?<w>\w+)|\s+|.", 0), (r"\A\w+|\w+\z|\w+\Z|.", 0), (r"\Q1+1\E|\h+|\v|.", 0), (r"\p{IsLatin}+|\p{script=Han}+|\p{sc=Cyrl}+|.", 0),
                       (r"(?<=a)b|(?<![0-9])[0-9]+|.", 0), (r"[a-z&&[^b]]+|[\p{L}&&[^\p{IsHan}]]|[0-9[x-z]]|.", 0), (r"\R|(?:ab){20}|.", 0)]:
        params = jt.GptBytePairEncodingParams("custom", jt.Pattern.compile(pat, flags), {b"a": 0}, {})
        if torch.cuda.is_available():
            jt.EncodingFactory.from_parameters(params).close()
        else:
            with pytest.raises(_capi.JtkError) as ei:
                jt.EncodingFactory.from_parameters(params)
            assert ei.value.code == _capi.JTK_E_CUDA  # the pattern compiled; only the device is missing


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under jtokkit_b200/ may import, include, link or load it."""
    loaders = ("import", "#include", "dlopen", "CDLL", "subprocess", "-l")
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jtokkit_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                for line in open(os.path.join(dirpath, f), errors="replace"):
                    if "oracle" in line.lower():
                        assert not any(tok in line for tok in loaders), (f, line)


def test_duplicate_ids_are_rejected_at_registration():
    """Two different byte sequences sharing one id: the device merge loop identifies parts by id, so registration refuses the
    vocabulary (JTK_E_ARG -> ValueError) instead of merging pairs the reference would not (the check runs before any
    device is touched, so it is testable without a GPU).  The same key twice is Map.put semantics and stays legal."""
    import jtokkit_b200 as jt
    pat = jt.Pattern.compile(jt.api.CL100K_PATTERN, jt.Pattern.UNICODE_CHARACTER_CLASS)
    vocab = {b"a": 0, b"b": 1, b"c": 2, b"d": 3, b"e": 4, b"ab": 5, b"cd": 5, b"abe": 6}
    with pytest.raises(ValueError) as ei:
        jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("dup", pat, vocab, {}))
    assert "two different byte sequences" in str(ei.value)
    # a single-byte id colliding with a multi-byte id is the same defect
    with pytest.raises(ValueError):
        jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("dup2", pat, {b"a": 0, b"b": 1, b"ab": 1}, {}))


def test_c_abi_builtin_loader_argument_errors():
    """jtk_encoding_create_builtin (EncodingFactory.java:139-164 for non-JVM callers): unknown names and unreadable files are
    argument errors with the reference's message shape; a good file gets as far as the device check."""
    import torch
    import jtokkit_b200 as jt
    from jtokkit_b200 import _capi
    good = os.path.join(ROOT, "jtokkit_b200", "data", "cl100k_base.tiktoken")
    with pytest.raises(ValueError) as ei:
        jt.Encoding.from_tiktoken_file("no_such_base", good)
    assert "unknown predefined encoding" in str(ei.value)
    with pytest.raises(ValueError) as ei:
        jt.Encoding.from_tiktoken_file("cl100k_base", os.path.join(ROOT, "does", "not", "exist.tiktoken"))
    assert "Could not find" in str(ei.value)
    bad = os.path.join(ROOT, "tests", "golden", "cl100k_base_encodings.csv")  # not a .tiktoken file
    with pytest.raises(ValueError):
        jt.Encoding.from_tiktoken_file("cl100k_base", bad)
    if not torch.cuda.is_available():
        with pytest.raises(_capi.JtkError) as ei:
            jt.Encoding.from_tiktoken_file("cl100k_base", good)
        assert ei.value.code == _capi.JTK_E_CUDA  # parsed fine, then: no device, no CPU fallback
