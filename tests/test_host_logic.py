"""Host-side logic that needs no GPU: static tables of the API mirror, vocabulary loader, synthetic corpus generator."""
import os

import numpy as np
import pytest


def test_enums_mirror_the_reference():
    import jtokkit_b200 as jt
    assert [t.get_name() for t in jt.EncodingType] == ["r50k_base", "p50k_base", "p50k_edit", "cl100k_base"]  # api/EncodingType.java:10-13
    assert jt.EncodingType.from_name("cl100k_base") is jt.EncodingType.CL100K_BASE and jt.EncodingType.from_name("x") is None
    assert len(list(jt.ModelType)) == 33  # api/ModelType.java:11-53
    assert jt.ModelType.GPT_4.get_encoding_type() is jt.EncodingType.CL100K_BASE and jt.ModelType.GPT_4_32K.get_max_context_length() == 32768
    assert jt.ModelType.from_name("text-davinci-003").get_encoding_type() is jt.EncodingType.P50K_BASE
    assert jt.ModelType.from_name("text-davinci-edit-001").get_encoding_type() is jt.EncodingType.P50K_EDIT
    assert jt.ModelType.from_name("gpt-5") is None


def test_load_mergeable_ranks():
    """EncodingFactory.loadMergeableRanks (EncodingFactory.java:139-164) on the three vendored resource files."""
    import jtokkit_b200 as jt
    for fname, n in [("r50k_base.tiktoken", 50256), ("p50k_base.tiktoken", 50280), ("cl100k_base.tiktoken", 100256)]:
        ranks = jt.EncodingFactory.load_mergeable_ranks(fname)
        assert len(ranks) == n
        assert sum(1 for k in ranks if len(k) == 1) == 256
    with pytest.raises(RuntimeError):
        jt.EncodingFactory.load_mergeable_ranks("missing.tiktoken")


def test_encoding_result_and_params_value_types():
    import jtokkit_b200 as jt
    r = jt.EncodingResult([1, 2], True)
    assert r.get_tokens() == [1, 2] and r.is_truncated() and "truncated=true" in repr(r)
    p = jt.GptBytePairEncodingParams("n", jt.Pattern.compile("x", jt.Pattern.UNICODE_CHARACTER_CLASS), {b"a": 1}, {"<s>": 2})
    assert p.get_name() == "n" and p.get_pattern().pattern() == "x" and p.get_pattern().flags() == 0x100
    assert p.get_encoder() == {b"a": 1} and p.get_special_tokens_encoder() == {"<s>": 2}


def test_pack_documents_is_string_getbytes_utf8():
    import jtokkit_b200 as jt
    blob, off = jt.pack_documents(["aé", "", "\ud800x"])  # a lone surrogate becomes '?' (ImmutableByteArray.java:16-19)
    assert bytes(blob) == "aé".encode() + b"?x" and off.tolist() == [0, 3, 3, 5]


def test_synthetic_corpora_have_the_named_shape():
    from jtokkit_b200 import synth
    data, off = synth.config3_multilingual("cpu", total=2 << 20)
    lens = np.diff(off.numpy())
    assert abs(data.numel() - (2 << 20)) < 70000 and off[0] == 0 and off[-1] == data.numel()
    assert lens.min() >= 900 and lens.max() <= 66000
    text = bytes(data.numpy()).decode("utf-8")  # valid UTF-8 throughout
    assert "<|" not in text
    data2, off2 = synth.config3_multilingual("cpu", total=2 << 20)
    assert np.array_equal(data.numpy(), data2.numpy()) and np.array_equal(off.numpy(), off2.numpy())  # deterministic
    chat, coff = synth.config4_chat("cpu", total=1 << 20)
    cl = np.diff(coff.numpy())
    assert 150 < cl.mean() < 400 and b"\n" not in bytes(chat.numpy())
    adv = synth.config5_adversarial(n=4096)
    assert len(adv) == 8 and all(len(b) in (4096, 4095) for b in adv)
