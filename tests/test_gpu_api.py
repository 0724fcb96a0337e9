"""GPU tests of the API mirror: registries (BaseEncodingRegistryTest.java), batch entry points, decode, error behaviour."""
import numpy as np
import pytest

from conftest import ENCODING_NAMES, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def registry():
    import jtokkit_b200 as jt
    return jt.Encodings.new_lazy_encoding_registry()


def test_registry_lookups(registry, gpu_encodings):
    """BaseEncodingRegistryTest.java:34-96: by type, name, model type, model name and the gpt-4 / gpt-3.5-turbo prefix rules."""
    import jtokkit_b200 as jt
    for t in jt.EncodingType:
        enc = registry.get_encoding(t)
        assert enc.get_name() == t.get_name()
        assert registry.get_encoding(t.get_name()) is enc
    for m in jt.ModelType:
        assert registry.get_encoding_for_model(m).get_name() == m.get_encoding_type().get_name()
        assert registry.get_encoding_for_model(m.get_name()).get_name() == m.get_encoding_type().get_name()
    for name, expect in [("gpt-4-32k-0314", "cl100k_base"), ("gpt-4-0314", "cl100k_base"), ("gpt-3.5-turbo-0301", "cl100k_base"),
                         ("gpt-3.5-turbo-16k-0613", "cl100k_base")]:
        assert registry.get_encoding_for_model(name).get_name() == expect
    assert registry.get_encoding("unknown") is None and registry.get_encoding_for_model("unknown-model") is None  # :136-141


def test_registry_registration(registry):
    """:98-134: custom encodings, registerGptBytePairEncoding with the predefined patterns, duplicates -> IllegalStateException."""
    import jtokkit_b200 as jt

    class DummyEncoding:
        def get_name(self):
            return "dummy"

    registry.register_custom_encoding(DummyEncoding())
    assert registry.get_encoding("dummy").get_name() == "dummy"
    with pytest.raises(RuntimeError):
        registry.register_custom_encoding(DummyEncoding())
    p = jt.EncodingFactory.predefined_params(jt.EncodingType.R50K_BASE)
    p.name = "my_r50k"
    registry.register_gpt_byte_pair_encoding(p)
    assert registry.get_encoding("my_r50k").encode("hello world") == registry.get_encoding(jt.EncodingType.R50K_BASE).encode("hello world")
    with pytest.raises(RuntimeError):
        registry.register_gpt_byte_pair_encoding(p)
    # empty maps are legal registration input (BaseEncodingRegistryTest.java:110-125 registers two empty maps)
    empty = jt.GptBytePairEncodingParams("empty_maps", jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE).get_pattern(), {}, {})
    registry.register_gpt_byte_pair_encoding(empty)
    with pytest.raises(ValueError):  # every byte is unknown: IllegalArgumentException from TokenEncoder.encode
        registry.get_encoding("empty_maps").encode("a")
    assert registry.get_encoding("empty_maps").encode("") == []
    # BaseEncodingRegistryTest.java:110-125: an arbitrary pattern with two empty maps registers (general pattern program)
    registry.register_gpt_byte_pair_encoding(jt.GptBytePairEncodingParams("test", jt.Pattern.compile("test"), {}, {}))
    enc = registry.get_encoding("test")
    assert enc.get_name() == "test"
    assert enc.encode("nothing matches here") == [] and enc.encode("") == []
    with pytest.raises(ValueError):  # "test" matches, but no byte of it is in the (empty) vocabulary
        enc.encode("a test")
    # constructs outside the compiler's subset fail at registration, nothing falls back to the CPU
    with pytest.raises(ValueError):
        registry.register_gpt_byte_pair_encoding(jt.GptBytePairEncodingParams("lookbehind", jt.Pattern.compile("(?<=ab)c"), {}, {}))


def test_default_registry_is_eager():
    import jtokkit_b200 as jt
    reg = jt.Encodings.new_default_encoding_registry()
    assert sorted(reg._encodings) == sorted(ENCODING_NAMES)  # DefaultEncodingRegistry.java:16-20
    lazy = jt.Encodings.new_lazy_encoding_registry()
    assert lazy._encodings == {}                             # LazyEncodingRegistryTest.java:17-23
    assert lazy.get_encoding_for_model("gpt-4").get_name() == "cl100k_base" and list(lazy._encodings) == ["cl100k_base"]


@pytest.mark.parametrize("name", ENCODING_NAMES)
def test_golden_roundtrip_and_counts(name, gpu_encodings):
    """<Enc>Test.java:31-37 (decode(encode(x)) == x) as batches; countTokens == encode(x).size() (GptBytePairEncoding.java:121-129)."""
    enc = gpu_encodings[name]
    rows = load_golden(name)
    texts = [r[0] for r in rows]
    res = enc.encode_batch(texts)
    back = enc.decode_bytes_batch(res.to_lists())
    assert [b.decode("utf-8") for b in back] == texts
    counts = enc.count_tokens_batch(texts)
    assert counts.tolist() == [len(r[1]) for r in rows]
    assert enc.count_tokens_ordinary(texts[5]) == len(rows[5][1])


def test_special_token_guard_per_document(gpu_encodings):
    """encodeInternal :52-56 - one bad document does not poison the batch; p50k_edit guards four strings, r50k one."""
    from jtokkit_b200 import _capi
    docs = ["plain", "has <|endoftext|> inside", "<|fim_prefix|>", "almost <|endoftext| >", "", "<|endofprompt|>", "tail <|endoftext|>"]
    expect = {"cl100k_base": [0, 1, 1, 0, 0, 1, 1], "p50k_edit": [0, 1, 1, 0, 0, 0, 1], "r50k_base": [0, 1, 0, 0, 0, 0, 1]}
    for name, exp in expect.items():
        res = gpu_encodings[name].encode_batch(docs)
        assert [int(bool(s & _capi.DOC_HAS_SPECIAL)) for s in res.doc_status] == exp, name
        ordinary = gpu_encodings[name].encode_ordinary_batch(docs)
        assert not ordinary.doc_status.any()
        assert ordinary.to_lists() == res.to_lists()  # the guard only flags; ids are the encodeOrdinary ids
    with pytest.raises(NotImplementedError):
        gpu_encodings["cl100k_base"].count_tokens("x <|fim_suffix|>")


def test_custom_vocabulary_unknown_bytes_and_duplicates(oracles):
    """A vocabulary without all single bytes: parts that are not tokens raise IllegalArgumentException (TokenEncoder.java:64-71);
    pairs still merge through byte parts that are not tokens themselves."""
    import jtokkit_b200 as jt
    from oracle import jo
    pat = jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE).get_pattern()
    vocab = {b"a": 5, b"b": 7, b"ab": 3, b"abc": 1, b" ": 9, b" a": 2, b"xy": 11, b"xyz": 4, b"-5": -5}
    enc = jt.Encoding(jt.GptBytePairEncodingParams("tiny", pat, vocab, {"<s>": 100}))
    orc = jo.OracleEncoding("tiny", pat.pattern(), pat.flags(), vocab, {"<s>": 100})
    for text in ["ab", "abab a", "abc", "aab", " a ab", "xy", "xyz", "xyzxy"]:
        assert enc.encode(text) == orc.encode(text), text
    for text in ["abd", "q", "x"]:
        with pytest.raises(ValueError) as ei:
            enc.encode(text)
        # TokenEncoder.java:67: "Unknown token for encoding: " + Arrays.toString(one byte)
        assert str(ei.value) == "Unknown token for encoding: [%d]" % ord(text[-1])
        with pytest.raises(ValueError):
            orc.encode(text)
    with pytest.raises(ValueError) as ei:
        enc.encode("dq")  # two different unknown bytes: the device reports the document, the payload is not guessed
    assert str(ei.value) == "Unknown token for encoding"
    assert enc.decode([3, 100, 1]) == "ab<s>abc"  # special tokens decode to their string (GptBytePairEncoding.java:307-310)
    with pytest.raises(NotImplementedError):
        enc.encode("a<s>")


def test_ragged_and_empty_batches(gpu_encodings, oracles):
    enc, orc = gpu_encodings["cl100k_base"], oracles["cl100k_base"]
    assert enc.encode_batch([]).to_lists() == []
    assert enc.encode_batch(["", "", ""]).to_lists() == [[], [], []]
    docs = ["", "a", "", "", "hello world " * 800, "", "x"]
    res = enc.encode_batch(docs)
    assert res.to_lists() == [orc.encode(d) for d in docs]
    assert res.token_offsets[0] == 0 and res.token_offsets[-1] == res.ids.size
    # documents that end exactly on tile boundaries (8 KiB) and a batch that ends on one
    docs = ["a" * 8192, "b c " * 2048, "", "d" * 8191, "e"]
    assert enc.encode_batch(docs).to_lists() == [orc.encode(d) for d in docs]
    docs = ["wor d" * 1638 + "xy"]
    assert len(docs[0]) == 8192 and enc.encode_batch(docs).to_lists() == [orc.encode(docs[0])]


def test_max_tokens_variants_match_oracle(gpu_encodings, oracles):
    """encode(text, maxTokens) incl. the back-off loop (:90-100): multi-byte characters split across tokens, limits <= 0."""
    enc, orc = gpu_encodings["cl100k_base"], oracles["cl100k_base"]
    texts = ["I love \U0001F355\U0001F355 pizza", "日本語のテキストです", "नमस्ते दुनिया", "a", "", "x" * 50, "�� mixed �"]
    for t in texts:
        for m in [-1, 0, 1, 2, 3, 5, 8, 1000]:
            r = enc.encode(t, m)
            assert (r.get_tokens(), r.is_truncated()) == orc.encode_max(t, m), (t, m)
            ro = enc.encode_ordinary(t, m)
            assert (ro.get_tokens(), ro.is_truncated()) == orc.encode_max(t, m, ordinary=True)
    r = enc.encode(None, 5)
    assert r.get_tokens() == [] and not r.is_truncated()


def test_special_token_encoding_matches_tiktoken_semantics(gpu_encodings, oracles):
    """special-token ENCODING (absent from the reference, README.md:46) behind its own method, against the
    oracle's restatement of tiktoken's encode(text, allowed_special="all") (itself checked against tiktoken in test_oracle.py)."""
    import random
    rng = random.Random(11)
    for name in ["cl100k_base", "p50k_edit", "r50k_base"]:
        g, o = gpu_encodings[name], oracles[name]
        from oracle import jo
        special = list(jo.BUILTIN[name][2].keys())
        units = special + ["hello", " world", " ", "\n", "<|", "|>", "<|endoftext", "endoftext|>", "<", "12345", "'s", "中文", "<|fim_", "  ", "x" * 300]
        texts = ["".join(rng.choice(units) for _ in range(rng.randint(0, 14))) for _ in range(400)]
        texts += ["", special[0], special[0] * 3, "a" + special[0], special[0] + "b", "tail <|endoftext|>", "x" * 9000 + special[0] + "y" * 9000]
        res = g.encode_with_special_tokens_batch(texts)
        assert not res.doc_status.any()
        for t, ids in zip(texts, res.to_lists()):
            assert ids == o.encode_with_special(t), (name, t[:80])
        assert g.encode_with_special_tokens("hello <|endoftext|> world") == o.encode_with_special("hello <|endoftext|> world")
        # the reference behaviour of encode() is unchanged
        with pytest.raises(NotImplementedError):
            g.encode("hello <|endoftext|> world")
    assert gpu_encodings["cl100k_base"].encode_with_special_tokens("<|endoftext|>") == [100257]


def test_c_abi_builtin_loader_matches_python_loader(gpu_encodings, oracles):
    """jtk_encoding_create_builtin parses each vendored .tiktoken file itself (EncodingFactory.loadMergeableRanks,
    EncodingFactory.java:139-164) and must give the same encoding as the handle built from the Python-side loader:
    same ids on golden inputs + fuzz text, same special-token guard, same decode."""
    import os
    import random
    import jtokkit_b200 as jt
    from conftest import load_golden
    rng = random.Random(5)
    files = {"r50k_base": "r50k_base.tiktoken", "p50k_base": "p50k_base.tiktoken", "p50k_edit": "p50k_base.tiktoken", "cl100k_base": "cl100k_base.tiktoken"}
    units = ["hello", " world", "'s", "'LL", " ", "  ", "\n", "\r\n", "1234567", "日本語", "é", "\U0001F355", "!!!", " x", "\t", "don't", "ſ"]
    for name, fname in files.items():
        enc = jt.Encoding.from_tiktoken_file(name, os.path.join(jt.api.DATA_DIR, fname))
        assert enc.get_name() == name
        texts = [row[0] for row in load_golden(name)[::3]] + ["".join(rng.choice(units) for _ in range(rng.randint(0, 40))) for _ in range(300)]
        got = enc.encode_ordinary_batch(texts)
        ref = gpu_encodings[name].encode_ordinary_batch(texts)
        assert np.array_equal(got.ids, ref.ids) and np.array_equal(got.token_offsets, ref.token_offsets)
        assert got.tokens(5) == oracles[name].encode_ordinary(texts[5])
        # predefined special tokens come from the library's own table (EncodingFactory.java:24-53)
        with pytest.raises(NotImplementedError):
            enc.encode("a <|endoftext|> b")
        assert enc.decode(ref.tokens(7)) == texts[7]
        if name in ("p50k_edit", "cl100k_base"):
            with pytest.raises(NotImplementedError):
                enc.encode("<|fim_middle|>")
        else:
            assert enc.encode("<|fim_middle|>") == gpu_encodings[name].encode("<|fim_middle|>")
        enc.close()  # deferred until `got` (a view of the encoding's pinned buffers) has been released
        got.close()


def test_encode_packed_rejects_inconsistent_offsets(gpu_encodings):
    """The C ABI takes plain pointers; the Python mirror checks the offsets against the buffer before the call."""
    enc = gpu_encodings["cl100k_base"]
    data = np.frombuffer(b"hello world", dtype=np.uint8)
    for off in ([0, 12], [1, 11], [0, 8, 4, 11], []):
        with pytest.raises(ValueError):
            enc.encode_packed(data, np.array(off, dtype=np.int64))
    assert enc.encode_packed(data, np.array([0, 5, 11], dtype=np.int64)).to_lists() == [[15339], [1917]]


def test_device_resident_decode_round_trip_and_unknown_ids(gpu_encodings):
    """jtk_decode_batch_device: decode(encode(x)) == x for a 64 MiB multilingual batch entirely on the device, byte
    offsets per document, unknown ids reported per document with the first offending id (GptBytePairEncoding.java:313), empty
    documents and the size query."""
    import torch
    from jtokkit_b200 import synth, _capi
    enc = gpu_encodings["cl100k_base"]
    data, off = synth.config3_multilingual("cuda", total=64 << 20, seed=21)
    n, nd = data.numel(), off.numel() - 1
    d_in = torch.zeros(n + 80, dtype=torch.uint8, device="cuda")
    d_in[:n] = data
    d_ids = torch.empty(n, dtype=torch.int32, device="cuda")
    d_tok = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
    d_st = torch.zeros(nd + 1, dtype=torch.int32, device="cuda")
    ntok, _, _, _ = enc.encode_device(d_in[:n], off, d_ids, d_tok, d_st)
    total, launches = enc.decode_device(d_ids[:ntok], d_tok, None, None, None, None)
    assert total == n and launches > 0
    d_out = torch.full((n + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    d_boff = torch.empty(nd + 1, dtype=torch.int64, device="cuda")
    d_dst = torch.zeros(nd, dtype=torch.int32, device="cuda")
    d_bad = torch.empty(nd, dtype=torch.int32, device="cuda")
    total, _ = enc.decode_device(d_ids[:ntok], d_tok, d_out[:n], d_boff, d_dst, d_bad)
    assert total == n and torch.equal(d_out[:n], data) and int(d_out[n:].min()) == 0xEE  # nothing written past the end
    assert torch.equal(d_boff, off) and int(d_dst.max()) == 0
    # capacity error
    with pytest.raises(_capi.JtkError) as ei:
        enc.decode_device(d_ids[:ntok], d_tok, d_out[:n - 1], d_boff, d_dst, d_bad)
    assert ei.value.code == _capi.JTK_E_CAPACITY
    # unknown ids: documents 3 and 5 get one each, document 5 two (the first one is reported); special tokens decode to their text
    ids = d_ids[:ntok].clone()
    tok = d_tok.cpu().numpy()
    ids[int(tok[3]) + 2] = 100261
    ids[int(tok[5]) + 1] = 2000000
    ids[int(tok[5]) + 4] = -7
    ids[int(tok[7])] = 100257
    d_dst.zero_()
    total2, _ = enc.decode_device(ids, d_tok, d_out, d_boff, d_dst, d_bad)
    st, bad = d_dst.cpu().numpy(), d_bad.cpu().numpy()
    assert st[3] == _capi.DOC_UNKNOWN_ID and bad[3] == 100261 and st[5] == _capi.DOC_UNKNOWN_ID and bad[5] == 2000000
    assert st.sum() == 2 * _capi.DOC_UNKNOWN_ID
    b7 = int(d_boff[7])
    assert bytes(d_out[b7:b7 + 13].cpu().numpy()) == b"<|endoftext|>"
    # empty batch, empty documents
    z = torch.zeros(4, dtype=torch.int32, device="cuda")
    t0 = torch.zeros(4, dtype=torch.int64, device="cuda")
    total3, _ = enc.decode_device(z[:0], t0, d_out, d_boff, d_dst, d_bad)
    assert total3 == 0 and int(d_boff[:4].abs().max()) == 0
