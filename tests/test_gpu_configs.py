"""GPU parity on the BASELINE.json configurations at their full sizes where the oracle finishes in seconds (config 2: 64 MiB,
config 4: over a million strings, config 5: all eight classes at 1 MiB; config 3 at 128 MiB here, its complete 1 GiB corpus is
compared inside bench.py) plus size-independent properties at larger sizes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def compare_with_oracle(enc, orc, data_np, off_np, ordinary, threads=None):
    res = enc.encode_packed(data_np, off_np, ordinary=ordinary)
    ids, tok_off, counts = orc.encode_batch_compact(data_np, off_np, threads or os.cpu_count() or 1, check_special=not ordinary)
    assert not res.doc_status.any()
    assert np.array_equal(res.token_offsets, tok_off)
    assert np.array_equal(res.ids, ids)
    return res


@pytest.mark.parametrize("name", ["r50k_base", "p50k_base"])
def test_config2_english_encode_ordinary(name, gpu_encodings, oracles):
    """configs[1] at full size: r50k_base / p50k_base encodeOrdinary on 64 MiB of synthetic English in 1 024 documents, bit-exact."""
    from jtokkit_b200 import synth
    data, off = synth.config2_english_64mib("cuda", total=64 << 20)
    assert data.numel() >= (64 << 20) - 65536 and off.numel() - 1 >= 1000
    compare_with_oracle(gpu_encodings[name], oracles[name], data.cpu().numpy(), off.cpu().numpy(), ordinary=True)


def test_config3_multilingual_encode(gpu_encodings, oracles):
    """configs[2]: cl100k_base encode of the multilingual corpus (128 MiB here, the first eighth of the bench corpus), bit-exact
    incl. document token offsets."""
    from jtokkit_b200 import synth
    data, off = synth.config3_multilingual("cuda", total=128 << 20)
    res = compare_with_oracle(gpu_encodings["cl100k_base"], oracles["cl100k_base"], data.cpu().numpy(), off.cpu().numpy(), ordinary=False)
    # round trip on the device decode path: decode(encode(x)) == x for the whole batch
    enc = gpu_encodings["cl100k_base"]
    sample = list(range(0, len(res), 97))
    back = enc.decode_bytes_batch([res.tokens(d) for d in sample])
    raw = bytes(data.cpu().numpy())
    offn = off.cpu().numpy()
    assert back == [raw[offn[d]:offn[d + 1]] for d in sample]


def test_config4_count_tokens_short_strings(gpu_encodings, oracles):
    """configs[3]: countTokens-only on short chat-length strings: 288 MiB = more than a million strings (a tenth of the
    10 M of BASELINE.json; every string's count is compared)."""
    from jtokkit_b200 import synth
    data, off = synth.config4_chat("cuda", total=288 << 20)
    d, o = data.cpu().numpy(), off.cpu().numpy()
    assert o.size - 1 >= 1_000_000
    res = gpu_encodings["cl100k_base"].encode_packed(d, o, count_only=True)
    _, _, counts = oracles["cl100k_base"].encode_batch_compact(d, o, os.cpu_count() or 1, check_special=True)
    assert res.ids is None
    assert np.array_equal(res.counts(), counts)


@pytest.mark.parametrize("name", ["cl100k_base", "r50k_base"])
def test_config5_adversarial_long_pieces(name, gpu_encodings, oracles):
    """configs[4] at full size: all eight classes of whitespace-free / repeated-byte 1 MiB documents (plus the 64 KiB versions);
    expected ids from the oracle's exact heap merge (the literal O(n^2) loop is checked against it in test_oracle)."""
    from jtokkit_b200 import synth
    from oracle import jo
    docs = synth.config5_adversarial(n=1 << 16) + synth.config5_adversarial(n=1 << 20)
    assert len(docs) == 16 and all(len(x.encode("utf-8") if isinstance(x, str) else x) >= (1 << 20) - 3 for x in docs[8:])
    enc, orc = gpu_encodings[name], oracles[name]
    res = enc.encode_ordinary_batch(docs)
    for d, doc in enumerate(docs):
        assert res.tokens(d) == orc.encode_ordinary(doc, jo.MERGE_HEAP), (name, d, len(doc))
    assert enc.decode_bytes_batch([res.tokens(0)])[0] == docs[0]


def test_multi_chunk_pipeline_matches_single_chunk(oracles):
    """The host-buffer call pipelines chunks of whole documents; chunking must not change anything."""
    import jtokkit_b200 as jt
    from jtokkit_b200 import synth
    data, off = synth.config3_multilingual("cuda", total=12 << 20, seed=99)
    d, o = data.cpu().numpy(), off.cpu().numpy()
    ref = jt.EncodingFactory.cl100k_base().encode_packed(d, o)
    os.environ["JTK_CHUNK_MB"] = "1"
    try:
        chunked = jt.EncodingFactory.cl100k_base().encode_packed(d, o)
    finally:
        del os.environ["JTK_CHUNK_MB"]
    assert np.array_equal(ref.ids, chunked.ids) and np.array_equal(ref.token_offsets, chunked.token_offsets)


def test_device_resident_call_and_properties_at_size():
    """jtk_encode_batch_device on 256 MiB resident in HBM: equals the host-buffer call; size-independent properties hold:
    token offsets are sorted, every id is a vocabulary id, decode(ids) reproduces the input bytes (checksum of checksums)."""
    import torch
    import jtokkit_b200 as jt
    from jtokkit_b200 import synth
    enc = jt.EncodingFactory.cl100k_base()
    data, off = synth.config3_multilingual("cuda", total=256 << 20, seed=5)
    n = data.numel()
    d_in = torch.zeros(n + 80, dtype=torch.uint8, device="cuda")
    d_in[:n] = data
    d_ids = torch.empty(n, dtype=torch.int32, device="cuda")
    d_tok = torch.empty(off.numel(), dtype=torch.int64, device="cuda")
    d_st = torch.zeros(off.numel(), dtype=torch.int32, device="cuda")
    ntok, nlong, launches, _ = enc.encode_device(d_in[:n], off, d_ids, d_tok, d_st)
    assert launches > 0 and int(d_st.max()) == 0
    tok = d_tok.cpu().numpy()
    assert tok[0] == 0 and tok[-1] == ntok and np.all(np.diff(tok) >= 0)
    ids = d_ids[:ntok]
    assert int(ids.min()) >= 0 and int(ids.max()) <= 100255
    host = enc.encode_packed(data.cpu().numpy(), off.cpu().numpy())
    assert np.array_equal(host.ids, ids.cpu().numpy()) and np.array_equal(host.token_offsets, tok)
    # total decoded length == input length and a strided sample of documents decodes to its bytes
    lens = torch.tensor([len(k) for k in jt.EncodingFactory.load_mergeable_ranks("cl100k_base.tiktoken").keys()])
    ranks = torch.tensor(list(jt.EncodingFactory.load_mergeable_ranks("cl100k_base.tiktoken").values()))
    table = torch.zeros(100256, dtype=torch.int64)
    table[ranks] = lens
    assert int(table.cuda()[ids.long()].sum()) == n
    raw, offn = bytes(data.cpu().numpy()), off.cpu().numpy()
    sample = list(range(0, offn.size - 1, 1999))
    back = enc.decode_bytes_batch([host.tokens(d) for d in sample])
    assert back == [raw[offn[d]:offn[d + 1]] for d in sample]


def test_result_buffer_estimate_too_small_is_recovered(oracles):
    """The one pinned result buffer of a batch is sized from the densest chunk seen so far (440 tokens per KiB on a fresh handle);
    a batch with one token per byte overruns it: the chunks that do not fit are kept aside and merged at the end.  Same ids."""
    import jtokkit_b200 as jt
    doc = ("1 2 3 4 5 6 7 8 9 0 " * 3277)[:65536]
    docs = [doc] * 64  # 4 MiB, ~4 M tokens
    os.environ["JTK_CHUNK_MB"] = "1"
    try:
        enc = jt.EncodingFactory.cl100k_base()
    finally:
        del os.environ["JTK_CHUNK_MB"]
    res = enc.encode_batch(docs)
    exp = oracles["cl100k_base"].encode(doc)
    assert len(exp) == 65536
    assert res.ids.size == 64 * 65536 and np.array_equal(res.token_offsets, np.arange(65) * 65536)
    for d in (0, 15, 16, 17, 40, 63):
        assert res.tokens(d) == exp
    again = enc.encode_batch(docs)  # the estimate has adapted: the direct path
    assert np.array_equal(again.ids, res.ids)
