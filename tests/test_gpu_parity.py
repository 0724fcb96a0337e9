"""First GPU parity tests: golden vectors, fuzz against the oracle, split flags, tile edges, long pieces."""
import os
import random

import numpy as np
import pytest

from conftest import ENCODING_NAMES, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ENCODING_NAMES)
def test_golden_encode(name, gpu_encodings):
    """<Enc>Test.java:19-29 (encodesCorrectly) and :62-72 (encodeOrdinary) as ONE batch through the C ABI."""
    enc = gpu_encodings[name]
    rows = load_golden(name)
    res = enc.encode_batch([r[0] for r in rows])
    res_o = enc.encode_ordinary_batch([r[0] for r in rows])
    bad = [(rows[d][0], res.tokens(d), rows[d][1]) for d in range(len(rows)) if res.tokens(d) != rows[d][1] or res_o.tokens(d) != rows[d][1]]
    assert not bad, bad[:3]
    assert not res.doc_status.any()


@pytest.mark.parametrize("name", ["cl100k_base", "r50k_base"])
def test_fuzz_vs_oracle(name, gpu_encodings, oracles):
    rng = random.Random(11)
    alph = list("abcdefghijklmnopqrstuvwxyzSTREVMLD     \n\n\r\t'''!!?.,;:-0123456789") + [
        "é", "ß", "ſ", "Ж", "я", "中", "文", "あ", "カ", " ", "　", " ", "١", "٢", "½", "🍕", "‍", "️", "्", "ा", "ก", "ั", "한", "😀", "ñ", "—", "“"]
    docs = []
    for _ in range(4000):
        n = rng.choice([0, 1, 5, 40, 200, 700, 3000])
        mode = rng.random()
        if mode < 0.7:
            s = "".join(rng.choice(alph) for _ in range(rng.randint(0, n)))
        elif mode < 0.8:
            s = rng.choice(["1", "\n", " ", "a", "!", "ab", "中", "\n ", " \n", "١"]) * rng.randint(0, n)
        else:
            s = "".join(rng.choice(["the ", " of", "ing", "tion", " 123", "'s", "'ll", "\n\n", "  ", "\t", "Hello", ", ", "中文", "!!!"]) for _ in range(rng.randint(0, n // 3)))
        docs.append(s)
    enc, orc = gpu_encodings[name], oracles[name]
    res = enc.encode_ordinary_batch(docs)
    bad = 0
    for d, s in enumerate(docs):
        exp = orc.encode_ordinary(s)
        if res.tokens(d) != exp:
            bad += 1
            if bad < 4:
                print("MISMATCH", repr(s[:80]), res.tokens(d)[:20], exp[:20])
    assert bad == 0
    assert res.token_offsets[-1] == res.ids.size


def test_split_flags(gpu_encodings, oracles):
    """jtk_split_batch_device vs the oracle's matcher.find() loop on a multi-tile buffer."""
    import ctypes as C
    import torch
    from jtokkit_b200 import _capi
    rng = random.Random(5)
    words = ["the", " of", "ing", "tion", " 123456", "'s", "'ll", "\n\n", "  ", "\t", "Hello", ", ", "中文", "!!!", "\n   ", "x" * 50, " ", "'", "a"]
    docs = ["".join(rng.choice(words) for _ in range(rng.randint(0, 4000))) for _ in range(30)]
    blobs = [d.encode() for d in docs]
    blob = b"".join(blobs)
    off = np.zeros(len(docs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in blobs])
    for name in ["cl100k_base", "r50k_base"]:
        enc, orc = gpu_encodings[name], oracles[name]
        d_in = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
        d_off = torch.from_numpy(off).cuda()
        d_flags = torch.zeros(len(blob), dtype=torch.uint8, device="cuda")
        _capi.check(_capi.lib().jtk_split_batch_device(enc._h, 0, d_in.data_ptr(), len(blob), d_off.data_ptr(), len(docs), d_flags.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream))
        got = d_flags.cpu().numpy()
        exp = np.zeros(len(blob), dtype=np.uint8)
        for d, b in enumerate(blobs):
            for (a, _) in orc.split(b):
                exp[off[d] + a] = 1
        diff = np.nonzero(got != exp)[0]
        assert diff.size == 0, (name, diff[:10], blob[max(0, diff[0] - 20):diff[0] + 20])


def test_long_pieces(gpu_encodings, oracles):
    """Pieces beyond the in-tile limit go through the long-piece kernels (config 5 of BASELINE.json, small)."""
    rng = random.Random(3)
    docs = ["a" * 5000, " " * 4097, "!" * 3000, "ab" * 2500, "\n" * 6000, "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(20000)),
            "hello " + "z" * 1500 + " world " + "中" * 1000 + "。ok", "x" * 1025, "y" * 1024, "1" * 5000]
    for name in ["cl100k_base", "r50k_base"]:
        enc, orc = gpu_encodings[name], oracles[name]
        res = enc.encode_ordinary_batch(docs)
        for d, s in enumerate(docs):
            assert res.tokens(d) == orc.encode_ordinary(s), (name, d, s[:20])


def test_api_single_calls(gpu_encodings, oracles):
    enc = gpu_encodings["cl100k_base"]
    assert enc.encode("hello world") == [15339, 1917]  # api/Encoding.java:18-19
    assert enc.encode_ordinary("hello <|endoftext|> world") == [15339, 83739, 8862, 728, 428, 91, 29, 1917]  # :73-74
    with pytest.raises(NotImplementedError):
        enc.encode("hello <|endoftext|> world")
    assert enc.count_tokens("This is a sample sentence.") == 6
    assert enc.decode([15339, 1917]) == "hello world"
    assert enc.decode_bytes([15339, 1917]) == b"hello world"
    with pytest.raises(ValueError):
        enc.decode([15339, 100261])
    r = enc.encode("This is a sample sentence.", 3)  # usage.md:90-91
    assert r.get_tokens() == [2028, 374, 264] and r.is_truncated()
    r = enc.encode("I love \U0001F355", 4)  # usage.md:96-97
    assert r.get_tokens() == [40, 3021] and r.is_truncated()
    assert enc.encode(None) == [] and enc.encode("") == []


@pytest.mark.parametrize("name", ENCODING_NAMES)
def test_golden_max_tokens(name, gpu_encodings):
    """<Enc>Test.java:39-60: encode(x, 10) tokens == column 3, truncated flag, decoded prefix - all 423 rows like the reference."""
    enc = gpu_encodings[name]
    rows = load_golden(name)
    assert len(rows) == 423
    for inp, full, ten in rows:
        r = enc.encode(inp, 10)
        assert r.get_tokens() == ten, inp
        assert r.is_truncated() == (len(full) > len(ten)), inp
        assert inp.startswith(enc.decode(r.get_tokens()))


def test_long_piece_rounds_fuzz_custom_vocabularies():
    """The long-piece kernel merges many pairs per round (every candidate up to a rank window that survives the (rank, position)
    priority) and checks the created pairs before changing anything.  Small alphabets, random vocabularies - including ones
    in which a concatenation ranks BELOW its parts, where the reference order differs from any naive parallel order - and
    pieces of 1.1 - 40 KiB, against the oracle's literal loop."""
    import random
    import jtokkit_b200 as jt
    from oracle import jo
    rng = random.Random(2024)
    pat = r"\S+|\s+"
    for trial in range(int(os.environ.get("JTK_TEST_TRIALS", "10"))):
        alphabet = [bytes([c]) for c in b"abcdeXY"[:rng.randint(2, 7)]]
        vocab = {b" ": 1000}
        for b in alphabet:
            vocab[b] = len(vocab)
        tokens = list(alphabet)
        for _ in range(rng.randint(3, 60)):
            a, b = rng.choice(tokens), rng.choice(tokens)
            if len(a + b) <= 12 and a + b not in vocab:
                vocab[a + b] = len(vocab)
                tokens.append(a + b)
        ranks = dict(vocab)
        if trial % 2 == 1:  # shuffle the ranks: no longer a valid merge order (longer tokens may rank below their parts)
            keys = [k for k in ranks if k != b" "]
            ids = [ranks[k] for k in keys]
            rng.shuffle(ids)
            ranks.update(zip(keys, ids))
        g = jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("long_fuzz_%d" % trial, jt.Pattern.compile(pat), ranks, {}))
        o = jo.OracleEncoding("long_fuzz_%d" % trial, pat, 0, ranks, {})
        docs = []
        for _ in range(6):
            n = rng.choice([1100, 2500, 9000, 40000])
            mode = rng.random()
            if mode < 0.4:
                s = b"".join(rng.choice(alphabet) for _ in range(n))
            elif mode < 0.7:
                unit = b"".join(rng.choice(alphabet) for _ in range(rng.randint(1, 5)))
                s = unit * (n // len(unit))
            else:
                s = b"".join(rng.choice(tokens) for _ in range(n // 3))
            docs.append(s + b" " + s[: n // 3])
        res = g.encode_ordinary_batch([d.decode() for d in docs])
        assert not res.doc_status.any()
        for d, ids in zip(docs, res.to_lists()):
            assert ids == o.encode_ordinary(d, jo.MERGE_LITERAL if len(d) < 12000 else jo.MERGE_AUTO), (trial, len(d))
