"""GPU parity of the general split-pattern path: registerGptBytePairEncoding with patterns other than the
two predefined ones (AbstractEncodingRegistry.java:63-66, BaseEncodingRegistryTest.java:110-125), against the oracle."""
import os
import random

import numpy as np
import pytest

from test_emu_cpu import GENERAL_PATTERNS, random_docs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cl100k_ranks():
    from oracle import jo
    jo.build()
    return jo.load_tiktoken(os.path.join(jo.DATA, jo.BUILTIN["cl100k_base"][1]))


def make_pair(name, pat, flags, ranks, special=None):
    import jtokkit_b200 as jt
    from oracle import jo
    special = special or {}
    g = jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams(name, jt.Pattern.compile(pat, flags), ranks, special))
    return g, jo.OracleEncoding(name, pat, flags, ranks, special)


@pytest.mark.parametrize("index", range(len(GENERAL_PATTERNS)))
def test_general_patterns_random_documents(index, cl100k_ranks):
    pat, flags = GENERAL_PATTERNS[index]
    g, o = make_pair("general_%d" % index, pat, flags, cl100k_ranks)
    rng = random.Random(500 + index)
    docs = []
    for it in range(400):
        docs += random_docs(rng, 3, ascii_only=(it % 3 == 1), utf8_letters=(it % 3 == 2))
    res = g.encode_ordinary_batch([d.decode("utf-8") for d in docs])
    assert len(res) == len(docs)
    flagged = 0
    for d, ids, st in zip(docs, res.to_lists(), res.doc_status):
        if st & 8:  # JTK_DOC_PATTERN_STACK: only group loops over long runs may hit it
            flagged += 1
            continue
        assert st == 0 and ids == o.encode_ordinary(d), (pat, d)
    assert flagged <= len(docs) // 50


def test_general_pattern_long_gaps_and_long_matches(cl100k_ranks):
    """Gaps and matches longer than a tile's forward halo (1040 bytes) and than a tile (8192 bytes), across tile edges."""
    g, o = make_pair("general_long", r"[a-z]+|\d{1,3}", 0, cl100k_ranks)
    rng = random.Random(3)
    docs = []
    for n_gap, n_word in [(10, 10), (1500, 20), (9000, 3000), (20000, 12000), (3, 40000), (8191, 1), (8192, 2), (1040, 1041)]:
        docs.append(("!" * n_gap + "".join(rng.choice("abcdefgh") for _ in range(n_word)) + " " * n_gap + "12345" + "?" * (n_gap // 2)).encode())
    docs.append(b"")
    docs.append(("#" * 70000).encode())
    res = g.encode_ordinary_batch([d.decode() for d in docs])
    assert not res.doc_status.any()
    for d, ids in zip(docs, res.to_lists()):
        assert ids == o.encode_ordinary(d), len(d)
    counts = g.count_tokens_batch([d.decode() for d in docs])
    assert list(counts) == [len(o.encode_ordinary(d)) for d in docs]


def test_general_pattern_special_tokens_and_split_flags(cl100k_ranks):
    import jtokkit_b200 as jt
    g, o = make_pair("general_special", r"\w+|\s+", 0, cl100k_ranks, {"<|endoftext|>": 100257})
    with pytest.raises(NotImplementedError):
        g.encode("a <|endoftext|> b")
    assert g.encode_ordinary("a <|endoftext|> b") == o.encode_ordinary(b"a <|endoftext|> b")
    assert g.encode("hello, world! 42") == o.encode(b"hello, world! 42")
    assert g.decode(g.encode("hello world")) == "hello world"
    # maxTokens goes through the same matches
    for mt in [0, 1, 2, 5]:
        r = g.encode("hello, wonderful world of tokens", mt)
        exp_ids, exp_trunc = o.encode_max(b"hello, wonderful world of tokens", mt)
        assert list(r.get_tokens()) == exp_ids and r.is_truncated() == exp_trunc


def test_general_pattern_stack_overflow_flags_the_document(cl100k_ranks):
    """A pattern without a DFA form ('$') runs as a backtracking program: a group loop over a long run exhausts its stack and the
    document is flagged (the JVM throws StackOverflowError).  The same loop as a DFA has no stack and simply matches."""
    g, _ = make_pair("general_deep", r"(?:a|b)+c$|.", 0, cl100k_ranks)
    with pytest.raises(RecursionError):
        g.encode_ordinary("ab" * 5000)
    assert g.encode_ordinary("abc") == g.encode_ordinary("abc")
    g2, o2 = make_pair("general_deep_dfa", r"(?:a|b)+c|.", 0, cl100k_ranks)
    for text in ["ab" * 5000, "ab" * 3000 + "c" + "ba" * 10, "abc"]:
        assert g2.encode_ordinary(text) == o2.encode_ordinary(text.encode())


def test_case_insensitive_predefined_pattern(cl100k_ranks):
    """Pattern.CASE_INSENSITIVE on the x50k pattern ('S / 'LL become contractions) takes the general program."""
    from oracle import jo
    pat = jo.BUILTIN["r50k_base"][0]
    g, o = make_pair("x50k_ci", pat, 0x102, cl100k_ranks)
    for text in ["I'LL GO, HE'S here and they'Re THERE'VE", "ſ'ſ x'S", "plain text 123"]:
        assert g.encode_ordinary(text) == o.encode_ordinary(text.encode())
