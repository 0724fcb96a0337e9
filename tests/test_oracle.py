"""The oracle (oracle/) against the reference's own golden vectors and against independent implementations.

This is what pins the oracle: tests/golden/*_encodings.csv are the reference's fixtures
(lib/src/test/resources, asserted by lib/src/test/java/com/knuddels/jtokkit/reference/*Test.java:19-111)."""
import os
import random

import numpy as np
import pytest

from conftest import ENCODING_NAMES, load_golden


@pytest.mark.parametrize("name", ENCODING_NAMES)
def test_golden_rows(name, oracles):
    """encode == column 2, encode(x, 10) == column 3 + truncated flag, decode(encode(x)) == x, encodeOrdinary likewise."""
    from oracle import jo
    o = oracles[name]
    rows = load_golden(name)
    assert len(rows) == 423
    for inp, full, ten in rows:
        assert o.encode(inp, jo.MERGE_LITERAL) == full
        assert o.encode_ordinary(inp, jo.MERGE_HEAP) == full
        got10, trunc = o.encode_max(inp, 10)
        assert got10 == ten and trunc == (len(full) > len(ten))
        got10o, trunco = o.encode_max(inp, 10, ordinary=True)
        assert got10o == ten and trunco == trunc
        assert o.decode_bytes(full).decode("utf-8") == inp
        assert inp.startswith(o.decode_bytes(got10).decode("utf-8"))


def test_documented_known_answers(oracles):
    o = oracles["cl100k_base"]
    assert o.encode("This is a sample sentence.") == [2028, 374, 264, 6205, 11914, 13]  # README.md:83-84
    assert o.encode("hello world") == [15339, 1917]  # api/Encoding.java:18-19
    assert o.encode_ordinary("hello <|endoftext|> world") == [15339, 83739, 8862, 728, 428, 91, 29, 1917]  # :73-74
    with pytest.raises(NotImplementedError):
        o.encode("hello <|endoftext|> world")
    assert o.encode_max("I love \U0001F355", 4) == ([40, 3021], True)  # usage.md:96-97
    assert o.encode_max("This is a sample sentence.", 3) == ([2028, 374, 264], True)  # usage.md:90-91
    assert o.decode_bytes([15339, 1917]) == b"hello world"  # api/Encoding.java:170-171
    with pytest.raises(ValueError):
        o.decode_bytes([15339, 100261])
    assert o.encode_max("abc", 0) == ([], True) and o.encode_max("", 5) == ([], False) and o.encode_max("abc", -3) == ([], True)
    for special in ["<|endoftext|>", "<|fim_prefix|>", "<|endofprompt|>"]:  # <Enc>Test.java:105-111
        assert o.decode_bytes(o.encode_ordinary(special)).decode() == special


@pytest.mark.parametrize("name", ["cl100k_base", "r50k_base"])
def test_heap_merge_equals_literal_loop(name, oracles):
    """The sub-quadratic merge used for long pieces is the same function as the reference's O(n^2) loop."""
    from oracle import jo
    o = oracles[name]
    rng = random.Random(4)
    vocab = [k for k in list(o.ranks.keys())[:20000] if k]
    for _ in range(1500):
        mode = rng.random()
        if mode < 0.3:
            piece = bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(1, 120)))
        elif mode < 0.6:
            piece = b"".join(rng.choice(vocab) for _ in range(rng.randint(1, 30)))
        elif mode < 0.8:
            piece = bytes(rng.randrange(256) for _ in range(rng.randint(1, 80)))
        else:
            piece = rng.choice([b"a", b"ab", b" ", b"\n", b"!", b"abc"]) * rng.randint(1, 200)
        assert o.merge_piece(piece, jo.MERGE_HEAP) == o.merge_piece(piece, jo.MERGE_LITERAL), piece


ALPH = list("abcdefghijklmnopqrstuvwxyzSTREVMLD   \n\r\t'''!!?.,;:-0123456789") + [
    "é", "ß", "ſ", "Ж", "я", "中", "文", "あ", "カ", " ", "　", " ", "١", "٢", "½", "🍕", "‍", "️", "्", "ा", "ก", "ั", "한", "😀", "ñ", "—", "“"]


@pytest.mark.parametrize("name", ["cl100k_base", "r50k_base"])
def test_differential_vs_tiktoken_and_regex(name, oracles):
    """Second opinions: tiktoken (the upstream JTokkit mirrors, benchmark/bench.py:16-28) built from the same vocabulary file and
    the reference's regex string, and the `regex` module for the split alone.  Code points are limited to old, stable ones."""
    tiktoken = pytest.importorskip("tiktoken")
    regex = pytest.importorskip("regex")
    from tiktoken.load import load_tiktoken_bpe
    from oracle import jo
    pat, fname, _ = jo.BUILTIN[name]
    tk = tiktoken.Encoding(name, pat_str=pat, mergeable_ranks=load_tiktoken_bpe(os.path.join(jo.DATA, fname)), special_tokens={})
    rx = regex.compile(pat)
    o = oracles[name]
    rng = random.Random(1)
    for _ in range(6000):
        s = "".join(rng.choice(ALPH) for _ in range(rng.randint(0, 40)))
        assert o.encode_ordinary(s, jo.MERGE_LITERAL) == tk.encode_ordinary(s), s
        cum = [0]
        for ch in s:
            cum.append(cum[-1] + len(ch.encode()))
        assert o.split(s) == [(cum[m.start()], cum[m.end()]) for m in rx.finditer(s)], s


@pytest.mark.parametrize("name", ["cl100k_base", "r50k_base"])
def test_special_token_encoding_vs_tiktoken(name, oracles):
    """jo_encode_with_special (used to check jtk_encode_batch_special; the reference has no special-token encoding) restates
    tiktoken's encode(text, allowed_special="all")."""
    tiktoken = pytest.importorskip("tiktoken")
    from tiktoken.load import load_tiktoken_bpe
    from oracle import jo
    pat, fname, special = jo.BUILTIN[name]
    tk = tiktoken.Encoding(name, pat_str=pat, mergeable_ranks=load_tiktoken_bpe(os.path.join(jo.DATA, fname)), special_tokens=special)
    o = oracles[name]
    rng = random.Random(3)
    units = list(special.keys()) + ["hello", " world", " ", "\n", "<|", "|>", "<|endoftext", "endoftext|>", "<", "12345", "'s", "中文", "<|fim_", "  "]
    for _ in range(2000):
        s = "".join(rng.choice(units) for _ in range(rng.randint(0, 12)))
        assert o.encode_with_special(s) == tk.encode(s, allowed_special="all"), s


def test_custom_patterns_follow_java_regex_semantics(oracles):
    """registerGptBytePairEncoding accepts arbitrary patterns (BaseEncodingRegistryTest.java:110-125): the oracle's matcher
    skips unmatched characters, advances past empty matches and honours ordered alternation."""
    from oracle import jo
    o = jo.OracleEncoding("t", "test", 0, {}, {})
    assert o.split("a test of tests") == [(2, 6), (10, 14)]
    assert o.encode_ordinary("no match here") == []
    o2 = jo.OracleEncoding("t2", "a|ab|b*", 0, {b"a": 0, b"b": 1, b"ab": 2}, {})
    assert o2.split("abb") == [(0, 1), (1, 3), (3, 3)]
    assert o2.encode_ordinary("abb") == [0, 1, 1]
    o3 = jo.OracleEncoding("t3", r"(?i:x+)(?!y)|\s+", 0, {}, {})
    assert o3.split("XXy xx") == [(0, 1), (3, 4), (4, 6)]
    o4 = jo.OracleEncoding("t4", r"(?<=a)b|.", 0, {}, {})  # look-behind over one character
    assert o4.split("abb") == [(0, 1), (1, 2), (2, 3)] and o4.split("ab") == [(0, 1), (1, 2)]
    with pytest.raises(ValueError):
        jo.OracleEncoding("bad", r"(?<=ab)c", 0, {}, {})


def test_batch_thread_pool_matches_single_calls(oracles):
    o = oracles["cl100k_base"]
    docs = ["hello world", "", "I love \U0001F355", "x" * 700, "a <|endoftext|> b", "tabs\tand\nnewlines\r\n  indented"]
    blobs = [d.encode() for d in docs]
    off = np.zeros(len(docs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in blobs])
    ids, tok_off, counts = o.encode_batch_compact(np.frombuffer(b"".join(blobs), dtype=np.uint8), off, 4, check_special=True)
    for d, s in enumerate(docs):
        if "<|endoftext|>" in s:
            assert counts[d] == -1
        else:
            assert ids[tok_off[d]:tok_off[d + 1]].tolist() == o.encode(s)


def test_java_utf16_decoding_of_truncated_sequences():
    """new String(bytes, UTF_8): a truncated trailing sequence is ONE U+FFFD (what the back-off loop compares)."""
    import ctypes as C
    from oracle import jo
    lib = jo.lib()
    lib.jo_java_utf8_to_utf16.restype = C.c_int64
    lib.jo_java_utf8_to_utf16.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]

    def dec(b):
        src = np.frombuffer(b, dtype=np.uint8)
        out = np.zeros(len(b) + 1, dtype=np.uint16)
        n = lib.jo_java_utf8_to_utf16(src.ctypes.data_as(C.c_void_p), len(b), out.ctypes.data_as(C.c_void_p))
        return out[:n].tolist()

    assert dec("aé".encode()) == [0x61, 0xE9]
    assert dec(b"a\xe2\x82") == [0x61, 0xFFFD]          # truncated 3-byte sequence
    assert dec(b"a\xf0\x9f\x8d") == [0x61, 0xFFFD]      # truncated 4-byte sequence
    assert dec("🍕".encode()) == [0xD83C, 0xDF55]       # surrogate pair
    assert dec(b"\x80a") == [0xFFFD, 0x61]              # stray continuation byte
    assert dec(b"\xe2\x28\xa1") == [0xFFFD, 0x28, 0xFFFD]


def test_oracle_unicode_properties_and_boundaries_against_the_regex_module():
    """The oracle's restatement of \\p{..}, \\w, \\d and \\b (java.util.regex semantics, Unicode 15.0 tables) against the `regex` module on
    stable code points.  \\b is compared under UNICODE_CHARACTER_CLASS only (without it Java asks Character.isLetterOrDigit, which
    no Python engine mirrors) and without stray combining marks (Java's hasBaseCharacter rule differs from UTS #18 there)."""
    import random
    import regex
    from oracle import jo
    pats = [(r"\p{Lu}\p{Ll}+|\p{Nd}+|\s+|.", 0x100, True), (r"\w+|\W+", 0x100, True), (r"\d+|\D", 0x100, True), (r"\w+|[^\w\s]+|\s+", 0, False),
            (r"\p{Lu}+|\p{Ll}+|\P{L}", 0, True), (r"[\p{Sc}\p{Sm}]+|\p{P}|\p{IsAlphabetic}+|\p{Z}+|.", 0x100, True), (r"\b\w+\b|\W", 0x100, True),
            (r"\B.|.", 0x100, True), (r"[^\W\d_]+|\d{1,3}|[\W_]", 0x100, True), (r"\p{gc=Mn}+|\p{IsLo}|\p{LC}+|\P{M}", 0, True)]
    alph = list("abcXYZ 019_-+$€£±.,;!?'\"()[]\n\t") + ["é", "ß", "Ж", "я", "中", "あ", "١", "२", "½", "Ⅷ", "ǅ", "ʰ", "　", " ", "—", "“", "𝐀", "🍕", "한", "é", "ा"]
    rng = random.Random(5)
    for pat, fl, unicode_mode in pats:
        o = jo.OracleEncoding("t", pat, fl, {bytes([b]): b for b in range(256)}, {})
        r = regex.compile(pat.replace(r"\p{IsAlphabetic}", r"\p{Alphabetic}").replace(r"\p{IsLo}", r"\p{Lo}"), regex.V0 | (regex.UNICODE if unicode_mode else regex.ASCII))
        for _ in range(300):
            t = "".join(rng.choice(alph) for _ in range(rng.randint(0, 30)))
            boff = [0]
            for ch in t:
                boff.append(boff[-1] + len(ch.encode()))
            exp = [(boff[m.start()], boff[m.end()]) for m in r.finditer(t) if m.end() > m.start()]
            assert [(a, e) for a, e in o.split(t.encode()) if e > a] == exp, (pat, t)


def test_oracle_groups_anchors_scripts_and_look_behind_against_the_regex_module():
    """Named groups, \\A \\Z \\z, \\Q..\\E, \\h \\v, Unicode scripts (\\p{IsHan}, \\p{script=..}, \\p{sc=..}) and one-character look-behind in the
    oracle's matcher against the `regex` module (Python spells \\z as \\Z, Java's \\Z as (?=\\n?\\Z), scripts without the Is prefix; its `.`
    also matches \\r, U+0085, U+2028 and U+2029, so the last alternative is written out)."""
    import random
    import regex
    from oracle import jo
    dot = "[^\\n\\r\\x85\\u2028\\u2029]"
    hs = "[ \\t\\xA0\\u1680\\u180e\\u2000-\\u200a\\u202f\\u205f\\u3000]"
    vs = "[\\n\\x0B\\f\\r\\x85\\u2028\\u2029]"
    pats = [(r"(?<word>\w+)|(?<sp>\s+)|.", 0, None, False), (r"\Aab|\w+\z|\w+|\W", 0, r"\Aab|\w+\Z|\w+|\W", False), (r"\w+\Z|\w+|\W", 0, r"\w+(?=\n?\Z)|\w+|\W", False),
            (r"\Qa.b\E+|\w+|.", 0, r"a\.b+|\w+|.", False), (r"\h+|\v+|\H", 0, hs + "+|" + vs + "+|[^" + hs[1:], True),
            (r"(?<=\d)[a-z]+|(?<![a-z])\d+|.", 0, None, False),
            (r"\p{IsHan}+|\p{script=Cyrillic}+|\p{sc=Latn}+|\P{IsHiragana}", 0, r"\p{Han}+|\p{Script=Cyrillic}+|\p{Script=Latin}+|\P{Hiragana}", True),
            (r"(?<!\p{L})\p{L}{1,3}|.", 0x100, None, True), (r"[\p{IsGreek}\p{IsHangul}]+|\p{IsCommon}|.", 0, r"[\p{Greek}\p{Hangul}]+|\p{Common}|.", True),
            (r"\R\n|\R+|\w+|.", 0, "(?:\\r\\n|" + vs + ")\\n|(?:\\r\\n|" + vs + ")+|\\w+|.", False)]
    alph = list("abcXYZ 019_-+$.,;!?'\n\r\t") + ["é", "ß", "Ж", "я", "中", "国", "あ", "カ", "١", "२", "　", " ", "—", "𝐀", "🍕", "한", "ा", "\u2028", "\x0b", "α", "Ω"]
    rng = random.Random(5)
    for pat, fl, rpat, uni in pats:
        o = jo.OracleEncoding("t", pat, fl, {bytes([b]): b for b in range(256)}, {})
        rp = rpat or pat
        if rp.endswith("|."):
            rp = rp[:-1] + dot
        r = regex.compile(rp, regex.V0 | (regex.UNICODE if (fl & 0x100 or uni) else regex.ASCII))
        for _ in range(300):
            t = "".join(rng.choice(alph) for _ in range(rng.randint(0, 24)))
            boff = [0]
            for ch in t:
                boff.append(boff[-1] + len(ch.encode()))
            exp = [(boff[m.start()], boff[m.end()]) for m in r.finditer(t) if m.end() > m.start()]
            assert [(a, e) for a, e in o.split(t.encode()) if e > a] == exp, (pat, t)


def test_oracle_nested_classes_and_intersections_against_the_regex_module():
    """[a[b-d]] (union), [a-z&&[^aeiou]] (intersection, subtraction), a leading ^ over the whole: against the `regex` module's V1 set
    operations (which want every operand bracketed)."""
    import random
    import regex
    from oracle import jo
    dot = "[^\\n\\r\\x85\\u2028\\u2029]"
    pats = [(r"[a-z&&[^aeiou]]+|[aeiou]+|.", r"[[a-z]&&[^aeiou]]+|[aeiou]+|.", True), (r"[\p{L}&&[^\p{IsHan}]]+|\p{IsHan}|.", r"[\p{L}&&[^\p{Han}]]+|\p{Han}|.", True),
            (r"[a-c[x-z]]+|[^a[0-9]]|.", r"[a-cx-z]+|[^a0-9]|.", True), (r"[^\w&&[^_]]+|.", r"[^\w&&[^_]]+|.", False),
            (r"[a-z&&b-y&&[^m]]+|.", r"[[a-z]&&[b-y]&&[^m]]+|.", True)]
    alph = list("abcmxyzXYZ 019_-+$.,;!?'") + ["é", "ß", "Ж", "я", "中", "国", "あ", "カ", "१"]
    rng = random.Random(6)
    for pat, rpat, uni in pats:
        o = jo.OracleEncoding("t", pat, 0, {bytes([b]): b for b in range(256)}, {})
        r = regex.compile(rpat[:-1] + dot, regex.V1 | (regex.UNICODE if uni else regex.ASCII))
        for _ in range(300):
            t = "".join(rng.choice(alph) for _ in range(rng.randint(0, 24)))
            boff = [0]
            for ch in t:
                boff.append(boff[-1] + len(ch.encode()))
            exp = [(boff[m.start()], boff[m.end()]) for m in r.finditer(t) if m.end() > m.start()]
            assert [(a, e) for a, e in o.split(t.encode()) if e > a] == exp, (pat, t)


def test_o200k_pattern_as_a_custom_pattern_vs_tiktoken():
    """The o200k_base split pattern (tiktoken_ext/openai_public.py), registered as a custom pattern over the cl100k vocabulary: the oracle's
    matcher against tiktoken's (fancy-regex) on the same pattern string and ranks."""
    tiktoken = pytest.importorskip("tiktoken")
    from tiktoken.load import load_tiktoken_bpe
    from oracle import jo
    pat = "|".join([r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?",
                    r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?",
                    r"\p{N}{1,3}", r" ?[^\s\p{L}\p{N}]+[\r\n/]*", r"\s*[\r\n]+", r"\s+(?!\S)", r"\s+"])
    ranks = load_tiktoken_bpe(os.path.join(jo.DATA, jo.BUILTIN["cl100k_base"][1]))
    tk = tiktoken.Encoding("o200k_pat", pat_str=pat, mergeable_ranks=ranks, special_tokens={})
    o = jo.OracleEncoding("o200k_pat", pat, 0x100, ranks, {})
    rng = random.Random(11)
    alph = [c for c in ALPH if c != "ſ"] + ["A", "B", "Z", "É", "Ö", "ǅ", "ʰ", "/"]  # (U+017F: the engines disagree on (?i) long s, as for cl100k)
    for _ in range(4000):
        s = "".join(rng.choice(alph) for _ in range(rng.randint(0, 40)))
        assert o.encode_ordinary(s) == tk.encode_ordinary(s), s
