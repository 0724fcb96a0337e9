"""CPU-side fuzz of the kernel's __host__ __device__ building blocks (tile emulator with 128-byte tiles) against the oracle."""
import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

ALPH = list("abcdefghijklmnopqrstuvwxyzSTREVMLD     \n\n\r\t'''!!?.,;:-0123456789") + [
    "é", "ß", "ſ", "Ж", "я", "中", "文", "あ", "カ", " ", "　", " ", "١", "٢", "½", "🍕", "‍", "️", "्", "ा", "ก", "ั", "한", "😀", "ñ", "—", "“"]


ASCII_ALPH = list("abcdefghijklmnopqrstuvwxyzSTREVMLDstrevmld          \n\n\n\r\t''''!!?.,;:-0123456789012")
ASCII_UNITS_A = ["'s", "'t", "'re", "'ve", "'m", "'ll", "'d", "'S", "'LL", "x", " ", "  ", "'", "!", "1", "\n", "re", "ll"]
ASCII_UNITS_B = ["\n", "\n", " ", " ", " ", "\t", "!", "a", "1", "12345678", "\r\n", "          "]


# non-ASCII letters, combining marks, multi-byte punctuation and emoji with ASCII whitespace / digits / apostrophes:
# the multi-byte side of the bit-parallel split rules
UTF8_ALPH = list("жяЖ中文あカ한नमけдоброеутро") + ["्", "ा", "ั", "ก", "é", "ß", "ñ", "—", "“", "。", "，", "🍕", "😀", "‍", "️"] + list(
    "      \n\n\t''!?.,-0123abcsStTdD")


def random_docs(rng, ndocs, ascii_only=False, utf8_letters=False):
    docs = []
    for _ in range(ndocs):
        n = rng.choice([0, 1, 5, 40, 200, 700])
        mode = rng.random()
        if utf8_letters:
            s = "".join(rng.choice(UTF8_ALPH) for _ in range(rng.randint(0, n)))
        elif ascii_only:
            # stresses the bit-parallel fast path: contractions, digit runs, whitespace / newline runs
            k = rng.randint(0, n)
            if mode < 0.5:
                s = "".join(rng.choice(ASCII_ALPH) for _ in range(k))
            elif mode < 0.75:
                s = "".join(rng.choice(ASCII_UNITS_A) for _ in range(k))
            else:
                s = "".join(rng.choice(ASCII_UNITS_B) for _ in range(k))
        elif mode < 0.7:
            s = "".join(rng.choice(ALPH) for _ in range(rng.randint(0, n)))
        elif mode < 0.8:
            s = rng.choice(["1", "\n", " ", "a", "!", "ab", "中", "\n ", " \n", "١"]) * rng.randint(0, n)
        else:
            s = "".join(rng.choice(["the ", " of", "ing", "tion", " 123", "'s", "'ll", "\n\n", "  ", "\t", "Hello", ", ", "中文", "!!!"]) for _ in range(rng.randint(0, n // 3)))
        docs.append(s.encode("utf-8"))
    return docs


@pytest.mark.parametrize("name", ["cl100k_base", "r50k_base"])
def test_tile_emulator_matches_oracle(name, oracles):
    import emu
    from oracle import jo
    pat, fname, special = jo.BUILTIN[name]
    ranks = jo.load_tiktoken(os.path.join(jo.DATA, fname))
    e = emu.EmuEncoding(name, pat, 0x100, ranks, special)
    o = oracles[name]
    rng = random.Random(7)
    for it in range(4500):
        docs = random_docs(rng, rng.randint(1, 4), ascii_only=(it % 3 == 1), utf8_letters=(it % 3 == 2))
        blob = b"".join(docs)
        off = np.zeros(len(docs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in docs])
        pf, ids, tok_off, status = e.run(np.frombuffer(blob, dtype=np.uint8), off)
        exp_ids, exp_off, exp_pf = [], [0], np.zeros(len(blob), dtype=np.uint8)
        for d, doc in enumerate(docs):
            for (a, _) in o.split(doc):
                exp_pf[off[d] + a] = 1
            exp_ids += o.encode_ordinary(doc)
            exp_off.append(len(exp_ids))
        assert np.array_equal(pf, exp_pf), docs
        assert ids.tolist() == exp_ids, docs
        assert tok_off.tolist() == exp_off


def test_table_statistics():
    """Pair table sizes quoted in SURVEY.md (a6): 233 378 (cl100k) / 108 299 (r50k) splits of tokens into two parts."""
    import emu
    from oracle import jo
    for name, npairs, ntok in [("cl100k_base", 233378, 100256), ("r50k_base", 108299, 50256)]:
        pat, fname, special = jo.BUILTIN[name]
        e = emu.EmuEncoding(name, pat, 0x100, jo.load_tiktoken(os.path.join(jo.DATA, fname)), special)
        st = e.stats()
        assert st[0] == ntok and st[1] == npairs
        assert st[2] <= 40 and st[4] <= 16  # longest probe sequences (piece table slots / pair table buckets)


# General split patterns: the device's backtracking program (jtk_regex.h + the compiler in jtk_regex.cpp)
# against the oracle's independently written matcher.  Flags: 0x02 CASE_INSENSITIVE, 0x40 UNICODE_CASE, 0x100 UNICODE_CHARACTER_CLASS.
GENERAL_PATTERNS = [
    ("test", 0), (r"\w+|\s+", 0), (r"[a-z]+|[0-9]{1,3}| ?[^a-z0-9 ]+", 0), (r"(?i:'s|'t|'re)|\p{L}+|\p{N}{1,3}|\s+(?!\S)|\s+|.", 0x100),
    (r"a*b|c", 0), (r"\s*[a-z]+?\d|[a-z]+", 0), (r"(?:ab|a)(?:bc|c)?|.", 0),
    (r"[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+", 0x100),
    (r"(?i)hello|the |[a-z]+", 0), (r"(?i)s+|k+", 0x42), (r"\p{L}++|\d*+x|\S", 0), (r"(?:[a-c]|de){2,3}|\S+?(?= )|\s", 0), (r"x*", 0),
    (r"^\w+|\w+$|\s", 0), (r"'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+", 0),
    (r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+", 0x102),
    (r"(?:\p{L}\p{N}?)+| (?! )|[^ ]", 0x100), (r"[一-鿿]+|\x41{2}|\x{1F355}|[a-zé-ü]+|\S", 0),
    # Unicode properties and word boundaries (general categories, Alphabetic, \d \w \b with and without UNICODE_CHARACTER_CLASS)
    (r"\p{Lu}\p{Ll}+|\p{Nd}+|\s+|.", 0x100), (r"\w+|\W+", 0x100), (r"\d+|\D", 0x100), (r"\p{Lu}+|\p{Ll}+|\P{L}", 0),
    (r"[\p{Sc}\p{Sm}]+|\p{P}|\p{IsAlphabetic}+|\p{Z}+|.", 0x100), (r"\b\w+\b|\W", 0x100), (r"\B.|.", 0x100), (r"\bfoo\b|\w+|\W", 0),
    (r"[^\W\d_]+|\d{1,3}|[\W_]", 0x100), (r"\p{gc=Mn}+|\p{IsLo}|\p{LC}+|\P{M}", 0),
    # named groups, \A \Z \z, \Q..\E, \h \v, scripts, one-character look-behind
    (r"(?<word>\w+)|(?<sp>\s+)|.", 0), (r"\Aab|\w+\z|\w+|\W", 0), (r"\w+\Z|\w+|\W", 0), (r"\Qa.b\E+|\w+|.", 0), (r"\h+|\v+|\H", 0),
    (r"(?<=\d)[a-z]+|(?<![a-z])\d+|.", 0), (r"\p{IsHan}+|\p{script=Cyrillic}+|\p{sc=Latn}+|\P{IsHiragana}", 0),
    (r"(?<!\p{L})\p{L}{1,3}|.", 0x100), (r"[\p{IsGreek}\p{IsHangul}]+|\p{IsCommon}|.", 0),
    # nested classes and class intersection
    (r"[a-z&&[^aeiou]]+|[aeiou]+|.", 0), (r"[\p{L}&&[^\p{IsHan}]]+|\p{IsHan}|.", 0), (r"[a-c[x-z]]+|[^a[0-9]]|.", 0), (r"[^\w&&[^_]]+|.", 0),
    (r"[a-z&&b-y&&[^m]]+|.", 0), (r"(?i)[a-f&&[^c]]+|\s+|.", 0),
    # CASE_INSENSITIVE and the cased-letter categories (JDK 9+: Lu, Ll, Lt each stand for all three)
    (r"\p{Lu}+|\P{L}+|.", 2), (r"(?i:\p{Ll})+|\p{Lu}|[\P{Lt}&&[^\s]]", 0),
    (r"\R+|\w+|[^\w\r\n]", 0), (r"(?:ab){20}c|(?:[a-c]\d){3,40}|.", 0),
    # split patterns of other tokenizers a JTokkit user may register: o200k_base (as tiktoken publishes it) and a DeepSeek-style one
    (r"[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]*[\p{Ll}\p{Lm}\p{Lo}\p{M}]+(?i:'s|'t|'re|'ve|'m|'ll|'d)?|[^\r\n\p{L}\p{N}]?[\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]+[\p{Ll}\p{Lm}\p{Lo}\p{M}]*(?i:'s|'t|'re|'ve|'m|'ll|'d)?|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n/]*|\s*[\r\n]+|\s+(?!\S)|\s+", 0x100),
    (r"[!\"#$%&'()*+,\-./:;<=>?@\[\\\]^_`{|}~][A-Za-z]+|[^\r\n\p{L}\p{P}\p{S}]?[\p{L}\p{M}]+| ?[\p{P}\p{S}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+|\p{N}{1,3}|[一-龥぀-ゟ゠-ヿ]+", 0x100),
]


def expected_piece_bits(o, docs, off, nbytes):
    """Piece-start and gap bits per byte from the oracle's Matcher.find() restatement (empty matches produce nothing)."""
    start, skip = np.zeros(nbytes, np.uint8), np.zeros(nbytes, np.uint8)
    for d, doc in enumerate(docs):
        prev = 0
        for (a, b) in o.split(doc):
            if b == a:
                continue
            if a > prev:
                start[off[d] + prev] = skip[off[d] + prev] = 1
            start[off[d] + a] = 1
            prev = b
        if prev < len(doc):
            start[off[d] + prev] = skip[off[d] + prev] = 1
    return start, skip


@pytest.mark.parametrize("index", range(len(GENERAL_PATTERNS)))
def test_general_pattern_program_matches_oracle_matcher(index):
    import emu
    from oracle import jo
    jo.build()
    pat, flags = GENERAL_PATTERNS[index]
    e = emu.EmuEncoding("g", pat, flags, {b"a": 0}, {})
    assert e.pattern_kind() == 3
    o = jo.OracleEncoding("g", pat, flags, {b"a": 0}, {})
    rng = random.Random(100 + index)
    for it in range(150):
        docs = random_docs(rng, rng.randint(1, 4), ascii_only=(it % 3 == 1), utf8_letters=(it % 3 == 2))
        blob = b"".join(docs)
        off = np.zeros(len(docs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in docs])
        exp_start, exp_skip = expected_piece_bits(o, docs, off, len(blob))
        # the DFA where the pattern has one (the default), and the backtracking program: both must reproduce Matcher.find()
        for no_dfa in ([False, True] if e.dfa_info()[0] else [True]):
            start, skip = e.general_split(np.frombuffer(blob, dtype=np.uint8), off, no_dfa=no_dfa)
            assert np.array_equal(start, exp_start) and np.array_equal(skip, exp_skip), (pat, docs, "vm" if no_dfa else "dfa")


def test_general_pattern_stack_overflow_is_reported():
    """The backtracking program keeps one frame per iteration of a group loop: a long run exhausts a small stack and the call says so
    instead of mis-splitting.  The DFA of the same pattern has no stack (nothing to overflow) and gives the oracle's pieces."""
    import emu
    from oracle import jo
    jo.build()
    pat = r"(?:a|b)+c|."
    e = emu.EmuEncoding("g", pat, 0, {b"a": 0}, {})
    assert e.dfa_info()[0] > 0
    doc = b"ab" * 400
    text = np.frombuffer(doc, dtype=np.uint8)
    off = np.array([0, text.size], dtype=np.int64)
    with pytest.raises(OverflowError):
        e.general_split(text, off, stack_cap=64, no_dfa=True)
    e.general_split(text, off, stack_cap=4096, no_dfa=True)
    start, skip = e.general_split(text, off, stack_cap=64)
    exp_start, exp_skip = expected_piece_bits(jo.OracleEncoding("g", pat, 0, {b"a": 0}, {}), [doc], off, len(doc))
    assert np.array_equal(start, exp_start) and np.array_equal(skip, exp_skip)
    # a pattern without a DFA form ('$') keeps the program and its stack
    e2 = emu.EmuEncoding("g", r"(?:a|b)+c$|.", 0, {b"a": 0}, {})
    assert e2.dfa_info()[0] == 0 and "$" in e2.dfa_info()[2]
    with pytest.raises(OverflowError):
        e2.general_split(text, off, stack_cap=64)


def test_dfa_run_shortcuts_on_long_runs_and_gaps():
    """The sequential DFA passes step over long ASCII runs four bytes at a time (a state that loops on all four; an attempt that dies at
    once on each of four start positions): documents made of long runs, long gaps and their borders at every alignment, against the
    oracle and against the backtracking program."""
    import emu
    from oracle import jo
    jo.build()
    rng = random.Random(9)
    pats = [(r"[a-z]+|\d{1,3}", 0), (r"\w+|\s+", 0), (r"a+b|a", 0), (r"x*", 0), (r"[a-z]+(?![a-z!])|\s+(?!\S)|\s", 0), (r"^[a-z]+|[0-9]++|!{2,5}", 0),
            (r"(?i:[a-c]+)z|[a-z]", 0), (r"\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+", 0x100)]
    units = ["a" * 3, "a" * 9, "a" * 70, "abc" * 30, "!" * 5, "!" * 40, "!" * 131, " " * 17, " " * 64, "123456789" * 7, "x", "b", "z", "é", "中" * 12, "\n", "aB" * 20, "Z" * 33]
    for pat, flags in pats:
        e = emu.EmuEncoding("g", pat, flags, {b"a": 0}, {})
        assert e.dfa_info()[0] > 0
        o = jo.OracleEncoding("g", pat, flags, {b"a": 0}, {})
        for it in range(60):
            docs = []
            for _ in range(rng.randint(1, 3)):
                docs.append(("".join(rng.choice(units) for _ in range(rng.randint(1, 8))))[rng.randint(0, 3):].encode())
            blob = b"".join(docs)
            off = np.zeros(len(docs) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(x) for x in docs])
            exp_start, exp_skip = expected_piece_bits(o, docs, off, len(blob))
            for no_dfa in (False, True):
                start, skip = e.general_split(np.frombuffer(blob, dtype=np.uint8), off, no_dfa=no_dfa)
                assert np.array_equal(start, exp_start) and np.array_equal(skip, exp_skip), (pat, docs, "vm" if no_dfa else "dfa")


def _random_pattern(rng, depth=0):
    """A random pattern from (a little more than) the supported grammar: literals, classes, properties, scripts, class algebra, greedy / lazy /
    possessive quantifiers, plain / named / capturing / (?i:) groups, alternation, look-ahead, look-behind, anchors, boundaries."""
    def atom(depth):
        r = rng.random()
        if r < 0.35:
            return rng.choice(["a", "b", "c", " ", "1", "é"])
        if r < 0.55:
            return rng.choice(["[ab]", "[^a]", r"\w", r"\s", r"\d", ".", "[a-c1]", r"\S", r"[^\s1]", r"\p{L}", r"\R", r"\h", r"\V", r"\p{IsLatin}", r"\p{Lu}", r"\P{Ll}",
                               "[a-c&&[^b]]", "[a[1 ]]", r"[\w&&[^\d]]", r"\p{IsGreek}", r"\Qa.\E", "[^é\\n]", r"\x41", r"[\p{L}&&[^\p{Lu}]]"])
        if r < 0.62:
            return rng.choice(["(?=a)", "(?!b)", r"(?!\S)", r"(?=\s)", "(?![ab])", "(?<=a)", "(?<!b)", r"(?<![\w])", r"(?<=\s)", "(?=ab)", "(?!a1)", r"\b", r"\B", "$", r"\Z"])
        if r < 0.66:
            return rng.choice(["^", r"\A", r"\z"])
        if depth > 2:
            return "a"
        return rng.choice(["(?:", "(?:", "(?<n>", "(", "(?i:"]) + _random_pattern(rng, depth + 1) + ")"

    def quantified(depth):
        a = atom(depth)
        if a.startswith("(?=") or a.startswith("(?!") or a.startswith("(?<=") or a.startswith("(?<!") or a in ("^", r"\A", r"\z", r"\b", r"\B", "$", r"\Z") or rng.random() < 0.5:
            return a
        q = rng.choice(["*", "+", "?", "{1,2}", "{2}", "{0,3}"])
        return a + q + (rng.choice(["", "", "?", "+"]) if not a.startswith("(") else rng.choice(["", "", "?"]))

    return "|".join("".join(quantified(depth) for _ in range(rng.randint(1, 3))) for _ in range(rng.randint(1, 3)))


def test_random_patterns_dfa_and_program_against_the_oracle():
    """Random patterns (what the compiler accepts of them), random short documents in groups that span several 32-byte slices: the DFA and
    the backtracking program under the sliced find() passes against the oracle's matcher.  (This fuzz found the speculative end bit at a
    join position reached over an empty match: jtk_rx_finish_word.)"""
    import emu
    from oracle import jo
    jo.build()
    rng = random.Random(2)
    alph = ["a", "b", "c", " ", "1", "\n", "é", "ab", "  ", "aa", "b1", "A", "B", "É", "ſ", "K", "\u212a", "\u00a0", "\u0661", "α", "Ω", "\r\n", "\r", "\t", "\x0b", "中"]
    accepted = with_dfa = 0
    for it in range(450):
        pat = _random_pattern(rng)
        # Pattern flags: none for most, then CASE_INSENSITIVE with and without UNICODE_CASE / UNICODE_CHARACTER_CLASS (the long s and the Kelvin sign
        # fold onto s and k for literal characters only, never for \\w or a property: this fuzz found the oracle folding those too)
        flags = 0 if it % 3 else rng.choice([2, 0x100, 0x102, 0x42])
        try:
            e = emu.EmuEncoding("g", pat, flags, {b"a": 0}, {})
        except ValueError:
            continue  # outside the subset (a loop over a nullable group, ...)
        o = jo.OracleEncoding("g", pat, flags, {b"a": 0}, {})
        accepted += 1
        has_dfa = e.dfa_info()[0] > 0
        with_dfa += has_dfa
        for _ in range(10):
            docs = [("".join(rng.choice(alph) for _ in range(rng.randint(0, 14)))).encode() for _ in range(rng.randint(1, 3))]
            blob = b"".join(docs)
            off = np.zeros(len(docs) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(x) for x in docs])
            exp_start, exp_skip = expected_piece_bits(o, docs, off, len(blob))
            for no_dfa in ([False, True] if has_dfa else [True]):
                try:
                    start, skip = e.general_split(np.frombuffer(blob, dtype=np.uint8), off, no_dfa=no_dfa)
                except OverflowError:
                    continue  # (the program's small test stack)
                assert np.array_equal(start, exp_start) and np.array_equal(skip, exp_skip), (pat, docs, "vm" if no_dfa else "dfa")
    assert accepted > 150 and with_dfa > 100


def test_join_after_an_empty_match_takes_no_speculative_end_bit():
    """A nullable pattern: the true run reaches position 44 by stepping over an empty match at 43, the speculative run of the slice by a
    match (40, 44).  The end bit of that speculative match must not become a piece start of the gap [43, 46)."""
    import emu
    from oracle import jo
    jo.build()
    pat = r"(?:[a-z1 \n]{4})?"
    docs = [b"c" * 27, b"x" * 16 + "bé".encode()]
    blob = b"".join(docs)
    off = np.array([0, 27, 27 + len(docs[1])], dtype=np.int64)
    e = emu.EmuEncoding("g", pat, 0, {b"a": 0}, {})
    exp_start, exp_skip = expected_piece_bits(jo.OracleEncoding("g", pat, 0, {b"a": 0}, {}), docs, off, len(blob))
    for no_dfa in (False, True):
        start, skip = e.general_split(np.frombuffer(blob, dtype=np.uint8), off, no_dfa=no_dfa)
        assert np.array_equal(start, exp_start) and np.array_equal(skip, exp_skip)
    assert np.nonzero(exp_skip)[0].tolist() == [24, 43]


def test_flat_dfa_loop_decodes_utf8_like_the_bytewise_decoder():
    import ctypes as C
    import emu
    L = emu.lib()
    L.emu_decode_word_check.restype = C.c_int64
    L.emu_decode_word_check.argtypes = [C.c_int64, C.c_uint64]
    assert L.emu_decode_word_check(2_000_000, 7) == 0


def test_dfa_is_built_for_the_predefined_pattern_strings_as_general_patterns():
    """The two predefined split patterns (possessive quantifiers, \\s+(?!\\S), (?i:...)) have a DFA form when they are registered
    with a flag that takes them off the rule path: small tables, and the same pieces as the oracle's matcher (checked above for
    look-alikes; here for the exact strings)."""
    import emu
    from oracle import jo
    jo.build()
    rng = random.Random(77)
    for name in ["cl100k_base", "r50k_base"]:
        pat = jo.BUILTIN[name][0]
        e = emu.EmuEncoding(name + "_ci", pat, 0x102, {b"a": 0}, {})
        nstates, nsym, why = e.dfa_info()
        assert 0 < nstates <= 64 and nsym <= 32, (nstates, nsym, why)
        o = jo.OracleEncoding(name + "_ci", pat, 0x102, {b"a": 0}, {})
        for it in range(60):
            docs = random_docs(rng, rng.randint(1, 4), ascii_only=(it % 3 == 1), utf8_letters=(it % 3 == 2))
            blob = b"".join(docs)
            off = np.zeros(len(docs) + 1, dtype=np.int64)
            off[1:] = np.cumsum([len(x) for x in docs])
            start, skip = e.general_split(np.frombuffer(blob, dtype=np.uint8), off)
            exp_start, exp_skip = expected_piece_bits(o, docs, off, len(blob))
            assert np.array_equal(start, exp_start) and np.array_equal(skip, exp_skip), (name, docs)


def test_predefined_patterns_do_not_take_the_general_path():
    import emu
    from oracle import jo
    for name in ["cl100k_base", "r50k_base"]:
        pat, fname, special = jo.BUILTIN[name]
        assert emu.EmuEncoding(name, pat, 0x100, {b"a": 0}, {}).pattern_kind() in (1, 2)
    pat = jo.BUILTIN["cl100k_base"][0]
    assert emu.EmuEncoding("ci", pat, 0x102, {b"a": 0}, {}).pattern_kind() == 3  # CASE_INSENSITIVE: general program


def test_segmentwise_merge_is_exact_for_random_vocabularies():
    """The merge kernels run bytePairMerge separately on the segments between positions whose byte bigram occurs in no token
    (jtk_safe_cut).  Random vocabularies over small alphabets - sparse and dense bigram tables, concatenations that rank below their
    parts, missing single bytes - against the oracle's literal loop, through the emulator's copy of that segment loop."""
    import emu
    from oracle import jo
    rng = random.Random(99)
    pat = jo.BUILTIN["cl100k_base"][0]
    for trial in range(60):
        alphabet = "abcdefgh"[:rng.randint(2, 8)]
        vocab = {}
        ids = list(range(1000))
        rng.shuffle(ids)
        for ch in alphabet:
            if rng.random() < 0.95:
                vocab[ch.encode()] = ids.pop()
        for _ in range(rng.randint(1, 40)):
            w = "".join(rng.choice(alphabet) for _ in range(rng.randint(2, 6))).encode()
            if w not in vocab:
                vocab[w] = ids.pop()
        e = emu.EmuEncoding("rnd", pat, 0x100, vocab, {})
        o = jo.OracleEncoding("rnd", pat, 0x100, vocab, {})
        docs = ["".join(rng.choice(alphabet) for _ in range(rng.choice([1, 2, 3, 7, 20, 60, 150]))).encode() for _ in range(40)]
        blob = b"".join(docs)
        off = np.zeros(len(docs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in docs])
        pf, got, tok_off, status = e.run(np.frombuffer(blob, dtype=np.uint8), off)
        for d, doc in enumerate(docs):
            try:
                exp = o.encode_ordinary(doc)
            except ValueError:
                assert status[d] & 2, (trial, doc)  # JTK_DOC_UNKNOWN_BYTES
                continue
            assert got[tok_off[d]:tok_off[d + 1]].tolist() == exp, (trial, vocab, doc)
