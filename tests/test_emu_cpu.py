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
