#!/usr/bin/env python3
"""Generate jtokkit_b200/csrc/unicode_ranges.inc from Python's unicodedata.

The split patterns of the reference (EncodingFactory.java:63,77,91,105) are compiled with
Pattern.UNICODE_CHARACTER_CLASS (EncodingFactory.java:129), so
  \\p{L} = general category L*   (Character.getType in the JVM)
  \\p{N} = general category N*   (Nd, Nl, No)
  \\s    = Unicode White_Space
The JVM's tables depend on its Unicode version; this build pins Unicode 15.0.0 (= Java 20/21,
= this image's Python unicodedata).  The file is DATA shared by the CUDA library and the CPU oracle;
both are cross-checked against the `regex` module and tiktoken in tests/.

Usage:  python tools/gen_unicode_tables.py   (rewrites the .inc in place)
"""
import os
import sys
import unicodedata

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "jtokkit_b200", "csrc", "unicode_ranges.inc")

# Unicode White_Space (PropList.txt); identical in every Unicode version >= 6.3.
WHITE_SPACE = [(0x09, 0x0D), (0x20, 0x20), (0x85, 0x85), (0xA0, 0xA0), (0x1680, 0x1680), (0x2000, 0x200A),
               (0x2028, 0x2029), (0x202F, 0x202F), (0x205F, 0x205F), (0x3000, 0x3000)]


GC_NAMES = ["Lu", "Ll", "Lt", "Lm", "Lo", "Mn", "Mc", "Me", "Nd", "Nl", "No", "Pc", "Pd", "Ps", "Pe", "Pi", "Pf", "Po", "Sm", "Sc", "Sk", "So",
            "Zs", "Zl", "Zp", "Cc", "Cf", "Cs", "Co", "Cn"]


# Unicode scripts (long name as Character.UnicodeScript spells it, ISO 15924 code); what the `regex` module does not know is skipped
SCRIPTS = [("Common", "Zyyy"), ("Inherited", "Zinh"), ("Latin", "Latn"), ("Greek", "Grek"), ("Cyrillic", "Cyrl"), ("Armenian", "Armn"), ("Hebrew", "Hebr"),
           ("Arabic", "Arab"), ("Syriac", "Syrc"), ("Thaana", "Thaa"), ("Devanagari", "Deva"), ("Bengali", "Beng"), ("Gurmukhi", "Guru"),
           ("Gujarati", "Gujr"), ("Oriya", "Orya"), ("Tamil", "Taml"), ("Telugu", "Telu"), ("Kannada", "Knda"), ("Malayalam", "Mlym"),
           ("Sinhala", "Sinh"), ("Thai", "Thai"), ("Lao", "Laoo"), ("Tibetan", "Tibt"), ("Myanmar", "Mymr"), ("Georgian", "Geor"),
           ("Hangul", "Hang"), ("Ethiopic", "Ethi"), ("Cherokee", "Cher"), ("Canadian_Aboriginal", "Cans"), ("Ogham", "Ogam"), ("Runic", "Runr"),
           ("Khmer", "Khmr"), ("Mongolian", "Mong"), ("Hiragana", "Hira"), ("Katakana", "Kana"), ("Bopomofo", "Bopo"), ("Han", "Hani"),
           ("Yi", "Yiii"), ("Old_Italic", "Ital"), ("Gothic", "Goth"), ("Deseret", "Dsrt"), ("Tagalog", "Tglg"), ("Hanunoo", "Hano"),
           ("Buhid", "Buhd"), ("Tagbanwa", "Tagb"), ("Limbu", "Limb"), ("Tai_Le", "Tale"), ("Linear_B", "Linb"), ("Ugaritic", "Ugar"),
           ("Shavian", "Shaw"), ("Osmanya", "Osma"), ("Cypriot", "Cprt"), ("Braille", "Brai"), ("Buginese", "Bugi"), ("Coptic", "Copt"),
           ("New_Tai_Lue", "Talu"), ("Glagolitic", "Glag"), ("Tifinagh", "Tfng"), ("Syloti_Nagri", "Sylo"), ("Old_Persian", "Xpeo"),
           ("Kharoshthi", "Khar"), ("Balinese", "Bali"), ("Cuneiform", "Xsux"), ("Phoenician", "Phnx"), ("Phags_Pa", "Phag"), ("Nko", "Nkoo"),
           ("Sundanese", "Sund"), ("Lepcha", "Lepc"), ("Ol_Chiki", "Olck"), ("Vai", "Vaii"), ("Saurashtra", "Saur"), ("Kayah_Li", "Kali"),
           ("Rejang", "Rjng"), ("Lycian", "Lyci"), ("Carian", "Cari"), ("Lydian", "Lydi"), ("Cham", "Cham"), ("Tai_Tham", "Lana"),
           ("Tai_Viet", "Tavt"), ("Avestan", "Avst"), ("Egyptian_Hieroglyphs", "Egyp"), ("Samaritan", "Samr"), ("Lisu", "Lisu"),
           ("Bamum", "Bamu"), ("Javanese", "Java"), ("Meetei_Mayek", "Mtei"), ("Imperial_Aramaic", "Armi"), ("Old_South_Arabian", "Sarb"),
           ("Inscriptional_Parthian", "Prti"), ("Inscriptional_Pahlavi", "Phli"), ("Old_Turkic", "Orkh"), ("Kaithi", "Kthi"), ("Batak", "Batk"),
           ("Brahmi", "Brah"), ("Mandaic", "Mand"), ("Chakma", "Cakm"), ("Meroitic_Cursive", "Merc"), ("Meroitic_Hieroglyphs", "Mero"),
           ("Miao", "Plrd"), ("Sharada", "Shrd"), ("Sora_Sompeng", "Sora"), ("Takri", "Takr"), ("Caucasian_Albanian", "Aghb"),
           ("Bassa_Vah", "Bass"), ("Duployan", "Dupl"), ("Elbasan", "Elba"), ("Grantha", "Gran"), ("Pahawh_Hmong", "Hmng"), ("Khojki", "Khoj"),
           ("Linear_A", "Lina"), ("Mahajani", "Mahj"), ("Manichaean", "Mani"), ("Mende_Kikakui", "Mend"), ("Modi", "Modi"), ("Mro", "Mroo"),
           ("Old_North_Arabian", "Narb"), ("Nabataean", "Nbat"), ("Palmyrene", "Palm"), ("Pau_Cin_Hau", "Pauc"), ("Old_Permic", "Perm"),
           ("Psalter_Pahlavi", "Phlp"), ("Siddham", "Sidd"), ("Khudawadi", "Sind"), ("Tirhuta", "Tirh"), ("Warang_Citi", "Wara"),
           ("Ahom", "Ahom"), ("Anatolian_Hieroglyphs", "Hluw"), ("Hatran", "Hatr"), ("Multani", "Mult"), ("Old_Hungarian", "Hung"),
           ("SignWriting", "Sgnw"), ("Adlam", "Adlm"), ("Bhaiksuki", "Bhks"), ("Marchen", "Marc"), ("Newa", "Newa"), ("Osage", "Osge"),
           ("Tangut", "Tang"), ("Masaram_Gondi", "Gonm"), ("Nushu", "Nshu"), ("Soyombo", "Soyo"), ("Zanabazar_Square", "Zanb"),
           ("Dogra", "Dogr"), ("Gunjala_Gondi", "Gong"), ("Makasar", "Maka"), ("Medefaidrin", "Medf"), ("Hanifi_Rohingya", "Rohg"),
           ("Sogdian", "Sogd"), ("Old_Sogdian", "Sogo"), ("Elymaic", "Elym"), ("Nandinagari", "Nand"), ("Nyiakeng_Puachue_Hmong", "Hmnp"),
           ("Wancho", "Wcho"), ("Chorasmian", "Chrs"), ("Dives_Akuru", "Diak"), ("Khitan_Small_Script", "Kits"), ("Yezidi", "Yezi"),
           ("Cypro_Minoan", "Cpmn"), ("Old_Uyghur", "Ougr"), ("Tangsa", "Tnsa"), ("Toto", "Toto"), ("Vithkuqi", "Vith"), ("Kawi", "Kawi"),
           ("Nag_Mundari", "Nagm")]


def ranges(pred):
    out, start = [], None
    for cp in range(0x110000):
        if pred(cp):
            if start is None:
                start = cp
        elif start is not None:
            out.append((start, cp - 1))
            start = None
    if start is not None:
        out.append((start, 0x10FFFF))
    return out


def emit(f, name, rs):
    f.write("static const uint32_t %s[][2] = {\n" % name)
    for i in range(0, len(rs), 6):
        f.write("  " + " ".join("{0x%X,0x%X}," % r for r in rs[i:i + 6]) + "\n")
    f.write("};\n")
    f.write("static const int %s_COUNT = %d;\n\n" % (name, len(rs)))


def main():
    ver = unicodedata.unidata_version
    if ver != "15.0.0":
        sys.stderr.write("warning: unicodedata is %s, tables were pinned at 15.0.0\n" % ver)
    L = ranges(lambda cp: unicodedata.category(chr(cp)).startswith("L"))
    N = ranges(lambda cp: unicodedata.category(chr(cp)).startswith("N"))
    with open(OUT, "w") as f:
        f.write("/* GENERATED by tools/gen_unicode_tables.py from Python unicodedata %s - do not edit.\n" % ver)
        f.write(" * Inclusive code point ranges.  JTK_UC_L = \\p{L}, JTK_UC_N = \\p{N}, JTK_UC_WS = \\s under\n")
        f.write(" * Pattern.UNICODE_CHARACTER_CLASS (reference: EncodingFactory.java:129). */\n")
        f.write("#define JTK_UNICODE_VERSION \"%s\"\n\n" % ver)
        emit(f, "JTK_UC_L", L)
        emit(f, "JTK_UC_N", N)
        emit(f, "JTK_UC_WS", WHITE_SPACE)
        # every general category (java.util.regex \\p{Lu}, \\p{IsLu}, \\p{gc=Lu}; Character.getType) as (lo, hi, category index) triples
        cats = {}
        for cp in range(0x110000):
            cats.setdefault(unicodedata.category(chr(cp)), []).append(cp)
        triples = []
        for gi, g in enumerate(GC_NAMES):
            if g == "Cn":
                continue  # unassigned: everything no other category covers
            cps = set(cats.get(g, []))
            for lo, hi in ranges(lambda cp: cp in cps):
                triples.append((lo, hi, gi))
        triples.sort()
        f.write("/* general categories in the order %s */\n" % " ".join(GC_NAMES))
        f.write("#define JTK_UC_GC_NAMES \"%s\"\n" % " ".join(GC_NAMES))
        f.write("static const uint32_t JTK_UC_GC[][3] = {\n")
        for i in range(0, len(triples), 5):
            f.write("  " + " ".join("{0x%X,0x%X,%d}," % t for t in triples[i:i + 5]) + "\n")
        f.write("};\nstatic const int JTK_UC_GC_COUNT = %d;\n\n" % len(triples))
        # the Alphabetic property (\\w and \\p{IsAlphabetic} under UNICODE_CHARACTER_CLASS): from the `regex` module, restricted to the
        # code points this Unicode version has assigned
        import regex
        alpha_re = regex.compile(r"\p{Alphabetic}")
        A = ranges(lambda cp: unicodedata.category(chr(cp)) not in ("Cn", "Cs") and alpha_re.match(chr(cp)) is not None)
        emit(f, "JTK_UC_ALPHA", A)
        # Unicode scripts (java.util.regex \\p{IsHan}, \\p{script=Han}, \\p{sc=Hani}; Character.UnicodeScript): from the `regex` module,
        # restricted to the code points this Unicode version has assigned; (lo, hi, script index) triples + the names (long name, ISO 15924 code)
        every = "".join(chr(cp) for cp in range(0x110000) if not 0xD800 <= cp <= 0xDFFF and unicodedata.category(chr(cp)) != "Cn")
        striples, snames = [], []
        for long_name, code in SCRIPTS:
            try:
                rx = regex.compile(r"\p{Script=%s}+" % long_name)
            except regex.error:
                continue
            rs = [(ord(m.group()[0]), ord(m.group()[-1])) for m in rx.finditer(every)]
            # (runs of `every` skip unassigned code points: split a run wherever the code points are not consecutive)
            fixed = []
            for m in rx.finditer(every):
                run = m.group()
                lo = prev = ord(run[0])
                for ch in run[1:]:
                    c = ord(ch)
                    if c != prev + 1:
                        fixed.append((lo, prev))
                        lo = c
                    prev = c
                fixed.append((lo, prev))
            if not fixed:
                continue
            si = len(snames)
            snames.append((long_name, code))
            striples += [(lo, hi, si) for lo, hi in fixed]
        striples.sort()
        f.write("/* Unicode scripts: JTK_UC_SCRIPT_NAMES[index] = \"LONG_NAME Code\" */\n")
        f.write("static const char *const JTK_UC_SCRIPT_NAMES[] = {\n")
        for i in range(0, len(snames), 6):
            f.write("  " + " ".join('"%s %s",' % (n.upper(), c) for n, c in snames[i:i + 6]) + "\n")
        f.write("};\nstatic const int JTK_UC_SCRIPT_NAME_COUNT = %d;\n" % len(snames))
        f.write("static const uint32_t JTK_UC_SCRIPT[][3] = {\n")
        for i in range(0, len(striples), 5):
            f.write("  " + " ".join("{0x%X,0x%X,%d}," % t for t in striples[i:i + 5]) + "\n")
        f.write("};\nstatic const int JTK_UC_SCRIPT_COUNT = %d;\n\n" % len(striples))
    print("wrote %s: %d L ranges, %d N ranges, %d category ranges, %d Alphabetic ranges, %d script ranges of %d scripts" %
          (OUT, len(L), len(N), len(triples), len(A), len(striples), len(snames)))


if __name__ == "__main__":
    main()
