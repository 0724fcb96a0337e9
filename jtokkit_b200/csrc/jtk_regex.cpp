/* Compiler for general split patterns: java.util.regex subset -> jtk_rx_inst program (see jtk_regex.h). */
#include "jtk_regex_compile.h"

#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "unicode_ranges.inc"

namespace {

enum { N_SET, N_CAT, N_ALT, N_REP, N_LOOK, N_EMPTY, N_BOL, N_EOL /* min = 1: \\z */, N_WORDB, N_LOOKB /* set: the one character before */ };

struct node {
	int type = N_EMPTY;
	int set = -1;                          /* N_SET; N_WORDB: the word-character set */
	int set2 = -1, set3 = -1;              /* N_WORDB: non-spacing marks, letters-or-digits */
	std::vector<std::unique_ptr<node>> kids; /* N_CAT, N_ALT */
	std::unique_ptr<node> sub;             /* N_REP, N_LOOK */
	int min = 0, max = 0, mode = 0;        /* N_REP */
	bool neg = false;                      /* N_LOOK */
};

struct set_builder {
	bool neg = false, dot = false;
	std::vector<std::pair<uint32_t, uint32_t>> ranges;
	std::vector<std::pair<uint32_t, uint32_t>> prop_ranges; /* from \\p{..} / \\d / \\w: Unicode properties, not subject to case folding */
	uint32_t flags = 0;
	bool ci = false;
};

struct parser {
	std::vector<uint32_t> p;
	size_t i = 0;
	int flags;
	std::string err;
	jtk_rx_compiled *out;
	int look_depth = 0, max_look_depth = 0;

	bool failed() const { return !err.empty(); }
	void fail(const std::string &m) {
		if (err.empty()) err = m + " (at pattern index " + std::to_string(i) + ")";
	}
	int peek() const { return i < p.size() ? (int) p[i] : -1; }
	bool eat(int c) {
		if (peek() == c) {
			i++;
			return true;
		}
		return false;
	}

	int finish_set(set_builder &b) {
		fold_case(b);
		if (failed()) return 0;
		b.ranges.insert(b.ranges.end(), b.prop_ranges.begin(), b.prop_ranges.end());
		return emit_set(b);
	}

	/* case-insensitive: close the literal ranges under ASCII case folding (+ U+017F / U+212A with UNICODE_CASE) */
	void fold_case(set_builder &b) {
		const bool ucase = (flags & (JTK_RE_UNICODE_CASE | JTK_RE_UNICODE_CHARACTER_CLASS)) != 0;
		if (b.ci) {
			std::vector<std::pair<uint32_t, uint32_t>> extra;
			for (auto &r : b.ranges) {
				for (uint32_t c = std::max<uint32_t>(r.first, 'a'); c <= std::min<uint32_t>(r.second, 'z'); c++) extra.emplace_back(c - 32, c - 32);
				for (uint32_t c = std::max<uint32_t>(r.first, 'A'); c <= std::min<uint32_t>(r.second, 'Z'); c++) extra.emplace_back(c + 32, c + 32);
				if (ucase) {
					if ((r.first <= 's' && 's' <= r.second) || (r.first <= 'S' && 'S' <= r.second)) extra.emplace_back(0x17F, 0x17F);
					if ((r.first <= 'k' && 'k' <= r.second) || (r.first <= 'K' && 'K' <= r.second)) extra.emplace_back(0x212A, 0x212A);
					if (r.first <= 0x17F && 0x17F <= r.second) {
						extra.emplace_back('s', 's');
						extra.emplace_back('S', 'S');
					}
					if (r.first <= 0x212A && 0x212A <= r.second) {
						extra.emplace_back('k', 'k');
						extra.emplace_back('K', 'K');
					}
				}
				if (ucase && r.second >= 0x80 && !(r.first == r.second && (r.first == 0x17F || r.first == 0x212A))) {
					fail("UNICODE_CASE matching of non-ASCII characters is not supported");
					return;
				}
			}
			b.ranges.insert(b.ranges.end(), extra.begin(), extra.end());
			b.ci = false; /* (folded) */
		}
	}

	typedef std::vector<std::pair<uint32_t, uint32_t>> rlist;
	static rlist norm(rlist v) { /* sorted, disjoint, adjacent ranges joined */
		std::sort(v.begin(), v.end());
		rlist m;
		for (auto &r : v) {
			if (!m.empty() && r.first <= m.back().second + 1) m.back().second = std::max(m.back().second, r.second);
			else m.push_back(r);
		}
		return m;
	}
	static rlist complement(const rlist &v) {
		rlist out;
		add_ranges(out, norm(v), true);
		return out;
	}
	static rlist intersect(const rlist &a0, const rlist &b0) {
		const rlist a = norm(a0), b = norm(b0);
		rlist out;
		size_t i = 0, j = 0;
		while (i < a.size() && j < b.size()) {
			const uint32_t lo = std::max(a[i].first, b[j].first), hi = std::min(a[i].second, b[j].second);
			if (lo <= hi) out.emplace_back(lo, hi);
			if (a[i].second < b[j].second) i++;
			else j++;
		}
		return out;
	}
	/* every code point the class accepts, as explicit ranges (case folding, property flags and the negation applied): what nested
	 * classes and intersections are computed on */
	rlist materialise(set_builder b) {
		fold_case(b);
		rlist r = b.ranges;
		r.insert(r.end(), b.prop_ranges.begin(), b.prop_ranges.end());
		const rlist L = table_ranges(JTK_UC_L, JTK_UC_L_COUNT), N = table_ranges(JTK_UC_N, JTK_UC_N_COUNT), S = table_ranges(JTK_UC_WS, JTK_UC_WS_COUNT);
		auto add = [&](const rlist &t, bool neg) { add_ranges(r, t, neg); };
		if (b.flags & JTK_RX_HAS_L) add(L, false);
		if (b.flags & JTK_RX_HAS_N) add(N, false);
		if (b.flags & JTK_RX_HAS_S) add(S, false);
		if (b.flags & JTK_RX_HAS_NOT_L) add(L, true);
		if (b.flags & JTK_RX_HAS_NOT_N) add(N, true);
		if (b.flags & JTK_RX_HAS_NOT_S) add(S, true);
		if (b.dot) r = complement({{'\n', '\n'}, {'\r', '\r'}, {0x85, 0x85}, {0x2028, 0x2029}});
		r = norm(r);
		return b.neg ? complement(r) : r;
	}

	/* the jtk_rx_set of a builder whose ranges are final */
	int emit_set(set_builder &b) {
		std::sort(b.ranges.begin(), b.ranges.end());
		std::vector<std::pair<uint32_t, uint32_t>> merged;
		for (auto &r : b.ranges) {
			if (!merged.empty() && r.first <= merged.back().second + 1) merged.back().second = std::max(merged.back().second, r.second);
			else merged.push_back(r);
		}
		jtk_rx_set s;
		memset(&s, 0, sizeof(s));
		s.flags = b.flags | (b.neg ? JTK_RX_NEG : 0) | (b.dot ? JTK_RX_DOT : 0);
		const bool ucc = (flags & JTK_RE_UNICODE_CHARACTER_CLASS) != 0;
		for (uint32_t c = 0; c < 128; c++) {
			bool in = false;
			for (auto &r : merged)
				if (c >= r.first && c <= r.second) in = true;
			const bool isl = (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
			const bool isn = c >= '0' && c <= '9';
			const bool iss = c == ' ' || (c >= 9 && c <= 13);
			(void) ucc;
			in = in || ((s.flags & JTK_RX_HAS_L) && isl) || ((s.flags & JTK_RX_HAS_N) && isn) || ((s.flags & JTK_RX_HAS_S) && iss) ||
			     ((s.flags & JTK_RX_HAS_NOT_L) && !isl) || ((s.flags & JTK_RX_HAS_NOT_N) && !isn) || ((s.flags & JTK_RX_HAS_NOT_S) && !iss);
			if (in) s.ascii[c >> 5] |= 1u << (c & 31);
		}
		s.range_begin = (int32_t) (out->ranges.size() / 2);
		for (auto &r : merged)
			if (r.second >= 128) {
				out->ranges.push_back(std::max<uint32_t>(r.first, 128));
				out->ranges.push_back(r.second);
				s.range_count++;
			}
		out->sets.push_back(s);
		return (int) out->sets.size() - 1;
	}

	/* Unicode general categories (unicode_ranges.inc: JTK_UC_GC) as a bit mask over the category indices; 0 = unknown name */
	static uint32_t gc_mask_of(const std::string &name) {
		static const char *const names = JTK_UC_GC_NAMES;
		if (name.size() == 2) {
			for (int g = 0; g < 30; g++)
				if (names[3 * g] == name[0] && names[3 * g + 1] == name[1]) return 1u << g;
			if (name == "LC") return 7u; /* Lu Ll Lt */
			return 0;
		}
		if (name.size() == 1) {
			uint32_t m = 0;
			for (int g = 0; g < 30; g++)
				if (names[3 * g] == name[0]) m |= 1u << g;
			return m;
		}
		return 0;
	}
	/* adds the code points whose category is in `mask` (or is not, for neg) */
	static void add_gc(std::vector<std::pair<uint32_t, uint32_t>> &dst, uint32_t mask, bool neg) {
		std::vector<std::pair<uint32_t, uint32_t>> in;
		uint32_t next = 0; /* first code point not yet covered by a listed (assigned) range: the gaps are Cn */
		const bool cn = (mask >> 29) & 1u;
		for (int k = 0; k < JTK_UC_GC_COUNT; k++) {
			const uint32_t lo = JTK_UC_GC[k][0], hi = JTK_UC_GC[k][1], g = JTK_UC_GC[k][2];
			if (cn && lo > next) in.emplace_back(next, lo - 1);
			if ((mask >> g) & 1u) in.emplace_back(lo, hi);
			next = hi + 1;
		}
		if (cn && next <= 0x10FFFF) in.emplace_back(next, 0x10FFFF);
		add_ranges(dst, in, neg);
	}
	/* dst += in, or its complement within [0, 0x10FFFF] (in: sorted, disjoint) */
	static void add_ranges(std::vector<std::pair<uint32_t, uint32_t>> &dst, std::vector<std::pair<uint32_t, uint32_t>> in, bool neg) {
		std::sort(in.begin(), in.end());
		if (!neg) {
			dst.insert(dst.end(), in.begin(), in.end());
			return;
		}
		uint32_t next = 0;
		for (auto &r : in) {
			if (r.first > next) dst.emplace_back(next, r.first - 1);
			next = std::max(next, r.second + 1);
		}
		if (next <= 0x10FFFF) dst.emplace_back(next, 0x10FFFF);
	}
	/* dst += the code points of the script `name` (or of all others, for neg); false when no such script is known */
	static bool add_script(std::vector<std::pair<uint32_t, uint32_t>> &dst, const std::string &name, bool neg) {
		int si = -1;
		for (int k = 0; k < JTK_UC_SCRIPT_NAME_COUNT && si < 0; k++) {
			const char *e = JTK_UC_SCRIPT_NAMES[k]; /* "LONG_NAME Code" */
			const char *sp = strchr(e, ' ');
			auto same = [](const char *a, size_t n, const std::string &b) {
				if (b.size() != n) return false;
				for (size_t i = 0; i < n; i++)
					if (tolower((unsigned char) a[i]) != tolower((unsigned char) b[i])) return false;
				return true;
			};
			if (same(e, (size_t) (sp - e), name) || same(sp + 1, strlen(sp + 1), name)) si = k;
		}
		if (si < 0) return false;
		std::vector<std::pair<uint32_t, uint32_t>> in;
		for (int k = 0; k < JTK_UC_SCRIPT_COUNT; k++)
			if ((int) JTK_UC_SCRIPT[k][2] == si) in.emplace_back(JTK_UC_SCRIPT[k][0], JTK_UC_SCRIPT[k][1]);
		add_ranges(dst, in, neg);
		return true;
	}
	static std::vector<std::pair<uint32_t, uint32_t>> table_ranges(const uint32_t (*t)[2], int n) {
		std::vector<std::pair<uint32_t, uint32_t>> v;
		for (int k = 0; k < n; k++) v.emplace_back(t[k][0], t[k][1]);
		return v;
	}
	/* \w under UNICODE_CHARACTER_CLASS: [\p{Alpha}\p{gc=Mn}\p{gc=Me}\p{gc=Mc}\p{Digit}\p{gc=Pc}\p{IsJoin_Control}] (java.util.regex.Pattern) */
	static std::vector<std::pair<uint32_t, uint32_t>> unicode_word() {
		std::vector<std::pair<uint32_t, uint32_t>> w = table_ranges(JTK_UC_ALPHA, JTK_UC_ALPHA_COUNT);
		add_gc(w, (1u << 5) | (1u << 6) | (1u << 7) | (1u << 8) | (1u << 11), false); /* Mn Mc Me Nd Pc */
		w.emplace_back(0x200C, 0x200D);
		std::sort(w.begin(), w.end());
		std::vector<std::pair<uint32_t, uint32_t>> m;
		for (auto &r : w) {
			if (!m.empty() && r.first <= m.back().second + 1) m.back().second = std::max(m.back().second, r.second);
			else m.push_back(r);
		}
		return m;
	}

	static int hexval(int c) {
		if (c >= '0' && c <= '9') return c - '0';
		if (c >= 'a' && c <= 'f') return c - 'a' + 10;
		if (c >= 'A' && c <= 'F') return c - 'A' + 10;
		return -1;
	}

	/* after a backslash: adds to the set or returns a literal (>= 0) */
	int escape(set_builder &b) {
		int c = peek();
		if (c < 0) {
			fail("dangling backslash");
			return -1;
		}
		i++;
		const bool ucc = (flags & JTK_RE_UNICODE_CHARACTER_CLASS) != 0;
		switch (c) {
		case 'r': return '\r';
		case 'n': return '\n';
		case 't': return '\t';
		case 'f': return '\f';
		case 'a': return 7;
		case 'e': return 27;
		case 's': /* [ \t\n\x0B\f\r], or White_Space under UNICODE_CHARACTER_CLASS */
			if (ucc) b.flags |= JTK_RX_HAS_S;
			else {
				b.ranges.emplace_back(9, 13);
				b.ranges.emplace_back(' ', ' ');
			}
			return -1;
		case 'S':
			if (ucc) b.flags |= JTK_RX_HAS_NOT_S;
			else {
				b.ranges.emplace_back(0, 8);
				b.ranges.emplace_back(14, 31);
				b.ranges.emplace_back(33, 0x10FFFF);
			}
			return -1;
		case 'h':
		case 'H': /* horizontal whitespace: [ \t\xA0\u1680\u180e\u2000-\u200a\u202f\u205f\u3000] */
			add_ranges(b.prop_ranges, {{9, 9}, {' ', ' '}, {0xA0, 0xA0}, {0x1680, 0x1680}, {0x180E, 0x180E}, {0x2000, 0x200A}, {0x202F, 0x202F}, {0x205F, 0x205F}, {0x3000, 0x3000}},
			           c == 'H');
			return -1;
		case 'v':
		case 'V': /* vertical whitespace: [\n\x0B\f\r\x85\u2028\u2029] */
			add_ranges(b.prop_ranges, {{0x0A, 0x0D}, {0x85, 0x85}, {0x2028, 0x2029}}, c == 'V');
			return -1;
		case 'd':
		case 'D': { /* [0-9], or \p{Nd} under UNICODE_CHARACTER_CLASS */
			std::vector<std::pair<uint32_t, uint32_t>> in;
			if (ucc) add_gc(in, 1u << 8, false);
			else in.emplace_back('0', '9');
			add_ranges(b.prop_ranges, in, c == 'D');
			return -1;
		}
		case 'w':
		case 'W': { /* [a-zA-Z_0-9], or the Unicode word characters under UNICODE_CHARACTER_CLASS */
			std::vector<std::pair<uint32_t, uint32_t>> in;
			if (ucc) in = unicode_word();
			else in = {{'0', '9'}, {'A', 'Z'}, {'_', '_'}, {'a', 'z'}};
			add_ranges(b.prop_ranges, in, c == 'W');
			return -1;
		}
		case 'x': {
			int v = 0;
			if (eat('{')) {
				int k = 0;
				while (hexval(peek()) >= 0) {
					v = v * 16 + hexval(peek());
					i++;
					k++;
				}
				if (!k || !eat('}') || v > 0x10FFFF) fail("bad \\x{...} escape");
			} else {
				for (int k = 0; k < 2; k++) {
					if (hexval(peek()) < 0) {
						fail("bad \\xhh escape");
						return -1;
					}
					v = v * 16 + hexval(peek());
					i++;
				}
			}
			return v;
		}
		case 'u': {
			int v = 0;
			for (int k = 0; k < 4; k++) {
				if (hexval(peek()) < 0) {
					fail("bad \\uhhhh escape");
					return -1;
				}
				v = v * 16 + hexval(peek());
				i++;
			}
			return v;
		}
		case 'p':
		case 'P': {
			bool neg = c == 'P';
			std::string name;
			if (eat('{')) {
				if (eat('^')) neg = !neg;
				while (peek() >= 0 && peek() != '}') name.push_back((char) p[i++]);
				if (!eat('}')) fail("unterminated \\p{...}");
			} else if (peek() >= 0) {
				name.push_back((char) p[i++]);
			}
			if (name == "L" || name == "IsL" || name == "gc=L" || name == "general_category=L") b.flags |= neg ? JTK_RX_HAS_NOT_L : JTK_RX_HAS_L;
			else if (name == "N" || name == "IsN" || name == "gc=N" || name == "general_category=N") b.flags |= neg ? JTK_RX_HAS_NOT_N : JTK_RX_HAS_N;
			else {
				std::string nm = name;
				bool is_prefix = false, script_only = false;
				for (const char *pre : {"Is", "gc=", "general_category=", "script=", "sc="})
					if (nm.rfind(pre, 0) == 0) {
						nm = nm.substr(strlen(pre));
						is_prefix = pre[0] == 'I';
						script_only = pre[0] == 's';
						break;
					}
				uint32_t mask = script_only ? 0u : gc_mask_of(nm);
				/* CharPredicates.forProperty(name, caseIns) of JDK 9+: under CASE_INSENSITIVE Lu, Ll and Lt each stand for all three */
				if (b.ci && (mask == 1u || mask == 2u || mask == 4u)) mask = 7u;
				if (script_only) { /* \p{script=Han}, \p{sc=Hani} (Character.UnicodeScript.forName: long name or ISO 15924 code, any case) */
					if (!add_script(b.prop_ranges, nm, neg)) fail("unknown script in \\p{...}");
				} else if (mask) add_gc(b.prop_ranges, mask, neg);
				else if (nm == "Alphabetic" || (ucc && nm == "Alpha")) add_ranges(b.prop_ranges, table_ranges(JTK_UC_ALPHA, JTK_UC_ALPHA_COUNT), neg);
				else if (nm == "Alpha") add_ranges(b.prop_ranges, {{'A', 'Z'}, {'a', 'z'}}, neg);
				else if (nm == "White_Space" || nm == "WhiteSpace" || (ucc && nm == "Space")) add_ranges(b.prop_ranges, table_ranges(JTK_UC_WS, JTK_UC_WS_COUNT), neg);
				else if (nm == "Space") add_ranges(b.prop_ranges, {{9, 13}, {' ', ' '}}, neg);
				else if (nm == "Digit") {
					std::vector<std::pair<uint32_t, uint32_t>> in;
					if (ucc) add_gc(in, 1u << 8, false);
					else in.emplace_back('0', '9');
					add_ranges(b.prop_ranges, in, neg);
				} else if (nm == "ASCII") add_ranges(b.prop_ranges, {{0, 127}}, neg);
				else if (is_prefix && add_script(b.prop_ranges, nm, neg)) { /* \p{IsHan}: binary properties and categories first, then scripts (Pattern.java) */
				} else fail("unsupported \\p{...} property (general categories, scripts, Alphabetic, White_Space, Alpha, Digit, Space and ASCII are supported)");
			}
			return -1;
		}
		default:
			if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9')) {
				fail("unsupported escape sequence");
				return -1;
			}
			return c;
		}
	}

	std::unique_ptr<node> set_node(set_builder &b) {
		auto nd = std::make_unique<node>();
		nd->type = N_SET;
		nd->set = finish_set(b);
		return nd;
	}

	/* after '[': the class up to its ']' (nested classes [a[b-d]] are members of the union, && intersects what stands on either side,
	 * a leading ^ negates the whole).  Classes without nesting and intersection keep their items as they are. */
	set_builder parse_class_body(bool ci) {
		set_builder b;
		b.ci = ci;
		const bool neg = eat('^');
		bool first = true, complex = false, have_acc = false;
		rlist acc, nested; /* intersection of the operands so far; nested classes of the current operand */
		auto close_operand = [&]() {
			rlist op = materialise(b);
			op.insert(op.end(), nested.begin(), nested.end());
			acc = have_acc ? intersect(acc, op) : norm(op);
			have_acc = true;
			b = set_builder();
			b.ci = ci;
			nested.clear();
		};
		for (;;) {
			int c = peek();
			if (c < 0) {
				fail("unterminated character class");
				break;
			}
			if (c == ']' && !first) {
				i++;
				break;
			}
			first = false;
			if (c == '[') {
				i++;
				set_builder in = parse_class_body(ci);
				if (failed()) break;
				const rlist m = materialise(in);
				nested.insert(nested.end(), m.begin(), m.end());
				complex = true;
				continue;
			}
			if (c == '&' && i + 1 < p.size() && p[i + 1] == '&') {
				i += 2;
				close_operand();
				complex = true;
				continue;
			}
			i++;
			int lit = c;
			if (c == '\\') {
				lit = escape(b);
				if (failed()) break;
				if (lit < 0) continue;
			}
			if (peek() == '-' && i + 1 < p.size() && p[i + 1] != ']') {
				i++;
				int hi = peek();
				i++;
				if (hi == '\\') {
					set_builder tmp;
					hi = escape(tmp);
					if (hi < 0) fail("bad range end in character class");
				}
				if (hi < lit) fail("illegal character range");
				if (failed()) break;
				b.ranges.emplace_back((uint32_t) lit, (uint32_t) hi);
			} else {
				b.ranges.emplace_back((uint32_t) lit, (uint32_t) lit);
			}
		}
		if (!complex || failed()) {
			b.neg = neg;
			return b;
		}
		close_operand();
		set_builder out_b;
		out_b.prop_ranges = acc; /* final: no further case folding */
		out_b.neg = neg;
		return out_b;
	}

	std::unique_ptr<node> parse_class(bool ci) {
		set_builder b = parse_class_body(ci);
		return set_node(b);
	}

	std::unique_ptr<node> parse_atom(bool &ci) {
		int c = peek();
		if (c == '(') {
			i++;
			bool sub_ci = ci;
			if (eat('?')) {
				if (eat(':')) {
				} else if (peek() == '!' || peek() == '=') {
					const bool neg = peek() == '!';
					i++;
					auto lk = std::make_unique<node>();
					lk->type = N_LOOK;
					lk->neg = neg;
					look_depth++;
					max_look_depth = std::max(max_look_depth, look_depth);
					lk->sub = parse_alt(sub_ci);
					look_depth--;
					if (!eat(')')) fail("missing ) after look-ahead");
					return lk;
				} else if (peek() == '<' && i + 1 < p.size() && (p[i + 1] == '=' || p[i + 1] == '!')) {
					/* look-behind over exactly one character: (?<=[set]) / (?<![set]) */
					const bool neg = p[i + 1] == '!';
					i += 2;
					auto sub = parse_alt(sub_ci);
					if (!eat(')')) fail("missing ) after look-behind");
					const node *one = sub.get();
					while (one->type == N_CAT && one->kids.size() == 1) one = one->kids[0].get();
					auto lb = std::make_unique<node>();
					lb->type = N_LOOKB;
					lb->neg = neg;
					if (one->type != N_SET) fail("look-behind over anything but one character is not supported");
					else lb->set = one->set;
					return lb;
				} else if (peek() == '<') { /* named group (?<name>X): the name is irrelevant here (no back-references) */
					i++;
					int k = 0;
					while (peek() >= 0 && peek() != '>') i++, k++;
					if (!k || !eat('>')) {
						fail("bad group name");
						return std::make_unique<node>();
					}
				} else {
					bool on = true, seen = false;
					while (peek() >= 0 && peek() != ')' && peek() != ':') {
						int f = peek();
						i++;
						seen = true;
						if (f == '-') on = false;
						else if (f == 'i') sub_ci = on;
						else if (f == 'u') flags |= on ? JTK_RE_UNICODE_CASE : 0;
						else {
							fail("unsupported inline flag");
							return std::make_unique<node>();
						}
					}
					if (!seen) {
						fail("bad group syntax");
						return std::make_unique<node>();
					}
					if (eat(')')) {
						ci = sub_ci;
						return std::make_unique<node>();
					}
					if (!eat(':')) {
						fail("bad inline flag group");
						return std::make_unique<node>();
					}
				}
			}
			auto g = parse_alt(sub_ci);
			if (!eat(')')) fail("missing )");
			return g;
		}
		if (c == '[') {
			i++;
			return parse_class(ci);
		}
		if (c == '.') {
			i++;
			set_builder b;
			b.dot = true;
			return set_node(b);
		}
		if (c == '^') {
			i++;
			auto nd = std::make_unique<node>();
			nd->type = N_BOL;
			return nd;
		}
		if (c == '$') {
			i++;
			auto nd = std::make_unique<node>();
			nd->type = N_EOL;
			return nd;
		}
		if (c == '\\') {
			i++;
			int nc = peek();
			if (nc == 'b' || nc == 'B') { /* word boundary (java.util.regex.Pattern.Bound) */
				i++;
				auto nd = std::make_unique<node>();
				nd->type = N_WORDB;
				nd->min = nc == 'B';
				const bool ucc2 = (flags & JTK_RE_UNICODE_CHARACTER_CLASS) != 0;
				set_builder w, mn, ld;
				/* isWord: the Unicode word characters under UNICODE_CHARACTER_CLASS, else '_' or Character.isLetterOrDigit */
				if (ucc2) w.prop_ranges = unicode_word();
				else {
					add_gc(w.prop_ranges, 0x1Fu | (1u << 8), false);
					w.prop_ranges.emplace_back('_', '_');
				}
				add_gc(mn.prop_ranges, 1u << 5, false);           /* NON_SPACING_MARK */
				add_gc(ld.prop_ranges, 0x1Fu | (1u << 8), false); /* Character.isLetterOrDigit: L* or Nd */
				nd->set = finish_set(w);
				nd->set2 = finish_set(mn);
				nd->set3 = finish_set(ld);
				return nd;
			}
			if (nc == 'A' || nc == 'Z' || nc == 'z') { /* \A: '^' without MULTILINE; \Z: '$' without MULTILINE; \z: the very end of the input */
				i++;
				auto nd = std::make_unique<node>();
				nd->type = nc == 'A' ? N_BOL : N_EOL;
				nd->min = nc == 'z';
				return nd;
			}
			if (nc == 'R') { /* linebreak matcher: \r\n|[\n\x0B\f\r\x85\u2028\u2029] (the documented equivalence; JDK 9+ backtracks into it like that) */
				i++;
				auto crlf = std::make_unique<node>();
				crlf->type = N_CAT;
				for (uint32_t c2 : {(uint32_t) '\r', (uint32_t) '\n'}) {
					set_builder b1;
					b1.ranges.emplace_back(c2, c2);
					crlf->kids.push_back(set_node(b1));
				}
				set_builder b2;
				b2.ranges = {{0x0A, 0x0D}, {0x85, 0x85}, {0x2028, 0x2029}};
				auto alt = std::make_unique<node>();
				alt->type = N_ALT;
				alt->kids.push_back(std::move(crlf));
				alt->kids.push_back(set_node(b2));
				return alt;
			}
			if (nc == 'G' || nc == 'X' || nc == 'Q' || nc == 'k' || (nc >= '1' && nc <= '9')) {
				fail("unsupported escape (boundary / back-reference / quoting)");
				return std::make_unique<node>();
			}
			set_builder b;
			b.ci = ci;
			int lit = escape(b);
			if (lit >= 0) b.ranges.emplace_back((uint32_t) lit, (uint32_t) lit);
			return set_node(b);
		}
		if (c == '*' || c == '+' || c == '?') {
			fail("dangling quantifier");
			return std::make_unique<node>();
		}
		i++;
		set_builder b;
		b.ci = ci;
		b.ranges.emplace_back((uint32_t) c, (uint32_t) c);
		return set_node(b);
	}

	std::unique_ptr<node> parse_quantified(bool &ci) {
		auto atom = parse_atom(ci);
		for (;;) {
			int c = peek(), mn, mx;
			if (c == '*') {
				mn = 0;
				mx = -1;
				i++;
			} else if (c == '+') {
				mn = 1;
				mx = -1;
				i++;
			} else if (c == '?') {
				mn = 0;
				mx = 1;
				i++;
			} else if (c == '{') {
				i++;
				int a = 0, k = 0;
				while (peek() >= '0' && peek() <= '9') {
					a = a * 10 + (peek() - '0');
					i++;
					k++;
				}
				mn = mx = a;
				if (eat(',')) {
					mx = -1;
					int b = 0, kb = 0;
					while (peek() >= '0' && peek() <= '9') {
						b = b * 10 + (peek() - '0');
						i++;
						kb++;
					}
					if (kb) mx = b;
				}
				if (!k || !eat('}') || (mx >= 0 && mx < mn)) {
					fail("bad {n,m} quantifier");
					return atom;
				}
			} else {
				return atom;
			}
			if (atom->type != N_SET && atom->type != N_CAT && atom->type != N_ALT && atom->type != N_REP) {
				fail("quantifier on a zero-width construct is not supported");
				return atom;
			}
			auto rep = std::make_unique<node>();
			rep->type = N_REP;
			rep->min = mn;
			rep->max = mx;
			rep->mode = eat('?') ? 1 : eat('+') ? 2 : 0;
			rep->sub = std::move(atom);
			atom = std::move(rep);
		}
	}

	std::unique_ptr<node> parse_cat(bool ci) {
		auto cat = std::make_unique<node>();
		cat->type = N_CAT;
		while (!failed()) {
			int c = peek();
			if (c < 0 || c == '|' || c == ')') break;
			cat->kids.push_back(parse_quantified(ci));
		}
		return cat;
	}

	std::unique_ptr<node> parse_alt(bool ci) {
		auto first = parse_cat(ci);
		if (peek() != '|') return first;
		auto alt = std::make_unique<node>();
		alt->type = N_ALT;
		alt->kids.push_back(std::move(first));
		while (!failed() && eat('|')) alt->kids.push_back(parse_cat(ci));
		return alt;
	}

	/* can the node match the empty string? */
	static bool nullable(const node *nd) {
		switch (nd->type) {
		case N_SET: return false;
		case N_CAT:
			for (auto &k : nd->kids)
				if (!nullable(k.get())) return false;
			return true;
		case N_ALT:
			for (auto &k : nd->kids)
				if (nullable(k.get())) return true;
			return false;
		case N_REP: return nd->min == 0 || nullable(nd->sub.get());
		default: return true;
		}
	}

	void emit(const node *nd) {
		auto &code = out->inst;
		switch (nd->type) {
		case N_EMPTY: break;
		case N_SET: code.push_back({JTK_RX_SET, nd->set, 0, 0, 0}); break;
		case N_BOL: code.push_back({JTK_RX_BOL, 0, 0, 0, 0}); break;
		case N_EOL: code.push_back({JTK_RX_EOL, nd->min, 0, 0, 0}); break;
		case N_LOOKB: code.push_back({JTK_RX_LOOKB, nd->neg ? 1 : 0, nd->set, 0, 0}); break;
		case N_WORDB: code.push_back({JTK_RX_WORDB, nd->min, nd->set, nd->set2, nd->set3}); break;
		case N_CAT:
			for (auto &k : nd->kids) emit(k.get());
			break;
		case N_ALT: {
			std::vector<size_t> jumps;
			for (size_t b = 0; b < nd->kids.size(); b++) {
				size_t split = 0;
				const bool last = b + 1 == nd->kids.size();
				if (!last) {
					split = code.size();
					code.push_back({JTK_RX_SPLIT, 0, 0, 0, 0});
				}
				if (!last) code[split].a = (int32_t) code.size();
				emit(nd->kids[b].get());
				if (!last) {
					jumps.push_back(code.size());
					code.push_back({JTK_RX_JMP, 0, 0, 0, 0});
					code[split].b = (int32_t) code.size();
				}
			}
			for (size_t j : jumps) code[j].a = (int32_t) code.size();
			break;
		}
		case N_LOOK: {
			/* LOOK jumps over the inlined sub-program */
			const size_t look = code.size();
			code.push_back({JTK_RX_LOOK, nd->neg ? 1 : 0, 0, 0, 0});
			const size_t jmp = code.size();
			code.push_back({JTK_RX_JMP, 0, 0, 0, 0});
			code[look].b = (int32_t) code.size();
			emit(nd->sub.get());
			code.push_back({JTK_RX_MATCH, 0, 0, 0, 0});
			code[jmp].a = (int32_t) code.size();
			/* LOOK continues at pc + 1 = the JMP, which skips the sub-program */
			break;
		}
		case N_REP: {
			const node *sub = nd->sub.get();
			if (sub->type == N_SET) {
				code.push_back({JTK_RX_REP, sub->set, nd->min, nd->max, nd->mode});
				break;
			}
			if (nullable(sub)) {
				fail("a loop over a sub-expression that can match the empty string is not supported");
				break;
			}
			if (nd->mode == 2) {
				fail("possessive quantifiers on groups are not supported");
				break;
			}
			if (nd->min > 64 || nd->max > 64) {
				fail("counted loops over groups beyond 64 are not supported");
				break;
			}
			for (int k = 0; k < nd->min; k++) emit(sub);
			if (nd->max < 0) { /* X* : L: SPLIT(body, out) body: X JMP L out: */
				const size_t l = code.size();
				code.push_back({JTK_RX_SPLIT, 0, 0, 0, 0});
				const size_t body = code.size();
				emit(sub);
				code.push_back({JTK_RX_JMP, (int32_t) l, 0, 0, 0});
				const size_t outp = code.size();
				code[l].a = nd->mode == 1 ? (int32_t) outp : (int32_t) body;
				code[l].b = nd->mode == 1 ? (int32_t) body : (int32_t) outp;
			} else {
				/* optional copies, nested so that each further copy is only tried after the previous one */
				std::vector<size_t> splits;
				for (int k = nd->min; k < nd->max; k++) {
					splits.push_back(code.size());
					code.push_back({JTK_RX_SPLIT, 0, 0, 0, 0});
					const size_t body = code.size();
					code[splits.back()].a = (int32_t) body;
					emit(sub);
				}
				for (size_t sidx : splits) {
					if (nd->mode == 1) {
						code[sidx].b = code[sidx].a;
						code[sidx].a = (int32_t) code.size();
					} else {
						code[sidx].b = (int32_t) code.size();
					}
				}
			}
			break;
		}
		}
	}
};

} /* namespace */

int jtk_rx_compile(const char *pattern, int flags, jtk_rx_compiled *out, std::string *err) {
	parser ps;
	ps.flags = flags;
	ps.out = out;
	out->inst.clear();
	out->sets.clear();
	out->ranges.clear();
	for (size_t i = 0, n = strlen(pattern); i < n;) {
		const unsigned char b = (unsigned char) pattern[i];
		uint32_t cp;
		int len;
		if (b < 0x80) {
			cp = b;
			len = 1;
		} else if ((b & 0xE0) == 0xC0) {
			cp = b & 0x1F;
			len = 2;
		} else if ((b & 0xF0) == 0xE0) {
			cp = b & 0x0F;
			len = 3;
		} else {
			cp = b & 0x07;
			len = 4;
		}
		for (int k = 1; k < len && i + (size_t) k < n; k++) cp = (cp << 6) | ((unsigned char) pattern[i + (size_t) k] & 0x3F);
		ps.p.push_back(cp);
		i += (size_t) len;
	}
	{ /* \Q...\E quoting, as Pattern.RemoveQEQuoting does it: the quoted characters become escaped literals (letters and digits as they are) */
		std::vector<uint32_t> q;
		bool quoted = false;
		for (size_t i = 0; i < ps.p.size(); i++) {
			const uint32_t c = ps.p[i];
			const bool bs = c == '\\' && i + 1 < ps.p.size();
			if (!quoted && bs && ps.p[i + 1] == 'Q') {
				quoted = true;
				i++;
			} else if (quoted && bs && ps.p[i + 1] == 'E') {
				quoted = false;
				i++;
			} else if (quoted) {
				if (!((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 128)) q.push_back('\\');
				q.push_back(c);
			} else {
				q.push_back(c);
				if (bs) q.push_back(ps.p[++i]); /* an escaped backslash does not begin a \Q */
			}
		}
		ps.p.swap(q);
	}
	bool ci = (flags & JTK_RE_CASE_INSENSITIVE) != 0;
	auto root = ps.parse_alt(ci);
	if (!ps.failed() && ps.i < ps.p.size()) ps.fail("unmatched )");
	if (!ps.failed() && ps.max_look_depth > 2) ps.fail("look-ahead nested more than twice is not supported");
	if (!ps.failed()) {
		out->nullable = parser::nullable(root.get());
		ps.emit(root.get());
		out->inst.push_back({JTK_RX_MATCH, 0, 0, 0, 0});
	}
	if (ps.failed()) {
		*err = ps.err;
		return JTK_E_PATTERN_UNSUPPORTED;
	}
	if (out->ranges.empty()) out->ranges.push_back(0);
	/* first bytes: walk the program from its entry over everything that consumes no character */
	memset(out->first, 0, sizeof(out->first));
	{
		bool any = out->nullable;
		std::vector<char> seen(out->inst.size(), 0);
		std::vector<int> todo(1, 0);
		while (!todo.empty() && !any) {
			const int pc = todo.back();
			todo.pop_back();
			if (pc < 0 || pc >= (int) out->inst.size() || seen[(size_t) pc]) continue;
			seen[(size_t) pc] = 1;
			const jtk_rx_inst &in = out->inst[(size_t) pc];
			switch (in.op) {
			case JTK_RX_SET:
			case JTK_RX_REP: {
				const jtk_rx_set &st = out->sets[(size_t) in.a];
				for (int w = 0; w < 4; w++) /* the bitmap holds the un-negated ASCII members; `.` ignores it */
					out->first[w] |= (st.flags & JTK_RX_DOT) ? ~0u : (st.flags & JTK_RX_NEG) ? ~st.ascii[w] : st.ascii[w];
				if (st.range_count > 0 || (st.flags & ~0u)) /* can hold non-ASCII characters (ranges, classes, negation, dot): any byte >= 0x80 */
					for (int w = 4; w < 8; w++) out->first[w] = ~0u;
				if (in.op == JTK_RX_REP && in.b == 0) todo.push_back(pc + 1); /* may take nothing */
				break;
			}
			case JTK_RX_SPLIT:
				todo.push_back(in.a);
				todo.push_back(in.b);
				break;
			case JTK_RX_JMP: todo.push_back(in.a); break;
			case JTK_RX_LOOK: /* zero width: the sub-program does not consume, matching goes on at pc + 1 */
			case JTK_RX_BOL:
			case JTK_RX_EOL:
			case JTK_RX_LOOKB:
			case JTK_RX_WORDB: todo.push_back(pc + 1); break;
			default: any = true; /* MATCH reached without consuming: the pattern can match the empty string */
			}
		}
		if (any) memset(out->first, 0xFF, sizeof(out->first));
	}
	return JTK_OK;
}
