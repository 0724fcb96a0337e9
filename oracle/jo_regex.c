/*
 * ORACLE - TEST INFRASTRUCTURE ONLY (see jo_regex.h).  Backtracking matcher with the
 * java.util.regex semantics used by GptBytePairEncoding.java:77-80.
 */
#include "jo_regex.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>

#include "../jtokkit_b200/csrc/unicode_ranges.inc"

/* ------------------------------------------------------------------ unicode classes */
static int in_ranges(const uint32_t (*r)[2], int n, uint32_t cp) {
	int lo = 0, hi = n - 1;
	while (lo <= hi) {
		int mid = (lo + hi) >> 1;
		if (cp < r[mid][0]) hi = mid - 1;
		else if (cp > r[mid][1]) lo = mid + 1;
		else return 1;
	}
	return 0;
}

/* one flag byte per code point (1 = \p{L}, 2 = \p{N}, 4 = White_Space), filled on first use */
static uint8_t *uc_flags;
static void uc_init(void) {
	uint8_t *t = (uint8_t *) calloc(0x110000, 1);
	for (uint32_t cp = 0; cp < 0x110000; cp++)
		t[cp] = (uint8_t) ((in_ranges(JTK_UC_L, JTK_UC_L_COUNT, cp) ? 1 : 0) | (in_ranges(JTK_UC_N, JTK_UC_N_COUNT, cp) ? 2 : 0) |
		                   (in_ranges(JTK_UC_WS, JTK_UC_WS_COUNT, cp) ? 4 : 0));
	if (!__sync_bool_compare_and_swap(&uc_flags, NULL, t)) free(t);
}
static inline uint8_t uc_get(uint32_t cp) {
	if (!uc_flags) uc_init();
	return cp < 0x110000 ? uc_flags[cp] : 0;
}
/* general category index per code point (order of JTK_UC_GC_NAMES; 29 = Cn) and the Alphabetic property, filled on first use */
static uint8_t *uc_gc;
static void uc_gc_init(void) {
	uint8_t *t = (uint8_t *) malloc(0x110000);
	memset(t, 29, 0x110000);
	for (int k = 0; k < JTK_UC_GC_COUNT; k++)
		for (uint32_t cp = JTK_UC_GC[k][0]; cp <= JTK_UC_GC[k][1]; cp++) t[cp] = (uint8_t) JTK_UC_GC[k][2];
	for (int k = 0; k < JTK_UC_ALPHA_COUNT; k++)
		for (uint32_t cp = JTK_UC_ALPHA[k][0]; cp <= JTK_UC_ALPHA[k][1]; cp++) t[cp] |= 0x80;
	if (!__sync_bool_compare_and_swap(&uc_gc, NULL, t)) free(t);
}
static inline int uc_category(uint32_t cp) {
	if (!uc_gc) uc_gc_init();
	return cp < 0x110000 ? (uc_gc[cp] & 0x7F) : 29;
}
static inline int uc_alphabetic(uint32_t cp) {
	if (!uc_gc) uc_gc_init();
	return cp < 0x110000 ? (uc_gc[cp] >> 7) : 0;
}
/* \w under UNICODE_CHARACTER_CLASS: [\p{Alpha}\p{gc=Mn}\p{gc=Me}\p{gc=Mc}\p{Digit}\p{gc=Pc}\p{IsJoin_Control}] */
static int uc_word(uint32_t cp) {
	const int g = uc_category(cp);
	return uc_alphabetic(cp) || g == 5 || g == 6 || g == 7 || g == 8 || g == 11 || cp == 0x200C || cp == 0x200D;
}
/* script index per code point (order of JTK_UC_SCRIPT_NAMES; 255 = no script / unassigned), filled on first use */
static uint8_t *uc_script;
static void uc_script_init(void) {
	uint8_t *t = (uint8_t *) malloc(0x110000);
	memset(t, 255, 0x110000);
	for (int k = 0; k < JTK_UC_SCRIPT_COUNT; k++)
		for (uint32_t cp = JTK_UC_SCRIPT[k][0]; cp <= JTK_UC_SCRIPT[k][1]; cp++) t[cp] = (uint8_t) JTK_UC_SCRIPT[k][2];
	if (!__sync_bool_compare_and_swap(&uc_script, NULL, t)) free(t);
}
static inline int uc_script_of(uint32_t cp) {
	if (!uc_script) uc_script_init();
	return cp < 0x110000 ? uc_script[cp] : 255;
}
/* Character.UnicodeScript.forName: the long name or the ISO 15924 code, case-insensitive; -1 = unknown */
static int script_index_of(const char *name) {
	for (int k = 0; k < JTK_UC_SCRIPT_NAME_COUNT; k++) {
		const char *e = JTK_UC_SCRIPT_NAMES[k]; /* "LONG_NAME Code" */
		const char *sp = strchr(e, ' ');
		const size_t ln = (size_t) (sp - e);
		if ((strlen(name) == ln && !strncasecmp(name, e, ln)) || !strcasecmp(name, sp + 1)) return k;
	}
	return -1;
}
int jo_uc_is_letter(uint32_t cp) { return uc_get(cp) & 1; }
int jo_uc_is_number(uint32_t cp) { return (uc_get(cp) >> 1) & 1; }
int jo_uc_is_space(uint32_t cp) { return (uc_get(cp) >> 2) & 1; }

/* ------------------------------------------------------------------ node tree */
enum { N_SET, N_ANY, N_CAT, N_ALT, N_REP, N_LOOK, N_EMPTY, N_BOL, N_EOL /* min = 1: \z, the very end only */, N_WORDB, N_LOOKB /* sub: one set */ };
enum { C_L = 1, C_N, C_SPACE, C_DIGIT, C_WORD, C_GC /* mask of general categories */, C_ALPHA, C_ASCII, C_SCRIPT /* mask = script index */, C_HSPACE, C_VSPACE };
enum { Q_GREEDY, Q_LAZY, Q_POSSESSIVE };

typedef struct cls_item {
	int kind; /* C_* */
	int neg;
	uint32_t mask; /* C_GC */
} cls_item;

typedef struct node {
	int type;
	/* N_SET */
	int neg, ci;
	int nranges, ncls;
	uint32_t (*ranges)[2];
	cls_item *cls;
	int nnest;             /* N_SET: nested classes [a[b-d]] - members of the union, each with its own negation */
	struct node **nest;
	struct node *and_next; /* N_SET: the next operand of an intersection [..&&..] (the negation of the whole class sits on the first) */
	/* N_CAT / N_ALT */
	int nkids;
	struct node **kids;
	/* N_REP / N_LOOK */
	struct node *sub;
	int min, max, mode; /* max < 0: unbounded */
	int look_neg;
	/* matcher-side caches (no semantic content): which ASCII characters the node can start with / match */
	uint32_t first[4];   /* ASCII characters a match of this node may begin with */
	int first_other;     /* may begin with a non-ASCII character */
	int nullable;        /* may match the empty string */
	uint32_t amap[4];    /* N_SET: ASCII characters the set accepts (case folding and negation applied) */
} node;

struct jo_regex {
	node *root;
	int flags;
	node **all;
	int nall, call;
};

typedef struct parser {
	const uint32_t *p;
	int n, i;
	int flags; /* pattern-level flags */
	jo_regex *re;
	char *err;
	int errlen;
	int failed;
} parser;

static void fail(parser *ps, const char *msg) {
	if (!ps->failed && ps->err && ps->errlen > 0) snprintf(ps->err, (size_t) ps->errlen, "%s (at pattern index %d)", msg, ps->i);
	ps->failed = 1;
}

static node *new_node(parser *ps, int type) {
	node *nd = (node *) calloc(1, sizeof(node));
	nd->type = type;
	jo_regex *re = ps->re;
	if (re->nall == re->call) {
		re->call = re->call ? re->call * 2 : 64;
		re->all = (node **) realloc(re->all, sizeof(node *) * (size_t) re->call);
	}
	re->all[re->nall++] = nd;
	return nd;
}

static void add_kid(node *nd, node *kid) {
	nd->kids = (node **) realloc(nd->kids, sizeof(node *) * (size_t) (nd->nkids + 1));
	nd->kids[nd->nkids++] = kid;
}
static void add_range(node *nd, uint32_t lo, uint32_t hi) {
	nd->ranges = (uint32_t(*)[2]) realloc(nd->ranges, sizeof(uint32_t[2]) * (size_t) (nd->nranges + 1));
	nd->ranges[nd->nranges][0] = lo;
	nd->ranges[nd->nranges][1] = hi;
	nd->nranges++;
}
static void add_cls(node *nd, int kind, int neg) {
	nd->cls = (cls_item *) realloc(nd->cls, sizeof(cls_item) * (size_t) (nd->ncls + 1));
	nd->cls[nd->ncls].kind = kind;
	nd->cls[nd->ncls].neg = neg;
	nd->cls[nd->ncls].mask = 0;
	nd->ncls++;
}
static void add_gc_cls(node *nd, uint32_t mask, int neg) {
	add_cls(nd, C_GC, neg);
	nd->cls[nd->ncls - 1].mask = mask;
}
/* mask of general categories for a one- or two-letter name (JTK_UC_GC_NAMES order), 0 = unknown */
static uint32_t gc_mask_of(const char *name) {
	static const char *const names = JTK_UC_GC_NAMES;
	uint32_t m = 0;
	if (!strcmp(name, "LC")) return 7u;
	for (int g = 0; g < 30; g++) {
		if (name[0] && name[1] && !name[2] && names[3 * g] == name[0] && names[3 * g + 1] == name[1]) m |= 1u << g;
		if (name[0] && !name[1] && names[3 * g] == name[0]) m |= 1u << g;
	}
	return m;
}

static int peek(parser *ps) { return ps->i < ps->n ? (int) ps->p[ps->i] : -1; }
static int eat(parser *ps, int c) {
	if (peek(ps) == c) {
		ps->i++;
		return 1;
	}
	return 0;
}

static node *parse_alt(parser *ps, int ci);
static void jo_regex_prepare(jo_regex *re);

static int hexval(int c) {
	if (c >= '0' && c <= '9') return c - '0';
	if (c >= 'a' && c <= 'f') return c - 'a' + 10;
	if (c >= 'A' && c <= 'F') return c - 'A' + 10;
	return -1;
}

/* Parses the part after a backslash.  Either adds a class item / range to `set` and returns 1,
 * or fails.  `*lit` receives a literal code point (>=0) when the escape denotes a single char. */
static int parse_escape(parser *ps, node *set, int *lit) {
	int c = peek(ps);
	*lit = -1;
	if (c < 0) {
		fail(ps, "dangling backslash");
		return 0;
	}
	ps->i++;
	switch (c) {
	case 'r': *lit = '\r'; return 1;
	case 'n': *lit = '\n'; return 1;
	case 't': *lit = '\t'; return 1;
	case 'f': *lit = '\f'; return 1;
	case 'a': *lit = 7; return 1;
	case 'e': *lit = 27; return 1;
	case '0': { /* octal \0n, \0nn, \0mnn */
		int v = 0, k = 0;
		while (k < 3 && peek(ps) >= '0' && peek(ps) <= '7') {
			int nv = v * 8 + (peek(ps) - '0');
			if (nv > 0377) break;
			v = nv;
			ps->i++;
			k++;
		}
		if (!k) {
			fail(ps, "bad octal escape");
			return 0;
		}
		*lit = v;
		return 1;
	}
	case 'x': {
		int v = 0;
		if (eat(ps, '{')) {
			int k = 0;
			while (hexval(peek(ps)) >= 0) {
				v = v * 16 + hexval(peek(ps));
				ps->i++;
				k++;
			}
			if (!k || !eat(ps, '}') || v > 0x10FFFF) {
				fail(ps, "bad \\x{...} escape");
				return 0;
			}
		} else {
			for (int k = 0; k < 2; k++) {
				if (hexval(peek(ps)) < 0) {
					fail(ps, "bad \\xhh escape");
					return 0;
				}
				v = v * 16 + hexval(peek(ps));
				ps->i++;
			}
		}
		*lit = v;
		return 1;
	}
	case 'u': {
		int v = 0;
		for (int k = 0; k < 4; k++) {
			if (hexval(peek(ps)) < 0) {
				fail(ps, "bad \\uhhhh escape");
				return 0;
			}
			v = v * 16 + hexval(peek(ps));
			ps->i++;
		}
		*lit = v;
		return 1;
	}
	case 's': add_cls(set, C_SPACE, 0); return 1;
	case 'S': add_cls(set, C_SPACE, 1); return 1;
	case 'h': add_cls(set, C_HSPACE, 0); return 1; /* horizontal / vertical whitespace (the same sets with and without UNICODE_CHARACTER_CLASS) */
	case 'H': add_cls(set, C_HSPACE, 1); return 1;
	case 'v': add_cls(set, C_VSPACE, 0); return 1;
	case 'V': add_cls(set, C_VSPACE, 1); return 1;
	case 'd':
	case 'D':
	case 'w':
	case 'W':
		add_cls(set, (c == 'd' || c == 'D') ? C_DIGIT : C_WORD, c == 'D' || c == 'W'); /* ASCII or Unicode meaning: cls_match */
		return 1;
	case 'p':
	case 'P': {
		int neg = (c == 'P');
		char name[48];
		int k = 0;
		if (eat(ps, '{')) {
			if (eat(ps, '^')) neg = !neg;
			while (peek(ps) >= 0 && peek(ps) != '}' && k < 47) name[k++] = (char) ps->p[ps->i++];
			if (!eat(ps, '}')) {
				fail(ps, "unterminated \\p{...}");
				return 0;
			}
		} else if (peek(ps) >= 0) {
			name[k++] = (char) ps->p[ps->i++];
		}
		name[k] = 0;
		if (!strcmp(name, "L") || !strcmp(name, "IsL") || !strcmp(name, "gc=L") || !strcmp(name, "general_category=L")) add_cls(set, C_L, neg);
		else if (!strcmp(name, "N") || !strcmp(name, "IsN") || !strcmp(name, "gc=N") || !strcmp(name, "general_category=N")) add_cls(set, C_N, neg);
		else {
			const char *nm = name;
			int is_prefix = 0, script_only = 0;
			if (!strncmp(nm, "Is", 2)) nm += 2, is_prefix = 1;
			else if (!strncmp(nm, "gc=", 3)) nm += 3;
			else if (!strncmp(nm, "general_category=", 17)) nm += 17;
			else if (!strncmp(nm, "script=", 7)) nm += 7, script_only = 1;
			else if (!strncmp(nm, "sc=", 3)) nm += 3, script_only = 1;
			if (script_only) { /* \p{script=Han}, \p{sc=Hani} */
				const int si = script_index_of(nm);
				if (si < 0) {
					fail(ps, "unknown script");
					return 0;
				}
				add_cls(set, C_SCRIPT, neg);
				set->cls[set->ncls - 1].mask = (uint32_t) si;
				return 1;
			}
			uint32_t mask = gc_mask_of(nm);
			/* JDK 9+ (CharPredicates.forProperty(name, caseIns)): with CASE_INSENSITIVE each of Lu, Ll, Lt means all cased letters */
			if (set->ci && (mask == 1u || mask == 2u || mask == 4u)) mask = 7u;
			const int ucc = (ps->flags & JO_RE_UNICODE_CHARACTER_CLASS) != 0;
			if (mask) add_gc_cls(set, mask, neg);
			else if (!strcmp(nm, "Alphabetic") || (ucc && !strcmp(nm, "Alpha"))) add_cls(set, C_ALPHA, neg);
			else if (!strcmp(nm, "Alpha")) {
				/* ASCII letters: as a nested item so that negation applies to the pair of ranges as a whole */
				add_gc_cls(set, 3u, neg);
				set->cls[set->ncls - 1].kind = C_ASCII; /* mask 3: letters */
			} else if (!strcmp(nm, "White_Space") || !strcmp(nm, "WhiteSpace")) {
				add_cls(set, C_SPACE, neg);
				set->cls[set->ncls - 1].mask = 1; /* always the Unicode property */
			} else if (!strcmp(nm, "Space")) add_cls(set, C_SPACE, neg);
			else if (!strcmp(nm, "Digit")) add_cls(set, C_DIGIT, neg);
			else if (!strcmp(nm, "ASCII")) {
				add_cls(set, C_ASCII, neg);
			} else if (is_prefix && script_index_of(nm) >= 0) { /* \p{IsHan}: binary properties and categories first, then scripts (Pattern.java) */
				add_cls(set, C_SCRIPT, neg);
				set->cls[set->ncls - 1].mask = (uint32_t) script_index_of(nm);
			} else {
				fail(ps, "unsupported \\p{...} property");
				return 0;
			}
		}
		return 1;
	}
	default:
		if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '1' && c <= '9')) {
			fail(ps, "unsupported escape sequence");
			return 0;
		}
		*lit = c; /* escaped punctuation */
		return 1;
	}
}

static node *parse_class(parser *ps, int ci) {
	node *top = new_node(ps, N_SET);
	node *set = top; /* the operand that currently collects items */
	set->ci = ci;
	if (eat(ps, '^')) set->neg = 1;
	int first = 1;
	for (;;) {
		int c = peek(ps);
		if (c < 0) {
			fail(ps, "unterminated character class");
			return top;
		}
		if (c == ']' && !first) {
			ps->i++;
			break;
		}
		first = 0;
		int lit = -1;
		if (c == '[') { /* nested class: a member of the union */
			ps->i++;
			node *in = parse_class(ps, ci);
			if (ps->failed) return top;
			set->nest = (node **) realloc(set->nest, sizeof(node *) * (size_t) (set->nnest + 1));
			set->nest[set->nnest++] = in;
			continue;
		}
		if (c == '&' && ps->i + 1 < ps->n && ps->p[ps->i + 1] == '&') { /* intersection: what follows is the next operand */
			ps->i += 2;
			node *op = new_node(ps, N_SET);
			op->ci = ci;
			set->and_next = op;
			set = op;
			continue;
		}
		ps->i++;
		if (c == '\\') {
			if (!parse_escape(ps, set, &lit)) return top;
			if (lit < 0) continue; /* class item added */
		} else {
			lit = c;
		}
		/* range? */
		if (peek(ps) == '-' && ps->i + 1 < ps->n && ps->p[ps->i + 1] != ']') {
			ps->i++;
			int hi = peek(ps);
			ps->i++;
			if (hi == '\\') {
				int l2 = -1;
				if (!parse_escape(ps, set, &l2)) return top;
				if (l2 < 0) {
					fail(ps, "bad range end in character class");
					return top;
				}
				hi = l2;
			}
			if (hi < lit) {
				fail(ps, "illegal character range");
				return top;
			}
			add_range(set, (uint32_t) lit, (uint32_t) hi);
		} else {
			add_range(set, (uint32_t) lit, (uint32_t) lit);
		}
	}
	return top;
}

static node *parse_atom(parser *ps, int *ci) {
	int c = peek(ps);
	if (c == '(') {
		ps->i++;
		int sub_ci = *ci;
		if (eat(ps, '?')) {
			if (eat(ps, ':')) {
				/* plain non-capturing */
			} else if (peek(ps) == '!' || peek(ps) == '=') {
				int neg = (peek(ps) == '!');
				ps->i++;
				node *lk = new_node(ps, N_LOOK);
				lk->look_neg = neg;
				lk->sub = parse_alt(ps, sub_ci);
				if (!eat(ps, ')')) fail(ps, "missing ) after look-ahead");
				return lk;
			} else if (peek(ps) == '<' && ps->i + 1 < ps->n && (ps->p[ps->i + 1] == '=' || ps->p[ps->i + 1] == '!')) {
				/* look-behind over exactly one character: (?<=[set]) / (?<![set]) */
				const int neg = ps->p[ps->i + 1] == '!';
				ps->i += 2;
				node *lb = new_node(ps, N_LOOKB);
				lb->look_neg = neg;
				node *sub = parse_alt(ps, sub_ci);
				if (!eat(ps, ')')) fail(ps, "missing ) after look-behind");
				while (sub->type == N_CAT && sub->nkids == 1) sub = sub->kids[0];
				if (sub->type != N_SET && sub->type != N_ANY) fail(ps, "look-behind over anything but one character is not supported");
				lb->sub = sub;
				return lb;
			} else if (peek(ps) == '<') { /* named group (?<name>X): the name is irrelevant here (no back-references) */
				ps->i++;
				int k = 0;
				while (peek(ps) >= 0 && peek(ps) != '>') ps->i++, k++;
				if (!k || !eat(ps, '>')) {
					fail(ps, "bad group name");
					return new_node(ps, N_EMPTY);
				}
			} else {
				/* inline flags: (?i) (?-i) (?iu:...) */
				int on = 1, seen = 0;
				while (peek(ps) >= 0 && peek(ps) != ')' && peek(ps) != ':') {
					int f = peek(ps);
					ps->i++;
					seen = 1;
					if (f == '-') on = 0;
					else if (f == 'i') sub_ci = on;
					else if (f == 'u' || f == 'U') { /* UNICODE_CASE / UNICODE_CHARACTER_CLASS inline: pattern-level only */
						if (on) ps->flags |= (f == 'u') ? JO_RE_UNICODE_CASE : (JO_RE_UNICODE_CHARACTER_CLASS | JO_RE_UNICODE_CASE);
					} else {
						fail(ps, "unsupported inline flag");
						return new_node(ps, N_EMPTY);
					}
				}
				if (!seen) {
					fail(ps, "bad group syntax");
					return new_node(ps, N_EMPTY);
				}
				if (eat(ps, ')')) { /* (?i) applies to the remainder of the enclosing group */
					*ci = sub_ci;
					return new_node(ps, N_EMPTY);
				}
				if (!eat(ps, ':')) {
					fail(ps, "bad inline flag group");
					return new_node(ps, N_EMPTY);
				}
			}
		}
		node *g = parse_alt(ps, sub_ci);
		if (!eat(ps, ')')) fail(ps, "missing )");
		return g;
	}
	if (c == '[') {
		ps->i++;
		return parse_class(ps, *ci);
	}
	if (c == '.') {
		ps->i++;
		return new_node(ps, N_ANY);
	}
	if (c == '^') {
		ps->i++;
		return new_node(ps, N_BOL);
	}
	if (c == '$') {
		ps->i++;
		return new_node(ps, N_EOL);
	}
	if (c == '\\') {
		ps->i++;
		int nc = peek(ps);
		if (nc == 'b' || nc == 'B') { /* word boundary */
			ps->i++;
			node *wb = new_node(ps, N_WORDB);
			wb->look_neg = nc == 'B';
			return wb;
		}
		if (nc == 'A') { /* the beginning of the input: '^' without MULTILINE */
			ps->i++;
			return new_node(ps, N_BOL);
		}
		if (nc == 'Z' || nc == 'z') { /* \Z: '$' without MULTILINE; \z: the very end */
			ps->i++;
			node *e = new_node(ps, N_EOL);
			e->min = nc == 'z';
			return e;
		}
		if (nc == 'R') { /* linebreak matcher = \r\n|[\n\x0B\f\r\x85\u2028\u2029] (Pattern javadoc; JDK 9+ matches it as that alternation) */
			ps->i++;
			node *alt = new_node(ps, N_ALT), *crlf = new_node(ps, N_CAT), *cr = new_node(ps, N_SET), *lf = new_node(ps, N_SET), *one = new_node(ps, N_SET);
			add_range(cr, '\r', '\r');
			add_range(lf, '\n', '\n');
			add_kid(crlf, cr);
			add_kid(crlf, lf);
			add_range(one, 0x0A, 0x0D);
			add_range(one, 0x85, 0x85);
			add_range(one, 0x2028, 0x2029);
			add_kid(alt, crlf);
			add_kid(alt, one);
			return alt;
		}
		if (nc == 'G' || nc == 'X' || nc == 'Q' || (nc >= '1' && nc <= '9') || nc == 'k') {
			fail(ps, "unsupported escape (boundary / back-reference / quoting)");
			return new_node(ps, N_EMPTY);
		}
		node *set = new_node(ps, N_SET);
		set->ci = *ci;
		int lit = -1;
		if (!parse_escape(ps, set, &lit)) return set;
		if (lit >= 0) add_range(set, (uint32_t) lit, (uint32_t) lit);
		return set;
	}
	if (c == '*' || c == '+' || c == '?') {
		fail(ps, "dangling quantifier");
		return new_node(ps, N_EMPTY);
	}
	/* literal */
	ps->i++;
	node *set = new_node(ps, N_SET);
	set->ci = *ci;
	add_range(set, (uint32_t) c, (uint32_t) c);
	return set;
}

static node *parse_quantified(parser *ps, int *ci) {
	node *atom = parse_atom(ps, ci);
	for (;;) {
		int c = peek(ps);
		int min, max;
		if (c == '*') {
			min = 0;
			max = -1;
			ps->i++;
		} else if (c == '+') {
			min = 1;
			max = -1;
			ps->i++;
		} else if (c == '?') {
			min = 0;
			max = 1;
			ps->i++;
		} else if (c == '{') {
			int save = ps->i;
			ps->i++;
			int a = 0, k = 0;
			while (peek(ps) >= '0' && peek(ps) <= '9') {
				a = a * 10 + (peek(ps) - '0');
				ps->i++;
				k++;
			}
			if (!k) {
				ps->i = save;
				fail(ps, "bad {n,m} quantifier");
				return atom;
			}
			min = a;
			max = a;
			if (eat(ps, ',')) {
				max = -1;
				int b = 0, kb = 0;
				while (peek(ps) >= '0' && peek(ps) <= '9') {
					b = b * 10 + (peek(ps) - '0');
					ps->i++;
					kb++;
				}
				if (kb) max = b;
			}
			if (!eat(ps, '}') || (max >= 0 && max < min)) {
				fail(ps, "bad {n,m} quantifier");
				return atom;
			}
		} else {
			return atom;
		}
		node *rep = new_node(ps, N_REP);
		rep->sub = atom;
		rep->min = min;
		rep->max = max;
		rep->mode = Q_GREEDY;
		if (eat(ps, '?')) rep->mode = Q_LAZY;
		else if (eat(ps, '+')) rep->mode = Q_POSSESSIVE;
		if (atom->type == N_LOOK || atom->type == N_LOOKB || atom->type == N_EMPTY || atom->type == N_BOL || atom->type == N_EOL || atom->type == N_WORDB) {
			fail(ps, "quantifier on a zero-width construct is not supported");
			return atom;
		}
		atom = rep;
	}
}

static node *parse_cat(parser *ps, int ci) {
	node *cat = new_node(ps, N_CAT);
	while (!ps->failed) {
		int c = peek(ps);
		if (c < 0 || c == '|' || c == ')') break;
		add_kid(cat, parse_quantified(ps, &ci));
	}
	return cat;
}

static node *parse_alt(parser *ps, int ci) {
	node *first = parse_cat(ps, ci);
	if (peek(ps) != '|') return first;
	node *alt = new_node(ps, N_ALT);
	add_kid(alt, first);
	while (!ps->failed && eat(ps, '|')) add_kid(alt, parse_cat(ps, ci));
	return alt;
}

jo_regex *jo_regex_compile(const char *pattern, int flags, char *err, int errlen) {
	size_t blen = strlen(pattern);
	uint32_t *cps = (uint32_t *) malloc(sizeof(uint32_t) * (blen + 1));
	int n = 0;
	for (size_t i = 0; i < blen;) {
		unsigned char b = (unsigned char) pattern[i];
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
		for (int k = 1; k < len && i + (size_t) k < blen; k++) cp = (cp << 6) | ((unsigned char) pattern[i + (size_t) k] & 0x3F);
		cps[n++] = cp;
		i += (size_t) len;
	}
	/* \Q...\E quoting, as Pattern.RemoveQEQuoting does it: the quoted characters become escaped literals (letters and digits as they are) */
	{
		uint32_t *q = (uint32_t *) malloc(sizeof(uint32_t) * (2 * (size_t) n + 1));
		int m2 = 0, quoted = 0;
		for (int i = 0; i < n; i++) {
			if (!quoted && cps[i] == '\\' && i + 1 < n && cps[i + 1] == 'Q') {
				quoted = 1;
				i++;
			} else if (quoted && cps[i] == '\\' && i + 1 < n && cps[i + 1] == 'E') {
				quoted = 0;
				i++;
			} else if (quoted) {
				const uint32_t c = cps[i];
				if (!((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 128)) q[m2++] = '\\';
				q[m2++] = c;
			} else {
				q[m2++] = cps[i];
				if (cps[i] == '\\' && i + 1 < n) q[m2++] = cps[++i]; /* an escaped backslash does not begin a \Q */
			}
		}
		free(cps);
		cps = q;
		n = m2;
	}
	jo_regex *re = (jo_regex *) calloc(1, sizeof(jo_regex));
	if (flags & JO_RE_UNICODE_CHARACTER_CLASS) flags |= JO_RE_UNICODE_CASE; /* Pattern.java: the flag implies UNICODE_CASE */
	parser ps;
	memset(&ps, 0, sizeof(ps));
	ps.p = cps;
	ps.n = n;
	ps.flags = flags;
	ps.re = re;
	ps.err = err;
	ps.errlen = errlen;
	re->root = parse_alt(&ps, (flags & JO_RE_CASE_INSENSITIVE) ? 1 : 0);
	if (!ps.failed && ps.i < ps.n) fail(&ps, "unmatched )");
	re->flags = ps.flags;
	free(cps);
	if (ps.failed) {
		jo_regex_free(re);
		return NULL;
	}
	jo_regex_prepare(re);
	return re;
}

void jo_regex_free(jo_regex *re) {
	if (!re) return;
	for (int i = 0; i < re->nall; i++) {
		free(re->all[i]->ranges);
		free(re->all[i]->cls);
		free(re->all[i]->nest);
		free(re->all[i]->kids);
		free(re->all[i]);
	}
	free(re->all);
	free(re);
}

/* ------------------------------------------------------------------ matching */
typedef struct mctx {
	const jo_regex *re;
	const uint32_t *s;
	int64_t n;
} mctx;

static int cls_match(const cls_item *it, uint32_t cp, int ucc) {
	switch (it->kind) {
	case C_L: return jo_uc_is_letter(cp);
	case C_N: return jo_uc_is_number(cp);
	case C_SPACE:
		if (ucc || it->mask) return jo_uc_is_space(cp);
		return cp == ' ' || (cp >= 0x09 && cp <= 0x0D);
	case C_DIGIT: /* [0-9], or general category Nd under UNICODE_CHARACTER_CLASS */
		if (ucc) return uc_category(cp) == 8;
		return cp >= '0' && cp <= '9';
	case C_WORD:
		if (ucc) return uc_word(cp);
		return (cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z') || (cp >= '0' && cp <= '9') || cp == '_';
	case C_GC: return (int) ((it->mask >> uc_category(cp)) & 1u);
	case C_ALPHA: return uc_alphabetic(cp);
	case C_ASCII:
		if (it->mask == 3u) return (cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z');
		return cp < 128;
	case C_SCRIPT: return uc_script_of(cp) == (int) it->mask;
	case C_HSPACE: /* [ \t\xA0\u1680\u180e\u2000-\u200a\u202f\u205f\u3000] */
		return cp == ' ' || cp == '\t' || cp == 0xA0 || cp == 0x1680 || cp == 0x180E || (cp >= 0x2000 && cp <= 0x200A) || cp == 0x202F || cp == 0x205F || cp == 0x3000;
	case C_VSPACE: /* [\n\x0B\f\r\x85\u2028\u2029] */
		return (cp >= 0x0A && cp <= 0x0D) || cp == 0x85 || cp == 0x2028 || cp == 0x2029;
	}
	return 0;
}

/* the literal characters and ranges of a class: what CASE_INSENSITIVE folds (java.util.regex folds single characters and ranges -
 * SingleI / SingleU / CIRange -, never the predefined classes and properties) */
static int set_match_ranges(const node *nd, uint32_t cp) {
	for (int i = 0; i < nd->nranges; i++)
		if (cp >= nd->ranges[i][0] && cp <= nd->ranges[i][1]) return 1;
	return 0;
}
static int set_match_one(const node *nd, uint32_t cp, int ucc) {
	if (set_match_ranges(nd, cp)) return 1;
	for (int i = 0; i < nd->ncls; i++)
		if (cls_match(&nd->cls[i], cp, ucc) != nd->cls[i].neg) return 1;
	return 0;
}

/* Case-insensitive comparison as java.util.regex does it for single characters: ASCII letters fold
 * onto each other; with UNICODE_CASE, Character.toLowerCase(Character.toUpperCase(c)) additionally
 * folds U+017F (long s) onto s and U+212A (Kelvin sign) onto k. */
static int set_match(const mctx *m, const node *nd, uint32_t cp);
/* one operand of a class: its own items (with case folding) or any nested class */
static int operand_match(const mctx *m, const node *nd, uint32_t cp) {
	int ucc = (m->re->flags & JO_RE_UNICODE_CHARACTER_CLASS) != 0;
	int r = set_match_one(nd, cp, ucc);
	if (!r && nd->ci) {
		if (cp >= 'a' && cp <= 'z') r = set_match_ranges(nd, cp - 32);
		else if (cp >= 'A' && cp <= 'Z') r = set_match_ranges(nd, cp + 32);
		if (!r && (m->re->flags & JO_RE_UNICODE_CASE)) {
			if (cp == 0x17F) r = set_match_ranges(nd, 's') || set_match_ranges(nd, 'S');
			else if (cp == 0x212A) r = set_match_ranges(nd, 'k') || set_match_ranges(nd, 'K');
			else if (cp == 's' || cp == 'S') r = set_match_ranges(nd, 0x17F);
			else if (cp == 'k' || cp == 'K') r = set_match_ranges(nd, 0x212A);
		}
	}
	for (int i = 0; i < nd->nnest && !r; i++) r = set_match(m, nd->nest[i], cp);
	return r;
}
/* the class: every operand of the intersection chain must accept the character; the negation applies to the whole */
static int set_match(const mctx *m, const node *nd, uint32_t cp) {
	int r = 1;
	for (const node *op = nd; op && r; op = op->and_next) r = operand_match(m, op, cp);
	return r != nd->neg;
}

static int is_line_term(uint32_t c) { return c == '\n' || c == '\r' || c == 0x85 || c == 0x2028 || c == 0x2029; }

/* continuation frames */
typedef struct kont {
	const node *nd; /* node to run next (plain frame) or the N_REP node (rep frame) */
	int is_rep;
	int count;      /* rep frame: iterations completed when this frame resumes */
	int64_t start;  /* rep frame: position where the last iteration began (empty-iteration guard) */
	const struct kont *next;
} kont;

static int64_t m_node(const mctx *m, const node *nd, int64_t i, const kont *k);
static int64_t m_rep(const mctx *m, const node *rep, int64_t i, int count, const kont *k);

static int64_t run_k(const mctx *m, const kont *k, int64_t i) {
	if (!k) return i;
	if (k->is_rep) {
		if (i == k->start) return run_k(m, k->next, i); /* an empty iteration ends the loop */
		return m_rep(m, k->nd, i, k->count, k->next);
	}
	return m_node(m, k->nd, i, k->next);
}

static int single_width(const node *nd) { return nd->type == N_SET || nd->type == N_ANY; }

static int sw_match(const mctx *m, const node *nd, int64_t i) {
	if (i >= m->n) return 0;
	const uint32_t c = m->s[i];
	if (nd->type == N_ANY) return !is_line_term(c);
	if (c < 128) return (nd->amap[c >> 5] >> (c & 31)) & 1u;
	return set_match(m, nd, c);
}

/* first-character sets: a pure prefilter, a branch that cannot start at the current character is skipped
 * exactly as it would have failed */
static void compute_first(const mctx *m, node *nd) {
	memset(nd->first, 0, sizeof(nd->first));
	nd->first_other = 0;
	nd->nullable = 0;
	switch (nd->type) {
	case N_SET:
		for (uint32_t c = 0; c < 128; c++)
			if (set_match(m, nd, c)) nd->amap[c >> 5] |= 1u << (c & 31);
		memcpy(nd->first, nd->amap, sizeof(nd->first));
		nd->first_other = 1;
		break;
	case N_ANY:
		memset(nd->first, 0xFF, sizeof(nd->first));
		nd->first_other = 1;
		break;
	case N_EMPTY:
	case N_BOL:
	case N_EOL:
	case N_WORDB:
		nd->nullable = 1;
		break;
	case N_LOOKB:
		compute_first(m, nd->sub); /* (fills the set's ASCII map) */
		nd->nullable = 1;
		break;
	case N_LOOK:
		compute_first(m, nd->sub);
		nd->nullable = 1;
		break;
	case N_REP:
		compute_first(m, nd->sub);
		memcpy(nd->first, nd->sub->first, sizeof(nd->first));
		nd->first_other = nd->sub->first_other;
		nd->nullable = nd->min == 0 || nd->sub->nullable;
		break;
	case N_ALT:
		for (int k = 0; k < nd->nkids; k++) {
			compute_first(m, nd->kids[k]);
			for (int w = 0; w < 4; w++) nd->first[w] |= nd->kids[k]->first[w];
			nd->first_other |= nd->kids[k]->first_other;
			nd->nullable |= nd->kids[k]->nullable;
		}
		break;
	case N_CAT:
		nd->nullable = 1;
		for (int k = 0; k < nd->nkids; k++) {
			compute_first(m, nd->kids[k]);
			if (nd->nullable) {
				for (int w = 0; w < 4; w++) nd->first[w] |= nd->kids[k]->first[w];
				nd->first_other |= nd->kids[k]->first_other;
			}
			if (!nd->kids[k]->nullable) nd->nullable = 0;
		}
		break;
	}
}

static int can_start(const mctx *m, const node *nd, int64_t i) {
	if (nd->nullable) return 1;
	if (i >= m->n) return 0;
	const uint32_t c = m->s[i];
	return c < 128 ? (int) ((nd->first[c >> 5] >> (c & 31)) & 1u) : nd->first_other;
}

static int64_t m_rep(const mctx *m, const node *rep, int64_t i, int count, const kont *k) {
	const node *atom = rep->sub;
	if (single_width(atom)) {
		/* count how many the atom can take, then hand back one at a time (greedy) */
		int64_t lim = rep->max < 0 ? m->n - i : rep->max;
		int64_t c = 0;
		if (rep->mode == Q_LAZY) {
			for (;;) {
				if (c >= rep->min) {
					int64_t r = run_k(m, k, i + c);
					if (r >= 0) return r;
				}
				if (c >= lim || !sw_match(m, atom, i + c)) return -1;
				c++;
			}
		}
		while (c < lim && sw_match(m, atom, i + c)) c++;
		if (c < rep->min) return -1;
		if (rep->mode == Q_POSSESSIVE) return run_k(m, k, i + c);
		for (; c >= rep->min; c--) {
			int64_t r = run_k(m, k, i + c);
			if (r >= 0) return r;
		}
		return -1;
	}
	/* general sub-expression */
	kont again;
	again.nd = rep;
	again.is_rep = 1;
	again.count = count + 1;
	again.start = i;
	again.next = k;
	if (rep->mode == Q_LAZY) {
		if (count >= rep->min) {
			int64_t r = run_k(m, k, i);
			if (r >= 0) return r;
		}
		if (rep->max < 0 || count < rep->max) return m_node(m, atom, i, &again);
		return -1;
	}
	if (rep->max < 0 || count < rep->max) {
		int64_t r = m_node(m, atom, i, &again);
		if (r >= 0) return r;
	}
	if (count >= rep->min) return run_k(m, k, i);
	return -1;
}

static int64_t m_cat(const mctx *m, const node *cat, int idx, int64_t i, const kont *k) {
	if (idx == cat->nkids) return run_k(m, k, i);
	if (idx == cat->nkids - 1) return m_node(m, cat->kids[idx], i, k);
	/* one plain continuation frame per later kid, on the stack (patterns are small) */
	kont frames[64];
	int nk = cat->nkids - idx - 1;
	if (nk > 64) return -1;
	for (int j = nk - 1; j >= 0; j--) {
		frames[j].nd = cat->kids[idx + 1 + j];
		frames[j].is_rep = 0;
		frames[j].count = 0;
		frames[j].start = 0;
		frames[j].next = (j == nk - 1) ? k : &frames[j + 1];
	}
	return m_node(m, cat->kids[idx], i, &frames[0]);
}

/* Bound.isWord(ch) || (getType(ch) == NON_SPACING_MARK && hasBaseCharacter): the character at index i */
static int word_char_at(const mctx *m, int64_t i) {
	const int ucc = (m->re->flags & JO_RE_UNICODE_CHARACTER_CLASS) != 0;
	const uint32_t cp = m->s[i];
	const int g = uc_category(cp);
	if (ucc ? uc_word(cp) : (cp == '_' || g <= 4 || g == 8)) return 1;
	if (g != 5) return 0;
	for (int64_t q = i - 1; q >= 0; q--) { /* hasBaseCharacter: back over non-spacing marks to a letter or digit */
		const int g2 = uc_category(m->s[q]);
		if (g2 <= 4 || g2 == 8) return 1;
		if (g2 != 5) return 0;
	}
	return 0;
}

static int64_t m_node(const mctx *m, const node *nd, int64_t i, const kont *k) {
	switch (nd->type) {
	case N_EMPTY: return run_k(m, k, i);
	case N_SET:
	case N_ANY:
		if (!sw_match(m, nd, i)) return -1;
		return run_k(m, k, i + 1);
	case N_BOL:
		if (i != 0) return -1;
		return run_k(m, k, i);
	case N_EOL: /* Java '$' without MULTILINE: at end, or before a final line terminator; \z (min = 1): at the end only */
		if (i == m->n || (!nd->min && ((i == m->n - 1 && is_line_term(m->s[i])) || (i == m->n - 2 && m->s[i] == '\r' && m->s[i + 1] == '\n')))) return run_k(m, k, i);
		return -1;
	case N_LOOKB: { /* the character before position i, if any, against the one set */
		const int hit = i > 0 && sw_match(m, nd->sub, i - 1);
		if (hit == (nd->look_neg != 0)) return -1;
		return run_k(m, k, i);
	}
	case N_WORDB: { /* java.util.regex.Pattern.Bound.check (JDK 11-18) */
		const int left = i > 0 && word_char_at(m, i - 1), right = i < m->n && word_char_at(m, i);
		if ((left != right) == (nd->look_neg != 0)) return -1;
		return run_k(m, k, i);
	}
	case N_CAT: return m_cat(m, nd, 0, i, k);
	case N_ALT:
		for (int b = 0; b < nd->nkids; b++) {
			if (!can_start(m, nd->kids[b], i)) continue;
			int64_t r = m_node(m, nd->kids[b], i, k);
			if (r >= 0) return r;
		}
		return -1;
	case N_REP: return m_rep(m, nd, i, 0, k);
	case N_LOOK: {
		int64_t r = m_node(m, nd->sub, i, NULL);
		if ((r >= 0) == (nd->look_neg != 0)) return -1;
		return run_k(m, k, i);
	}
	}
	return -1;
}

static void jo_regex_prepare(jo_regex *re) {
	mctx m;
	m.re = re;
	m.s = NULL;
	m.n = 0;
	compute_first(&m, re->root);
}

int jo_regex_search(const jo_regex *re, const uint32_t *cps, int64_t n, int64_t from, int64_t *ms, int64_t *me) {
	mctx m;
	m.re = re;
	m.s = cps;
	m.n = n;
	for (int64_t st = from; st <= n; st++) {
		if (!can_start(&m, re->root, st)) continue;
		int64_t r = m_node(&m, re->root, st, NULL);
		if (r >= 0) {
			*ms = st;
			*me = r;
			return 1;
		}
	}
	return 0;
}
