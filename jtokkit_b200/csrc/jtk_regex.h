/*
 * General split patterns: a java.util.regex subset compiled at registration into a small backtracking program that the
 * GPU runs with one thread per document (jtk_general_split_kernel).  This is what makes
 * EncodingRegistry.registerGptBytePairEncoding (AbstractEncodingRegistry.java:63-66) accept patterns other than the two
 * predefined ones, e.g. Pattern.compile("test") in BaseEncodingRegistryTest.java:110-125.  The predefined patterns never
 * take this path (they compile to class tables + bit-parallel rules, jtk_device.cuh); it is the slow, general one.
 *
 * Semantics restated (the JDK is not part of /root/reference): matching over code points, ordered alternation, greedy /
 * lazy / possessive quantifiers with backtracking, (?i) / (?i:...) with ASCII case folding (+ U+017F / U+212A under
 * UNICODE_CASE), positive / negative look-ahead, ^ $ ., Matcher.find() resumption (after an empty match the search
 * advances by one character), characters matched by no alternative are skipped.
 * Not supported (registration fails with JTK_E_PATTERN_UNSUPPORTED, nothing falls back to the CPU): look-behind,
 * back-references, \b, named groups, class intersection, \d / \w under UNICODE_CHARACTER_CLASS, \p{..} other than L and N,
 * loops over sub-expressions that can match the empty string, counted loops beyond 16.
 */
#ifndef JTK_REGEX_H
#define JTK_REGEX_H

#include <stdint.h>

#include "jtk_common.h"

enum {
	JTK_RX_SET = 1,    /* a = set index: one code point in the set */
	JTK_RX_REP,        /* a = set index, b = min, c = max (-1 unbounded), d = mode: 0 greedy, 1 lazy, 2 possessive */
	JTK_RX_SPLIT,      /* a = preferred pc, b = alternative pc */
	JTK_RX_JMP,        /* a = pc */
	JTK_RX_LOOK,       /* a = 1 negative / 0 positive, b = pc of the sub-program (ends in MATCH); continues at pc + 1 */
	JTK_RX_BOL,
	JTK_RX_EOL,
	JTK_RX_MATCH
};

struct jtk_rx_inst {
	int32_t op, a, b, c, d;
};

/* A set of code points: ASCII bitmap + class flags + ranges for the rest.  Case-insensitive variants are expanded at compile time. */
struct jtk_rx_set {
	uint32_t ascii[4];
	uint32_t flags;      /* bit 0 negated; bits 1-3: contains \p{L}, \p{N}, \s; bits 4-6: contains \P{L}, \P{N}, \S; bit 7: any (.) */
	int32_t range_begin; /* into the ranges array: pairs (lo, hi), sorted, non-ASCII matters only */
	int32_t range_count;
};

#define JTK_RX_NEG 1u
#define JTK_RX_HAS_L 2u
#define JTK_RX_HAS_N 4u
#define JTK_RX_HAS_S 8u
#define JTK_RX_HAS_NOT_L 16u
#define JTK_RX_HAS_NOT_N 32u
#define JTK_RX_HAS_NOT_S 64u
#define JTK_RX_DOT 128u
#define JTK_RX_STACK 1024 /* backtrack frames per thread on the device (global memory, 16 KiB); a document that needs more is flagged JTK_DOC_PATTERN_STACK */

struct jtk_rx_program {
	const jtk_rx_inst *inst;
	int32_t ninst;
	const jtk_rx_set *sets;
	const uint32_t *ranges;
};

JTK_HD jtk_rx_program jtk_rx_program_of(const jtk_tables &T) {
	jtk_rx_program P;
	P.inst = static_cast<const jtk_rx_inst *>(T.rx_inst);
	P.ninst = T.rx_ninst;
	P.sets = static_cast<const jtk_rx_set *>(T.rx_sets);
	P.ranges = T.rx_ranges;
	return P;
}

/* code point at byte position p of s[0..n) (UTF-8, as jtk_decode_char: malformed bytes are one-byte characters) */
JTK_HD uint32_t jtk_rx_decode(const uint8_t *s, int64_t p, int64_t n, int *len) {
	const uint32_t b0 = s[p];
	*len = 1;
	if (b0 < 0x80) return b0;
	if (b0 < 0xC0 || b0 >= 0xF8) return 0xFFFD;
	const int k = b0 < 0xE0 ? 2 : b0 < 0xF0 ? 3 : 4;
	if (p + k > n) return 0xFFFD;
	uint32_t cp = b0 & (0xFFu >> (k + 1));
	for (int i = 1; i < k; i++) {
		const uint32_t b = s[p + i];
		if ((b & 0xC0) != 0x80) return 0xFFFD;
		cp = (cp << 6) | (b & 0x3F);
	}
	*len = k;
	return cp;
}

/* start of the character that ends at byte position p (p > lo) */
JTK_HD int64_t jtk_rx_prev(const uint8_t *s, int64_t p, int64_t lo) {
	int64_t q = p - 1;
	int k = 0;
	while (k < 3 && q > lo && (s[q] & 0xC0) == 0x80) {
		q--;
		k++;
	}
	int len;
	jtk_rx_decode(s, q, p, &len);
	return q + len == p ? q : p - 1;
}

JTK_HD bool jtk_rx_in_set(const jtk_rx_program &P, const jtk_tables &T, int set, uint32_t cp) {
	const jtk_rx_set &S = P.sets[set];
	bool in;
	if (S.flags & JTK_RX_DOT) {
		in = !(cp == '\n' || cp == '\r' || cp == 0x85 || cp == 0x2028 || cp == 0x2029);
	} else if (cp < 128) {
		in = (S.ascii[cp >> 5] >> (cp & 31)) & 1u;
	} else {
		in = false;
		if (S.flags & (JTK_RX_HAS_L | JTK_RX_HAS_N | JTK_RX_HAS_S | JTK_RX_HAS_NOT_L | JTK_RX_HAS_NOT_N | JTK_RX_HAS_NOT_S)) {
			const int c = cp < 0x110000u ? (int) T.cp_stage2[((uint32_t) T.cp_stage1[cp >> 8] << 8) | (cp & 255u)] : (int) JTK_C_O;
			const bool isl = c >= JTK_C_L, isn = c == JTK_C_N, iss = c >= JTK_C_SP && c <= JTK_C_WO;
			in = ((S.flags & JTK_RX_HAS_L) && isl) || ((S.flags & JTK_RX_HAS_N) && isn) || ((S.flags & JTK_RX_HAS_S) && iss) ||
			     ((S.flags & JTK_RX_HAS_NOT_L) && !isl) || ((S.flags & JTK_RX_HAS_NOT_N) && !isn) || ((S.flags & JTK_RX_HAS_NOT_S) && !iss);
		}
		if (!in) {
			int lo = 0, hi = S.range_count - 1;
			while (lo <= hi) {
				const int mid = (lo + hi) >> 1;
				const uint32_t a = P.ranges[2 * (S.range_begin + mid)], b = P.ranges[2 * (S.range_begin + mid) + 1];
				if (cp < a) hi = mid - 1;
				else if (cp > b) lo = mid + 1;
				else {
					in = true;
					break;
				}
			}
		}
	}
	return in != ((S.flags & JTK_RX_NEG) != 0);
}

struct jtk_rx_frame {
	int32_t pc;
	int32_t count; /* REP frames: characters currently taken; -1 for plain alternatives */
	int64_t pos;   /* alternative: position to resume at; REP: position where the run started */
};

/* Runs the program from instruction `pc0` at byte position `start` of the document s[lo..n).  Returns the end of the match
 * or -1; *overflow is set when the backtrack stack st[0..cap) was too small (the caller flags the document).  DEPTH: look-ahead
 * nesting; a look-ahead runs on the unused rest of the same stack. */
template <int DEPTH>
JTK_HD int64_t jtk_rx_run(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t lo, int64_t n, int64_t start, int pc0, jtk_rx_frame *st, int cap,
                          bool *overflow) {
	int sp = 0;
	int pc = pc0;
	int64_t pos = start;
	for (;;) {
		bool fail = false;
		const jtk_rx_inst in = P.inst[pc];
		switch (in.op) {
		case JTK_RX_MATCH: return pos;
		case JTK_RX_SET: {
			if (pos >= n) {
				fail = true;
				break;
			}
			int len;
			const uint32_t cp = jtk_rx_decode(s, pos, n, &len);
			if (!jtk_rx_in_set(P, T, in.a, cp)) fail = true;
			else {
				pos += len;
				pc++;
			}
			break;
		}
		case JTK_RX_REP: {
			int count = 0;
			int64_t q = pos;
			if (in.d == 1) { /* lazy: take the minimum, remember how to take one more */
				while (count < in.b && q < n) {
					int len;
					const uint32_t cp = jtk_rx_decode(s, q, n, &len);
					if (!jtk_rx_in_set(P, T, in.a, cp)) break;
					q += len;
					count++;
				}
				if (count < in.b) {
					fail = true;
					break;
				}
				if (sp >= cap) {
					*overflow = true;
					return -1;
				}
				st[sp].pc = pc;
				st[sp].count = count;
				st[sp].pos = q; /* lazy frames keep the current end */
				sp++;
				pos = q;
				pc++;
				break;
			}
			while ((in.c < 0 || count < in.c) && q < n) {
				int len;
				const uint32_t cp = jtk_rx_decode(s, q, n, &len);
				if (!jtk_rx_in_set(P, T, in.a, cp)) break;
				q += len;
				count++;
			}
			if (count < in.b) {
				fail = true;
				break;
			}
			if (in.d == 0 && count > in.b) { /* greedy: one frame that hands characters back one at a time */
				if (sp >= cap) {
					*overflow = true;
					return -1;
				}
				st[sp].pc = pc;
				st[sp].count = count;
				st[sp].pos = q; /* current end; giving back = stepping one character to the left */
				sp++;
			}
			pos = q;
			pc++;
			break;
		}
		case JTK_RX_SPLIT:
			if (sp >= cap) {
				*overflow = true;
				return -1;
			}
			st[sp].pc = in.b;
			st[sp].count = -1;
			st[sp].pos = pos;
			sp++;
			pc = in.a;
			break;
		case JTK_RX_JMP: pc = in.a; break;
		case JTK_RX_BOL:
			if (pos != lo) fail = true;
			else pc++;
			break;
		case JTK_RX_EOL: { /* Java '$' without MULTILINE: at the end, or before a final line terminator */
			bool ok = pos == n;
			if (!ok && pos < n) {
				int len;
				const uint32_t cp = jtk_rx_decode(s, pos, n, &len);
				const bool term = cp == '\n' || cp == '\r' || cp == 0x85 || cp == 0x2028 || cp == 0x2029;
				if (term && pos + len == n) ok = true;
				if (cp == '\r' && pos + 2 == n && s[pos + 1] == '\n') ok = true;
			}
			if (!ok) fail = true;
			else pc++;
			break;
		}
		case JTK_RX_LOOK: {
			int64_t r = -1;
			if (DEPTH < 2) r = jtk_rx_run<(DEPTH < 2 ? DEPTH + 1 : 2)>(P, T, s, lo, n, pos, in.b, st + sp, cap - sp, overflow); /* look-ahead nests at most twice (checked at compile time of the pattern) */
			else *overflow = true;
			if ((r >= 0) == (in.a != 0)) fail = true;
			else pc++;
			break;
		}
		default: fail = true;
		}
		if (!fail) continue;
		/* backtrack */
		for (;;) {
			if (sp == 0) return -1;
			jtk_rx_frame &f = st[sp - 1];
			if (f.count < 0) { /* plain alternative */
				pc = f.pc;
				pos = f.pos;
				sp--;
				break;
			}
			const jtk_rx_inst rep = P.inst[f.pc];
			if (rep.d == 0) { /* greedy run: hand one character back */
				if (f.count > rep.b) {
					f.count--;
					f.pos = jtk_rx_prev(s, f.pos, lo);
					pos = f.pos;
					pc = f.pc + 1;
					if (f.count == rep.b) sp--;
					break;
				}
				sp--;
			} else { /* lazy run: take one more character */
				bool took = false;
				if ((rep.c < 0 || f.count < rep.c) && f.pos < n) {
					int len;
					const uint32_t cp = jtk_rx_decode(s, f.pos, n, &len);
					if (jtk_rx_in_set(P, T, rep.a, cp)) {
						f.pos += len;
						f.count++;
						pos = f.pos;
						pc = f.pc + 1;
						took = true;
					}
				}
				if (took) break;
				sp--;
			}
		}
	}
}

/* Matcher.find() over one document s[lo..n): calls emit(match_start, match_end) for every match in order. */
template <typename Emit>
JTK_HD void jtk_rx_find_all(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t lo, int64_t n, jtk_rx_frame *st, int cap, Emit emit, bool *overflow) {
	int64_t first = -1, last = lo;
	for (;;) {
		int64_t from = last;
		if (from == first) { /* the previous match was empty: advance by one character */
			if (from >= n) return;
			int len;
			jtk_rx_decode(s, from, n, &len);
			from += len;
		}
		int64_t ms = -1, me = -1;
		for (int64_t stp = from; stp <= n;) {
			const int64_t r = jtk_rx_run<0>(P, T, s, lo, n, stp, 0, st, cap, overflow);
			if (*overflow) return;
			if (r >= 0) {
				ms = stp;
				me = r;
				break;
			}
			if (stp >= n) break;
			int len;
			jtk_rx_decode(s, stp, n, &len);
			stp += len;
		}
		if (ms < 0) return;
		emit(ms, me);
		first = ms;
		last = me;
	}
}

/* One document -> piece bits.  set_start(g) marks a piece start at global byte g, set_skip(g) marks the piece that starts
 * at g as a gap (text no alternative matched: Matcher.find() skips it, it produces no tokens).  Empty matches produce
 * nothing either (bytePairMerge of an empty piece is an empty list, GptBytePairEncoding.java:205-209; a vocabulary with
 * an empty key is rejected at registration when the pattern can match the empty string). */
template <typename SetStart, typename SetSkip>
JTK_HD void jtk_rx_split_document(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t lo, int64_t n, jtk_rx_frame *st, int cap, SetStart set_start,
                                  SetSkip set_skip, bool *overflow) {
	int64_t prev_end = lo;
	jtk_rx_find_all(
	    P, T, s, lo, n, st, cap,
	    [&](int64_t ms, int64_t me) {
		    if (me == ms) return;
		    if (ms > prev_end) {
			    set_start(prev_end);
			    set_skip(prev_end);
		    }
		    set_start(ms);
		    prev_end = me;
	    },
	    overflow);
	if (*overflow) return;
	if (prev_end < n) {
		set_start(prev_end);
		set_skip(prev_end);
	}
}

#endif /* JTK_REGEX_H */
