/*
 * General split patterns: a java.util.regex subset compiled at registration into a small program, determinised into a DFA
 * over code-point classes where the pattern allows it (jtk_dfa.cpp: everything but '$', \\b and multi-character look-ahead) and
 * run as a backtracking program otherwise; the GPU matches per 512-byte slice and stitches per document
 * (jtk_general_slice / stitch / finish kernels, see below).  This is what makes
 * EncodingRegistry.registerGptBytePairEncoding (AbstractEncodingRegistry.java:63-66) accept patterns other than the two
 * predefined ones, e.g. Pattern.compile("test") in BaseEncodingRegistryTest.java:110-125.  The predefined patterns never
 * take this path (they compile to class tables + bit-parallel rules, jtk_device.cuh); it is the slow, general one.
 *
 * Semantics restated (the JDK is not part of /root/reference): matching over code points, ordered alternation, greedy /
 * lazy / possessive quantifiers with backtracking, (?i) / (?i:...) with ASCII case folding (+ U+017F / U+212A under
 * UNICODE_CASE), positive / negative look-ahead, ^ $ ., Matcher.find() resumption (after an empty match the search
 * advances by one character), characters matched by no alternative are skipped.
 * Unicode properties: every general category (\p{Lu}, \p{IsLu}, \p{gc=Lu}, one-letter groups, LC), Alphabetic, White_Space and
 * the POSIX names Alpha / Digit / Space / ASCII; \d \D \w \W and \b \B in their ASCII and UNICODE_CHARACTER_CLASS meanings
 * (tables of Unicode 15.0, unicode_ranges.inc; \b as java.util.regex.Pattern.Bound of JDK 11-18 defines it).
 * Also: named groups (as plain groups), \A \Z \z, \Q..\E, \h \H \v \V, \R, Unicode scripts (\p{IsHan}, \p{script=Han}, \p{sc=Hani}), look-behind
 * over exactly one character ((?<=[set]) / (?<![set])), nested classes and class intersection ([a[b-d]], [a-z&&[^aeiou]]).
 * Not supported (registration fails with JTK_E_PATTERN_UNSUPPORTED, nothing falls back to the CPU): longer look-behind,
 * back-references, atomic groups, block properties, \G \X,
 * loops over sub-expressions that can match the empty string, counted loops over groups beyond 64.
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
	JTK_RX_EOL,        /* a = 0: '$' / \\Z without MULTILINE (the end, or before a final line terminator); a = 1: \\z (the end only) */
	JTK_RX_MATCH,
	JTK_RX_WORDB,      /* a = 1 for \\B; b, c, d = sets: word characters, non-spacing marks, letters-or-digits (java.util.regex.Pattern.Bound) */
	JTK_RX_LOOKB       /* look-behind over one character: a = 1 negative / 0 positive, b = set the character before the position is tested against */
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
#define JTK_RX_STACK 1024 /* backtrack frames per thread of the per-document pass (global memory, 16 KiB); a document that needs more is flagged JTK_DOC_PATTERN_STACK */
#define JTK_RX_STACK_SMALL 128 /* frames per thread of the per-slice pass (eight times as many threads share the same memory); what overflows here is redone by the per-document pass */

struct jtk_rx_program {
	const jtk_rx_inst *inst;
	int32_t ninst;
	const jtk_rx_set *sets;
	const uint32_t *ranges;
	const uint32_t *first; /* 256-bit map of the bytes a match can begin with */
	/* the program as a DFA over code-point classes (jtk_dfa.cpp); null when the pattern has no DFA form: the backtracking program runs */
	const uint16_t *dfa_trans;  /* state * dfa_nsym + symbol -> next state (0 = dead) | 0x8000: a match ends before the character read */
	const uint16_t *dfa_stage1; /* code point >> 8 -> block */
	const uint8_t *dfa_stage2;  /* block * 256 + (code point & 255) -> class */
	const uint8_t *dfa_ascii;   /* the block of U+0000..U+00FF */
	const uint8_t *dfa_stay;    /* run codes per (state, ASCII byte), null: runs are matched character by character (jtk_rx_chain_dfa) */
	int32_t dfa_nsym, dfa_start, dfa_start_bol, dfa_acc_lo;
};

JTK_HD jtk_rx_program jtk_rx_program_of(const jtk_tables &T) {
	jtk_rx_program P;
	P.first = T.rx_first;
	P.inst = static_cast<const jtk_rx_inst *>(T.rx_inst);
	P.ninst = T.rx_ninst;
	P.sets = static_cast<const jtk_rx_set *>(T.rx_sets);
	P.ranges = T.rx_ranges;
	P.dfa_trans = T.rx_dfa_trans;
	P.dfa_stage1 = T.rx_dfa_stage1;
	P.dfa_stage2 = T.rx_dfa_stage2;
	P.dfa_ascii = T.rx_dfa_trans ? T.rx_dfa_stage2 + ((size_t) T.rx_dfa_stage1[0] << 8) : nullptr;
	P.dfa_stay = T.rx_dfa_trans ? T.rx_dfa_stay : nullptr;
	P.dfa_nsym = T.rx_dfa_nsym;
	P.dfa_start = T.rx_dfa_start;
	P.dfa_start_bol = T.rx_dfa_start_bol;
	P.dfa_acc_lo = T.rx_dfa_acc_lo;
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
 * nesting; a look-ahead runs on the unused rest of the same stack.  *hit_end is set when the end of the text n was observed (a
 * speculative run over a truncated view of the document cannot be trusted then). */
template <int DEPTH>
JTK_HD int64_t jtk_rx_run(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t lo, int64_t n, int64_t start, int pc0, jtk_rx_frame *st, int cap,
                          bool *overflow, bool *hit_end) {
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
				*hit_end = true;
				fail = true;
				break;
			}
			if (pos + 4 > n) *hit_end = true;
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
				if (q + 4 > n) *hit_end = true;
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
			if (q + 4 > n) *hit_end = true; /* (within a character of the end: the last character may have been decoded differently) */
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
			if (pos + 4 > n) *hit_end = true;
			bool ok = pos == n;
			if (!ok && pos < n && in.a == 0) {
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
		case JTK_RX_WORDB: {
			/* Pattern.Bound.check: a character is a word character if isWord(ch), or if it is a non-spacing mark that follows a
			 * letter or digit (possibly across other non-spacing marks) */
			if (pos + 4 > n) *hit_end = true;
			auto word_at = [&](int64_t at) { /* character starting at byte position `at` (lo <= at < n) */
				int len;
				const uint32_t cp = jtk_rx_decode(s, at, n, &len);
				if (jtk_rx_in_set(P, T, in.b, cp)) return true;
				if (!jtk_rx_in_set(P, T, in.c, cp)) return false;
				int64_t q = at;
				while (q > lo) { /* hasBaseCharacter: walk back over non-spacing marks */
					q = jtk_rx_prev(s, q, lo);
					const uint32_t c2 = jtk_rx_decode(s, q, n, &len);
					if (jtk_rx_in_set(P, T, in.d, c2)) return true;
					if (!jtk_rx_in_set(P, T, in.c, c2)) return false;
				}
				return false;
			};
			const bool left = pos > lo && word_at(jtk_rx_prev(s, pos, lo));
			const bool right = pos < n && word_at(pos);
			if ((left != right) == (in.a != 0)) fail = true;
			else pc++;
			break;
		}
		case JTK_RX_LOOKB: { /* the character before the position (none at the document start) against one set */
			bool hit = false;
			if (pos > lo) {
				int len;
				const uint32_t cp = jtk_rx_decode(s, jtk_rx_prev(s, pos, lo), n, &len);
				hit = jtk_rx_in_set(P, T, in.b, cp);
			}
			if (hit == (in.a != 0)) fail = true;
			else pc++;
			break;
		}
		case JTK_RX_LOOK: {
			int64_t r = -1;
			if (DEPTH < 2) r = jtk_rx_run<(DEPTH < 2 ? DEPTH + 1 : 2)>(P, T, s, lo, n, pos, in.b, st + sp, cap - sp, overflow, hit_end); /* look-ahead nests at most twice (checked at compile time of the pattern) */
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
				if (f.pos + 4 > n) *hit_end = true;
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

/* The same question answered by the DFA (jtk_dfa.cpp): end of the match that begins at `start`, or -1.  One table lookup per
 * character; the last position at which the most preferred surviving thread matched is the answer (leftmost-first). */
JTK_HD int64_t jtk_rx_dfa_run(const jtk_rx_program &P, const uint8_t *s, int64_t lo, int64_t n, int64_t start, bool *hit_end) {
	uint32_t state = (uint32_t) (start == lo ? P.dfa_start_bol : P.dfa_start);
	const uint32_t nsym = (uint32_t) P.dfa_nsym;
	int64_t pos = start, last = -1;
	for (;;) {
		if (pos >= n) { /* end of the text: the symbol after the last class */
			*hit_end = true;
			if (P.dfa_trans[state * nsym + nsym - 1] & 0x8000u) last = pos;
			return last;
		}
		if (pos + 4 > n) *hit_end = true;
		int len = 1;
		const uint32_t b0 = s[pos];
		uint32_t sym;
		if (b0 < 0x80) {
			sym = P.dfa_ascii[b0];
		} else {
			const uint32_t cp = jtk_rx_decode(s, pos, n, &len);
			sym = P.dfa_stage2[((uint32_t) P.dfa_stage1[cp >> 8] << 8) | (cp & 255u)];
		}
		const uint32_t t = P.dfa_trans[state * nsym + sym];
		if (t & 0x8000u) last = pos;
		state = t & 0x7FFFu;
		if (state == 0) return last;
		pos += len;
		if (state >= (uint32_t) P.dfa_acc_lo) return pos;
	}
}

/* match at `start`: the DFA when the pattern has one, else the backtracking program */
JTK_HD int64_t jtk_rx_match_at(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t lo, int64_t n, int64_t start, jtk_rx_frame *st, int cap,
                               bool *overflow, bool *hit_end) {
	if (P.dfa_trans) return jtk_rx_dfa_run(P, s, lo, n, start, hit_end);
	return jtk_rx_run<0>(P, T, s, lo, n, start, 0, st, cap, overflow, hit_end);
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
			bool hit_end = false;
			const int64_t r = jtk_rx_match_at(P, T, s, lo, n, stp, st, cap, overflow, &hit_end);
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

/* ---------------------------------------------------------------------------------------------
 * Parallel form of Matcher.find() over long documents.
 *
 * What find() does next depends only on the position `from` at which it resumes (the end of the previous match, one
 * character further after an empty match): two runs that reach the same `from` are identical from there on.  So the byte
 * array is cut into slices of JTK_RX_SLICE bytes and every slice is matched SPECULATIVELY from its first byte, recording
 * the `from` positions it passes.  A cheap sequential pass per document then follows the true chain from the document
 * start: inside a slice it re-matches only until it lands on a position the speculative run has passed too, accepts the
 * rest of that slice's result and jumps to the slice's exit.  Runs started inside a piece re-align with the true sequence
 * within a few matches (the next whitespace, typically), so almost all matching happens in the parallel pass.
 *
 * Bits: ms / me = starts / ends of non-empty matches.  A document start counts as a match end.  Piece starts are then
 * ms | me, and a piece that starts at an end without a start is a gap (text no alternative matched: no tokens).
 * Empty matches produce nothing (bytePairMerge of an empty piece is an empty list, GptBytePairEncoding.java:205-209; a
 * vocabulary with an empty key is rejected at registration when the pattern can match the empty string).
 * ------------------------------------------------------------------------------------------- */
#ifndef JTK_RX_SLICE
#define JTK_RX_SLICE 512
#endif
#define JTK_RX_SPEC_LIMIT (8 * JTK_RX_SLICE) /* a speculative run sees this much of the document beyond its slice; it gives up when the matcher reaches the end of that view */
#define JTK_RX_NO_EXIT (-2)                  /* slice record: the speculative run gave up */

struct jtk_rx_split_buffers {
	uint32_t *ms, *me;             /* final bits (after jtk_rx_finish_word: piece starts / gap flags) */
	uint32_t *s_ms, *s_me, *s_from; /* speculative bits per slice */
	int64_t *exit_slice;           /* `from` with which the speculative run left the slice (>= slice end), or JTK_RX_NO_EXIT */
	int64_t *last_ms, *last_me;    /* the match that crosses the slice end (its bits lie outside the slice), or -1 */
	int64_t *join;                 /* first position of the slice from which the speculative bits are valid (none: >= slice end) */
	int64_t *exit_doc;             /* per document: `from` with which the run from the document start left its first slice */
	int64_t nwords, nslices;
};

enum { JTK_RX_FOUND = 0, JTK_RX_NONE = 1, JTK_RX_GAVE_UP = 2 };

/* next match at or after `from`.  `view` <= n is the end of the text as this search may see it: when it is short of n and the
 * matcher gets to see it, the search gives up (with the whole text the outcome might differ). */
JTK_HD int jtk_rx_find_next(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t lo, int64_t n, int64_t from, int64_t view, int64_t *ms, int64_t *me,
                            jtk_rx_frame *st, int cap, bool *overflow) {
	bool hit_end = false;
	for (int64_t stp = from; stp <= view;) {
		/* positions whose first byte cannot begin a match are stepped over character by character without running the program */
		while (stp < view && !((P.first[s[stp] >> 5] >> (s[stp] & 31)) & 1u)) {
			int len;
			jtk_rx_decode(s, stp, view, &len);
			stp += len;
			if (stp + 4 > view) hit_end = true; /* the last character of a truncated view may have been cut */
		}
		if (hit_end && view < n) return JTK_RX_GAVE_UP;
		const int64_t r = jtk_rx_match_at(P, T, s, lo, view, stp, st, cap, overflow, &hit_end);
		if (*overflow) return JTK_RX_NONE;
		if (hit_end && view < n) return JTK_RX_GAVE_UP;
		if (r >= 0) {
			*ms = stp;
			*me = r;
			return JTK_RX_FOUND;
		}
		if (stp >= view) break;
		int len;
		jtk_rx_decode(s, stp, view, &len);
		stp += len;
	}
	return view < n ? JTK_RX_GAVE_UP : JTK_RX_NONE;
}

JTK_HD bool jtk_rx_bit(const uint32_t *b, int64_t g) { return (b[g >> 5] >> (g & 31)) & 1u; }

/* Four bytes of s starting at byte position pos (little endian; bytes at or beyond hi read as anything), through a two-word window
 * over the aligned words of s: sequential matching loads each word once instead of once per byte.  s is 4-byte aligned on the device
 * (the inputs of the kernels are 16-byte aligned); an aligned word that holds a valid byte lies inside the allocation. */
struct jtk_rx_window {
	int64_t idx; /* index of the aligned word in w0; w1 is the next one */
	uint32_t w0, w1;
};
JTK_HD uint32_t jtk_rx_fetch4(const uint8_t *s, int64_t pos, int64_t hi, jtk_rx_window &win) {
#if defined(__CUDA_ARCH__)
	const int64_t idx = pos >> 2;
	if (idx != win.idx) {
		const uint32_t *wp = reinterpret_cast<const uint32_t *>(s) + idx;
		const uint32_t a = idx == win.idx + 1 ? win.w1 : wp[0];
		win.w1 = 4 * (idx + 1) < hi ? wp[1] : 0u;
		win.w0 = a;
		win.idx = idx;
	}
	return __funnelshift_r(win.w0, win.w1, (uint32_t) (pos & 3) * 8u);
#else
	uint32_t x = 0;
	for (int k = 0; k < 4 && pos + k < hi; k++) x |= (uint32_t) s[pos + k] << (8 * k);
	return x;
#endif
}
/* jtk_rx_decode on the four bytes x = s[pos .. pos + 4) with avail = n - pos bytes of text left (>= 1): the same code point and length */
JTK_HD uint32_t jtk_rx_decode_word(uint32_t x, int64_t avail, int *len) {
	const uint32_t b0 = x & 0xFFu;
	*len = 1;
	if (b0 < 0x80) return b0;
	if (b0 < 0xC0 || b0 >= 0xF8) return 0xFFFD;
	const int k = b0 < 0xE0 ? 2 : b0 < 0xF0 ? 3 : 4;
	if (k > avail) return 0xFFFD;
	const int sh = 8 * (4 - k);
	if ((x & ((0xC0C0C000u << sh) >> sh)) != ((0x80808000u << sh) >> sh)) return 0xFFFD; /* bytes 1 .. k-1 must be 10xxxxxx */
	uint32_t cp = b0 & (0xFFu >> (k + 1));
	cp = (cp << 6) | ((x >> 8) & 0x3Fu);
	if (k > 2) cp = (cp << 6) | ((x >> 16) & 0x3Fu);
	if (k > 3) cp = (cp << 6) | ((x >> 24) & 0x3Fu);
	*len = k;
	return cp;
}

/* jtk_rx_chain (below) for a pattern with a DFA, as ONE flat loop: every iteration reads one character and takes one transition,
 * whatever the lane is in the middle of - a match attempt, the step over a position where nothing matches, the look-ahead past the
 * end of a match - so the lanes of a warp stay converged; what happens when an attempt ends (record the match, move `from`, give up)
 * is a short predicated tail.  A position whose first byte cannot begin a match costs one iteration (the start state dies on it), so
 * the first-byte filter of the program search is not needed.  Same results as jtk_rx_chain with jtk_rx_dfa_run, call by call. */
template <typename Or>
JTK_HD int64_t jtk_rx_chain_dfa(const jtk_rx_program &P, const uint8_t *s, int64_t lo, int64_t hi, int64_t from, int64_t end_run, int64_t limit, uint32_t *ms_bits,
                                uint32_t *me_bits, uint32_t *from_bits, const uint32_t *stop_bits, int64_t *cross_ms, int64_t *cross_me, bool *joined, Or bor) {
	*cross_ms = *cross_me = -1;
	*joined = false;
	if (from >= end_run) return from;
	const int64_t view = (limit < 0 || end_run + limit > hi) ? hi : end_run + limit;
	const bool trunc = view < hi;
	const uint32_t nsym = (uint32_t) P.dfa_nsym, acc_lo = (uint32_t) P.dfa_acc_lo;
	if (from_bits) bor(from_bits + (from >> 5), 1u << (from & 31));
	int64_t stp = from, pos = from, last = -1; /* the attempt in progress began at stp, has read up to pos, last match end seen */
	uint32_t state = (uint32_t) (stp == lo ? P.dfa_start_bol : P.dfa_start);
	bool hit_end = false;
	int len0 = 1; /* length of the character at stp (read by the first step of the attempt) */
	jtk_rx_window win;
	win.idx = -2; /* (neither this word nor the one before it) */
	for (;;) {
		/* ---- runs, four ASCII bytes at a time (sequential runs only: the speculative runs of the slice pass switch it off): a state that loops on
		 * all four bytes swallows them; at the start of an attempt, four bytes on each of which the attempt dies at once are stepped
		 * over.  Exactly what four single steps would do, well clear of the end of the view. ---- */
		if (P.dfa_stay && pos + 8 <= view) {
			const uint32_t x4 = jtk_rx_fetch4(s, pos, hi, win);
			if ((x4 & 0x80808080u) == 0) {
				const uint8_t *row = P.dfa_stay + state * 128u;
				const uint32_t c = row[x4 & 0x7Fu];
				if (c && c == row[(x4 >> 8) & 0x7Fu] && c == row[(x4 >> 16) & 0x7Fu] && c == row[x4 >> 24]) {
					if (c != 3) {
						if (pos == stp) len0 = 1;
						if (c == 2) last = pos + 3; /* a match ends before each of the four; the last one counts */
						pos += 4;
						continue;
					}
					if (pos == stp && state == (uint32_t) P.dfa_start && stp != lo) { /* nothing can begin at these four positions */
						stp += 4;
						pos = stp;
						continue;
					}
				}
			}
		}
		/* ---- one character, one transition ---- */
		uint32_t sym = nsym - 1; /* end of the (visible) text */
		int len = 1;
		if (pos >= view) {
			hit_end = true;
		} else {
			if (pos + 4 > view) hit_end = true;
			const uint32_t x = jtk_rx_fetch4(s, pos, hi, win);
			const uint32_t b0 = x & 0xFFu;
			if (b0 < 0x80) {
				sym = P.dfa_ascii[b0];
			} else {
				const uint32_t cp = jtk_rx_decode_word(x, view - pos, &len);
				sym = P.dfa_stage2[((uint32_t) P.dfa_stage1[cp >> 8] << 8) | (cp & 255u)];
			}
			if (pos == stp) len0 = len;
		}
		const uint32_t t = P.dfa_trans[state * nsym + sym];
		if (t & 0x8000u) last = pos;
		const uint32_t ns = t & 0x7FFFu;
		if (sym != nsym - 1 && ns != 0) {
			pos += len;
			state = ns;
			if (ns < acc_lo) continue;
			last = pos; /* nothing but MATCH is left: the attempt ends here */
		}
		/* ---- the attempt at stp is over ---- */
		if (hit_end && trunc) return JTK_RX_NO_EXIT; /* the matcher saw the end of a truncated view: with the whole text the outcome might differ */
		if (last < 0) { /* nothing matches at stp: the search moves on by one character */
			if (stp >= view) return hi;
			stp += len0;
		} else {
			if (last > stp) {
				if (last >= end_run) { /* its end bit lies outside this run's range: the caller places it */
					*cross_ms = stp;
					*cross_me = last;
					return last;
				}
				bor(ms_bits + (stp >> 5), 1u << (stp & 31));
				bor(me_bits + (last >> 5), 1u << (last & 31));
				from = last;
			} else { /* empty match: nothing to emit, the search moves on by one character */
				if (stp >= hi) return hi;
				from = stp + len0;
			}
			/* the next find() begins at `from` */
			if (from >= end_run) return from;
			if (stop_bits && jtk_rx_bit(stop_bits, from)) {
				*joined = true;
				return from;
			}
			if (from_bits) bor(from_bits + (from >> 5), 1u << (from & 31));
			stp = from;
		}
		pos = stp;
		last = -1;
		state = (uint32_t) (stp == lo ? P.dfa_start_bol : P.dfa_start);
	}
}

/* Follows the chain from `from` inside the document [lo, hi) until it leaves [.., end_run).  Match bits go to ms / me through
 * `bor` (matches that end inside the run only; the one that crosses end_run is returned in *cross_ms / *cross_me).  from_bits
 * (nullable): every `from` passed is recorded.  stop_bits (nullable): the run stops as soon as it lands on a recorded `from`
 * of another run and returns it with *joined = true.  Returns the `from` at which it stopped (>= end_run, hi when nothing
 * matches any more) or JTK_RX_NO_EXIT when a search gave up (limit >= 0: bytes of the document beyond end_run that a search may see).  On a
 * backtrack-stack overflow *overflow is set and the `from` whose search overflowed is returned (everything before it is done). */
template <typename Or>
JTK_HD int64_t jtk_rx_chain(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t lo, int64_t hi, int64_t from, int64_t end_run, int64_t limit,
                            uint32_t *ms_bits, uint32_t *me_bits, uint32_t *from_bits, const uint32_t *stop_bits, int64_t *cross_ms, int64_t *cross_me, bool *joined,
                            jtk_rx_frame *st, int cap, bool *overflow, Or bor) {
	if (P.dfa_trans) return jtk_rx_chain_dfa(P, s, lo, hi, from, end_run, limit, ms_bits, me_bits, from_bits, stop_bits, cross_ms, cross_me, joined, bor);
	*cross_ms = *cross_me = -1;
	*joined = false;
	bool first = true;
	while (from < end_run) {
		if (stop_bits && !first && jtk_rx_bit(stop_bits, from)) {
			*joined = true;
			return from;
		}
		first = false;
		if (from_bits) bor(from_bits + (from >> 5), 1u << (from & 31));
		int64_t ms, me;
		const int64_t view = (limit < 0 || end_run + limit > hi) ? hi : end_run + limit;
		const int r = jtk_rx_find_next(P, T, s, lo, hi, from, view, &ms, &me, st, cap, overflow);
		if (*overflow) return from;
		if (r == JTK_RX_GAVE_UP) return JTK_RX_NO_EXIT;
		if (r == JTK_RX_NONE) return hi;
		if (me > ms) {
			if (me >= end_run) { /* its end bit lies outside this run's range: the caller places it */
				*cross_ms = ms;
				*cross_me = me;
				return me;
			}
			bor(ms_bits + (ms >> 5), 1u << (ms & 31));
			bor(me_bits + (me >> 5), 1u << (me & 31));
			from = me;
		} else { /* empty match: nothing to emit, the search moves on by one character */
			if (ms >= hi) return hi;
			int len;
			jtk_rx_decode(s, ms, hi, &len);
			from = ms + len;
		}
	}
	return from;
}

/* Pass 1, one call per slice: the speculative run from the slice start (when the slice starts inside a document) and the true
 * runs of the documents that start inside the slice.  A run that overflows the (small) backtrack stack is left to pass 2. */
template <typename Or>
JTK_HD bool jtk_rx_slice_pass(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t total, const int64_t *doc_off, int64_t ndocs, int64_t slice,
                              const jtk_rx_split_buffers &B, jtk_rx_frame *st, int cap, int64_t *bad_doc, Or bor) {
	const int64_t s0 = slice * JTK_RX_SLICE, s1 = s0 + JTK_RX_SLICE < total ? s0 + JTK_RX_SLICE : total;
	/* d = last document that starts at or before s0 */
	int64_t a = 0, b = ndocs - 1;
	while (a < b) {
		const int64_t mid = (a + b + 1) >> 1;
		if (doc_off[mid] <= s0) a = mid;
		else b = mid - 1;
	}
	int64_t d = a;
	bool overflow = false, joined;
	int64_t cms, cme;
	B.exit_slice[slice] = JTK_RX_NO_EXIT;
	B.last_ms[slice] = B.last_me[slice] = -1;
	B.join[slice] = s1;
	if (ndocs > 0 && doc_off[d] < s0 && doc_off[d + 1] > s0) { /* the slice starts inside document d */
		const int64_t lo = doc_off[d], hi = doc_off[d + 1], end_run = hi < s1 ? hi : s1;
		jtk_rx_program Ps = P;
		Ps.dfa_stay = nullptr; /* (all lanes of a warp run this loop side by side: the run shortcut would only add instructions to it) */
		const int64_t e = jtk_rx_chain(Ps, T, s, lo, hi, s0, end_run, JTK_RX_SPEC_LIMIT, B.s_ms, B.s_me, B.s_from, nullptr, &cms, &cme, &joined, st, cap, &overflow, bor);
		if (overflow) { /* only the true chain may condemn a document: this slice is simply matched again by pass 2 */
			overflow = false;
		} else {
			B.exit_slice[slice] = e;
			B.last_ms[slice] = cms;
			B.last_me[slice] = cme;
		}
	}
	/* documents that start inside the slice: the true chain from their start (a document start counts as a match end) */
	while (d < ndocs && doc_off[d] < s0) d++;
	for (; d < ndocs && doc_off[d] < s1; d++) {
		const int64_t lo = doc_off[d], hi = doc_off[d + 1], end_run = hi < s1 ? hi : s1;
		bor(B.me + (lo >> 5), 1u << (lo & 31));
		const int64_t e = jtk_rx_chain(P, T, s, lo, hi, lo, end_run, -1, B.ms, B.me, nullptr, nullptr, &cms, &cme, &joined, st, cap, &overflow, bor);
		overflow = false; /* (small stack) pass 2 resumes at e with the large one: everything before e is done */
		if (cms >= 0) {
			bor(B.ms + (cms >> 5), 1u << (cms & 31));
			bor(B.me + (cme >> 5), 1u << (cme & 31));
		}
		B.exit_doc[d] = e;
	}
	return true;
}

/* Pass 2, one call per document: follows the true chain through the slices after the document's first one. */
template <typename Or>
JTK_HD bool jtk_rx_stitch_doc(const jtk_rx_program &P, const jtk_tables &T, const uint8_t *s, int64_t total, const int64_t *doc_off, int64_t d,
                              const jtk_rx_split_buffers &B, jtk_rx_frame *st, int cap, Or bor) {
	const int64_t lo = doc_off[d], hi = doc_off[d + 1];
	int64_t cur = B.exit_doc[d];
	bool overflow = false, joined;
	int64_t cms, cme;
	while (cur < hi) {
		const int64_t slice = cur / JTK_RX_SLICE, s0 = slice * JTK_RX_SLICE, s1 = s0 + JTK_RX_SLICE < total ? s0 + JTK_RX_SLICE : total;
		const int64_t end_run = hi < s1 ? hi : s1;
		const bool usable = B.exit_slice[slice] != JTK_RX_NO_EXIT && s0 > lo; /* the slice's speculative run belongs to this document and completed */
		if (!(usable && jtk_rx_bit(B.s_from, cur))) {
			/* match from the true position until the speculative run's trail is hit (or the slice ends) */
			const int64_t e = jtk_rx_chain(P, T, s, lo, hi, cur, end_run, -1, B.ms, B.me, nullptr, usable ? B.s_from : nullptr, &cms, &cme, &joined, st, cap, &overflow, bor);
			if (overflow) return false;
			if (!joined) {
				if (cms >= 0) {
					bor(B.ms + (cms >> 5), 1u << (cms & 31));
					bor(B.me + (cme >> 5), 1u << (cme & 31));
				}
				cur = e;
				continue;
			}
			cur = e;
		}
		/* from `cur` on the slice's speculative result is the true one */
		B.join[slice] = cur;
		if (B.last_ms[slice] >= 0) {
			bor(B.ms + (B.last_ms[slice] >> 5), 1u << (B.last_ms[slice] & 31));
			bor(B.me + (B.last_me[slice] >> 5), 1u << (B.last_me[slice] & 31));
		}
		cur = B.exit_slice[slice];
	}
	return true;
}

/* Pass 3, one call per 32-bit word: accepted speculative bits join the final ones; then ms / me become piece starts / gap flags. */
JTK_HD void jtk_rx_finish_word(const jtk_rx_split_buffers &B, int64_t w, int64_t total) {
	const int64_t slice = (w * 32) / JTK_RX_SLICE;
	uint32_t ms = B.ms[w], me = B.me[w];
	if (slice < B.nslices) {
		/* speculative START bits at word positions >= j are valid, END bits at positions > j: an end bit at the join position itself belongs
		 * to the speculative run's past (the match with which IT arrived there); the true run sets its own - and may have arrived by stepping
		 * over an empty match, without any match ending there */
		const int64_t j = B.join[slice] - w * 32;
		const uint32_t mask = j <= 0 ? 0xFFFFFFFFu : j >= 32 ? 0u : (0xFFFFFFFFu << j);
		const uint32_t mask_end = j < 0 ? 0xFFFFFFFFu : j >= 31 ? 0u : (0xFFFFFFFEu << j);
		ms |= B.s_ms[w] & mask;
		me |= B.s_me[w] & mask_end;
	}
	if ((total >> 5) == w) me |= 1u << (total & 31); /* the end of the input ends the last piece */
	B.ms[w] = ms | me;
	B.me[w] = me & ~ms;
}

#endif /* JTK_REGEX_H */
