/*
 * Per-tile building blocks of the encode kernel, written as __host__ __device__ functions over a
 * jtk_tile_ctx so that the same code runs inside the CUDA kernel (shared memory arrays) and inside the
 * host-side tile emulator that tests/ use to fuzz the split rules against the oracle.
 *
 * What each block replaces in the reference (GptBytePairEncoding.java):
 *   jtk_classify_chunk / jtk_boundary_chunk : the matcher.find() loop, :77-80 (split patterns EncodingFactory.java:63,105)
 *   jtk_special_at                          : text.contains(specialToken), :52-56
 *   jtk_lookup_piece                        : encoder.containsDecodedToken / encode fast path, :81-83
 *   jtk_merge_short                         : bytePairMerge + getRank, :200-300
 */
#ifndef JTK_DEVICE_CUH
#define JTK_DEVICE_CUH

#include "jtk_common.h"

struct jtk_tile_ctx {
	/* region arrays (shared memory on the device); region index r <-> global byte g0 + r */
	uint8_t *sb;     /* JTK_REGION + 32 input bytes (zero padded) */
	uint32_t *bmask; /* piece-start bits, (JTK_REGION + 32) / 32 + 1 words */
	uint32_t *dmask; /* document-start bits, same size; the end of the input counts as a document start */
	uint32_t *planes; /* per 16-byte chunk four words: bit planes 0 | 1 << 16 and 2 | 3 << 16 of the class codes, continuation-byte flags |
	                   * byte high bits << 16 (sixteen positions each), one spare.  This is the ONLY product of the classification
	                   * step: the bit-parallel split rules read the planes directly, the per-position rules through jtk_clsb */
	int32_t *tok;    /* JTK_TILE + JTK_FWD_HALO token staging, indexed by r - JTK_BACK_HALO */
	int32_t *rk;     /* same size: pair ranks during merging, then per-piece token counts */
	/* geometry */
	int64_t g0;      /* global position of region index 0 (negative for the first tile) */
	int64_t total;   /* total input bytes */
	int32_t rs;      /* first usable region index (>= -g0, not in the middle of a character) */
	int32_t carry_n; /* \p{N} characters immediately before region index rs in the same document */
	/* global inputs (for the rare walks that leave the region) */
	const uint8_t *gbytes;
	const int64_t *doc_off;
	int64_t ndocs;
	const jtk_tables *T;
	const uint32_t *lut_sp; /* T->lut_sp or a shared-memory copy of it */
	const uint8_t *cls2;    /* T->cls2 or a shared-memory copy of it */
};

#define JTK_MASK_WORDS ((JTK_REGION + 32) / 32 + 1)

JTK_HD bool jtk_docstart(const jtk_tile_ctx &c, int r) { return (c.dmask[r >> 5] >> (r & 31)) & 1u; }
/* class code | JTK_CONT of region index r, read back from the chunk's bit planes */
JTK_HD int jtk_clsb(const jtk_tile_ctx &c, int r) {
	const uint32_t *o = c.planes + 4 * (r >> 4);
	const int b = r & 15;
	const uint32_t p01 = o[0] >> b, p23 = o[1] >> b, pc = o[2] >> b;
	return (int) ((p01 & 1u) | ((p01 >> 15) & 2u) | ((p23 & 1u) << 2) | ((p23 >> 13) & 8u) | ((pc & 1u) << 7));
}
JTK_HD int jtk_cls(const jtk_tile_ctx &c, int r) { return jtk_clsb(c, r) & JTK_CLS_MASK; }

#if defined(__CUDA_ARCH__)
#define JTK_SMEM_OR(p, v) atomicOr((p), (v))
JTK_HD int jtk_ctz(uint32_t m) { return __ffs((int) m) - 1; }
JTK_HD int jtk_clz(uint32_t m) { return __clz((int) m); }
#else
#define JTK_SMEM_OR(p, v) (*(p) |= (v))
JTK_HD int jtk_ctz(uint32_t m) { return __builtin_ctz(m); }
JTK_HD int jtk_clz(uint32_t m) { return __builtin_clz(m); }
#endif

/* ---------------------------------------------------------------------------------------------
 * region set-up
 * ------------------------------------------------------------------------------------------- */
/* Copies the 16 input bytes of region chunk `chunk` (the pad chunk JTK_REGION_CHUNKS included) into c.sb;
 * positions outside [0, total) read as zero.  c.gbytes must be 16-byte aligned. */
JTK_HD void jtk_load_chunk(jtk_tile_ctx &c, int chunk) {
	const int64_t g = c.g0 + 16 * (int64_t) chunk;
	uint8_t *dst = c.sb + 16 * chunk;
	if (g >= 0 && g + 16 <= c.total) {
#if defined(__CUDA_ARCH__)
		*reinterpret_cast<uint4 *>(dst) = __ldg(reinterpret_cast<const uint4 *>(c.gbytes + g));
#else
		for (int i = 0; i < 16; i++) dst[i] = c.gbytes[g + i];
#endif
	} else {
		for (int i = 0; i < 16; i++) dst[i] = (g + i >= 0 && g + i < c.total) ? c.gbytes[g + i] : (uint8_t) 0;
	}
}

/* Sets the document-start bits of documents first_doc + tid, first_doc + tid + nthreads, ... that fall into
 * the region.  doc_off[ndocs] == total is marked too: the end of the input ends the last piece. */
JTK_HD void jtk_mark_docstarts(jtk_tile_ctx &c, int64_t first_doc, int tid, int nthreads) {
	for (int64_t d = first_doc + tid; d <= c.ndocs; d += nthreads) {
		const int64_t r = c.doc_off[d] - c.g0;
		if (r > JTK_REGION + 16) break;
		if (r >= 0) JTK_SMEM_OR(&c.dmask[r >> 5], 1u << (r & 31));
	}
}

/* First usable region index: not before the input, not inside a character that began before the region. */
JTK_HD int jtk_region_first(const jtk_tile_ctx &c) {
	int rs = c.g0 < 0 ? (int) -c.g0 : 0;
	const int lim = rs + 3;
	while (rs < lim && (c.sb[rs] & 0xC0) == 0x80 && !jtk_docstart(c, rs)) rs++;
	return rs;
}

/* \\p{N} carry into the region (see jtk_count_n_before); call after classification. */
JTK_HD int jtk_global_count_n_before(const jtk_tile_ctx &c, int64_t g);
JTK_HD int jtk_region_carry_n(const jtk_tile_ctx &c) {
	if (c.T->pattern_kind != JTK_PAT_CL100K) return 0;
	if (jtk_docstart(c, c.rs) || c.g0 + c.rs <= 0) return 0;
	if (jtk_cls(c, c.rs) != JTK_C_N) return 0;
	return jtk_global_count_n_before(c, c.g0 + c.rs);
}

/* ---------------------------------------------------------------------------------------------
 * code point classification
 * ------------------------------------------------------------------------------------------- */
JTK_HD int jtk_cp_class(const jtk_tables &T, uint32_t cp) {
	if (cp < 128) return T.ascii_cls[cp];
	if (cp >= 0x110000u) return JTK_C_O;
	return T.cp_stage2[((uint32_t) T.cp_stage1[cp >> 8] << 8) | (cp & 255u)];
}

/* Length a UTF-8 lead byte announces (1 for ASCII, stray continuation bytes and invalid leads). */
JTK_HD int jtk_lead_len(uint8_t b) {
	if (b < 0xC0) return 1;
	if (b < 0xE0) return 2;
	if (b < 0xF0) return 3;
	if (b < 0xF8) return 4;
	return 1;
}

/* Character whose lead byte is at `p` in the byte array `s` (positions < limit are readable, document starts
 * are given by `is_start(p)`).  Returns its class; *len receives its byte length.  A lead byte whose
 * continuation bytes are missing is a one-byte "other" character, as is a stray continuation byte
 * (input from String.getBytes(UTF_8) never contains either). */
template <typename StartFn>
JTK_HD int jtk_decode_char(const jtk_tables &T, const uint8_t *s, int64_t p, int64_t limit, StartFn is_start, int *len) {
	uint8_t b = s[p];
	*len = 1;
	if (b < 0x80) return T.ascii_cls[b];
	int n = jtk_lead_len(b);
	if (n == 1) return JTK_C_O;
	uint32_t cp = b & (0xFFu >> (n + 1));
	for (int k = 1; k < n; k++) {
		if (p + k >= limit || is_start(p + k)) return JTK_C_O;
		uint8_t cb = s[p + k];
		if ((cb & 0xC0) != 0x80) return JTK_C_O;
		cp = (cp << 6) | (cb & 0x3Fu);
	}
	*len = n;
	return jtk_cp_class(T, cp);
}

/* read-only global load (non-coherent path on the device: lets the compiler batch independent loads) */
#if defined(__CUDA_ARCH__)
#define JTK_LDG(ptr) __ldg(ptr)
#else
#define JTK_LDG(ptr) (*(ptr))
#endif

/* bit b of each of the four bytes of w -> 4 mask bits */
JTK_HD uint32_t jtk_plane4(uint32_t w, int b) { return (((w >> b) & 0x01010101u) * 0x01020408u) >> 24; }

/* result byte i = byte number (nibble i of sel) of the eight bytes {x, y} (the PRMT instruction on the device) */
JTK_HD uint32_t jtk_byte_perm(uint32_t x, uint32_t y, uint32_t sel) {
#if defined(__CUDA_ARCH__)
	return __byte_perm(x, y, sel);
#else
	const uint64_t v = x | ((uint64_t) y << 32);
	uint32_t r = 0;
	for (int i = 0; i < 4; i++) r |= (uint32_t) ((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
	return r;
#endif
}

/* the 32-bit word that starts at byte address p (p + 3 readable, any alignment): two aligned loads and a funnel shift */
JTK_HD uint32_t jtk_load_u32(const uint8_t *p) {
	const uint32_t *aw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t) 3);
	const int sh = (int) (reinterpret_cast<uintptr_t>(p) & 3) * 8;
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(aw[0], aw[1], sh);
#else
	return sh ? (aw[0] >> sh) | (aw[1] << (32 - sh)) : aw[0];
#endif
}

/* Classification of the 16 bytes of chunk `chunk` (region indices 16 * chunk ..) into the chunk's bit planes.
 * Pass 1, every chunk: per byte one read of the 256-entry table lut_sp (the four class bits of an ASCII byte spread over the four
 * bytes of a word, zero for bytes >= 0x80) accumulated with a shift by the byte's position - two accumulators of eight
 * positions, two byte permutes turn them into the plane words.  Pass 2, chunks with non-ASCII bytes only: one loop iteration
 * per multi-byte CHARACTER whose lead byte lies in the chunk or in the three bytes before it (a character may straddle the
 * chunk boundary; the previous chunk's thread clips it, this one adds the rest): UTF-8 decoding from one unaligned word,
 * class from a flat table (two-byte characters: 2 KiB table in shared memory; three-byte: 32 KiB nibble table; four-byte:
 * the two-level table), planes updated with one multiply-add per plane word.  A lead byte whose continuation bytes are
 * missing (or cut off by a document start) is a one-byte "other" character, as is a stray continuation byte (input from
 * String.getBytes(UTF_8) never contains either). */
JTK_HD void jtk_classify_chunk(jtk_tile_ctx &c, int chunk) {
	const int r0 = chunk * 16;
	const uint32_t *w = reinterpret_cast<const uint32_t *>(c.sb + r0);
	const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
	const uint32_t *lut = c.lut_sp;
#define JTK_SP(x, k) lut[((x) >> (8 * (k))) & 0xFFu]
	const uint32_t lo = JTK_SP(w0, 0) + (JTK_SP(w0, 1) << 1) + (JTK_SP(w0, 2) << 2) + (JTK_SP(w0, 3) << 3) + (JTK_SP(w1, 0) << 4) + (JTK_SP(w1, 1) << 5) +
	                    (JTK_SP(w1, 2) << 6) + (JTK_SP(w1, 3) << 7);
	const uint32_t hi = JTK_SP(w2, 0) + (JTK_SP(w2, 1) << 1) + (JTK_SP(w2, 2) << 2) + (JTK_SP(w2, 3) << 3) + (JTK_SP(w3, 0) << 4) + (JTK_SP(w3, 1) << 5) +
	                    (JTK_SP(w3, 2) << 6) + (JTK_SP(w3, 3) << 7);
#undef JTK_SP
	uint32_t p01 = jtk_byte_perm(lo, hi, 0x5140u), p23 = jtk_byte_perm(lo, hi, 0x7362u), pch = 0;
	if (((w0 | w1 | w2 | w3) & 0x80808080u) != 0) {
		const jtk_tables &T = *c.T;
		const uint32_t HB = jtk_plane4(w0, 7) | (jtk_plane4(w1, 7) << 4) | (jtk_plane4(w2, 7) << 8) | (jtk_plane4(w3, 7) << 12);
		/* lead bytes (>= 0xC0) of the window [r0 - 3, r0 + 16): bit j <-> region index r0 - 3 + j */
		uint32_t L = (jtk_plane4(w0 & (w0 << 1), 7) | (jtk_plane4(w1 & (w1 << 1), 7) << 4) | (jtk_plane4(w2 & (w2 << 1), 7) << 8) | (jtk_plane4(w3 & (w3 << 1), 7) << 12)) << 3;
		/* document-start bits of region indices r0 - 3 .. r0 + 28 (bit 3 = r0) */
		uint32_t dwin;
		{
			const int lo3 = r0 - 3;
			if (lo3 < 0) {
				dwin = c.dmask[0] << 3;
			} else {
				const int wi = lo3 >> 5, sh = lo3 & 31;
				dwin = c.dmask[wi] >> sh;
				if (sh) dwin |= c.dmask[wi + 1] << (32 - sh);
			}
		}
		if (chunk > 0 && (w0 & 0xC0u) == 0x80u && !((dwin >> 3) & 1u)) {
			/* the chunk begins inside a character: its lead byte is the closest lead byte among the three bytes before the chunk */
			const uint32_t wp = w[-1];
			const uint32_t lp = jtk_plane4(wp & (wp << 1), 7) >> 1;
			if (lp) L |= 1u << (31 - jtk_clz(lp));
		}
		uint32_t cont = 0;
		for (uint32_t m = L; m; m &= m - 1) {
			const int j = jtk_ctz(m);
			const uint32_t x = jtk_load_u32(c.sb + r0 - 3 + j);
			const uint32_t b0 = x & 0xFFu;
			if (b0 >= 0xF8u) continue; /* invalid lead: stays a one-byte "other" character */
			const int n = b0 < 0xE0u ? 2 : b0 < 0xF0u ? 3 : 4;
			if (j + n <= 3) continue;  /* a character of the previous chunk that ends before this one */
			const uint32_t cm = (0x00C0C0C0u >> (8 * (4 - n))) << 8;
			if ((x & cm) != (cm & 0x80808080u) || ((dwin >> (j + 1)) & ((1u << (n - 1)) - 1u)) != 0) continue;
			uint32_t k;
			if (n == 2) {
				k = c.cls2[((x & 0x1Fu) << 6) | ((x >> 8) & 0x3Fu)];
			} else if (n == 3) {
				const uint32_t cp = ((x & 0x0Fu) << 12) | ((x >> 2) & 0xFC0u) | ((x >> 16) & 0x3Fu);
				k = (JTK_LDG(T.bmp_nib + (cp >> 1)) >> ((cp & 1u) * 4)) & 0xFu;
			} else {
				k = (uint32_t) jtk_cp_class(T, ((x & 0x07u) << 18) | ((x << 4) & 0x3F000u) | ((x >> 10) & 0xFC0u) | ((x >> 24) & 0x3Fu));
			}
			const uint32_t mm = ((((1u << n) - 1u) << j) >> 3) & 0xFFFFu; /* the character's bytes inside this chunk */
			p01 += ((k & 1u) | ((k & 2u) << 15)) * mm;
			p23 += (((k >> 2) & 1u) | ((k & 8u) << 13)) * mm;
			cont |= mm & ~((1u << j) >> 3);
		}
		pch = cont | (HB << 16);
	}
	uint32_t *o = c.planes + 4 * chunk;
#if defined(__CUDA_ARCH__)
	*reinterpret_cast<uint4 *>(o) = make_uint4(p01, p23, pch, 0u);
#else
	o[0] = p01;
	o[1] = p23;
	o[2] = pch;
	o[3] = 0;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * walks that may leave the region (rare): plain global-memory versions
 * ------------------------------------------------------------------------------------------- */
/* largest document start <= g */
JTK_HD int64_t jtk_doc_floor(const jtk_tile_ctx &c, int64_t g) {
	int64_t lo = 0, hi = c.ndocs; /* doc_off[lo] <= g */
	while (lo < hi) {
		int64_t mid = (lo + hi + 1) >> 1;
		if (c.doc_off[mid] <= g) lo = mid;
		else hi = mid - 1;
	}
	return c.doc_off[lo];
}
/* smallest document start > g (the end of the input is one) */
JTK_HD int64_t jtk_doc_ceil(const jtk_tile_ctx &c, int64_t g) {
	int64_t lo = 0, hi = c.ndocs; /* answer index in [lo, hi] */
	while (lo < hi) {
		int64_t mid = (lo + hi) >> 1;
		if (c.doc_off[mid] > g) hi = mid;
		else lo = mid + 1;
	}
	return c.doc_off[lo] > g ? c.doc_off[lo] : c.total;
}

struct jtk_never_start {
	JTK_HD bool operator()(int64_t) const { return false; }
};

/* Class of the character that ends at global position g (exclusive) inside the document starting at doc_lo;
 * *lead receives its first byte.  g > doc_lo required. */
JTK_HD int jtk_global_char_before(const jtk_tile_ctx &c, int64_t g, int64_t doc_lo, int64_t *lead) {
	int64_t p = g - 1;
	int k = 0;
	while (k < 3 && p > doc_lo && (c.gbytes[p] & 0xC0) == 0x80) {
		p--;
		k++;
	}
	int len;
	int cl = jtk_decode_char(*c.T, c.gbytes, p, g, jtk_never_start(), &len);
	if (p + len != g) { /* malformed tail: the last byte stands alone */
		*lead = g - 1;
		return jtk_decode_char(*c.T, c.gbytes, g - 1, g, jtk_never_start(), &len);
	}
	*lead = p;
	return cl;
}

/* \p{N} characters immediately before global position g in its document (mod 3 is all that matters). */
JTK_HD int jtk_global_count_n_before(const jtk_tile_ctx &c, int64_t g) {
	int64_t lo = jtk_doc_floor(c, g);
	int k = 0;
	while (g > lo) {
		if ((g & 7) == 0 && g - 8 >= lo) {
			/* eight ASCII digits at once (a run of a million digits would otherwise be walked byte by byte from every tile) */
			const uint64_t t = *reinterpret_cast<const uint64_t *>(c.gbytes + g - 8) ^ 0x3030303030303030ull;
			if ((((t + 0x7676767676767676ull) | t) & 0x8080808080808080ull) == 0) {
				k = (k + 8) % 3;
				g -= 8;
				continue;
			}
		}
		int64_t lead;
		if (jtk_global_char_before(c, g, lo, &lead) != JTK_C_N) break;
		k = (k + 1) % 3;
		g = lead;
	}
	return k;
}

/* ---------------------------------------------------------------------------------------------
 * split rules
 * ------------------------------------------------------------------------------------------- */
/* lead index of the character that ends right before region index r (r > c.rs, no document start at r) */
JTK_HD int jtk_prev_lead(const jtk_tile_ctx &c, int r) {
	int p = r - 1;
	while (p > c.rs && (jtk_clsb(c, p) & JTK_CONT)) p--;
	return p;
}
JTK_HD int jtk_char_len(const jtk_tile_ctx &c, int r) {
	int len = 1;
	while (len < 4 && (jtk_clsb(c, r + len) & JTK_CONT)) len++;
	return len;
}

/* An "other" character (incl. the apostrophe) at lead index r starts a piece unless it continues an
 * other-run or is taken by the optional leading space of ` ?[^\s\p{L}\p{N}]+`. */
JTK_HD bool jtk_other_is_start(const jtk_tile_ctx &c, int r) {
	if (jtk_docstart(c, r)) return true;
	int pc = jtk_cls(c, r - 1);
	return !(jtk_is_other(pc) || pc == JTK_C_SP);
}

/* cur and prev are letters: did one of 's 't 'm 'd 're 've 'll end exactly before r? */
JTK_HD bool jtk_ends_contraction(const jtk_tile_ctx &c, int r) {
	const int k1 = jtk_cls(c, r - 1); /* continuation bytes carry their character's class */
	const bool single = (k1 >= JTK_C_LS && k1 <= JTK_C_LD);
	const bool second = (k1 == JTK_C_LE || k1 == JTK_C_LL);
	if (!single && !second) return false;
	const int r1 = jtk_prev_lead(c, r);
	if (jtk_docstart(c, r1)) return false;
	int r2 = jtk_prev_lead(c, r1);
	int k2 = jtk_cls(c, r2);
	if (single) return k2 == JTK_C_AP && jtk_other_is_start(c, r2);
	if (!((k1 == JTK_C_LE && (k2 == JTK_C_LR || k2 == JTK_C_LV)) || (k1 == JTK_C_LL && k2 == JTK_C_LL))) return false;
	if (jtk_docstart(c, r2)) return false;
	int r3 = jtk_prev_lead(c, r2);
	return jtk_cls(c, r3) == JTK_C_AP && jtk_other_is_start(c, r3);
}

/* x50k: does a contraction alternative match at the apostrophe at lead index r? */
JTK_HD bool jtk_contraction_starts(const jtk_tile_ctx &c, int r) {
	if (jtk_docstart(c, r + 1)) return false;
	int k1 = jtk_cls(c, r + 1);
	if (k1 >= JTK_C_LS && k1 <= JTK_C_LD) return true;
	if (k1 != JTK_C_LR && k1 != JTK_C_LV && k1 != JTK_C_LL) return false;
	int n1 = jtk_char_len(c, r + 1);
	if (jtk_docstart(c, r + 1 + n1)) return false;
	int k2 = jtk_cls(c, r + 1 + n1);
	return (k1 == JTK_C_LL) ? (k2 == JTK_C_LL) : (k2 == JTK_C_LE);
}

/* cl100k: scanning forward from the character at lead index r (a non-NL whitespace character), does a
 * \r or \n occur before the whitespace run ends?  (`\s*[\r\n]+` takes the run up to its last NL.) */
JTK_HD bool jtk_nl_ahead(const jtk_tile_ctx &c, int r) {
	int p = r + jtk_char_len(c, r);
	for (;;) {
		if (p >= JTK_REGION) break;
		if (jtk_docstart(c, p)) return false;
		int k = jtk_cls(c, p);
		if (k == JTK_C_NL) return true;
		if (k != JTK_C_SP && k != JTK_C_WO) return false;
		p += jtk_char_len(c, p);
	}
	/* left the region: continue in global memory */
	int64_t g = c.g0 + p;
	int64_t hi = jtk_doc_ceil(c, g - 1);
	while (g < hi) {
		int len;
		int k = jtk_decode_char(*c.T, c.gbytes, g, hi, jtk_never_start(), &len);
		if (k == JTK_C_NL) return true;
		if (k != JTK_C_SP && k != JTK_C_WO) return false;
		g += len;
	}
	return false;
}

/* cl100k: walking back from lead index r over \r\n characters, is the first other character an "other"
 * one?  (Then those NLs were taken by the `[\r\n]*` tail of ` ?[^\s\p{L}\p{N}]+[\r\n]*`.) */
JTK_HD bool jtk_nl_absorbed(const jtk_tile_ctx &c, int r) {
	int p = r;
	for (;;) {
		if (jtk_docstart(c, p)) return false;
		if (p <= c.rs) break;
		p = jtk_prev_lead(c, p);
		int k = jtk_cls(c, p);
		if (k != JTK_C_NL) return jtk_is_other(k);
	}
	int64_t g = c.g0 + p;
	int64_t lo = jtk_doc_floor(c, g);
	while (g > lo) {
		int64_t lead;
		int k = jtk_global_char_before(c, g, lo, &lead);
		if (k != JTK_C_NL) return jtk_is_other(k);
		g = lead;
	}
	return false;
}

/* \p{N} characters immediately before lead index r in the same document, mod 3 */
JTK_HD int jtk_count_n_before(const jtk_tile_ctx &c, int r) {
	int k = 0, p = r;
	for (;;) {
		if (jtk_docstart(c, p)) return k;
		if (p <= c.rs) return (k + c.carry_n) % 3;
		p = jtk_prev_lead(c, p);
		if (jtk_cls(c, p) != JTK_C_N) return k;
		k = (k + 1) % 3;
	}
}

/* Is region index r (a lead byte that is not a document start) the first byte of a piece?
 * nrun: running count (mod 3) of \p{N} characters right before r, or -1 when not yet known. */
JTK_HD bool jtk_is_piece_start(const jtk_tile_ctx &c, int r, int cur, int *nrun) {
	const int prev = jtk_cls(c, r - 1);
	const bool cl100k = c.T->pattern_kind == JTK_PAT_CL100K;
	if (jtk_is_letter(cur)) {
		if (jtk_is_letter(prev)) return jtk_ends_contraction(c, r);
		if (cl100k) {
			if (prev == JTK_C_N || prev == JTK_C_NL) return true;
			if (jtk_is_space(prev)) return false; /* [^\r\n\p{L}\p{N}]?\p{L}+ takes the whitespace character */
			return !jtk_other_is_start(c, jtk_prev_lead(c, r));
		}
		if (prev == JTK_C_SP) return false; /* ` ?\p{L}+` */
		if (prev == JTK_C_AP) {
			int rp = r - 1;
			return !(jtk_other_is_start(c, rp) && jtk_contraction_starts(c, rp));
		}
		return true;
	}
	if (cur == JTK_C_N) {
		if (cl100k) { /* \p{N}{1,3}: runs are cut in threes from their start */
			if (prev != JTK_C_N) return true;
			if (*nrun < 0) *nrun = jtk_count_n_before(c, r);
			return *nrun == 0;
		}
		return !(prev == JTK_C_N || prev == JTK_C_SP); /* ` ?\p{N}+` */
	}
	if (jtk_is_other(cur)) return !(jtk_is_other(prev) || prev == JTK_C_SP);
	if (cur == JTK_C_NL) return jtk_is_letter(prev) || prev == JTK_C_N; /* cl100k only: after "other" it belongs to [\r\n]*, inside whitespace to \s*[\r\n]+ */
	/* SP / WO */
	if (!jtk_is_space(prev)) return true;
	int nr = r + jtk_char_len(c, r);
	if (!jtk_docstart(c, nr) && !jtk_is_space(jtk_cls(c, nr))) return true; /* \s+(?!\S) gives the last whitespace character back */
	if (prev == JTK_C_NL) {
		if (!jtk_nl_ahead(c, r)) return true; /* \s*[\r\n]+ ended right before r */
		return jtk_nl_absorbed(c, r);         /* ... or the NLs before r belong to the preceding "other" piece */
	}
	return false;
}

/* ---------------------------------------------------------------------------------------------
 * bit-parallel split rules
 * The window is the 32 bytes [r0 - 8, r0 + 24) around chunk r0 .. r0 + 15; bit i of every mask stands for window byte
 * i.  Continuation bytes carry their character's class, so "class of the previous character" is a left shift by one
 * even in multi-byte text; piece starts are only reported on lead bytes.  Preconditions (else the per-position rules
 * below run): the window lies inside the usable region and the input, contains no document start, and its whitespace,
 * digits and contraction letters are all ASCII (U+3000, Arabic-Indic digits, U+017F ... take the slow path).
 * The formulas are the rules of jtk_is_piece_start written as boolean algebra; the emulator test
 * (tests/test_emu_cpu.py) fuzzes both paths against the oracle.
 * ------------------------------------------------------------------------------------------- */

/* spread `seed` forward (towards higher positions) through runs of `run`: result has every position reachable from a
 * seed bit by stepping +1 while staying inside run */
JTK_HD uint32_t jtk_spread_up(uint32_t seed, uint32_t run) {
	uint32_t f = seed | ((seed << 1) & run);
	uint32_t m = run & (run << 1);
	f |= (f << 2) & m;
	m &= m << 2;
	f |= (f << 4) & m;
	m &= m << 4;
	f |= (f << 8) & m;
	m &= m << 8;
	f |= (f << 16) & m;
	return f;
}
/* same towards lower positions */
JTK_HD uint32_t jtk_spread_down(uint32_t seed, uint32_t run) {
	uint32_t f = seed | ((seed >> 1) & run);
	uint32_t m = run & (run >> 1);
	f |= (f >> 2) & m;
	m &= m >> 2;
	f |= (f >> 4) & m;
	m &= m >> 4;
	f |= (f >> 8) & m;
	m &= m >> 8;
	f |= (f >> 16) & m;
	return f;
}

struct jtk_planes {
	uint32_t P0, P1, P2, P3, CONT, HB; /* bit planes of the class codes, continuation-byte flags and high bits of the 32 window bytes */
};

/* The split rules as boolean algebra over the window positions in `valid` (a contiguous range of bits): positions outside
 * `valid` do not exist, exactly as if the text began at the lowest valid position and ended after the highest one.
 * starts_doc: the lowest valid position is a document start (otherwise the text continues before the window);
 * top_is_end: the text ends after the highest valid position (otherwise it continues beyond the window).
 * Returns false when a rule needs context from outside the window (the caller falls back to the per-position rules). */
JTK_HD bool jtk_boundary_eval(const jtk_tile_ctx &c, int ws, const jtk_planes &pl, uint32_t valid, bool starts_doc, bool top_is_end, uint32_t *Bout) {
	const uint32_t P0 = pl.P0, P1 = pl.P1, P2 = pl.P2, P3 = pl.P3, CONT = pl.CONT & valid;
	/* class masks from the four bit planes of the class code (jtk_common.h) */
	const uint32_t O_ = ~P3 & ~P2 & ~P1 & valid;                  /* 0, 1 */
	const uint32_t AP = O_ & P0;                                  /* 1 */
	const uint32_t SP = ~P3 & ~P2 & P1 & ~P0 & valid;             /* 2 */
	const uint32_t NL = ~P3 & ~P2 & P1 & P0 & valid;              /* 3 */
	const uint32_t WO = ~P3 & P2 & ~P1 & ~P0 & valid;             /* 4 */
	const uint32_t N = ~P3 & P2 & ~P1 & P0 & valid;               /* 5 */
	const uint32_t L = (P3 | (P2 & P1)) & valid;                  /* 6 .. 14 */
	const uint32_t S1 = ((~P3 & P2 & P1 & P0) | (P3 & ~P2 & ~(P1 & P0))) & valid; /* 7 .. 10: s t m d */
	const uint32_t RV = ((P3 & ~P2 & P1 & P0) | (P3 & P2 & ~P1 & ~P0)) & valid;   /* 11, 12: r v */
	const uint32_t LL = P3 & P2 & ~P1 & P0 & valid;               /* 13 */
	const uint32_t LE = P3 & P2 & P1 & ~P0 & valid;               /* 14 */
	if ((SP | NL | WO | N | S1 | RV | LL | LE) & pl.HB) return false; /* multi-byte whitespace / digit / contraction letter */
	const uint32_t low = starts_doc ? 0u : (valid & 1u); /* window position 0 exists and has text before it */
	const uint32_t LEAD = ~CONT;
	const uint32_t Ostart = O_ & LEAD & ~((O_ | SP) << 1);
	const uint32_t OstartAll = jtk_spread_up(Ostart, CONT & O_); /* ... on every byte of that character */
	const uint32_t APs = AP & Ostart;
	const uint32_t CE = (L << 1) & (((S1 << 1) & (APs << 2)) | ((((LE << 1) & (RV << 2)) | ((LL << 1) & (LL << 2))) & (APs << 3)));
	uint32_t B;
	if (c.T->pattern_kind == JTK_PAT_CL100K) {
		const uint32_t Wn = SP | WO, W = Wn | NL;
		const uint32_t BL = L & ((N << 1) | (NL << 1) | ((O_ << 1) & ~(OstartAll << 1)) | CE);
		const uint32_t BNL = NL & ((L | N) << 1);
		uint32_t BW = Wn & (~(W << 1) | ((~W & valid) >> 1)); /* start of a run, or its last character when a non-whitespace character follows */
		const uint32_t cand = Wn & (NL << 1) & ~BW & 0xFFFF00u; /* after an NL, followed by whitespace: needs the two run scans */
		if (cand) {
			/* does a \r\n come before the whitespace run ends?  unknown when the run leaves the window */
			const uint32_t hit = jtk_spread_down(NL, Wn), edge = top_is_end ? 0u : jtk_spread_down(Wn & 0x80000000u, Wn);
			const uint32_t nla = hit >> 1;
			if (cand & (edge >> 1) & ~nla) return false;
			BW |= cand & ~nla;
			const uint32_t cand2 = cand & nla;
			if (cand2) {
				/* are the NLs before it the tail of an "other" piece?  unknown when the NL run reaches the window start */
				const uint32_t tail = jtk_spread_up(O_, NL) & NL, open = jtk_spread_up(NL & low, NL);
				if (cand2 & (open << 1)) return false;
				BW |= cand2 & (tail << 1);
			}
		}
		/* \p{N}{1,3}: every third position of a digit run starts a piece */
		uint32_t BN = 0;
		if (N & 0xFFFF00u) {
			uint32_t T0 = N & ~(N << 1);
			if (N & low) { /* the run of window position 0 began earlier: its phase comes from a walk */
				T0 &= ~1u;
				const int k = jtk_count_n_before(c, ws);
				const int o = (3 - k) % 3;
				const uint32_t run0 = jtk_spread_up(1u, N);
				T0 |= (1u << o) & run0;
			}
			const uint32_t C3 = N & (N << 1) & (N << 2) & (N << 3);
			uint32_t T = T0 | ((T0 << 3) & C3);
			const uint32_t C6 = C3 & (C3 << 3);
			T |= (T << 6) & C6;
			const uint32_t C12 = C6 & (C6 << 6);
			T |= (T << 12) & C12;
			BN = T & N;
		}
		B = BL | Ostart | BNL | BW | BN;
	} else {
		const uint32_t W = SP | WO | NL;
		const uint32_t CS = APs & ((S1 >> 1) | ((RV >> 1) & (LE >> 2)) | ((LL >> 1) & (LL >> 2)));
		const uint32_t BL = L & (CE | ~((L | SP | CS) << 1));
		const uint32_t BN = N & ~((N | SP) << 1);
		const uint32_t BW = W & (~(W << 1) | ((~W & valid) >> 1));
		B = BL | Ostart | BN | BW;
	}
	*Bout = B & LEAD & valid;
	return true;
}

JTK_HD bool jtk_boundary_fast(const jtk_tile_ctx &c, int chunk, uint32_t *out) {
	const int r0 = chunk * 16;
	const int ws = r0 - 8; /* window start */
	if (ws < c.rs || c.g0 + r0 + 24 > c.total) return false;
	/* document starts at window positions 0 .. 31 (D) and right after the window (Dtop) */
	uint32_t D;
	{
		const int wi = ws >> 5, sh = ws & 31;
		D = c.dmask[wi] >> sh;
		if (sh) D |= c.dmask[wi + 1] << (32 - sh);
	}
	const uint32_t Dtop = (c.dmask[(ws + 32) >> 5] >> ((ws + 32) & 31)) & 1u;
	/* the window [ws, ws + 32) = last eight positions of the previous chunk, this chunk, first eight of the next one */
	jtk_planes pl;
	{
		const uint32_t *pp = c.planes + 4 * (chunk - 1);
#if defined(__CUDA_ARCH__)
		const uint4 qa = *reinterpret_cast<const uint4 *>(pp), qb = *reinterpret_cast<const uint4 *>(pp + 4), qc = *reinterpret_cast<const uint4 *>(pp + 8);
		const uint32_t a0 = qa.x, a1 = qa.y, a2 = qa.z, b0 = qb.x, b1 = qb.y, b2 = qb.z, c0 = qc.x, c1 = qc.y, c2 = qc.z;
#else
		const uint32_t a0 = pp[0], a1 = pp[1], a2 = pp[2], b0 = pp[4], b1 = pp[5], b2 = pp[6], c0 = pp[8], c1 = pp[9], c2 = pp[10];
#endif
#define JTK_WIN_LO(a, b, c) ((((a) >> 8) & 0xFFu) | (((b) & 0xFFFFu) << 8) | (((c) & 0xFFu) << 24))
#define JTK_WIN_HI(a, b, c) ((((a) >> 24) & 0xFFu) | (((b) >> 16) << 8) | ((((c) >> 16) & 0xFFu) << 24))
		pl.P0 = JTK_WIN_LO(a0, b0, c0);
		pl.P1 = JTK_WIN_HI(a0, b0, c0);
		pl.P2 = JTK_WIN_LO(a1, b1, c1);
		pl.P3 = JTK_WIN_HI(a1, b1, c1);
		pl.CONT = JTK_WIN_LO(a2, b2, c2);
		pl.HB = JTK_WIN_HI(a2, b2, c2);
#undef JTK_WIN_LO
#undef JTK_WIN_HI
	}
	uint32_t B;
	if ((D | Dtop) == 0) {
		if (!jtk_boundary_eval(c, ws, pl, 0xFFFFFFFFu, false, false, &B)) return false;
	} else {
		/* one document start in or right after the window: the rules are evaluated once for the text that ends before it and once
		 * for the text that starts at it (short documents otherwise send every eighth chunk through the per-position rules) */
		if ((D & (D - 1)) || (D && Dtop)) return false;
		const int d = D ? jtk_ctz(D) : 32;
		B = 0;
		if (d > 0) {
			uint32_t Bl;
			if (!jtk_boundary_eval(c, ws, pl, d >= 32 ? 0xFFFFFFFFu : ((1u << d) - 1u), false, true, &Bl)) return false;
			B |= Bl;
		}
		if (d < 32) {
			uint32_t Br;
			if (!jtk_boundary_eval(c, ws, pl, ~((1u << d) - 1u), true, false, &Br)) return false;
			B |= Br | (1u << d);
		}
	}
	*out = (B >> 8) & 0xFFFFu;
	return true;
}

/* Piece-start bits for the 16 positions of chunk `chunk`. */
JTK_HD uint32_t jtk_boundary_generic(const jtk_tile_ctx &c, int chunk) {
	const int r0 = chunk * 16;
	uint32_t bits = 0;
	int nrun = -1;
	for (int i = 0; i < 16; i++) {
		const int r = r0 + i;
		if (r < c.rs) continue;
		if (jtk_docstart(c, r)) {
			bits |= 1u << i;
			nrun = 0;
			if (c.g0 + r >= c.total) break;
			if (jtk_clsb(c, r) == JTK_C_N) nrun = 1; /* class N on a lead byte */
			continue;
		}
		if (c.g0 + r >= c.total) break;
		const int cb = jtk_clsb(c, r);
		if (cb & JTK_CONT) continue;
		const int cur = cb & JTK_CLS_MASK;
		if (jtk_is_piece_start(c, r, cur, &nrun)) bits |= 1u << i;
		if (cur == JTK_C_N) {
			if (nrun < 0) nrun = jtk_count_n_before(c, r);
			nrun = (nrun + 1) % 3;
		} else {
			nrun = 0;
		}
	}
	return bits;
}

JTK_HD uint32_t jtk_boundary_chunk(const jtk_tile_ctx &c, int chunk) {
	uint32_t bits = 0;
	if (jtk_boundary_fast(c, chunk, &bits)) return bits;
	return jtk_boundary_generic(c, chunk);
}

/* ---------------------------------------------------------------------------------------------
 * safe cuts: positions inside a piece where bytePairMerge can be run separately on both sides (jtk_safe_cut below: the byte
 * bigram around the position occurs in no token, so no merge ever crosses it and the piece as a whole cannot be a token
 * either).  The split kernel lists such positions as additional piece starts: the unit of the merge kernels and of the memo
 * becomes the SEGMENT - in scripts the vocabulary covers poorly a character or two instead of a run of letters.  Cuts are
 * optional, so only chunks that contain a non-ASCII byte are examined (the 16-byte chunk grid is the same in every tile).
 * ------------------------------------------------------------------------------------------- */
JTK_HD bool jtk_safe_cut(const uint32_t *bits, uint32_t b0, uint32_t b1);
/* The finer test with one byte of context on either side: the first merge that crosses the position between b1 and b2 joins a part
 * that ends with b1 and a part that begins with b2 into a token; that token is b1 b2 itself, or it contains b0 b1 b2 or b1 b2 b3.  If
 * none of the three exists no merge crosses the position.  (b0 / b3 may belong to a neighbouring document or be padding: the test
 * can then only say "may cross" too often, never too rarely.) */
JTK_HD bool jtk_safe_cut4(const jtk_tables &T, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
	if (jtk_safe_cut(T.bigram_bits, b1, b2)) return true;
	if (JTK_LDG(T.bytepair + ((b1 << 8) | b2)) != JTK_RANK_MAX) return false;
	const uint32_t s0 = jtk_trigram_slot(b0, b1, b2), s1 = jtk_trigram_slot(b1, b2, b3);
	return !(((JTK_LDG(T.trigram_bits + (s0 >> 5)) >> (s0 & 31)) | (JTK_LDG(T.trigram_bits + (s1 >> 5)) >> (s1 & 31))) & 1u);
}
/* cut bits of the 16 positions of chunk `chunk` (chunk >= 1); the caller masks document starts / the end of the input.  Only character
 * boundaries next to a non-ASCII character are examined: a third of the positions in CJK text, almost none in English. */
JTK_HD uint32_t jtk_cut_chunk(const jtk_tile_ctx &c, int chunk) {
#ifdef JTK_NO_CUTS
	return 0; /* development switch: no safe cuts at all (the pieces of the split pattern are the unit everywhere) */
#endif
	const uint32_t pch = c.planes[4 * chunk + 2];
	if ((pch >> 16) == 0) return 0; /* all ASCII */
	const uint8_t *p = c.sb + 16 * chunk;
	uint32_t bits = 0;
	/* candidates: character boundaries with a non-ASCII character on at least one side (ASCII runs are left to the split pattern) */
	const uint32_t hb = pch >> 16, side = hb | (hb << 1) | (uint32_t) (p[-1] >> 7);
	for (uint32_t m = side & ~pch & 0xFFFFu; m; m &= m - 1) {
		const int i = jtk_ctz(m);
		if (jtk_safe_cut4(*c.T, p[i - 2], p[i - 1], p[i], p[i + 1])) bits |= 1u << i;
	}
	return bits;
}

/* ---------------------------------------------------------------------------------------------
 * special-token guard
 * ------------------------------------------------------------------------------------------- */
/* Does any special token start at global position g (document ends at doc_hi)? */
JTK_HD bool jtk_special_at(const jtk_tables &T, const uint8_t *gbytes, int64_t g, int64_t doc_hi) {
	for (int s = 0; s < T.nspecial; s++) {
		uint32_t a = T.special_off[s], b = T.special_off[s + 1];
		uint32_t len = b - a;
		if (len == 0 || g + len > doc_hi) continue;
		uint32_t i = 0;
		while (i < len && gbytes[g + i] == T.special_bytes[a + i]) i++;
		if (i == len) return true;
	}
	return false;
}

/* The longest special token that starts at global position g and ends at or before hi: its index and *len, or -1.
 * (Special-token ENCODING, jtk_encode_batch_special: not in the reference, semantics of tiktoken's allowed_special="all".) */
JTK_HD int jtk_special_match(const jtk_tables &T, const uint8_t *gbytes, int64_t g, int64_t hi, int *len) {
	int best = -1;
	uint32_t best_len = 0;
	for (int s = 0; s < T.nspecial; s++) {
		const uint32_t a = T.special_off[s], n = T.special_off[s + 1] - a;
		if (n <= best_len || g + n > hi) continue;
		uint32_t i = 0;
		while (i < n && gbytes[g + i] == T.special_bytes[a + i]) i++;
		if (i == n) {
			best = s;
			best_len = n;
		}
	}
	*len = (int) best_len;
	return best;
}

/* ---------------------------------------------------------------------------------------------
 * table lookups
 * ------------------------------------------------------------------------------------------- */
/* whole-piece lookup for keys of 2..24 bytes (six zero-padded key words, h = jtk_hash6(k, len)); the rank or JTK_RANK_MAX */
JTK_HD int32_t jtk_lookup_a(const jtk_tables &T, const uint32_t *k, uint32_t len, uint32_t h) {
	uint32_t b = h & T.mask_a;
	for (;;) {
		const jtk_slot_a s = T.tab_a[b]; /* 32 bytes, one sector: two 16-byte loads */
		if (s.len == 0) return JTK_RANK_MAX;
		if (s.len == len && s.k01[0] == k[0] && s.k01[1] == k[1] && s.k25[0] == k[2] && s.k25[1] == k[3] && s.k25[2] == k[4] && s.k25[3] == k[5]) return (int32_t) s.rank;
		b = (b + 1) & T.mask_a;
	}
}

/* whole-piece lookup for keys of 1..8 bytes (k0, k1 zero padded): the first half of a slot holds everything */
JTK_HD uint32_t jtk_hash6_short(uint32_t k0, uint32_t k1, uint32_t len) {
	const uint32_t k[6] = {k0, k1, 0u, 0u, 0u, 0u}; /* the four zero words fold into constants */
	return jtk_hash6(k, len);
}
JTK_HD int32_t jtk_lookup_a8(const jtk_tables &T, uint32_t k0, uint32_t k1, uint32_t len, uint32_t h) {
	uint32_t b = h & T.mask_a;
	for (;;) {
#if defined(__CUDA_ARCH__)
		const uint4 s = JTK_LDG(reinterpret_cast<const uint4 *>(T.tab_a + b));
		const uint32_t s0 = s.x, s1 = s.y, sl = s.z, sr = s.w;
#else
		const uint32_t s0 = T.tab_a[b].k01[0], s1 = T.tab_a[b].k01[1], sl = T.tab_a[b].len, sr = T.tab_a[b].rank;
#endif
		if (sl == 0) return JTK_RANK_MAX;
		if (sl == len && s0 == k0 && s1 == k1) return (int32_t) sr;
		b = (b + 1) & T.mask_a;
	}
}

/* Six zero-padded key words of the n <= 24 bytes at p: unaligned words from seven aligned loads (the staging buffer is
 * padded; p + 27 must be readable), masked to n bytes. */
JTK_HD void jtk_build_key(const uint8_t *p, int n, uint32_t *k) {
	const uint32_t *aw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t) 3);
	const int sh = (int) (reinterpret_cast<uintptr_t>(p) & 3) * 8;
	uint32_t a[7];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int i = 0; i < 7; i++) a[i] = aw[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int i = 0; i < 6; i++) {
		const uint32_t w = sh ? (a[i] >> sh) | (a[i + 1] << (32 - sh)) : a[i];
		const int rem = n - 4 * i; /* key bytes that belong to word i */
		k[i] = rem >= 4 ? w : rem <= 0 ? 0u : (w & ((1u << (8 * rem)) - 1u));
	}
}

/* the first four zero-padded key words of the n <= 16 bytes at p (five aligned loads) */
JTK_HD void jtk_build_key4(const uint8_t *p, int n, uint32_t *k) {
	const uint32_t *aw = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t) 3);
	const int sh = (int) (reinterpret_cast<uintptr_t>(p) & 3) * 8;
	uint32_t a[5];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int i = 0; i < 5; i++) a[i] = aw[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
	for (int i = 0; i < 4; i++) {
		const uint32_t w = sh ? (a[i] >> sh) | (a[i + 1] << (32 - sh)) : a[i];
		const int rem = n - 4 * i;
		k[i] = rem >= 4 ? w : rem <= 0 ? 0u : (w & ((1u << (8 * rem)) - 1u));
	}
}

/* whole-piece lookup for keys of 25..max_token_len bytes, verified byte by byte */
JTK_HD int32_t jtk_lookup_b(const jtk_tables &T, const uint8_t *p, uint32_t n) {
	uint64_t h = jtk_hash_bytes_init();
	for (uint32_t i = 0; i < n; i++) h = jtk_hash_bytes_step(h, p[i]);
	h = jtk_hash_bytes_final(h, n);
	const uint32_t lo = (uint32_t) h, hi = (uint32_t) (h >> 32);
	uint32_t b = lo & T.mask_b;
	for (;;) {
		for (int j = 0; j < 2; j++) {
			const jtk_slot s = T.tab_b[2 * b + j];
			if (s.w == 0) return JTK_RANK_MAX;
			if (s.x == lo && s.y == hi) {
				const uint32_t a = T.tok_off[s.w - 1], e = T.tok_off[s.w];
				if (e - a == n) {
					uint32_t i = 0;
					while (i < n && T.tok_bytes[a + i] == p[i]) i++;
					if (i == n) return (int32_t) s.z;
				}
			}
		}
		b = (b + 1) & T.mask_b;
	}
}

/* rank of the concatenation of two parts, or JTK_RANK_MAX (getRank, GptBytePairEncoding.java:285-300) */
JTK_HD int32_t jtk_pair_resolve(const jtk_tables &T, uint32_t b, const jtk_slot &f0, const jtk_slot &f1, int32_t l, int32_t r) {
	jtk_slot s0 = f0, s1 = f1;
	for (;;) {
		if (s0.w == 0) return JTK_RANK_MAX;
		if (s0.x == (uint32_t) l && s0.y == (uint32_t) r) return (int32_t) s0.z;
		if (s1.w == 0) return JTK_RANK_MAX;
		if (s1.x == (uint32_t) l && s1.y == (uint32_t) r) return (int32_t) s1.z;
		b = (b + 1) & T.mask_p;
		s0 = T.pair[2 * b];
		s1 = T.pair[2 * b + 1];
	}
}

JTK_HD int32_t jtk_lookup_pair(const jtk_tables &T, int32_t l, int32_t r) {
	const uint32_t b = jtk_hash_pair(l, r) & T.mask_p;
	return jtk_pair_resolve(T, b, T.pair[2 * b], T.pair[2 * b + 1], l, r);
}

/* The two rank probes of one merge step (:254-257) with all four slot loads in flight together. */
JTK_HD void jtk_lookup_pair2(const jtk_tables &T, int32_t l1, int32_t r1, int32_t l2, int32_t r2, int32_t *o1, int32_t *o2) {
	const uint32_t b1 = jtk_hash_pair(l1, r1) & T.mask_p, b2 = jtk_hash_pair(l2, r2) & T.mask_p;
	const jtk_slot a0 = T.pair[2 * b1], a1 = T.pair[2 * b1 + 1];
	const jtk_slot c0 = T.pair[2 * b2], c1 = T.pair[2 * b2 + 1];
	*o1 = jtk_pair_resolve(T, b1, a0, a1, l1, r1);
	*o2 = jtk_pair_resolve(T, b2, c0, c1, l2, r2);
}

/* Up to four rank probes with all slot loads in flight together; probe i is skipped (JTK_RANK_MAX) when bit i of `valid` is clear. */
JTK_HD void jtk_lookup_pair4(const jtk_tables &T, const int32_t *l, const int32_t *r, uint32_t valid, int32_t *out) {
	uint32_t b[4];
	jtk_slot s[8];
	for (int i = 0; i < 4; i++) {
		b[i] = jtk_hash_pair(l[i], r[i]) & T.mask_p;
		if ((valid >> i) & 1u) {
			s[2 * i] = T.pair[2 * b[i]];
			s[2 * i + 1] = T.pair[2 * b[i] + 1];
		}
	}
	for (int i = 0; i < 4; i++) out[i] = ((valid >> i) & 1u) ? jtk_pair_resolve(T, b[i], s[2 * i], s[2 * i + 1], l[i], r[i]) : JTK_RANK_MAX;
}

/* keys of 25..max_token_len bytes: filter on (first eight bytes, length) first, the byte-wise hash only for survivors */
JTK_HD int32_t jtk_lookup_long(const jtk_tables &T, const uint8_t *p, int n) {
	uint32_t k[6];
	jtk_build_key(p, 8, k);
	const uint32_t f = jtk_hash3(k[0], k[1], (uint32_t) n) & 0xFFFFu;
	if (!((T.long_filter[f >> 5] >> (f & 31)) & 1u)) return JTK_RANK_MAX;
	return jtk_lookup_b(T, p, (uint32_t) n);
}

/* Whole-piece lookup of the n bytes at p (n >= 1): rank or JTK_RANK_MAX. */
JTK_HD int32_t jtk_lookup_piece(const jtk_tables &T, const uint8_t *p, int n) {
	if (n == 1) {
		int32_t id = T.byte_id[p[0]];
		return id < JTK_PSEUDO_BASE + 256 ? JTK_RANK_MAX : id;
	}
	if (n <= JTK_INLINE_KEY_MAX) {
		uint32_t k[6];
		jtk_build_key(p, n, k);
		return jtk_lookup_a(T, k, (uint32_t) n, jtk_hash6(k, (uint32_t) n));
	}
	if (n > T.max_token_len) return JTK_RANK_MAX;
	return jtk_lookup_long(T, p, n);
}

/* ---------------------------------------------------------------------------------------------
 * bytePairMerge for one piece of 2..JTK_SHORT_PIECE bytes by one thread (GptBytePairEncoding.java:200-275)
 * tok / rk: n staging slots each, element k at index k * stride (the kernel interleaves the slots of the threads of
 * a CTA so that slot k of every thread falls into a different bank).  Returns the token count; tokens end up in slots 0..count-1.
 * *unknown is set when a final part is a byte that is not in the vocabulary.
 * ------------------------------------------------------------------------------------------- */
JTK_HD int jtk_ctz64(uint64_t m) {
#if defined(__CUDA_ARCH__)
	return __ffsll((long long) m) - 1;
#else
	return __builtin_ctzll(m);
#endif
}
JTK_HD int jtk_clz64(uint64_t m) {
#if defined(__CUDA_ARCH__)
	return __clzll((long long) m);
#else
	return __builtin_clzll(m);
#endif
}

/* May bytePairMerge be run separately on the bytes before and after a position whose neighbouring bytes are b0, b1?  Yes when no
 * token contains b0 b1 next to each other: every part the loop ever forms is a token (or a single byte), so no part can span the
 * position, the pair across it never has a rank, and a merge on one side only changes ranks on that side - the global leftmost-minimum
 * order (GptBytePairEncoding.java:232-240) restricted to either side is that side's own order.  `bits` = jtk_tables::bigram_bits. */
JTK_HD bool jtk_safe_cut(const uint32_t *bits, uint32_t b0, uint32_t b1) {
	const uint32_t g = (b0 << 8) | b1;
	return !((JTK_LDG(bits + (g >> 5)) >> (g & 31)) & 1u);
}

/* MaskT = uint32_t for pieces up to 32 bytes, uint64_t up to 64 bytes */
template <typename MaskT>
JTK_HD int jtk_merge_short_t(const jtk_tables &T, const uint8_t *p, int n, int32_t *tok, int32_t *rk, int stride, bool *unknown) {
	constexpr int BITS = (int) sizeof(MaskT) * 8;
	const MaskT ONE = 1;
	auto ctz = [](MaskT m) { return sizeof(MaskT) == 8 ? jtk_ctz64((uint64_t) m) : jtk_ctz((uint32_t) m); };
	auto top = [](MaskT m) { return sizeof(MaskT) == 8 ? 63 - jtk_clz64((uint64_t) m) : 31 - jtk_clz((uint32_t) m); };
	for (int k = 0; k < n; k++) {
		tok[k * stride] = T.byte_id[p[k]];
		rk[k * stride] = (k + 1 < n) ? T.bytepair[((uint32_t) p[k] << 8) | p[k + 1]] : JTK_RANK_MAX;
	}
	MaskT alive = (n >= BITS) ? ~(MaskT) 0 : ((ONE << n) - ONE);
	for (;;) {
		/* leftmost strict minimum (:232-240), and the leftmost minimum over the other positions: the candidate for the NEXT
		 * iteration of the reference loop, merged in the same round trip to the pair table when that is provably what the
		 * sequential loop would do */
		int32_t mr = JTK_RANK_MAX, mr2 = JTK_RANK_MAX;
		int mi = -1, mi2 = -1;
		for (int half = 0; half < BITS / 32; half++) { /* 32 bits at a time: find-first-set on 64 bits is several instructions */
			for (uint32_t m = (uint32_t) (alive >> (half ? BITS - 32 : 0)); m;) {
				const int k = 32 * half + jtk_ctz(m);
				m &= m - 1;
				const int32_t r = rk[k * stride];
				if (r < mr) {
					mr2 = mr;
					mi2 = mi;
					mr = r;
					mi = k;
				} else if (r < mr2) {
					mr2 = r;
					mi2 = k;
				}
			}
		}
		if (mi < 0) break; /* :247,260-262 */
		const MaskT above = alive & ~(((ONE << mi) << 1) - ONE); /* parts after mi */
		const int nx = ctz(above);
		const MaskT above2 = above & (above - 1);
		const MaskT below = alive & ((ONE << mi) - ONE);
		const int pv = below ? top(below) : -1;
		tok[mi * stride] = mr; /* rank == id of the merged token */
		alive &= ~(ONE << nx);
		rk[nx * stride] = JTK_RANK_MAX;
		/* The second merge is what the reference does next iff its pair is untouched by the first merge (not a neighbour of
		 * it) and neither of the two ranks the first merge recomputes comes before it in (rank, position) order.  The id of
		 * a merged token is the rank just found, so all four probes can be issued before any of them returns. */
		const bool spec = mi2 >= 0 && mi2 != nx && mi2 != pv;
		int nx2 = -1, nn2 = -1, pv2 = -1;
		if (spec) {
			const MaskT aboveB = alive & ~(((ONE << mi2) << 1) - ONE);
			nx2 = ctz(aboveB);
			const MaskT above2B = aboveB & (aboveB - 1);
			const MaskT belowB = alive & ((ONE << mi2) - ONE);
			if (above2B) nn2 = ctz(above2B);
			if (belowB) pv2 = top(belowB);
		}
		int32_t pl[4], pr[4], res[4];
		uint32_t valid = 0;
		pl[0] = mr, pr[0] = 0;
		if (above2) { /* :254 */
			pr[0] = tok[ctz(above2) * stride];
			valid |= 1u;
		}
		pl[1] = 0, pr[1] = mr;
		if (pv >= 0) { /* :255-257 */
			pl[1] = tok[pv * stride];
			valid |= 2u;
		}
		pl[2] = mr2, pr[2] = 0;
		if (nn2 >= 0) {
			pr[2] = tok[nn2 * stride];
			valid |= 4u;
		}
		pl[3] = 0, pr[3] = mr2;
		if (pv2 >= 0) {
			pl[3] = tok[pv2 * stride];
			valid |= 8u;
		}
		jtk_lookup_pair4(T, pl, pr, valid, res);
		rk[mi * stride] = res[0];
		if (pv >= 0) rk[pv * stride] = res[1];
		if (spec) {
			const bool left_first = pv >= 0 && (res[1] < mr2 || (res[1] == mr2 && pv < mi2));
			const bool right_first = res[0] < mr2 || (res[0] == mr2 && mi < mi2);
			if (!left_first && !right_first) {
				tok[mi2 * stride] = mr2;
				alive &= ~(ONE << nx2);
				rk[nx2 * stride] = JTK_RANK_MAX;
				rk[mi2 * stride] = res[2];
				if (pv2 >= 0) rk[pv2 * stride] = res[3];
			}
		}
	}
	int cnt = 0;
	for (MaskT m = alive; m;) {
		const int k = ctz(m);
		m &= m - 1;
		const int32_t t = tok[k * stride];
		if (t < JTK_PSEUDO_BASE + 256) *unknown = true;
		tok[(cnt++) * stride] = t;
	}
	return cnt;
}

JTK_HD int jtk_merge_short(const jtk_tables &T, const uint8_t *p, int n, int32_t *tok, int32_t *rk, int stride, bool *unknown) {
	if (n <= 32) return jtk_merge_short_t<uint32_t>(T, p, n, tok, rk, stride, unknown);
	return jtk_merge_short_t<uint64_t>(T, p, n, tok, rk, stride, unknown);
}

/* Sequential bytePairMerge for any length (used by the host-side emulator for pieces the device handles
 * with the warp-cooperative loop); nxt is an n+1 scratch array. */
JTK_HD int jtk_merge_seq(const jtk_tables &T, const uint8_t *p, int n, int32_t *tok, int32_t *rk, int32_t *nxt, bool *unknown) {
	for (int k = 0; k < n; k++) {
		tok[k] = T.byte_id[p[k]];
		rk[k] = (k + 1 < n) ? T.bytepair[((uint32_t) p[k] << 8) | p[k + 1]] : JTK_RANK_MAX;
		nxt[k] = k + 1;
	}
	for (;;) {
		int32_t mr = JTK_RANK_MAX;
		int mi = -1, pv = -1, mpv = -1;
		for (int k = 0; k < n; pv = k, k = nxt[k])
			if (rk[k] < mr) {
				mr = rk[k];
				mi = k;
				mpv = pv;
			}
		if (mi < 0) break;
		int nx = nxt[mi], nn = nxt[nx];
		tok[mi] = mr;
		nxt[mi] = nn;
		rk[mi] = (nn < n) ? jtk_lookup_pair(T, mr, tok[nn]) : JTK_RANK_MAX;
		if (mpv >= 0) rk[mpv] = jtk_lookup_pair(T, tok[mpv], mr);
	}
	int cnt = 0;
	for (int k = 0; k < n;) {
		int32_t t = tok[k];
		int nk = nxt[k];
		if (t < JTK_PSEUDO_BASE + 256) *unknown = true;
		tok[cnt++] = t;
		k = nk;
	}
	return cnt;
}

#endif /* JTK_DEVICE_CUH */
