/*
 * ORACLE - TEST INFRASTRUCTURE ONLY (see jo_oracle.h for the rules and the parity pin).
 *
 * Restates, function by function, the reference's encode path:
 *   GptBytePairEncoding.java:47-59   encodeInternal (special-token guard)     -> jo_encode(check_special)
 *   GptBytePairEncoding.java:71-103  encodeOrdinaryInternal (find loop, fast path, back-off) -> encode_core / jo_encode_max
 *   GptBytePairEncoding.java:110-119 addTokens                               -> inside encode_core
 *   GptBytePairEncoding.java:136-151 decodeBytes, :302-314 decodeToken         -> jo_decode_bytes
 *   GptBytePairEncoding.java:200-275 bytePairMerge, :285-300 getRank           -> merge_literal
 *   TokenEncoder.java:38-45,53-82    the two hash maps                        -> vocab_* (byte-string keyed)
 *   ImmutableByteArray.java:62-79    getBytesBetween (slice copy per probe)   -> slice copy in rank_of_slice
 * It keeps the reference's algorithmic shape on purpose (hash map keyed by byte strings, one slice
 * copy + hash per rank probe, O(n) removal from the parts list), because it doubles as the timed
 * CPU baseline.  merge_heap is the exact sub-quadratic variant used for very long pieces.
 */
#define _GNU_SOURCE
#include "jo_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jo_regex.h"

#define RANK_MAX 0x7fffffff /* Integer.MAX_VALUE sentinel, GptBytePairEncoding.java:208 */

/* ------------------------------------------------------------------ byte-string keyed map */
typedef struct vocab {
	uint8_t *bytes;   /* concatenated keys */
	int64_t *off;     /* nkeys + 1 */
	int32_t *rank;    /* nkeys */
	int64_t nkeys;
	int64_t *slots;   /* bucket heads: index into keys or -1 (chained like java.util.HashMap) */
	int64_t *chain;   /* next key in the same bucket or -1 */
	uint32_t *khash;  /* cached hash per key, as HashMap.Node.hash */
	uint64_t mask;
	int64_t maxlen;
	/* inverse map: rank -> key index (sorted by rank) */
	int64_t *by_rank;
} vocab;

static uint32_t bytes_hash(const uint8_t *p, int64_t n) { /* Arrays.hashCode(byte[]) as ImmutableByteArray.java:104-107 */
	uint32_t h = 1;
	for (int64_t i = 0; i < n; i++) h = 31u * h + (uint32_t) (int32_t) (int8_t) p[i];
	return h;
}
static uint64_t spread(uint32_t h) { /* HashMap.hash() spreading */ return (uint64_t) (h ^ (h >> 16)); }

static int64_t vocab_find(const vocab *v, const uint8_t *p, int64_t n) {
	if (v->nkeys == 0 || n > v->maxlen) return -1;
	const uint32_t h = (uint32_t) spread(bytes_hash(p, n));
	for (int64_t k = v->slots[h & v->mask]; k >= 0; k = v->chain[k]) {
		if (v->khash[k] != h) continue;
		int64_t len = v->off[k + 1] - v->off[k];
		if (len == n && memcmp(v->bytes + v->off[k], p, (size_t) n) == 0) return k;
	}
	return -1;
}

static int cmp_rank_idx(const void *a, const void *b, void *ctx) {
	const vocab *v = (const vocab *) ctx;
	int32_t ra = v->rank[*(const int64_t *) a], rb = v->rank[*(const int64_t *) b];
	return ra < rb ? -1 : ra > rb ? 1 : 0;
}

static void vocab_build(vocab *v, const uint8_t *keys, const int64_t *key_off, const int32_t *ranks, int64_t nkeys) {
	memset(v, 0, sizeof(*v));
	int64_t total = nkeys ? key_off[nkeys] : 0;
	v->bytes = (uint8_t *) malloc((size_t) (total ? total : 1));
	if (total) memcpy(v->bytes, keys, (size_t) total);
	v->off = (int64_t *) malloc(sizeof(int64_t) * (size_t) (nkeys + 1));
	v->rank = (int32_t *) malloc(sizeof(int32_t) * (size_t) (nkeys ? nkeys : 1));
	v->off[0] = 0;
	uint64_t cap = 16;
	while (cap * 3 < (uint64_t) nkeys * 4 + 4) cap <<= 1; /* HashMap load factor 0.75 */
	v->mask = cap - 1;
	v->slots = (int64_t *) malloc(sizeof(int64_t) * cap);
	v->chain = (int64_t *) malloc(sizeof(int64_t) * (size_t) (nkeys ? nkeys : 1));
	v->khash = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (nkeys ? nkeys : 1));
	for (uint64_t i = 0; i < cap; i++) v->slots[i] = -1;
	/* Map.put semantics: a later duplicate key replaces the earlier value (TokenEncoder.java:41-44). */
	int64_t m = 0, pos = 0;
	for (int64_t i = 0; i < nkeys; i++) {
		const uint8_t *p = keys + key_off[i];
		int64_t len = key_off[i + 1] - key_off[i];
		v->nkeys = m;
		if (len > v->maxlen) v->maxlen = len;
		int64_t k = vocab_find(v, p, len);
		if (k >= 0) {
			v->rank[k] = ranks[i];
			continue;
		}
		memcpy(v->bytes + pos, p, (size_t) len);
		v->off[m] = pos;
		pos += len;
		v->off[m + 1] = pos;
		v->rank[m] = ranks[i];
		const uint32_t h = (uint32_t) spread(bytes_hash(p, len));
		v->khash[m] = h;
		v->chain[m] = v->slots[h & v->mask];
		v->slots[h & v->mask] = m;
		m++;
	}
	v->nkeys = m;
	v->by_rank = (int64_t *) malloc(sizeof(int64_t) * (size_t) (m ? m : 1));
	for (int64_t i = 0; i < m; i++) v->by_rank[i] = i;
	qsort_r(v->by_rank, (size_t) m, sizeof(int64_t), cmp_rank_idx, v);
}

static int64_t vocab_find_rank(const vocab *v, int32_t rank) {
	int64_t lo = 0, hi = v->nkeys - 1;
	while (lo <= hi) {
		int64_t mid = (lo + hi) >> 1;
		int32_t r = v->rank[v->by_rank[mid]];
		if (r < rank) lo = mid + 1;
		else if (r > rank) hi = mid - 1;
		else {
			/* encodedToDecoded.put: the last key put for a value wins; keys are unique here so any hit is fine,
			 * but prefer the last one inserted for determinism */
			while (mid + 1 < v->nkeys && v->rank[v->by_rank[mid + 1]] == rank) mid++;
			return v->by_rank[mid];
		}
	}
	return -1;
}

static void vocab_free(vocab *v) {
	free(v->bytes);
	free(v->off);
	free(v->rank);
	free(v->slots);
	free(v->chain);
	free(v->khash);
	free(v->by_rank);
}

/* ------------------------------------------------------------------ encoding object */
struct jo_encoding {
	jo_regex *re;
	vocab enc;
	vocab spec; /* special tokens: string bytes -> id */
};

jo_encoding *jo_create(const char *pattern, int flags, const uint8_t *keys, const int64_t *key_off, const int32_t *ranks, int64_t nkeys,
                       const uint8_t *spec, const int64_t *spec_off, const int32_t *spec_ids, int64_t nspec, char *err, int errlen) {
	jo_regex *re = jo_regex_compile(pattern, flags, err, errlen);
	if (!re) return NULL;
	jo_encoding *e = (jo_encoding *) calloc(1, sizeof(*e));
	e->re = re;
	vocab_build(&e->enc, keys, key_off, ranks, nkeys);
	vocab_build(&e->spec, spec, spec_off, spec_ids, nspec);
	return e;
}

void jo_destroy(jo_encoding *e) {
	if (!e) return;
	jo_regex_free(e->re);
	vocab_free(&e->enc);
	vocab_free(&e->spec);
	free(e);
}

/* ------------------------------------------------------------------ UTF-8 helpers */
/* Decode to code points keeping byte offsets.  Input comes from String.getBytes(UTF_8)
 * (ImmutableByteArray.java:16-19) and is therefore well formed; a stray byte is kept as its own
 * "character" of class other so that the function is total. */
static int64_t utf8_decode_all(const uint8_t *s, int64_t n, uint32_t *cps, int64_t *boff) {
	int64_t m = 0, i = 0;
	while (i < n) {
		uint8_t b = s[i];
		uint32_t cp = 0xFFFD;
		int len = 1;
		if (b < 0x80) {
			cp = b;
		} else if (b >= 0xC2 && b <= 0xDF && i + 1 < n && (s[i + 1] & 0xC0) == 0x80) {
			cp = ((uint32_t) (b & 0x1F) << 6) | (s[i + 1] & 0x3F);
			len = 2;
		} else if (b >= 0xE0 && b <= 0xEF && i + 2 < n && (s[i + 1] & 0xC0) == 0x80 && (s[i + 2] & 0xC0) == 0x80) {
			cp = ((uint32_t) (b & 0x0F) << 12) | ((uint32_t) (s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F);
			len = 3;
		} else if (b >= 0xF0 && b <= 0xF4 && i + 3 < n && (s[i + 1] & 0xC0) == 0x80 && (s[i + 2] & 0xC0) == 0x80 && (s[i + 3] & 0xC0) == 0x80) {
			cp = ((uint32_t) (b & 0x07) << 18) | ((uint32_t) (s[i + 1] & 0x3F) << 12) | ((uint32_t) (s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F);
			len = 4;
		}
		cps[m] = cp;
		boff[m] = i;
		m++;
		i += len;
	}
	boff[m] = n;
	return m;
}

static int not_cont(uint8_t b) { return (b & 0xC0) != 0x80; }

/* new String(bytes, StandardCharsets.UTF_8) as the JDK decodes it (String.decodeUTF8_UTF16 with
 * replacement; JDK source knowledge, the JDK is not under /root/reference): every malformed or
 * truncated sequence becomes one U+FFFD. */
int64_t jo_java_utf8_to_utf16(const uint8_t *src, int64_t sl, uint16_t *dst) {
	int64_t sp = 0, dp = 0;
	const uint16_t REPL = 0xFFFD;
	while (sp < sl) {
		uint8_t b1 = src[sp++];
		if (b1 < 0x80) {
			dst[dp++] = b1;
		} else if ((b1 & 0xE0) == 0xC0 && (b1 & 0x1E) != 0) {
			if (sp < sl) {
				uint8_t b2 = src[sp++];
				if (not_cont(b2)) {
					dst[dp++] = REPL;
					sp--;
				} else {
					dst[dp++] = (uint16_t) (((b1 & 0x1F) << 6) | (b2 & 0x3F));
				}
				continue;
			}
			dst[dp++] = REPL;
			break;
		} else if ((b1 & 0xF0) == 0xE0) {
			if (sp + 1 < sl) {
				uint8_t b2 = src[sp++], b3 = src[sp++];
				if ((b1 == 0xE0 && (b2 & 0xE0) == 0x80) || not_cont(b2) || not_cont(b3)) {
					dst[dp++] = REPL;
					sp -= 3;
					/* malformed3 */
					sp += ((b1 == 0xE0 && (src[sp + 1] & 0xE0) == 0x80) || not_cont(src[sp + 1])) ? 1 : 2;
				} else {
					uint16_t c = (uint16_t) (((b1 & 0x0F) << 12) | ((b2 & 0x3F) << 6) | (b3 & 0x3F));
					dst[dp++] = (c >= 0xD800 && c <= 0xDFFF) ? REPL : c;
				}
				continue;
			}
			if (sp < sl && ((b1 == 0xE0 && (src[sp] & 0xE0) == 0x80) || not_cont(src[sp]))) {
				dst[dp++] = REPL;
				continue;
			}
			dst[dp++] = REPL;
			break;
		} else if ((b1 & 0xF8) == 0xF0) {
			if (sp + 2 < sl) {
				uint8_t b2 = src[sp++], b3 = src[sp++], b4 = src[sp++];
				uint32_t uc = ((uint32_t) (b1 & 0x07) << 18) | ((uint32_t) (b2 & 0x3F) << 12) | ((uint32_t) (b3 & 0x3F) << 6) | (b4 & 0x3F);
				if (not_cont(b2) || not_cont(b3) || not_cont(b4) || uc < 0x10000 || uc > 0x10FFFF) {
					dst[dp++] = REPL;
					sp -= 4;
					/* malformed4 */
					uint8_t c1 = src[sp], c2 = src[sp + 1];
					if (c1 > 0xF4 || (c1 == 0xF0 && (c2 < 0x90 || c2 > 0xBF)) || (c1 == 0xF4 && (c2 & 0xF0) != 0x80) || not_cont(c2)) sp += 1;
					else if (not_cont(src[sp + 2])) sp += 2;
					else sp += 3;
				} else {
					uc -= 0x10000;
					dst[dp++] = (uint16_t) (0xD800 + (uc >> 10));
					dst[dp++] = (uint16_t) (0xDC00 + (uc & 0x3FF));
				}
				continue;
			}
			if (b1 > 0xF4 || (sp < sl && ((b1 == 0xF0 && (src[sp] < 0x90 || src[sp] > 0xBF)) || (b1 == 0xF4 && (src[sp] & 0xF0) != 0x80) || not_cont(src[sp])))) {
				dst[dp++] = REPL;
				continue;
			}
			sp++;
			if (sp < sl && not_cont(src[sp])) {
				dst[dp++] = REPL;
				continue;
			}
			dst[dp++] = REPL;
			break;
		} else {
			dst[dp++] = REPL;
		}
	}
	return dp;
}

/* ------------------------------------------------------------------ bytePairMerge, literal */
/* getRank (GptBytePairEncoding.java:285-300): copies the slice, hashes it, probes the map. */
static int32_t rank_of_slice(const vocab *v, const uint8_t *piece, int64_t from, int64_t to, uint8_t *scratch) {
	int64_t len = to - from;
	if (len > v->maxlen) return RANK_MAX;
	memcpy(scratch, piece + from, (size_t) len); /* ImmutableByteArray.getBytesBetween copies */
	int64_t k = vocab_find(v, scratch, len);
	return k < 0 ? RANK_MAX : v->rank[k];
}

static int64_t emit_parts(const vocab *v, const uint8_t *piece, const int64_t *idx, int64_t nparts, int32_t *out, int64_t cap) {
	/* :270-274 - encoder.encode throws for a part that is not in the vocabulary */
	if (nparts - 1 > cap) return JO_E_CAPACITY;
	for (int64_t i = 0; i + 1 < nparts; i++) {
		int64_t k = vocab_find(v, piece + idx[i], idx[i + 1] - idx[i]);
		if (k < 0) return JO_E_UNKNOWN_BYTES;
		out[i] = v->rank[k];
	}
	return nparts - 1;
}

static int64_t merge_literal(const vocab *v, const uint8_t *piece, int64_t n, int32_t *out, int64_t cap) {
	int64_t nparts = n + 1;
	int64_t idx_small[130];
	int32_t rk_small[130];
	uint8_t scratch_small[264];
	const int small = nparts <= 130 && v->maxlen < 264;
	int64_t *idx = small ? idx_small : (int64_t *) malloc(sizeof(int64_t) * (size_t) nparts);
	int32_t *rk = small ? rk_small : (int32_t *) malloc(sizeof(int32_t) * (size_t) nparts);
	uint8_t *scratch = small ? scratch_small : (uint8_t *) malloc((size_t) (v->maxlen + 1));
	for (int64_t i = 0; i < nparts; i++) {
		idx[i] = i;
		rk[i] = RANK_MAX;
	}
	for (int64_t i = 0; i < nparts - 2; i++) rk[i] = rank_of_slice(v, piece, idx[i], idx[i + 2], scratch);
	while (nparts > 1) {
		int64_t mi = 0;
		int32_t mr = RANK_MAX;
		for (int64_t i = 0; i < nparts - 1; i++)
			if (rk[i] < mr) {
				mr = rk[i];
				mi = i;
			}
		if (mr == RANK_MAX) break;
		/* both affected ranks are recomputed with skip = 1 before the removal (:254-257) */
		rk[mi] = (mi + 3 >= nparts) ? RANK_MAX : rank_of_slice(v, piece, idx[mi], idx[mi + 3], scratch);
		if (mi > 0) rk[mi - 1] = (mi + 2 >= nparts) ? RANK_MAX : rank_of_slice(v, piece, idx[mi - 1], idx[mi + 2], scratch);
		memmove(idx + mi + 1, idx + mi + 2, sizeof(int64_t) * (size_t) (nparts - mi - 2)); /* parts.remove(minRankIndex + 1) */
		memmove(rk + mi + 1, rk + mi + 2, sizeof(int32_t) * (size_t) (nparts - mi - 2));
		nparts--;
	}
	int64_t r = emit_parts(v, piece, idx, nparts, out, cap);
	if (!small) {
		free(idx);
		free(rk);
		free(scratch);
	}
	return r;
}

/* ------------------------------------------------------------------ bytePairMerge, exact heap variant */
typedef struct hent {
	int32_t rank;
	int64_t pos;
} hent;
static int hless(hent a, hent b) { return a.rank < b.rank || (a.rank == b.rank && a.pos < b.pos); }
static void hpush(hent *h, int64_t *hn, hent e) {
	int64_t i = (*hn)++;
	while (i > 0) {
		int64_t p = (i - 1) >> 1;
		if (!hless(e, h[p])) break;
		h[i] = h[p];
		i = p;
	}
	h[i] = e;
}
static hent hpop(hent *h, int64_t *hn) {
	hent top = h[0], last = h[--(*hn)];
	int64_t i = 0, n = *hn;
	for (;;) {
		int64_t c = 2 * i + 1;
		if (c >= n) break;
		if (c + 1 < n && hless(h[c + 1], h[c])) c++;
		if (!hless(h[c], last)) break;
		h[i] = h[c];
		i = c;
	}
	if (n > 0) h[i] = last;
	return top;
}

static int64_t merge_heap(const vocab *v, const uint8_t *piece, int64_t n, int32_t *out, int64_t cap) {
	/* parts are identified by their start byte; nxt/prv form the list; sentinel start n closes it */
	int64_t *nxt = (int64_t *) malloc(sizeof(int64_t) * (size_t) (n + 2));
	int64_t *prv = (int64_t *) malloc(sizeof(int64_t) * (size_t) (n + 2));
	int32_t *cur = (int32_t *) malloc(sizeof(int32_t) * (size_t) (n + 2));
	uint8_t *alive = (uint8_t *) malloc((size_t) (n + 2));
	hent *heap = (hent *) malloc(sizeof(hent) * (size_t) (3 * n + 4));
	uint8_t *scratch = (uint8_t *) malloc((size_t) (v->maxlen + 1));
	int64_t hn = 0;
	for (int64_t i = 0; i <= n; i++) {
		nxt[i] = i + 1;
		prv[i] = i - 1;
		alive[i] = 1;
		cur[i] = RANK_MAX;
	}
	for (int64_t i = 0; i + 2 <= n; i++) {
		cur[i] = rank_of_slice(v, piece, i, i + 2, scratch);
		if (cur[i] != RANK_MAX) hpush(heap, &hn, (hent) {cur[i], i});
	}
	while (hn > 0) {
		hent e = hpop(heap, &hn);
		int64_t i = e.pos;
		if (!alive[i] || cur[i] != e.rank) continue; /* stale entry */
		int64_t j = nxt[i];     /* part to absorb */
		int64_t k = nxt[j];     /* new right neighbour start (<= n) */
		alive[j] = 0;
		nxt[i] = k;
		prv[k] = i;
		cur[i] = (k < n) ? rank_of_slice(v, piece, i, nxt[k], scratch) : RANK_MAX;
		if (cur[i] != RANK_MAX) hpush(heap, &hn, (hent) {cur[i], i});
		int64_t p = prv[i];
		if (p >= 0) {
			cur[p] = rank_of_slice(v, piece, p, k, scratch);
			if (cur[p] != RANK_MAX) hpush(heap, &hn, (hent) {cur[p], p});
		}
	}
	int64_t cnt = 0, r = 0;
	for (int64_t i = 0; i < n; i = nxt[i]) {
		if (cnt >= cap) {
			r = JO_E_CAPACITY;
			break;
		}
		int64_t kx = vocab_find(v, piece + i, nxt[i] - i);
		if (kx < 0) {
			r = JO_E_UNKNOWN_BYTES;
			break;
		}
		out[cnt++] = v->rank[kx];
	}
	free(nxt);
	free(prv);
	free(cur);
	free(alive);
	free(heap);
	free(scratch);
	return r < 0 ? r : cnt;
}

int64_t jo_merge_piece(const jo_encoding *e, const uint8_t *piece, int64_t n, int merge_algo, int32_t *out, int64_t cap) {
	if (merge_algo == JO_MERGE_HEAP || (merge_algo == JO_MERGE_AUTO && n > 256)) return merge_heap(&e->enc, piece, n, out, cap);
	return merge_literal(&e->enc, piece, n, out, cap);
}

/* ------------------------------------------------------------------ the find loop */
typedef struct decoded_text {
	uint32_t *cps;
	int64_t *boff;
	int64_t ncp;
} decoded_text;

static void decode_text(const uint8_t *text, int64_t n, decoded_text *d) {
	d->cps = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (n + 1));
	d->boff = (int64_t *) malloc(sizeof(int64_t) * (size_t) (n + 2));
	d->ncp = utf8_decode_all(text, n, d->cps, d->boff);
}
static void free_text(decoded_text *d) {
	free(d->cps);
	free(d->boff);
}

/* Matcher.find(): resume at the end of the previous match; after an empty match advance by one. */
static int next_match(const jo_encoding *e, const decoded_text *d, int64_t *first, int64_t *last) {
	int64_t from = *last;
	if (from == *first) from++;
	if (from > d->ncp) return 0;
	int64_t ms, me;
	if (!jo_regex_search(e->re, d->cps, d->ncp, from, &ms, &me)) return 0;
	*first = ms;
	*last = me;
	return 1;
}

int64_t jo_split(const jo_encoding *e, const uint8_t *text, int64_t n, int64_t *starts, int64_t *ends, int64_t cap) {
	decoded_text d;
	decode_text(text, n, &d);
	int64_t first = -1, last = 0, cnt = 0;
	while (next_match(e, &d, &first, &last)) {
		if (cnt >= cap) {
			cnt = JO_E_CAPACITY;
			break;
		}
		starts[cnt] = d.boff[first];
		ends[cnt] = d.boff[last];
		cnt++;
	}
	free_text(&d);
	return cnt;
}

int jo_contains_special(const jo_encoding *e, const uint8_t *text, int64_t n) {
	for (int64_t s = 0; s < e->spec.nkeys; s++) {
		const uint8_t *p = e->spec.bytes + e->spec.off[s];
		int64_t len = e->spec.off[s + 1] - e->spec.off[s];
		if (len == 0) return 1; /* "".contains -> true */
		for (int64_t i = 0; i + len <= n; i++)
			if (text[i] == p[0] && memcmp(text + i, p, (size_t) len) == 0) return 1;
	}
	return 0;
}

/* encodeOrdinaryInternal up to (not including) the back-off loop.  max_tokens < 0 means null. */
static int64_t encode_core(const jo_encoding *e, const uint8_t *text, int64_t n, int has_max, int max_tokens, int merge_algo, int32_t *out, int64_t cap) {
	decoded_text d;
	decode_text(text, n, &d);
	int64_t first = -1, last = 0, cnt = 0, status = 0;
	int32_t *tmp = NULL;
	int64_t tmpcap = 0;
	while (next_match(e, &d, &first, &last) && !(has_max && (int64_t) max_tokens <= cnt)) {
		const uint8_t *piece = text + d.boff[first];
		int64_t plen = d.boff[last] - d.boff[first];
		int64_t k = vocab_find(&e->enc, piece, plen); /* containsDecodedToken + encode, :81-83 */
		if (k >= 0) {
			if (cnt >= cap) {
				status = JO_E_CAPACITY;
				break;
			}
			out[cnt++] = e->enc.rank[k];
			continue;
		}
		if (plen + 1 > tmpcap) {
			tmpcap = plen + 64;
			tmp = (int32_t *) realloc(tmp, sizeof(int32_t) * (size_t) tmpcap);
		}
		int64_t m = jo_merge_piece(e, piece, plen, merge_algo, tmp, tmpcap);
		if (m < 0) {
			status = m;
			break;
		}
		if (has_max && m > (int64_t) max_tokens - cnt) m = (int64_t) max_tokens - cnt; /* addTokens, :110-119 */
		if (cnt + m > cap) {
			status = JO_E_CAPACITY;
			break;
		}
		memcpy(out + cnt, tmp, sizeof(int32_t) * (size_t) m);
		cnt += m;
	}
	free(tmp);
	free_text(&d);
	return status < 0 ? status : cnt;
}

/* Special-token ENCODING.  The reference does not implement it (README.md:46 "not started"; encodeInternal throws, :52-56), so
 * this restates the upstream the reference mirrors - tiktoken's encode(text, allowed_special="all") - with the reference's
 * pieces: the text is cut at every occurrence of a special token (leftmost first, non-overlapping; the longest token when
 * several start at the same position), the stretches between them are encoded like encodeOrdinary each on its own, and every
 * occurrence contributes the special token's id.  Empty special tokens never match. */
int64_t jo_encode_with_special(const jo_encoding *e, const uint8_t *text, int64_t n, int merge_algo, int32_t *out, int64_t cap) {
	int64_t cnt = 0, seg = 0, i = 0;
	while (i <= n) {
		int64_t best = -1, best_len = 0;
		if (i < n)
			for (int64_t s = 0; s < e->spec.nkeys; s++) {
				const uint8_t *p = e->spec.bytes + e->spec.off[s];
				int64_t len = e->spec.off[s + 1] - e->spec.off[s];
				if (len > best_len && i + len <= n && memcmp(text + i, p, (size_t) len) == 0) {
					best = s;
					best_len = len;
				}
			}
		if (best < 0 && i < n) {
			i++;
			continue;
		}
		int64_t m = encode_core(e, text + seg, i - seg, 0, 0, merge_algo, out + cnt, cap - cnt);
		if (m < 0) return m;
		cnt += m;
		if (best < 0) break;
		if (cnt >= cap) return JO_E_CAPACITY;
		out[cnt++] = e->spec.rank[best];
		i += best_len;
		seg = i;
	}
	return cnt;
}

int64_t jo_encode(const jo_encoding *e, const uint8_t *text, int64_t n, int check_special, int merge_algo, int32_t *out, int64_t cap) {
	if (check_special && jo_contains_special(e, text, n)) return JO_E_SPECIAL;
	return encode_core(e, text, n, 0, 0, merge_algo, out, cap);
}

int64_t jo_decode_bytes(const jo_encoding *e, const int32_t *ids, int64_t n, uint8_t *out, int64_t cap, int32_t *bad_id) {
	int64_t pos = 0;
	for (int64_t i = 0; i < n; i++) {
		const vocab *v = &e->enc;
		int64_t k = vocab_find_rank(v, ids[i]);
		if (k < 0) {
			v = &e->spec;
			k = vocab_find_rank(v, ids[i]);
		}
		if (k < 0) {
			if (bad_id) *bad_id = ids[i];
			return JO_E_UNKNOWN_ID;
		}
		int64_t len = v->off[k + 1] - v->off[k];
		if (pos + len > cap) return JO_E_CAPACITY;
		memcpy(out + pos, v->bytes + v->off[k], (size_t) len);
		pos += len;
	}
	return pos;
}

int64_t jo_encode_max(const jo_encoding *e, const uint8_t *text, int64_t n, int check_special, int max_tokens, int32_t *out, int64_t cap,
                      int *truncated) {
	*truncated = 0;
	if (check_special && jo_contains_special(e, text, n)) return JO_E_SPECIAL;
	int64_t cnt = encode_core(e, text, n, 1, max_tokens, JO_MERGE_AUTO, out, cap);
	if (cnt < 0) return cnt;
	/* :90-100 - drop trailing tokens until decode(tokens) is a prefix of the text (UTF-16 comparison) */
	uint16_t *t16 = (uint16_t *) malloc(sizeof(uint16_t) * (size_t) (n + 1));
	int64_t tlen = jo_java_utf8_to_utf16(text, n, t16);
	int64_t dcap = 0;
	for (int64_t i = 0; i < cnt; i++) dcap += e->enc.maxlen > e->spec.maxlen ? e->enc.maxlen : e->spec.maxlen;
	uint8_t *dbytes = (uint8_t *) malloc((size_t) (dcap + 1));
	uint16_t *d16 = (uint16_t *) malloc(sizeof(uint16_t) * (size_t) (dcap + 1));
	int64_t result = 0;
	for (int64_t drop = 0; drop <= cnt; drop++) {
		int32_t bad;
		int64_t nb = jo_decode_bytes(e, out, cnt - drop, dbytes, dcap, &bad);
		if (nb < 0) {
			result = nb;
			break;
		}
		int64_t dlen = jo_java_utf8_to_utf16(dbytes, nb, d16);
		if (dlen <= tlen && memcmp(t16, d16, sizeof(uint16_t) * (size_t) dlen) == 0) {
			*truncated = tlen > dlen;
			result = cnt - drop;
			break;
		}
	}
	free(t16);
	free(dbytes);
	free(d16);
	return result;
}

/* ------------------------------------------------------------------ per-document thread pool */
typedef struct batch_job {
	const jo_encoding *e;
	const uint8_t *utf8;
	const int64_t *doc_off;
	int64_t ndocs;
	int check_special, merge_algo;
	int32_t *ids;
	int64_t *counts;
	volatile int64_t next; /* shared work counter */
} batch_job;

static void *batch_worker(void *arg) {
	batch_job *j = (batch_job *) arg;
	for (;;) {
		int64_t d = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
		if (d >= j->ndocs) break;
		int64_t b = j->doc_off[d], n = j->doc_off[d + 1] - b;
		j->counts[d] = jo_encode(j->e, j->utf8 + b, n, j->check_special, j->merge_algo, j->ids + b, n);
	}
	return NULL;
}

int64_t jo_encode_batch(const jo_encoding *e, const uint8_t *utf8, const int64_t *doc_off, int64_t ndocs, int nthreads, int check_special,
                        int merge_algo, int32_t *ids, int64_t *counts) {
	batch_job j;
	j.e = e;
	j.utf8 = utf8;
	j.doc_off = doc_off;
	j.ndocs = ndocs;
	j.check_special = check_special;
	j.merge_algo = merge_algo;
	j.ids = ids;
	j.counts = counts;
	j.next = 0;
	if (nthreads < 1) nthreads = 1;
	pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t) * (size_t) nthreads);
	for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &j);
	batch_worker(&j);
	for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
	free(th);
	int64_t total = 0;
	for (int64_t d = 0; d < ndocs; d++)
		if (counts[d] > 0) total += counts[d];
	return total;
}
