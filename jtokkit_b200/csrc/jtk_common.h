/*
 * Shared between host table construction and the device kernels.
 *
 * Everything here is the device-side replacement of the reference's L0 layer:
 *   TokenEncoder.java:16-17     two HashMaps                 -> piece tables A/B, pair table, token byte store
 *   java.util.regex.Pattern     compiled split pattern       -> code point class tables + a rule kind
 */
#ifndef JTK_COMMON_H
#define JTK_COMMON_H

#include <stdint.h>

#if defined(__CUDACC__)
#define JTK_HD __host__ __device__ __forceinline__
#else
#define JTK_HD inline
struct uint4 { /* host builds without the CUDA headers (tests/emu) */
	unsigned int x, y, z, w;
};
#endif

/* ---- code point classes (4 bits) ---------------------------------------------------------------
 * The split patterns (EncodingFactory.java:63,105) only distinguish \p{L}, \p{N}, \s (with ' ' and
 * [\r\n] singled out), the apostrophe, "everything else", and - for the contraction alternatives
 * 's|'t|'re|'ve|'m|'ll|'d - eight letters.  A continuation byte carries its lead byte's class plus
 * JTK_CONT so that "class of the previous character" is a single byte read. */
enum {
	JTK_C_O = 0,   /* other: [^\s\p{L}\p{N}] */
	JTK_C_AP = 1,  /* apostrophe (also "other") */
	JTK_C_SP = 2,  /* U+0020 */
	JTK_C_NL = 3,  /* \r \n (cl100k only; folded into WO for the x50k pattern) */
	JTK_C_WO = 4,  /* any other White_Space */
	JTK_C_N = 5,   /* \p{N} */
	JTK_C_L = 6,   /* \p{L} without a contraction role */
	JTK_C_LS = 7,  /* s (cl100k: also S and U+017F) */
	JTK_C_LT = 8,
	JTK_C_LM = 9,
	JTK_C_LD = 10,
	JTK_C_LR = 11,
	JTK_C_LV = 12,
	JTK_C_LL = 13,
	JTK_C_LE = 14,
	JTK_C_NONE = 15 /* no character: before a document's first / after its last */
};
#define JTK_CONT 0x80
#define JTK_CLS_MASK 0x0F

JTK_HD bool jtk_is_letter(int c) { return c >= JTK_C_L && c <= JTK_C_LE; }
JTK_HD bool jtk_is_space(int c) { return c >= JTK_C_SP && c <= JTK_C_WO; }
JTK_HD bool jtk_is_other(int c) { return c <= JTK_C_AP; }

/* ---- split rule kinds ("compiled pattern") ---------------------------------------------------- */
enum {
	JTK_PAT_X50K = 1,  /* r50k_base / p50k_base / p50k_edit, EncodingFactory.java:63,77,91 */
	JTK_PAT_CL100K = 2, /* cl100k_base, EncodingFactory.java:105 */
	JTK_PAT_GENERAL = 3 /* any other pattern: backtracking program (jtk_regex.h), one thread per document */
};

#define JTK_RANK_MAX 0x7fffffff                  /* Integer.MAX_VALUE sentinel, GptBytePairEncoding.java:208 */
#define JTK_PSEUDO_BASE ((int32_t) 0x80000000)   /* id of a single byte that is not in the vocabulary: PSEUDO_BASE + byte */
#define JTK_INLINE_KEY_MAX 24                    /* piece table A stores keys of up to 24 bytes inline */

/* ---- tile geometry (overridable so that the host-side emulator in tests/ can use tiny tiles) ----------
 * The region a CTA stages = back halo + tile + forward halo.  With the default numbers it is 511 chunks of 16 bytes plus one
 * pad chunk = 512: one chunk per thread of a 512-thread CTA in the classification and split steps, no partial second round. */
#ifndef JTK_TILE
#define JTK_TILE 7072       /* bytes owned by one tile (a multiple of 32) */
#endif
#ifndef JTK_BACK_HALO
#define JTK_BACK_HALO 64    /* context bytes before the tile */
#endif
#ifndef JTK_LONG_PIECE
#define JTK_LONG_PIECE 1024 /* pieces longer than this go to the long-piece kernels */
#endif
#define JTK_FWD_HALO (JTK_LONG_PIECE + 16) /* a piece starting in the tile ends inside the halo or is long */
#define JTK_REGION (JTK_BACK_HALO + JTK_TILE + JTK_FWD_HALO)
#define JTK_REGION_CHUNKS (JTK_REGION / 16)
#ifndef JTK_NT
/* threads per CTA of the split+lookup kernel: one per 16-byte chunk of the region incl. the pad chunk, rounded up to warps */
#define JTK_NT ((JTK_REGION_CHUNKS + 1 + 31) / 32 * 32)
#endif
#ifndef JTK_SHORT_PIECE
#define JTK_SHORT_PIECE 64  /* thread-per-piece merge up to this length, lane groups above */
#endif

/* 16-byte table slots */
struct jtk_slot {
	uint32_t x, y, z, w;
};
/* 32-byte slot of piece table A: 24 key bytes (zero padded), key length, rank; one slot per 32-byte sector.  The first 16 bytes
 * hold everything a key of up to eight bytes needs (90 % of the pieces of English text): one 16-byte load, three compares. */
struct alignas(16) jtk_slot_a {
	uint32_t k01[2]; /* key bytes 0..7 */
	uint32_t len;    /* 0 = empty */
	uint32_t rank;
	uint32_t k25[4]; /* key bytes 8..23 */
};

/* Per-call piece memo (direct mapped): the tokens bytePairMerge produced for a short piece, so that later occurrences of
 * the same piece in the same call cost one probe instead of a merge loop.  Filled by the merge kernel, read by the
 * split+lookup kernels of later sub-batches / chunks of the SAME call (stream ordered, never concurrently); entries of
 * earlier calls are recognised by their epoch and treated as empty.  Results are identical with or without it. */
#define JTK_MEMO_MAX_PIECE 16
#define JTK_MEMO_MAX_TOKENS 10
struct alignas(64) jtk_memo_entry {
	uint32_t key[4];
	uint32_t meta;  /* epoch << 8 | state (0 empty, 1 being written, 2 valid) */
	uint32_t n_cnt; /* piece length | token count << 8 */
	int32_t tok[JTK_MEMO_MAX_TOKENS];
};

/* Device tables of one encoding on one device (all pointers are device memory). */
struct jtk_tables {
	int32_t pattern_kind;
	int32_t max_token_len;
	/* code point -> class: ASCII direct, the rest two-level (cp >> 8 -> block, block * 256 + low byte) */
	const uint8_t *ascii_cls;   /* 128 */
	const uint16_t *cp_stage1;  /* 0x1100 */
	const uint8_t *cp_stage2;   /* nblocks * 256 */
	/* the same classes in the forms the tile kernel reads (jtk_classify_chunk) */
	const uint32_t *lut_sp;     /* 256: class bits 0..3 of an ASCII byte in bits 0, 8, 16, 24; zero for bytes >= 0x80 */
	const uint8_t *cls2;        /* 2048: class of the code points below U+0800 (everything a two-byte sequence can encode) */
	const uint8_t *bmp_nib;     /* 32768: class of the code points below U+10000, two per byte (low nibble = even code point) */
	/* whole-piece lookup (GptBytePairEncoding.java:81-83), keys <= 24 bytes inline */
	const jtk_slot_a *tab_a;
	uint32_t mask_a;            /* slot count - 1 (linear probing) */
	/* whole-piece lookup, keys of 25..max_token_len bytes: slot = {hash lo, hash hi, rank, token index}, verified against tok_bytes */
	const jtk_slot *tab_b;
	uint32_t mask_b;
	const uint32_t *long_filter; /* 65 536-bit filter over (first eight bytes, length) of the keys of table B: most long pieces skip the byte-wise hash */
	const uint8_t *tok_bytes;   /* concatenated token bytes, by token index */
	const uint32_t *tok_off;    /* ntokens + 1 */
	/* merge loop (GptBytePairEncoding.java:200-300) re-keyed on (id left, id right) -> rank of the concatenation */
	const int32_t *byte_id;     /* 256: rank of the single byte or JTK_PSEUDO_BASE + byte */
	const int32_t *bytepair;    /* 65536: rank of the two-byte token b0 b1, or JTK_RANK_MAX */
	const jtk_slot *pair;       /* slot = {id left, id right, rank, 1} ; empty slot has w == 0 */
	uint32_t mask_p;
	const uint32_t *trigram_bits; /* 2^20 bits, hashed (jtk_trigram_slot): set for every three bytes that occur next to each other in some token (a set bit may
	                               * also be a hash collision: the test errs on the side of "occurs") */
	const uint32_t *bigram_bits; /* 65536 bits: bit (b0 << 8 | b1) is set when some token contains the bytes b0 b1 next to each other.  Where it is
	                              * clear no merge can ever join the two bytes, so bytePairMerge runs independently on both sides (jtk_safe_cut) */
	/* special-token guard (GptBytePairEncoding.java:52-56) */
	int32_t nspecial;
	int32_t special_has_empty;  /* "".contains: every document is flagged */
	const uint8_t *special_bytes;
	const uint32_t *special_off; /* nspecial + 1 */
	const int32_t *special_ids;  /* nspecial (special-token encoding, jtk_encode_batch_special) */
	uint32_t special_first[8];   /* bitmap of first bytes */
	uint32_t special_first_single; /* that byte when exactly one non-zero first byte occurs, else 0 */
	/* decode (GptBytePairEncoding.java:136-151,302-314): id -> token index, open addressing {id, token index + 1} */
	const uint32_t *dec_keys;    /* pairs (id, index + 1), 0 in the second word = empty */
	uint32_t mask_d;
	const uint8_t *dec_bytes;    /* ordinary tokens followed by special-token strings */
	const uint32_t *dec_off;
	const uint4 *dec_direct;     /* id -> {offset into dec_bytes << 8 | length (0xFFFFFFFF = unknown id), first twelve bytes} when all ids are small and non-negative, else null */
	uint32_t dec_direct_size;
	/* JTK_PAT_GENERAL only: the compiled split program (jtk_regex.h: jtk_rx_inst / jtk_rx_set / code point ranges) */
	const void *rx_inst;
	const void *rx_sets;
	const uint32_t *rx_ranges;
	int32_t rx_ninst;
	uint32_t rx_first[8]; /* bytes with which a match can begin (all ones when the pattern can match the empty string) */
	/* the same program as a DFA over code-point classes (jtk_dfa.cpp), null when the pattern has no DFA form */
	const uint16_t *rx_dfa_trans;
	const uint16_t *rx_dfa_stage1;
	const uint8_t *rx_dfa_stage2;
	const uint8_t *rx_dfa_stay; /* nstates x 128 run codes (jtk_regex_compile.h), or null */
	int32_t rx_dfa_nsym, rx_dfa_nstates, rx_dfa_start, rx_dfa_start_bol, rx_dfa_acc_lo;
};

/* ---- hashing (identical on host and device) ---------------------------------------------------- */
JTK_HD uint32_t jtk_hash3(uint32_t a, uint32_t b, uint32_t c) {
	uint32_t h = a * 0x9E3779B1u;
	h ^= (b + 0x7F4A7C15u) * 0x85EBCA77u;
	h ^= (c + 0x165667B1u) * 0xC2B2AE3Du;
	h ^= h >> 15;
	h *= 0x2C1B3C6Du;
	h ^= h >> 13;
	return h;
}

/* hash of an inline key (six words; unused words are zero) */
JTK_HD uint32_t jtk_hash6(const uint32_t *k, uint32_t len) {
	uint32_t h = (k[0] + len) * 0x9E3779B1u;
	h ^= (k[1] + 0x7F4A7C15u) * 0x85EBCA77u;
	h = (h << 13) | (h >> 19);
	h ^= (k[2] + 0x165667B1u) * 0xC2B2AE3Du;
	h ^= (k[3] + 0x27D4EB2Fu) * 0x9E3779B1u;
	h = (h << 11) | (h >> 21);
	h ^= (k[4] + 0x85EBCA6Bu) * 0x85EBCA77u;
	h ^= (k[5] + 0xC2B2AE35u) * 0xC2B2AE3Du;
	h ^= h >> 15;
	h *= 0x2C1B3C6Du;
	h ^= h >> 13;
	return h;
}

JTK_HD uint32_t jtk_hash_pair(int32_t l, int32_t r) {
	uint32_t h = (uint32_t) l * 0x9E3779B1u;
	h ^= ((uint32_t) r + 0x7F4A7C15u) * 0x85EBCA77u;
	h ^= h >> 15;
	h *= 0x2C1B3C6Du;
	h ^= h >> 13;
	return h;
}

/* slot of a byte trigram in jtk_tables::trigram_bits */
JTK_HD uint32_t jtk_trigram_slot(uint32_t b0, uint32_t b1, uint32_t b2) { return (((b0 << 16) | (b1 << 8) | b2) * 0x9E3779B1u) >> 12; }

/* 64-bit byte-string hash for table B (FNV-1a over bytes, then a finaliser) */
JTK_HD uint64_t jtk_hash_bytes_step(uint64_t h, uint8_t b) { return (h ^ b) * 0x100000001B3ull; }
JTK_HD uint64_t jtk_hash_bytes_init() { return 0xCBF29CE484222325ull; }
JTK_HD uint64_t jtk_hash_bytes_final(uint64_t h, uint32_t len) {
	h ^= (uint64_t) len * 0x9E3779B97F4A7C15ull;
	h ^= h >> 29;
	h *= 0xBF58476D1CE4E5B9ull;
	h ^= h >> 32;
	return h;
}

#endif /* JTK_COMMON_H */
