/* Host-side table construction; see jtk_tables.h. */
#include "jtk_tables.h"
#include "jtk_regex_compile.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_map>

#include "unicode_ranges.inc"

#ifndef JTK_TABLE_A_LOAD
#define JTK_TABLE_A_LOAD 0.25 /* upper bound on the load of piece table A (slot count rounded up to a power of two): misses walk to the first empty slot, one dependent L2 load per step (measured: 0.5 -> 0.25 takes 5 % off the split+lookup kernel) */
#endif

static const char *const X50K_PATTERN = "'s|'t|'re|'ve|'m|'ll|'d| ?\\p{L}+| ?\\p{N}+| ?[^\\s\\p{L}\\p{N}]+|\\s+(?!\\S)|\\s+";
static const char *const CL100K_PATTERN =
    "(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\\r\\n\\p{L}\\p{N}]?\\p{L}+|\\p{N}{1,3}| ?[^\\s\\p{L}\\p{N}]+[\\r\\n]*|\\s*[\\r\\n]+|\\s+(?!\\S)|\\s+";

static const jtk_builtin_def BUILTINS[] = {
    {"r50k_base", X50K_PATTERN, 1, {"<|endoftext|>"}, {50256}},
    {"p50k_base", X50K_PATTERN, 1, {"<|endoftext|>"}, {50256}},
    {"p50k_edit", X50K_PATTERN, 4, {"<|endoftext|>", "<|fim_prefix|>", "<|fim_middle|>", "<|fim_suffix|>"}, {50256, 50281, 50282, 50283}},
    {"cl100k_base",
     CL100K_PATTERN,
     5,
     {"<|endoftext|>", "<|fim_prefix|>", "<|fim_middle|>", "<|fim_suffix|>", "<|endofprompt|>"},
     {100257, 100258, 100259, 100260, 100276}},
};

const char *jtk_unicode_version() { return JTK_UNICODE_VERSION; }

const jtk_builtin_def *jtk_find_builtin(const char *name) {
	for (const jtk_builtin_def &b : BUILTINS)
		if (!strcmp(b.name, name)) return &b;
	return nullptr;
}

/* ------------------------------------------------------------------ .tiktoken loader */
static int b64val(int c) {
	if (c >= 'A' && c <= 'Z') return c - 'A';
	if (c >= 'a' && c <= 'z') return c - 'a' + 26;
	if (c >= '0' && c <= '9') return c - '0' + 52;
	if (c == '+') return 62;
	if (c == '/') return 63;
	return -1;
}

int jtk_load_tiktoken_file(const char *path, std::vector<uint8_t> *bytes, std::vector<int64_t> *off, std::vector<int32_t> *ranks, std::string *err) {
	std::ifstream in(path, std::ios::binary);
	if (!in) {
		*err = std::string("Could not find ") + path;
		return JTK_E_ARG;
	}
	bytes->clear();
	off->assign(1, 0);
	ranks->clear();
	std::string line;
	while (std::getline(in, line)) {
		if (!line.empty() && line.back() == '\r') line.pop_back();
		if (line.empty()) continue;
		size_t sp = line.find_first_of(" \t");
		if (sp == std::string::npos) {
			*err = std::string("Invalid line in ") + path + ": " + line;
			return JTK_E_ARG;
		}
		uint32_t acc = 0;
		int nbits = 0;
		for (size_t i = 0; i < sp; i++) {
			if (line[i] == '=') break;
			int v = b64val((unsigned char) line[i]);
			if (v < 0) {
				*err = std::string("Invalid base64 in ") + path + ": " + line;
				return JTK_E_ARG;
			}
			acc = (acc << 6) | (uint32_t) v;
			nbits += 6;
			if (nbits >= 8) {
				nbits -= 8;
				bytes->push_back((uint8_t) ((acc >> nbits) & 0xFF));
			}
		}
		size_t rs = line.find_first_not_of(" \t", sp);
		if (rs == std::string::npos) {
			*err = std::string("Invalid line in ") + path + ": " + line;
			return JTK_E_ARG;
		}
		ranks->push_back((int32_t) strtol(line.c_str() + rs, nullptr, 10));
		off->push_back((int64_t) bytes->size());
	}
	return JTK_OK;
}

/* ------------------------------------------------------------------ pattern -> class tables */
static bool in_ranges(const uint32_t (*r)[2], int n, uint32_t cp) {
	int lo = 0, hi = n - 1;
	while (lo <= hi) {
		int mid = (lo + hi) >> 1;
		if (cp < r[mid][0]) hi = mid - 1;
		else if (cp > r[mid][1]) lo = mid + 1;
		else return true;
	}
	return false;
}

static int contraction_kind(uint32_t cp, bool fold_case, bool unicode_case) {
	if (fold_case) {
		if (cp >= 'A' && cp <= 'Z') cp += 32;
		if (unicode_case && cp == 0x17F) cp = 's'; /* toLowerCase(toUpperCase(U+017F)) == 's' */
	}
	switch (cp) {
	case 's': return JTK_C_LS;
	case 't': return JTK_C_LT;
	case 'm': return JTK_C_LM;
	case 'd': return JTK_C_LD;
	case 'r': return JTK_C_LR;
	case 'v': return JTK_C_LV;
	case 'l': return JTK_C_LL;
	case 'e': return JTK_C_LE;
	}
	return JTK_C_L;
}

static uint8_t classify_cp(uint32_t cp, int kind, int flags) {
	const bool ucc = (flags & JTK_RE_UNICODE_CHARACTER_CLASS) != 0;
	const bool ucase = ucc || (flags & JTK_RE_UNICODE_CASE);
	/* \p{L} and \p{N} are Unicode general categories with or without the flag; \s depends on it */
	if (in_ranges(JTK_UC_L, JTK_UC_L_COUNT, cp)) return (uint8_t) contraction_kind(cp, kind == JTK_PAT_CL100K, ucase);
	if (in_ranges(JTK_UC_N, JTK_UC_N_COUNT, cp)) return JTK_C_N;
	bool ws = ucc ? in_ranges(JTK_UC_WS, JTK_UC_WS_COUNT, cp) : (cp == ' ' || (cp >= 0x09 && cp <= 0x0D));
	if (ws) {
		if (cp == ' ') return JTK_C_SP;
		if (kind == JTK_PAT_CL100K && (cp == '\r' || cp == '\n')) return JTK_C_NL;
		return JTK_C_WO;
	}
	if (cp == '\'') return JTK_C_AP;
	return JTK_C_O;
}

static void build_class_tables(jtk_host_tables *t, int flags) {
	t->ascii_cls.resize(128);
	for (uint32_t cp = 0; cp < 128; cp++) t->ascii_cls[cp] = classify_cp(cp, t->pattern_kind, flags);
	t->cp_stage1.assign(0x1100, 0);
	t->cp_stage2.clear();
	std::unordered_map<std::string, uint16_t> seen;
	std::string block(256, '\0');
	for (uint32_t b = 0; b < 0x1100; b++) {
		for (uint32_t i = 0; i < 256; i++) block[i] = (char) classify_cp((b << 8) | i, t->pattern_kind, flags);
		auto it = seen.find(block);
		if (it == seen.end()) {
			uint16_t idx = (uint16_t) (t->cp_stage2.size() / 256);
			t->cp_stage2.insert(t->cp_stage2.end(), block.begin(), block.end());
			it = seen.emplace(block, idx).first;
		}
		t->cp_stage1[b] = it->second;
	}
	/* flat forms for the tile kernel */
	t->lut_sp.assign(256, 0);
	for (uint32_t b = 0; b < 128; b++) {
		const uint32_t k = t->ascii_cls[b];
		t->lut_sp[b] = (k & 1u) | (((k >> 1) & 1u) << 8) | (((k >> 2) & 1u) << 16) | (((k >> 3) & 1u) << 24);
	}
	t->cls2.resize(2048);
	for (uint32_t cp = 0; cp < 2048; cp++) t->cls2[cp] = classify_cp(cp, t->pattern_kind, flags);
	t->bmp_nib.assign(32768, 0);
	for (uint32_t cp = 0; cp < 65536; cp++) t->bmp_nib[cp >> 1] |= (uint8_t) (classify_cp(cp, t->pattern_kind, flags) << ((cp & 1u) * 4));
}

/* ------------------------------------------------------------------ hash tables */
static uint32_t pow2_at_least(uint64_t n) {
	uint32_t p = 16;
	while (p < n) p <<= 1;
	return p;
}

/* bucketed insert: a bucket is two consecutive slots; `empty` tells whether a slot is free */
template <typename EmptyFn>
static int bucket_insert(std::vector<jtk_slot> &tab, uint32_t mask, uint32_t hash, const jtk_slot &s, EmptyFn empty) {
	uint32_t b = hash & mask;
	for (int probe = 1;; probe++) {
		if (empty(tab[2 * b])) {
			tab[2 * b] = s;
			return probe;
		}
		if (empty(tab[2 * b + 1])) {
			tab[2 * b + 1] = s;
			return probe;
		}
		b = (b + 1) & mask;
	}
}

static void pack_inline_key(const uint8_t *p, int n, uint32_t w[6]) {
	uint8_t buf[24];
	memset(buf, 0, sizeof(buf));
	memcpy(buf, p, (size_t) n);
	memcpy(w, buf, 24);
}

int jtk_build_host_tables(const jtk_params *p, jtk_host_tables *t, std::string *err) {
	if (!p || !p->pattern) {
		*err = "params / pattern is null";
		return JTK_E_ARG;
	}
	if ((p->vocab_size > 0 && (!p->vocab_bytes || !p->vocab_off || !p->vocab_ranks)) || p->vocab_size < 0 || p->special_size < 0 ||
	    (p->special_size > 0 && (!p->special_off || !p->special_ids))) {
		*err = "vocabulary / special token arrays are inconsistent";
		return JTK_E_ARG;
	}
	t->name = p->name ? p->name : "";

	/* ---- the split pattern: recognised patterns are compiled to a class table + rule kind ---- */
	if (!strcmp(p->pattern, X50K_PATTERN)) t->pattern_kind = JTK_PAT_X50K;
	else if (!strcmp(p->pattern, CL100K_PATTERN)) t->pattern_kind = JTK_PAT_CL100K;
	else t->pattern_kind = JTK_PAT_GENERAL;
	if (p->pattern_flags & JTK_RE_CASE_INSENSITIVE)
		t->pattern_kind = JTK_PAT_GENERAL; /* the rule kinds restate the patterns as EncodingFactory compiles them (case-sensitive) */
	jtk_rx_compiled general_prog;
	if (t->pattern_kind == JTK_PAT_GENERAL) {
		/* any other pattern: compiled to a backtracking program (jtk_regex.h); constructs outside the subset fail here */
		jtk_rx_compiled &prog = general_prog;
		std::string rerr;
		int rc = jtk_rx_compile(p->pattern, p->pattern_flags, &prog, &rerr);
		if (rc != JTK_OK) {
			*err = std::string("split pattern is not supported by the device pattern compiler: ") + rerr + ": " + p->pattern;
			return rc;
		}
		if (prog.nullable)
			for (int64_t i = 0; i < p->vocab_size; i++)
				if (p->vocab_off[i + 1] == p->vocab_off[i]) {
					*err = "a pattern that can match the empty string together with an empty vocabulary key is not supported";
					return JTK_E_PATTERN_UNSUPPORTED;
				}
		t->rx_inst.resize(prog.inst.size() * sizeof(jtk_rx_inst));
		memcpy(t->rx_inst.data(), prog.inst.data(), t->rx_inst.size());
		t->rx_sets.resize(prog.sets.size() * sizeof(jtk_rx_set));
		if (!prog.sets.empty()) memcpy(t->rx_sets.data(), prog.sets.data(), t->rx_sets.size());
		t->rx_ranges = prog.ranges;
		t->rx_ninst = (int32_t) prog.inst.size();
		memcpy(t->rx_first, prog.first, sizeof(t->rx_first));
	}
	/* the general program reads the class table for \p{L}, \p{N} and the Unicode \s only: always White_Space there */
	build_class_tables(t, t->pattern_kind == JTK_PAT_GENERAL ? (p->pattern_flags | JTK_RE_UNICODE_CHARACTER_CLASS) : p->pattern_flags);
	if (t->pattern_kind == JTK_PAT_GENERAL) {
		/* the program determinised (leftmost-first DFA over code-point classes) where the pattern allows it; JTK_RX_DFA=0 keeps the
		 * backtracking program (development switch: both give the same pieces) */
		const char *sw = getenv("JTK_RX_DFA");
		jtk_rx_dfa_host dfa;
		jtk_tables v;
		memset(&v, 0, sizeof(v));
		v.cp_stage1 = t->cp_stage1.data();
		v.cp_stage2 = t->cp_stage2.data();
		if (sw && sw[0] == '0') {
			t->rx_dfa_why = "switched off (JTK_RX_DFA=0)";
		} else if (jtk_rx_build_dfa(general_prog, v, &dfa, &t->rx_dfa_why)) {
			t->rx_dfa_trans = std::move(dfa.trans);
			t->rx_dfa_stage1 = std::move(dfa.stage1);
			t->rx_dfa_stage2 = std::move(dfa.stage2);
			t->rx_dfa_stay = std::move(dfa.stay);
			t->rx_dfa_nsym = dfa.nsym;
			t->rx_dfa_nstates = dfa.nstates;
			t->rx_dfa_start = dfa.start;
			t->rx_dfa_start_bol = dfa.start_bol;
			t->rx_dfa_acc_lo = dfa.acc_lo;
		}
	}

	/* ---- vocabulary: Map.put semantics (a later duplicate key replaces the value), TokenEncoder.java:41-44 ---- */
	std::unordered_map<std::string, int32_t> index_of; /* key bytes -> token index */
	t->tok_bytes.clear();
	t->tok_off.assign(1, 0);
	t->tok_rank.clear();
	for (int64_t i = 0; i < p->vocab_size; i++) {
		int64_t a = p->vocab_off[i], b = p->vocab_off[i + 1];
		if (b < a) {
			*err = "vocab_off is not monotone";
			return JTK_E_ARG;
		}
		int32_t rank = p->vocab_ranks[i];
		if (rank >= JTK_RANK_MAX - 1 || rank < -(1 << 30)) {
			*err = "token ids must lie in [-2^30, Integer.MAX_VALUE - 2] (the rest of the int range encodes device-side records)";
			return JTK_E_ARG;
		}
		std::string key((const char *) p->vocab_bytes + a, (size_t) (b - a));
		auto it = index_of.find(key);
		if (it != index_of.end()) {
			t->tok_rank[(size_t) it->second] = rank;
			continue;
		}
		index_of.emplace(key, (int32_t) t->tok_rank.size());
		t->tok_bytes.insert(t->tok_bytes.end(), key.begin(), key.end());
		t->tok_off.push_back((uint32_t) t->tok_bytes.size());
		t->tok_rank.push_back(rank);
		if ((int32_t) key.size() > t->max_token_len) t->max_token_len = (int32_t) key.size();
	}
	const int64_t ntok = (int64_t) t->tok_rank.size();
	t->n_tokens = ntok;
	{
		/* The device merge loop identifies a part by its token id (pair table keyed on (id left, id right), the id of a merged part is
		 * the rank just found), so two different byte sequences must not share an id.  The reference looks ranks up by bytes
		 * (GptBytePairEncoding.getRank :285-300) and tolerates such a vocabulary, but cannot decode it unambiguously either
		 * (TokenEncoder.java:41-44: the last key put for a value wins); none of the predefined vocabularies has one. */
		std::unordered_map<int32_t, int64_t> first_with_rank;
		first_with_rank.reserve((size_t) ntok * 2);
		for (int64_t k = 0; k < ntok; k++) {
			auto ins = first_with_rank.emplace(t->tok_rank[(size_t) k], k);
			if (!ins.second) {
				char buf[160];
				snprintf(buf, sizeof(buf), "token id %d is assigned to two different byte sequences (vocabulary entries %lld and %lld): ids must be unique per byte sequence",
				         (int) t->tok_rank[(size_t) k], (long long) ins.first->second, (long long) k);
				*err = buf;
				return JTK_E_ARG;
			}
		}
	}
	if (t->tok_bytes.empty()) t->tok_bytes.push_back(0);

	t->byte_id.resize(256);
	for (int b = 0; b < 256; b++) t->byte_id[(size_t) b] = JTK_PSEUDO_BASE + b;
	t->bytepair.assign(65536, JTK_RANK_MAX);

	/* byte bigrams that occur inside some token (see jtk_tables::bigram_bits) */
	t->bigram_bits.assign(2048, 0);
	t->trigram_bits.assign(32768, 0);
	for (int64_t k = 0; k < ntok; k++) {
		const uint8_t *kb = t->tok_bytes.data() + t->tok_off[(size_t) k];
		const uint32_t len = t->tok_off[(size_t) k + 1] - t->tok_off[(size_t) k];
		for (uint32_t i = 0; i + 1 < len; i++) {
			const uint32_t g = (uint32_t) kb[i] << 8 | kb[i + 1];
			t->bigram_bits[g >> 5] |= 1u << (g & 31);
		}
		for (uint32_t i = 0; i + 2 < len; i++) {
			const uint32_t g = jtk_trigram_slot(kb[i], kb[i + 1], kb[i + 2]);
			t->trigram_bits[g >> 5] |= 1u << (g & 31);
		}
	}

	int64_t n_a = 0, n_b = 0;
	for (int64_t k = 0; k < ntok; k++) {
		uint32_t len = t->tok_off[(size_t) k + 1] - t->tok_off[(size_t) k];
		const uint8_t *kb = t->tok_bytes.data() + t->tok_off[(size_t) k];
		if (len == 1) t->byte_id[kb[0]] = t->tok_rank[(size_t) k];
		if (len == 2) t->bytepair[(size_t) kb[0] << 8 | kb[1]] = t->tok_rank[(size_t) k];
		if (len >= 1 && len <= JTK_INLINE_KEY_MAX) n_a++;
		else if (len > JTK_INLINE_KEY_MAX) n_b++;
	}

	/* table A: inline keys, one 32-byte slot per probe, load <= 0.5 */
	t->mask_a = pow2_at_least((uint64_t) (n_a / JTK_TABLE_A_LOAD) + 1) - 1;
	t->tab_a.assign((size_t) t->mask_a + 1, jtk_slot_a{{0, 0}, 0, 0, {0, 0, 0, 0}});
	/* table B: hashed long keys, w = token index + 1 (0 = empty) */
	t->mask_b = pow2_at_least((uint64_t) (n_b / 0.8) + 1) - 1;
	t->tab_b.assign(2 * (size_t) (t->mask_b + 1), jtk_slot{0, 0, 0, 0});
	t->long_filter.assign(2048, 0);
	for (int64_t k = 0; k < ntok; k++) {
		uint32_t len = t->tok_off[(size_t) k + 1] - t->tok_off[(size_t) k];
		const uint8_t *kb = t->tok_bytes.data() + t->tok_off[(size_t) k];
		if (len == 0) continue; /* an empty key can only match an empty piece, which emits nothing on this path */
		if (len <= JTK_INLINE_KEY_MAX) {
			jtk_slot_a s;
			uint32_t kw[6];
			pack_inline_key(kb, (int) len, kw);
			s.k01[0] = kw[0], s.k01[1] = kw[1];
			s.k25[0] = kw[2], s.k25[1] = kw[3], s.k25[2] = kw[4], s.k25[3] = kw[5];
			s.len = len;
			s.rank = (uint32_t) t->tok_rank[(size_t) k];
			uint32_t b = jtk_hash6(kw, len) & t->mask_a;
			int pr = 1;
			while (t->tab_a[b].len != 0) {
				b = (b + 1) & t->mask_a;
				pr++;
			}
			t->tab_a[b] = s;
			t->max_probe_a = std::max(t->max_probe_a, pr);
		} else {
			uint64_t h = jtk_hash_bytes_init();
			for (uint32_t i = 0; i < len; i++) h = jtk_hash_bytes_step(h, kb[i]);
			h = jtk_hash_bytes_final(h, len);
			{
				uint32_t w[6];
				pack_inline_key(kb, 8, w);
				const uint32_t f = jtk_hash3(w[0], w[1], len) & 0xFFFFu;
				t->long_filter[f >> 5] |= 1u << (f & 31);
			}
			jtk_slot s{(uint32_t) h, (uint32_t) (h >> 32), (uint32_t) t->tok_rank[(size_t) k], (uint32_t) k + 1};
			int pr = bucket_insert(t->tab_b, t->mask_b, (uint32_t) h, s, [](const jtk_slot &x) { return x.w == 0; });
			t->max_probe_b = std::max(t->max_probe_b, pr);
		}
	}

	/* pair table: every split of every token into two parts (a part is a single byte or a token) */
	std::vector<jtk_slot> entries;
	for (int64_t k = 0; k < ntok; k++) {
		uint32_t len = t->tok_off[(size_t) k + 1] - t->tok_off[(size_t) k];
		if (len < 2) continue;
		const char *kb = (const char *) t->tok_bytes.data() + t->tok_off[(size_t) k];
		for (uint32_t cut = 1; cut < len; cut++) {
			int32_t idl, idr;
			if (cut == 1) idl = t->byte_id[(uint8_t) kb[0]];
			else {
				auto it = index_of.find(std::string(kb, cut));
				if (it == index_of.end()) continue;
				idl = t->tok_rank[(size_t) it->second];
			}
			if (len - cut == 1) idr = t->byte_id[(uint8_t) kb[len - 1]];
			else {
				auto it = index_of.find(std::string(kb + cut, len - cut));
				if (it == index_of.end()) continue;
				idr = t->tok_rank[(size_t) it->second];
			}
			entries.push_back(jtk_slot{(uint32_t) idl, (uint32_t) idr, (uint32_t) t->tok_rank[(size_t) k], 1});
		}
	}
	t->n_pairs = (int64_t) entries.size();
	t->mask_p = pow2_at_least((uint64_t) (entries.size() / 0.8) + 1) - 1;
	t->pair.assign(2 * (size_t) (t->mask_p + 1), jtk_slot{0, 0, 0, 0});
	for (const jtk_slot &e : entries) {
		int pr = bucket_insert(t->pair, t->mask_p, jtk_hash_pair((int32_t) e.x, (int32_t) e.y), e, [](const jtk_slot &x) { return x.w == 0; });
		t->max_probe_p = std::max(t->max_probe_p, pr);
	}

	/* ---- special tokens ---- */
	t->nspecial = (int32_t) p->special_size;
	t->special_bytes.clear();
	t->special_off.assign(1, 0);
	t->special_ids.clear();
	for (int64_t i = 0; i < p->special_size; i++) {
		int64_t a = p->special_off[i], b = p->special_off[i + 1];
		if (b < a) {
			*err = "special_off is not monotone";
			return JTK_E_ARG;
		}
		if (a == b) t->special_has_empty = 1;
		else t->special_first[p->special_bytes[a] >> 5] |= 1u << (p->special_bytes[a] & 31);
		t->special_bytes.insert(t->special_bytes.end(), p->special_bytes + a, p->special_bytes + b);
		t->special_off.push_back((uint32_t) t->special_bytes.size());
		t->special_ids.push_back(p->special_ids[i]);
	}
	if (t->special_bytes.empty()) t->special_bytes.push_back(0);

	/* ---- decode table: ordinary map first, then the special-token map (GptBytePairEncoding.java:302-314) ---- */
	t->dec_bytes.assign(t->tok_bytes.begin(), t->tok_bytes.begin() + t->tok_off.back());
	t->dec_off = t->tok_off;
	const int64_t ndec_ord = ntok;
	for (int64_t i = 0; i < p->special_size; i++) {
		t->dec_bytes.insert(t->dec_bytes.end(), t->special_bytes.begin() + t->special_off[(size_t) i], t->special_bytes.begin() + t->special_off[(size_t) i + 1]);
		t->dec_off.push_back((uint32_t) t->dec_bytes.size());
	}
	if (t->dec_bytes.empty()) t->dec_bytes.push_back(0);
	const int64_t ndec = ndec_ord + p->special_size;
	t->mask_d = pow2_at_least((uint64_t) (ndec / 0.5) + 1) - 1;
	t->dec_keys.assign(2 * (size_t) (t->mask_d + 1), 0);
	auto dec_put = [&](int32_t id, uint32_t index, bool replace) {
		uint32_t s = jtk_hash_pair(id, 0) & t->mask_d;
		for (;;) {
			if (t->dec_keys[2 * s + 1] == 0) {
				t->dec_keys[2 * s] = (uint32_t) id;
				t->dec_keys[2 * s + 1] = index + 1;
				return;
			}
			if (t->dec_keys[2 * s] == (uint32_t) id) {
				if (replace) t->dec_keys[2 * s + 1] = index + 1; /* encodedToDecoded.put: last key for a value wins */
				return;
			}
			s = (s + 1) & t->mask_d;
		}
	};
	for (int64_t k = 0; k < ntok; k++) dec_put(t->tok_rank[(size_t) k], (uint32_t) k, true);
	{
		/* special ids only decode through the special map when the ordinary map misses; among specials the last put wins */
		std::unordered_map<int32_t, uint32_t> last;
		for (int64_t i = 0; i < p->special_size; i++) last[t->special_ids[(size_t) i]] = (uint32_t) (ndec_ord + i);
		std::unordered_map<int32_t, bool> ordinary;
		for (int64_t k = 0; k < ntok; k++) ordinary[t->tok_rank[(size_t) k]] = true;
		for (auto &kv : last)
			if (!ordinary.count(kv.first)) dec_put(kv.first, kv.second, true);
	}
	/* direct form of the same map for dense id spaces (every predefined encoding): id -> (offset, length) */
	{
		int64_t max_id = -1;
		bool neg = false;
		for (int64_t k = 0; k < ntok; k++) {
			max_id = std::max<int64_t>(max_id, t->tok_rank[(size_t) k]);
			neg |= t->tok_rank[(size_t) k] < 0;
		}
		for (int64_t i = 0; i < p->special_size; i++) {
			max_id = std::max<int64_t>(max_id, t->special_ids[(size_t) i]);
			neg |= t->special_ids[(size_t) i] < 0;
		}
		t->dec_direct.clear();
		if (!neg && max_id >= 0 && max_id < (2 << 20) && t->dec_bytes.size() < (1u << 24) && t->max_token_len < 255) {
			t->dec_direct.assign(4 * (size_t) (max_id + 1), 0xFFFFFFFFu);
			auto put = [&](int32_t id, int64_t index) { /* {offset << 8 | length, first twelve bytes} */
				const uint32_t o = t->dec_off[(size_t) index], l = t->dec_off[(size_t) index + 1] - o;
				uint32_t f[3] = {0, 0, 0};
				for (uint32_t k = 0; k < 12 && k < l; k++) f[k >> 2] |= (uint32_t) t->dec_bytes[o + k] << (8 * (k & 3));
				t->dec_direct[4 * (size_t) id] = (o << 8) | l;
				t->dec_direct[4 * (size_t) id + 1] = f[0];
				t->dec_direct[4 * (size_t) id + 2] = f[1];
				t->dec_direct[4 * (size_t) id + 3] = f[2];
			};
			/* the same precedence as the hash form: among ordinary tokens the last key put for an id wins, special tokens only where no ordinary token has the id */
			for (int64_t i = 0; i < p->special_size; i++) put(t->special_ids[(size_t) i], ndec_ord + i);
			for (int64_t k = 0; k < ntok; k++) put(t->tok_rank[(size_t) k], k);
		}
	}
	return JTK_OK;
}

jtk_tables jtk_host_view(const jtk_host_tables &h) {
	jtk_tables v;
	memset(&v, 0, sizeof(v));
	v.pattern_kind = h.pattern_kind;
	v.max_token_len = h.max_token_len;
	v.ascii_cls = h.ascii_cls.data();
	v.cp_stage1 = h.cp_stage1.data();
	v.cp_stage2 = h.cp_stage2.data();
	v.lut_sp = h.lut_sp.data();
	v.cls2 = h.cls2.data();
	v.bmp_nib = h.bmp_nib.data();
	v.tab_a = h.tab_a.data();
	v.mask_a = h.mask_a;
	v.tab_b = h.tab_b.data();
	v.mask_b = h.mask_b;
	v.long_filter = h.long_filter.data();
	v.tok_bytes = h.tok_bytes.data();
	v.tok_off = h.tok_off.data();
	v.byte_id = h.byte_id.data();
	v.bytepair = h.bytepair.data();
	v.pair = h.pair.data();
	v.mask_p = h.mask_p;
	v.bigram_bits = h.bigram_bits.data();
	v.trigram_bits = h.trigram_bits.data();
	v.nspecial = h.nspecial;
	v.special_has_empty = h.special_has_empty;
	v.special_bytes = h.special_bytes.data();
	v.special_off = h.special_off.data();
	v.special_ids = h.special_ids.data();
	memcpy(v.special_first, h.special_first, sizeof(v.special_first));
	v.dec_keys = h.dec_keys.data();
	v.mask_d = h.mask_d;
	v.dec_bytes = h.dec_bytes.data();
	v.dec_off = h.dec_off.data();
	v.dec_direct = h.dec_direct.empty() ? nullptr : reinterpret_cast<const uint4 *>(h.dec_direct.data());
	v.dec_direct_size = (uint32_t) (h.dec_direct.size() / 4);
	v.rx_inst = h.rx_inst.data();
	v.rx_sets = h.rx_sets.data();
	v.rx_ranges = h.rx_ranges.data();
	v.rx_ninst = h.rx_ninst;
	memcpy(v.rx_first, h.rx_first, sizeof(v.rx_first));
	v.rx_dfa_trans = h.rx_dfa_trans.empty() ? nullptr : h.rx_dfa_trans.data();
	v.rx_dfa_stage1 = h.rx_dfa_stage1.data();
	v.rx_dfa_stage2 = h.rx_dfa_stage2.data();
	v.rx_dfa_stay = h.rx_dfa_stay.empty() ? nullptr : h.rx_dfa_stay.data();
	v.rx_dfa_nsym = h.rx_dfa_nsym;
	v.rx_dfa_nstates = h.rx_dfa_nstates;
	v.rx_dfa_start = h.rx_dfa_start;
	v.rx_dfa_start_bol = h.rx_dfa_start_bol;
	v.rx_dfa_acc_lo = h.rx_dfa_acc_lo;
	return v;
}
