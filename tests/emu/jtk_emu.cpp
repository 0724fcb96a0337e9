/*
 * TEST INFRASTRUCTURE ONLY: a host-side emulator of the tile kernel's per-tile logic.
 * It drives the very same __host__ __device__ building blocks (jtk_device.cuh) that the CUDA kernel
 * uses - region set-up, classification, split rules, table lookups, thread-level merge - tile by tile
 * on the CPU, so the split rules and the halo / carry logic can be fuzzed against the oracle without a
 * GPU.  It is compiled with tiny tiles (see tests/emu/Makefile) to hit tile edges constantly.
 * The product never links this file.
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../jtokkit_b200/csrc/jtk_device.cuh"
#include "../../jtokkit_b200/csrc/jtk_regex.h"
#include "../../jtokkit_b200/csrc/jtk_tables.h"

struct emu_encoding {
	jtk_host_tables host;
	jtk_tables view;
};

extern "C" {

emu_encoding *emu_create(const jtk_params *p, char *err, int errlen) {
	emu_encoding *e = new emu_encoding();
	std::string msg;
	int rc = jtk_build_host_tables(p, &e->host, &msg);
	if (rc != JTK_OK) {
		snprintf(err, (size_t) errlen, "%d: %s", rc, msg.c_str());
		delete e;
		return nullptr;
	}
	e->view = jtk_host_view(e->host);
	return e;
}

void emu_destroy(emu_encoding *e) { delete e; }

int emu_tile() { return JTK_TILE; }

/* stats for tests: table sizes and probe lengths */
void emu_stats(const emu_encoding *e, int64_t *out) {
	out[0] = e->host.n_tokens;
	out[1] = e->host.n_pairs;
	out[2] = e->host.max_probe_a;
	out[3] = e->host.max_probe_b;
	out[4] = e->host.max_probe_p;
	out[5] = (int64_t) e->host.cp_stage2.size() / 256;
	out[6] = e->host.max_token_len;
}

struct emu_tile_state {
	std::vector<uint32_t> sbw; /* word aligned staging buffer */
	std::vector<uint32_t> bmask, dmask, planes, cmask; /* cmask: safe-cut bits (jtk_cut_chunk), kept apart from the pattern's piece starts */
	std::vector<int32_t> tok, rk;
	jtk_tile_ctx c;
	emu_tile_state() : sbw((JTK_REGION + 32) / 4 + 1), bmask(JTK_MASK_WORDS), dmask(JTK_MASK_WORDS), cmask(JTK_MASK_WORDS), planes(4 * (JTK_REGION_CHUNKS + 2)), tok(JTK_TILE + JTK_FWD_HALO), rk(JTK_TILE + JTK_FWD_HALO) {
		c.sb = reinterpret_cast<uint8_t *>(sbw.data());
		c.bmask = bmask.data();
		c.dmask = dmask.data();
		c.planes = planes.data();
		c.tok = tok.data();
		c.rk = rk.data();
	}
	/* the kernel's phases P1-P3 for one tile */
	void run(int64_t tile) {
		c.g0 = tile * JTK_TILE - JTK_BACK_HALO;
		c.rs = 0;
		c.carry_n = 0;
		std::fill(bmask.begin(), bmask.end(), 0u);
		std::fill(dmask.begin(), dmask.end(), 0u);
		for (int ch = 0; ch <= JTK_REGION_CHUNKS; ch++) jtk_load_chunk(c, ch);
		int64_t first_doc = 0;
		while (first_doc <= c.ndocs && c.doc_off[first_doc] < c.g0) first_doc++;
		for (int t = 0; t < JTK_NT; t++) jtk_mark_docstarts(c, first_doc, t, JTK_NT);
		c.rs = jtk_region_first(c);
		for (int ch = 0; ch <= JTK_REGION_CHUNKS; ch++) jtk_classify_chunk(c, ch);
		c.carry_n = jtk_region_carry_n(c);
		std::fill(cmask.begin(), cmask.end(), 0u);
		for (int ch = JTK_BACK_HALO / 16; ch < JTK_REGION_CHUNKS; ch++) {
			const uint32_t rb = jtk_boundary_chunk(c, ch);
			((uint16_t *) bmask.data())[ch] = (uint16_t) rb;
			((uint16_t *) cmask.data())[ch] = (uint16_t) (jtk_cut_chunk(c, ch) & ~rb);
		}
	}
	bool bit(int r) const { return (bmask[(size_t) (r >> 5)] >> (r & 31)) & 1u; }
	bool cut(int r) const { return (cmask[(size_t) (r >> 5)] >> (r & 31)) & 1u; }
};

/*
 * Runs every tile.  Pass 1 computes the piece-start flags tile by tile (region set-up, classification,
 * carries, split rules - exactly the kernel's phases P1-P3) and checks that the flags a tile computes for its
 * forward halo agree with the owner tile's; pass 2 encodes every piece with the kernel's lookup and merge
 * functions.  piece_flags / ids / tok_off / status are nullable.  Returns the token count, or a negative
 * number when a halo view disagrees with the owner tile.
 */
int64_t emu_run(const emu_encoding *e, const uint8_t *bytes, int64_t total, const int64_t *doc_off, int64_t ndocs, uint32_t flags,
                uint8_t *piece_flags, int32_t *ids, int64_t *tok_off, int32_t *status) {
	std::vector<uint8_t> in((size_t) total + 64, 0); /* 16-byte aligned copy of the input, as the device buffer is */
	if (total) memcpy(in.data(), bytes, (size_t) total);
	std::vector<uint8_t> start((size_t) total + 1, 0), halo((size_t) total + 1, 2), cutat((size_t) total + 1, 0), hcut((size_t) total + 1, 2);
	const int64_t ntiles = (total + JTK_TILE - 1) / JTK_TILE;
	emu_tile_state t;
	t.c.total = total;
	t.c.gbytes = in.data();
	t.c.doc_off = doc_off;
	t.c.ndocs = ndocs;
	t.c.T = &e->view;
	t.c.lut_sp = e->view.lut_sp;
	t.c.cls2 = e->view.cls2;
	int64_t halo_mismatch = 0;
	for (int64_t tile = 0; tile < ntiles; tile++) {
		t.run(tile);
		for (int r = JTK_BACK_HALO; r < JTK_BACK_HALO + JTK_TILE + JTK_LONG_PIECE + 1; r++) {
			const int64_t g = t.c.g0 + r;
			if (g >= total) break;
			const uint8_t b = t.bit(r) ? 1 : 0, cb = t.cut(r) ? 1 : 0;
			if (r < JTK_BACK_HALO + JTK_TILE) {
				start[(size_t) g] = b;
				cutat[(size_t) g] = cb;
				if (halo[(size_t) g] != 2 && halo[(size_t) g] != b) halo_mismatch++;
				if (hcut[(size_t) g] != 2 && hcut[(size_t) g] != cb) halo_mismatch++; /* the cuts a tile sees in its halo are the owner's */
			} else if (halo[(size_t) g] == 2) {
				halo[(size_t) g] = b;
				hcut[(size_t) g] = cb;
			} else if (halo[(size_t) g] != b || hcut[(size_t) g] != cb) {
				halo_mismatch++;
			}
		}
	}
	if (halo_mismatch) return -1000000 - halo_mismatch;
	start[(size_t) total] = 1;
	if (piece_flags)
		for (int64_t g = 0; g < total; g++) piece_flags[g] = start[(size_t) g];

	int64_t out_pos = 0, next_doc = 0;
	t.c.g0 = 0;
	for (int64_t g = 0; g < total; g++) {
		if ((flags & JTK_CHECK_SPECIAL) && status) {
			uint8_t by = in[(size_t) g];
			if ((e->view.special_first[by >> 5] >> (by & 31)) & 1u) {
				int64_t hi = jtk_doc_ceil(t.c, g);
				if (jtk_special_at(e->view, in.data(), g, hi)) {
					int64_t d = 0;
					while (d + 1 <= ndocs && doc_off[d + 1] <= g) d++;
					status[d] |= JTK_DOC_HAS_SPECIAL;
				}
			}
		}
		while (tok_off && next_doc <= ndocs && doc_off[next_doc] <= g) tok_off[next_doc++] = out_pos;
		if (!start[(size_t) g] || !ids) continue;
		int64_t eg = g + 1;
		while (!start[(size_t) eg]) eg++;
		const int64_t n = eg - g;
		const uint8_t *p = in.data() + g;
		bool unknown = false;
		bool has_cut = false;
		for (int64_t i = 1; i < n; i++) has_cut |= cutat[(size_t) (g + i)] != 0;
		if (n == 1) {
			int32_t id = e->view.byte_id[p[0]];
			if (id < JTK_PSEUDO_BASE + 256) unknown = true;
			ids[out_pos++] = id;
		} else {
			/* like the split kernel: a piece with a safe cut inside is never looked up as a whole (it cannot be a token), its segments are
			 * merged one by one; a piece without cuts takes the whole-piece shortcut first (GptBytePairEncoding.java:81-83) */
			int32_t whole = has_cut ? JTK_RANK_MAX : jtk_lookup_piece(e->view, p, (int) n);
			if (has_cut && jtk_lookup_piece(e->view, p, (int) n) != JTK_RANK_MAX) return -2000000; /* the argument above must hold */
			if (whole != JTK_RANK_MAX) {
				ids[out_pos++] = whole;
			} else {
				std::vector<int32_t> t2((size_t) n), r2((size_t) n), nx((size_t) n + 1);
				int64_t sa = 0;
				for (int64_t i = 1; i <= n; i++) {
					if (i < n && !cutat[(size_t) (g + i)]) continue;
					const int64_t len = i - sa;
					int cnt;
					if (len <= JTK_SHORT_PIECE) cnt = jtk_merge_short(e->view, p + sa, (int) len, t2.data(), r2.data(), 1, &unknown);
					else cnt = jtk_merge_seq(e->view, p + sa, (int) len, t2.data(), r2.data(), nx.data(), &unknown);
					for (int k = 0; k < cnt; k++) ids[out_pos++] = t2[(size_t) k];
					sa = i;
				}
			}
		}
		if (unknown && status) {
			int64_t d = 0;
			while (d + 1 <= ndocs && doc_off[d + 1] <= g) d++;
			status[d] |= JTK_DOC_UNKNOWN_BYTES;
		}
	}
	while (tok_off && next_doc <= ndocs) tok_off[next_doc++] = out_pos;
	return out_pos;
}
}

/* General split patterns: runs the three passes of the sliced Matcher.find() (jtk_regex.h) the way the kernels do, with
 * JTK_RX_SLICE-byte slices (tests/emu/Makefile makes them tiny).  start[g] = 1 at piece starts, skip[g] = 1 where the piece
 * starting at g is a gap.  Returns 0, 1 on stack overflow, -1 if the encoding has no general program. */
extern "C" int emu_general_split(const emu_encoding *e, const uint8_t *bytes, const int64_t *doc_off, int64_t ndocs, uint8_t *start, uint8_t *skip, int stack_cap, int no_dfa) {
	if (e->view.pattern_kind != JTK_PAT_GENERAL) return -1;
	jtk_rx_program P = jtk_rx_program_of(e->view);
	if (no_dfa) P.dfa_trans = nullptr; /* the backtracking program even where the pattern has a DFA */
	const int64_t total = doc_off[ndocs];
	const int64_t nslices = (total + JTK_RX_SLICE - 1) / JTK_RX_SLICE, nwords = total / 32 + 2;
	std::vector<uint32_t> bits((size_t) (5 * nwords), 0);
	std::vector<int64_t> rec((size_t) (4 * nslices + ndocs + 1), 0);
	jtk_rx_split_buffers B;
	B.ms = bits.data();
	B.me = B.ms + nwords;
	B.s_ms = B.me + nwords;
	B.s_me = B.s_ms + nwords;
	B.s_from = B.s_me + nwords;
	B.exit_slice = rec.data();
	B.last_ms = B.exit_slice + nslices;
	B.last_me = B.last_ms + nslices;
	B.join = B.last_me + nslices;
	B.exit_doc = B.join + nslices;
	B.nwords = nwords;
	B.nslices = nslices;
	std::vector<jtk_rx_frame> st((size_t) stack_cap);
	auto bor = [](uint32_t *w, uint32_t m) { *w |= m; };
	int64_t bad = -1;
	for (int64_t s = 0; s < nslices; s++)
		jtk_rx_slice_pass(P, e->view, bytes, total, doc_off, ndocs, s, B, st.data(), stack_cap < 10 ? stack_cap : 10, &bad, bor); /* tiny stack: overflows are left to pass 2, as on the device */
	for (int64_t d = 0; d < ndocs; d++) {
		if (doc_off[d + 1] == doc_off[d]) continue;
		if (!jtk_rx_stitch_doc(P, e->view, bytes, total, doc_off, d, B, st.data(), stack_cap, bor)) return 1;
	}
	for (int64_t w = 0; w < nwords; w++) jtk_rx_finish_word(B, w, total);
	for (int64_t g = 0; g < total; g++) {
		start[g] = (B.ms[g >> 5] >> (g & 31)) & 1u;
		skip[g] = (B.me[g >> 5] >> (g & 31)) & 1u;
	}
	return 0;
}

extern "C" int emu_pattern_kind(const emu_encoding *e) { return e->view.pattern_kind; }

/* jtk_rx_decode_word (the flat DFA loop's decoder over a fetched word) against jtk_rx_decode (the byte-wise one) on pseudo-random
 * four-byte windows, every lead-byte class and every amount of text left; returns the number of disagreements */
extern "C" int64_t emu_decode_word_check(int64_t samples, uint64_t seed) {
	int64_t bad = 0;
	uint64_t x = seed * 2862933555777941757ull + 3037000493ull;
	for (int64_t i = 0; i < samples; i++) {
		x = x * 6364136223846793005ull + 1442695040888963407ull;
		uint8_t b[4];
		for (int k = 0; k < 4; k++) {
			const uint32_t r = (uint32_t) (x >> (8 + 12 * k)) & 0xFFFu;
			/* mostly continuation bytes after the first, all classes of lead byte in front */
			b[k] = (k > 0 && (r & 0x300u)) ? (uint8_t) (0x80u | (r & 0x3Fu)) : (uint8_t) r;
		}
		for (int avail = 1; avail <= 4; avail++) {
			int l0, l1;
			const uint32_t c0 = jtk_rx_decode(b, 0, avail, &l0);
			uint32_t w = 0;
			for (int k = 0; k < 4; k++) w |= (uint32_t) b[k] << (8 * k); /* (bytes beyond avail are garbage on purpose) */
			const uint32_t c1 = jtk_rx_decode_word(w, avail, &l1);
			if (c0 != c1 || l0 != l1) bad++;
		}
	}
	return bad;
}

/* DFA of a general pattern: states (0 = the pattern has no DFA form), symbols; why not, if not */
extern "C" int emu_dfa_info(const emu_encoding *e, int *nsym, char *why, int why_cap) {
	*nsym = e->view.rx_dfa_nsym;
	snprintf(why, (size_t) why_cap, "%s", e->host.rx_dfa_why.c_str());
	return e->view.rx_dfa_trans ? e->view.rx_dfa_nstates : 0;
}
