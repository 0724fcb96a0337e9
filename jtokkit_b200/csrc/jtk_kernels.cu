/*
 * sm_100a kernels of the JTokkit encode path.
 *
 * jtk_encode_tiles_kernel is the hot path: ONE pass over the input.  A persistent CTA takes 8 KiB tiles by
 * ticket, stages the tile plus halos in shared memory with 16-byte loads, classifies code points, evaluates
 * the split rules into a piece-start bitmask, resolves ~94 % of the pieces with one probe of the byte-keyed
 * piece table (L2 resident), merges the rest (thread per short piece, warp per medium piece) against the
 * pair table, and places the tile's ids with a decoupled look-back scan over per-tile token counts, so the
 * only HBM traffic is the input bytes in and the ids / document token offsets out.
 *
 * Reference code this replaces: GptBytePairEncoding.encodeOrdinaryInternal / bytePairMerge / getRank
 * (GptBytePairEncoding.java:71-103,200-300) and the special-token guard of encodeInternal (:52-56).
 */
#include "jtk_kernels.cuh"

#include <algorithm>

#include "jtk_device.cuh"

namespace {

constexpr int NT = JTK_NT;
constexpr int NWARPS = NT / 32;
constexpr int TOKN = JTK_TILE + JTK_FWD_HALO;
constexpr int QCAP = TOKN / 2;
constexpr int BH = JTK_BACK_HALO;

/* indices into the small shared "misc" array */
enum { M_TILE = 0, M_NSHORT, M_NMED, M_SHORT_NEXT, M_MED_NEXT, M_RS, M_CARRY, M_TOTAL, M_WSUM = 16 };

/* first set bit in [from, limit] of a bit array, or -1 */
__device__ __forceinline__ int next_bit(const uint32_t *bm, int from, int limit) {
	int w = from >> 5;
	uint32_t x = bm[w] & (0xFFFFFFFFu << (from & 31));
	for (;;) {
		if (x) {
			int b = (w << 5) + __ffs((int) x) - 1;
			return b <= limit ? b : -1;
		}
		w++;
		if ((w << 5) > limit) return -1;
		x = bm[w];
	}
}

/* document that contains global position g (the last document starting at or before g) */
__device__ int64_t doc_of(const jtk_encode_args &a, int64_t g) {
	int64_t lo = 0, hi = a.ndocs - 1;
	while (lo < hi) {
		int64_t mid = (lo + hi + 1) >> 1;
		if (a.doc_off[mid] <= g) lo = mid;
		else hi = mid - 1;
	}
	return lo;
}

__device__ void flag_doc(const jtk_encode_args &a, int64_t g, int bit) {
	if (a.doc_status && a.ndocs > 0) atomicOr(a.doc_status + doc_of(a, g), bit);
}

/* ---------------------------------------------------------------------------------------------
 * bytePairMerge for one piece of 33..JTK_LONG_PIECE bytes by one warp (GptBytePairEncoding.java:200-275).
 * Position k is owned by lane k & 31 (row k >> 5), so rank scans are bank-conflict free; the argmin over
 * adjacent pair ranks is a REDUX min over per-lane minima, neighbours are found with REDUX over per-lane
 * alive masks, and the two rank probes of a merge (:254-257) are issued by two lanes in parallel.
 * ------------------------------------------------------------------------------------------- */
__device__ int merge_warp(const jtk_tables &T, const uint8_t *p, int n, int32_t *tok, int32_t *rk, bool *unknown) {
	const unsigned full = 0xFFFFFFFFu;
	const int lane = threadIdx.x & 31;
	for (int k = lane; k < n; k += 32) {
		tok[k] = T.byte_id[p[k]];
		rk[k] = (k + 1 < n) ? T.bytepair[((uint32_t) p[k] << 8) | p[k + 1]] : JTK_RANK_MAX;
	}
	__syncwarp();
	const int rows = (n + 31) >> 5;
	uint32_t alive = 0;
	for (int j = 0; j < rows; j++)
		if ((j << 5) + lane < n) alive |= 1u << j;
	int32_t lr = JTK_RANK_MAX;
	int lk = 0x7fffffff;
	auto rescan = [&]() {
		lr = JTK_RANK_MAX;
		lk = 0x7fffffff;
		for (uint32_t m = alive; m;) {
			int j = __ffs((int) m) - 1;
			m &= m - 1;
			int k = (j << 5) + lane;
			int32_t r = rk[k];
			if (r < lr) {
				lr = r;
				lk = k;
			}
		}
	};
	rescan();
	for (;;) {
		const int32_t mr = __reduce_min_sync(full, lr);
		if (mr == JTK_RANK_MAX) break;
		const int mi = __reduce_min_sync(full, lr == mr ? lk : 0x7fffffff); /* leftmost minimum (:232-240) */
		/* next alive position after k / previous alive position before k */
		auto next_alive = [&](int k) {
			int j0 = (k >> 5) + (lane <= (k & 31) ? 1 : 0);
			uint32_t m = j0 >= 32 ? 0u : (alive & (0xFFFFFFFFu << j0));
			int cand = m ? (((__ffs((int) m) - 1) << 5) + lane) : 0x7fffffff;
			return __reduce_min_sync(full, cand);
		};
		auto prev_alive = [&](int k) {
			int j1 = (k >> 5) - (lane < (k & 31) ? 0 : 1);
			uint32_t m = j1 < 0 ? 0u : (alive & (j1 >= 31 ? 0xFFFFFFFFu : ((2u << j1) - 1u)));
			int cand = m ? (((31 - __clz((int) m)) << 5) + lane) : -1;
			return __reduce_max_sync(full, cand);
		};
		const int nx = next_alive(mi);
		const int nn = next_alive(nx);
		const int pv = prev_alive(mi);
		if (lane == (mi & 31)) tok[mi] = mr;
		if (lane == (nx & 31)) {
			alive &= ~(1u << (nx >> 5));
			rk[nx] = JTK_RANK_MAX;
		}
		__syncwarp();
		if (lane == 0) rk[mi] = (nn != 0x7fffffff) ? jtk_lookup_pair(T, mr, tok[nn]) : JTK_RANK_MAX;
		if (lane == 1 && pv >= 0) rk[pv] = jtk_lookup_pair(T, tok[pv], mr);
		__syncwarp();
		if (lane == (mi & 31) || lane == (nx & 31) || (pv >= 0 && lane == (pv & 31))) rescan();
	}
	/* compact the surviving parts to the front of tok[] */
	int out = 0;
	bool unk = false;
	for (int j = 0; j < rows; j++) {
		const bool a = (alive >> j) & 1u;
		const int32_t v = a ? tok[(j << 5) + lane] : 0;
		const unsigned b = __ballot_sync(full, a);
		__syncwarp();
		if (a) {
			tok[out + __popc(b & ((1u << lane) - 1u))] = v;
			if (v < JTK_PSEUDO_BASE + 256) unk = true;
		}
		out += __popc(b);
		__syncwarp();
	}
	if (__any_sync(full, unk)) *unknown = true;
	return out;
}

/* ---------------------------------------------------------------------------------------------
 * the tile kernel
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(JTK_NT, 2) jtk_encode_tiles_kernel(const __grid_constant__ jtk_encode_args a) {
	extern __shared__ __align__(16) uint8_t smem[];
	uint8_t *sb = smem;
	uint8_t *cls = sb + (JTK_REGION + 16);
	uint32_t *bmask = reinterpret_cast<uint32_t *>(cls + (JTK_REGION + 16));
	uint32_t *dmask = bmask + JTK_MASK_WORDS;
	int32_t *tok = reinterpret_cast<int32_t *>(dmask + JTK_MASK_WORDS);
	int32_t *rk = tok + TOKN;
	uint16_t *slowq = reinterpret_cast<uint16_t *>(rk + TOKN);
	uint32_t *chunk_pref = reinterpret_cast<uint32_t *>(slowq + QCAP);
	uint32_t *misc = chunk_pref + NT;

	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;
	const jtk_tables &T = a.T;

	jtk_tile_ctx c;
	c.sb = sb;
	c.cls = cls;
	c.bmask = bmask;
	c.dmask = dmask;
	c.tok = tok;
	c.rk = rk;
	c.total = a.total;
	c.gbytes = a.bytes;
	c.doc_off = a.doc_off;
	c.ndocs = a.ndocs;
	c.T = &T;

	for (;;) {
		/* ---- P0: take a tile ---- */
		__syncthreads(); /* previous iteration's readers of misc / staging are done */
		if (tid == 0) {
			misc[M_TILE] = atomicAdd(&a.hdr->ticket, 1u);
			misc[M_NSHORT] = 0;
			misc[M_NMED] = 0;
			misc[M_SHORT_NEXT] = 0;
			misc[M_MED_NEXT] = 0;
		}
		__syncthreads();
		const long long tile = misc[M_TILE];
		if (tile >= a.ntiles) break;
		const int64_t tb = tile * (int64_t) JTK_TILE;
		c.g0 = tb - BH;
		c.rs = 0;
		c.carry_n = 0;

		/* ---- P1: stage bytes (16-byte loads), clear masks, mark document starts ---- */
		for (int w = tid; w < JTK_MASK_WORDS; w += NT) {
			bmask[w] = 0;
			dmask[w] = 0;
		}
		for (int ch = tid; ch <= JTK_REGION_CHUNKS; ch += NT) jtk_load_chunk(c, ch);
		__syncthreads();
		const int64_t first_doc = a.tile_first_doc[tile];
		jtk_mark_docstarts(c, first_doc, tid, NT);
		__syncthreads();
		if (tid == 0) misc[M_RS] = (uint32_t) jtk_region_first(c);
		__syncthreads();
		c.rs = (int) misc[M_RS];

		/* ---- P2: code point classes ---- */
		for (int ch = tid; ch <= JTK_REGION_CHUNKS; ch += NT) jtk_classify_chunk(c, ch);
		__syncthreads();
		if (tid == 0) misc[M_CARRY] = (uint32_t) jtk_region_carry_n(c);
		__syncthreads();
		c.carry_n = (int) misc[M_CARRY];

		/* ---- P3: split rules -> piece-start bits; special-token guard ---- */
		for (int ch = BH / 16 + tid; ch < JTK_REGION_CHUNKS; ch += NT) reinterpret_cast<uint16_t *>(bmask)[ch] = (uint16_t) jtk_boundary_chunk(c, ch);
		if ((a.flags & JTK_CHECK_SPECIAL) && T.nspecial > 0) {
			const int r0 = BH + tid * 16;
			for (int i = 0; i < 16; i++) {
				const int64_t g = c.g0 + r0 + i;
				if (g >= a.total) break;
				const uint8_t b = sb[r0 + i];
				if ((T.special_first[b >> 5] >> (b & 31)) & 1u) {
					if (jtk_special_at(T, a.bytes, g, jtk_doc_ceil(c, g))) flag_doc(a, g, JTK_DOC_HAS_SPECIAL);
				}
			}
		}
		__syncthreads();

		/* ---- P4: whole-piece lookup for the pieces that start in this thread's 16 bytes ---- */
		const int r0 = BH + tid * 16;
		const int64_t gbase = tb + tid * 16;
		uint32_t mybits = reinterpret_cast<const uint16_t *>(bmask)[BH / 16 + tid];
		if (gbase >= a.total) mybits = 0;
		else if (gbase + 16 > a.total) mybits &= (1u << (int) (a.total - gbase)) - 1u;
		if (a.piece_flags) {
			for (int i = 0; i < 16 && gbase + i < a.total; i++) a.piece_flags[gbase + i] = (mybits >> i) & 1u;
		}
		for (uint32_t m = mybits; m;) {
			const int i = __ffs((int) m) - 1;
			m &= m - 1;
			const int r = r0 + i;
			const int s = r - BH;
			const int e = next_bit(bmask, r + 1, r + JTK_LONG_PIECE);
			if (e < 0) { /* longer than JTK_LONG_PIECE: deferred to the long-piece kernels */
				rk[s] = -1;
				continue;
			}
			const int n = e - r;
			const uint8_t *p = sb + r;
			if (n == 1) {
				const int32_t id = T.byte_id[p[0]];
				if (id < JTK_PSEUDO_BASE + 256) flag_doc(a, gbase + i, JTK_DOC_UNKNOWN_BYTES);
				tok[s] = id;
				rk[s] = 1;
				continue;
			}
			const int32_t id = jtk_lookup_piece(T, p, n);
			if (id != JTK_RANK_MAX) {
				tok[s] = id;
				rk[s] = 1;
			} else {
				rk[s] = n;
				if (n <= JTK_SHORT_PIECE) slowq[atomicAdd(&misc[M_NSHORT], 1u)] = (uint16_t) s;
				else slowq[QCAP - 1 - atomicAdd(&misc[M_NMED], 1u)] = (uint16_t) s;
			}
		}
		__syncthreads();

		/* ---- P5: merge loops.  Medium pieces one per warp, short pieces one per thread, both handed out dynamically ---- */
		{
			const unsigned nmed = misc[M_NMED], nshort = misc[M_NSHORT];
			for (;;) {
				unsigned j = 0;
				if (lane == 0) j = atomicAdd(&misc[M_MED_NEXT], 1u);
				j = __shfl_sync(0xFFFFFFFFu, j, 0);
				if (j >= nmed) break;
				const int s = slowq[QCAP - 1 - j];
				const int n = rk[s];
				__syncwarp();
				bool unk = false;
				const int cnt = merge_warp(T, sb + BH + s, n, tok + s, rk + s, &unk);
				__syncwarp();
				if (lane == 0) {
					rk[s] = cnt;
					if (unk) flag_doc(a, tb + s, JTK_DOC_UNKNOWN_BYTES);
				}
			}
			for (;;) {
				unsigned base = 0;
				if (lane == 0) base = atomicAdd(&misc[M_SHORT_NEXT], 32u);
				base = __shfl_sync(0xFFFFFFFFu, base, 0);
				if (base >= nshort) break;
				const unsigned j = base + lane;
				if (j < nshort) {
					const int s = slowq[j];
					const int n = rk[s];
					bool unk = false;
					const int cnt = jtk_merge_short(T, sb + BH + s, n, tok + s, rk + s, &unk);
					rk[s] = cnt;
					if (unk) flag_doc(a, tb + s, JTK_DOC_UNKNOWN_BYTES);
				}
			}
		}
		__syncthreads();

		/* ---- P6: token counts -> block scan -> tile-local placement.  Tokens go to this tile's slice of the staging
		 * buffer; jtk_tile_scan_kernel + jtk_gather_kernel place them globally, so no CTA ever waits for another. ---- */
		int mycount = 0;
		for (uint32_t m = mybits; m;) {
			const int i = __ffs((int) m) - 1;
			m &= m - 1;
			const int cnt = rk[r0 + i - BH];
			mycount += cnt > 0 ? cnt : 0;
		}
		int x = mycount;
		for (int o = 1; o < 32; o <<= 1) {
			int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31) misc[M_WSUM + warp] = (uint32_t) x;
		__syncthreads();
		if (warp == 0) {
			int v = lane < NWARPS ? (int) misc[M_WSUM + lane] : 0;
			for (int o = 1; o < 32; o <<= 1) {
				int y = __shfl_up_sync(0xFFFFFFFFu, v, o);
				if (lane >= o) v += y;
			}
			if (lane < NWARPS) misc[M_WSUM + lane] = (uint32_t) v;
			if (lane == NWARPS - 1) a.tile_count[tile] = v;
			/* first piece start of the tile, for the long-piece bounds kernel */
			int fb = 0x7fffffff;
			for (int w = BH / 32 + lane; w < (BH + JTK_TILE) / 32; w += 32) {
				uint32_t bits = bmask[w];
				if (bits) fb = min(fb, (w << 5) + __ffs((int) bits) - 1);
			}
			fb = __reduce_min_sync(0xFFFFFFFFu, fb);
			if (lane == 0) a.tile_first_b[tile] = fb == 0x7fffffff ? -1 : c.g0 + fb;
		}
		__syncthreads();
		const int excl = x - mycount + (warp ? (int) misc[M_WSUM + warp - 1] : 0);
		chunk_pref[tid] = (uint32_t) excl;
		const bool write_ids = !(a.flags & JTK_COUNT_ONLY) && a.stage != nullptr;
		{
			int32_t *dst = a.stage + tile * (long long) TOKN;
			int pos = excl;
			for (uint32_t m = mybits; m;) {
				const int i = __ffs((int) m) - 1;
				m &= m - 1;
				const int s = r0 + i - BH;
				const int cnt = rk[s];
				if (cnt < 0) {
					const unsigned idx = atomicAdd(&a.hdr->n_long, 1u);
					if ((int64_t) idx < a.long_cap) {
						jtk_long_piece lp;
						lp.start = gbase + i;
						const int e = next_bit(bmask, r0 + i + 1, JTK_REGION - 1);
						lp.end = e < 0 ? -1 : c.g0 + e;
						lp.insert_at = pos; /* tile-local; jtk_long_bounds_kernel adds the tile's base */
						lp.count = 0;
						lp.scratch = 0;
						lp.doc = 0;
						lp.flags = 0;
						a.long_list[idx] = lp;
					}
					continue;
				}
				if (write_ids)
					for (int k = 0; k < cnt; k++) dst[pos + k] = tok[s + k];
				pos += cnt;
			}
		}
		__syncthreads(); /* chunk_pref complete */
		/* tile-local token offsets of the documents that start in this tile (the end of the input included);
		 * jtk_gather_kernel adds the tile's base */
		if (a.tok_off) {
			for (int64_t d = first_doc + tid; d <= a.ndocs; d += NT) {
				const int64_t g = a.doc_off[d];
				if (g >= tb + JTK_TILE) break;
				if (g < tb) continue;
				const int s = (int) (g - tb);
				const int ch = s >> 4;
				long long off = chunk_pref[ch];
				uint32_t bits = reinterpret_cast<const uint16_t *>(bmask)[BH / 16 + ch] & ((1u << (s & 15)) - 1u);
				for (; bits;) {
					const int i = __ffs((int) bits) - 1;
					bits &= bits - 1;
					const int cnt = rk[(ch << 4) + i];
					off += cnt > 0 ? cnt : 0;
				}
				a.tok_off[d] = off;
			}
		}
	}
}

/* first document whose start is >= the tile's region start */
__global__ void jtk_tile_first_doc_kernel(const int64_t *doc_off, int64_t ndocs, int64_t ntiles, int32_t *out) {
	const int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (t >= ntiles) return;
	const int64_t g0 = t * (int64_t) JTK_TILE - BH;
	int64_t lo = 0, hi = ndocs + 1; /* first d in [0, ndocs] with doc_off[d] >= g0, else ndocs + 1 */
	while (lo < hi) {
		int64_t mid = (lo + hi) >> 1;
		if (doc_off[mid] >= g0) hi = mid;
		else lo = mid + 1;
	}
	out[t] = (int32_t) lo;
}

/* exclusive scan of the per-tile token counts (one block; ntiles is ~131k per GiB) + batch totals */
__global__ void __launch_bounds__(1024, 1) jtk_tile_scan_kernel(const jtk_encode_args a) {
	__shared__ long long s_w[32];
	__shared__ long long s_carry;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_carry = 0;
	__syncthreads();
	for (int64_t base = 0; base < a.ntiles; base += 1024) {
		const int64_t i = base + tid;
		const long long val = i < a.ntiles ? a.tile_count[i] : 0;
		long long x = val;
		for (int o = 1; o < 32; o <<= 1) {
			long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31) s_w[warp] = x;
		__syncthreads();
		if (warp == 0) {
			long long w = s_w[lane];
			for (int o = 1; o < 32; o <<= 1) {
				long long y = __shfl_up_sync(0xFFFFFFFFu, w, o);
				if (lane >= o) w += y;
			}
			s_w[lane] = w;
		}
		__syncthreads();
		const long long excl = s_carry + x - val + (warp ? s_w[warp - 1] : 0);
		if (i < a.ntiles) a.tile_base[i] = excl;
		__syncthreads();
		if (tid == 1023) s_carry = excl + val;
		__syncthreads();
	}
	if (tid == 0) {
		a.tile_base[a.ntiles] = s_carry;
		a.hdr->total_tokens = (unsigned long long) s_carry;
		if (!(a.flags & JTK_COUNT_ONLY) && a.ids && s_carry > a.ids_cap) a.hdr->overflow = 1;
	}
}

/* staging -> final ids (one CTA per tile, coalesced both ways) and tile-local -> global document token offsets */
__global__ void __launch_bounds__(256) jtk_gather_kernel(const jtk_encode_args a) {
	const bool write_ids = !(a.flags & JTK_COUNT_ONLY) && a.ids != nullptr && a.stage != nullptr && !a.hdr->overflow;
	if (write_ids) {
		for (int64_t t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
			const int32_t *src = a.stage + t * (long long) TOKN;
			int32_t *dst = a.ids + a.tile_base[t];
			const int n = a.tile_count[t];
			for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
		}
	}
	if (a.tok_off) {
		const long long total_tokens = a.tile_base[a.ntiles];
		for (int64_t d = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; d <= a.ndocs; d += (int64_t) gridDim.x * blockDim.x) {
			const int64_t t = a.doc_off[d] / JTK_TILE;
			a.tok_off[d] = t >= a.ntiles ? total_tokens : a.tok_off[d] + a.tile_base[t];
		}
	}
}

/* =============================================================================================
 * long pieces (> JTK_LONG_PIECE bytes): exact bytePairMerge by rounds, one CTA per piece, state in global
 * scratch.  A round merges every pair whose rank equals the current global minimum (taking every second one
 * in a chain of adjacent equal-rank pairs, which is what leftmost-first does); this equals the sequential
 * loop as long as no new pair ranks below the round's minimum.  That is checked every round; on a
 * violation the piece restarts in strict mode (one merge per round = the reference loop verbatim).
 * ============================================================================================= */
constexpr int LNT = 1024;
constexpr int32_t DEAD = JTK_RANK_MAX; /* rk of a removed part */

__global__ void jtk_long_bounds_kernel(const jtk_encode_args a, unsigned int n_long) {
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_long) return;
	jtk_long_piece lp = a.long_list[i];
	a.long_list[i].insert_at = lp.insert_at + a.tile_base[lp.start / JTK_TILE];
	if (lp.end < 0) {
		int64_t end = a.total;
		for (int64_t t = lp.start / JTK_TILE + 1; t < a.ntiles; t++) {
			const int64_t fb = a.tile_first_b[t];
			if (fb >= 0) {
				end = fb;
				break;
			}
		}
		a.long_list[i].end = end;
	}
}

struct long_scan_state {
	int run;   /* number of trailing flagged parts (parity is what matters) */
	int reset; /* 1 when a live unflagged part occurs in the span */
};

__global__ void __launch_bounds__(LNT, 1)
    jtk_long_merge_kernel(const jtk_encode_args a, unsigned int n_long, int32_t *scr_tok, int32_t *scr_rk, int32_t *scr_nxt, int32_t *scr_prv) {
	__shared__ int32_t s_red[LNT / 32];
	__shared__ int s_scan_run[LNT / 32];
	__shared__ int s_scan_reset[LNT / 32];
	__shared__ int32_t s_min;
	__shared__ int s_flag, s_piece, s_carry_run, s_first;
	const jtk_tables &T = a.T;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (;;) {
		__syncthreads();
		if (tid == 0) s_piece = (int) atomicAdd(&a.hdr->long_next, 1u);
		__syncthreads();
		const unsigned pi = (unsigned) s_piece;
		if (pi >= n_long) break;
		const jtk_long_piece lp = a.long_list[pi];
		const int64_t n64 = lp.end - lp.start;
		const int n = (int) n64;
		const uint8_t *p = a.bytes + lp.start;
		int32_t *tok = scr_tok + lp.scratch, *rk = scr_rk + lp.scratch, *nxt = scr_nxt + lp.scratch, *prv = scr_prv + lp.scratch;
		bool strict = false;
	restart:
		for (int k = tid; k < n; k += LNT) {
			tok[k] = T.byte_id[p[k]];
			rk[k] = (k + 1 < n) ? T.bytepair[((uint32_t) p[k] << 8) | p[k + 1]] : JTK_RANK_MAX;
			nxt[k] = k + 1;
			prv[k] = k - 1;
		}
		__syncthreads();
		for (;;) {
			/* global minimum rank (and, in strict mode, its leftmost position) */
			int32_t mr = JTK_RANK_MAX;
			for (int k = tid; k < n; k += LNT) mr = min(mr, rk[k]);
			mr = __reduce_min_sync(0xFFFFFFFFu, mr);
			if (lane == 0) s_red[warp] = mr;
			__syncthreads();
			if (warp == 0) {
				int32_t v = s_red[lane];
				v = __reduce_min_sync(0xFFFFFFFFu, v);
				if (lane == 0) {
					s_min = v;
					s_flag = 0;
					s_carry_run = 0;
					s_first = 0x7fffffff;
				}
			}
			__syncthreads();
			mr = s_min;
			if (mr == JTK_RANK_MAX) break;
			if (strict) {
				int first = 0x7fffffff;
				for (int k = tid; k < n; k += LNT)
					if (rk[k] == mr) {
						first = k;
						break;
					}
				first = __reduce_min_sync(0xFFFFFFFFu, first);
				if (lane == 0) atomicMin(&s_first, first);
				__syncthreads();
				if (tid == 0) {
					const int k = s_first, j = nxt[k], nn = nxt[j];
					tok[k] = mr;
					rk[j] = DEAD;
					nxt[j] = -1;
					nxt[k] = nn;
					if (nn < n) prv[nn] = k;
					rk[k] = nn < n ? jtk_lookup_pair(T, mr, tok[nn]) : JTK_RANK_MAX;
					const int pv = prv[k];
					if (pv >= 0) rk[pv] = jtk_lookup_pair(T, tok[pv], mr);
				}
				__syncthreads();
				continue;
			}
			/* selection: in list order, a flagged part (rk == mr) is selected iff an even number of flagged parts
			 * directly precede it.  Blocked scan over positions, LNT * 8 positions per step, removed parts are
			 * transparent.  Selected parts get rk = -mark (SEL) so that the merge pass can find them. */
			for (int base = 0; base < n; base += LNT * 8) {
				const int k0 = base + tid * 8;
				int run = 0, reset = 0; /* summary of this thread's 8 positions */
				int loc_run[8] = {0, 0, 0, 0, 0, 0, 0, 0};
				/* removed parts are marked by nxt[k] = -1 and are transparent */
				for (int i = 0; i < 8; i++) {
					const int k = k0 + i;
					if (k >= n || nxt[k] < 0) continue;
					loc_run[i] = 0;
					if (rk[k] == mr) {
						loc_run[i] = run; /* flagged parts directly before it inside this thread's span */
						run++;
					} else {
						run = 0;
						reset = 1;
					}
				}
				/* exclusive scan of (run, reset) across threads */
				int xr = run, xs = reset;
				for (int o = 1; o < 32; o <<= 1) {
					int yr = __shfl_up_sync(0xFFFFFFFFu, xr, o), ys = __shfl_up_sync(0xFFFFFFFFu, xs, o);
					if (lane >= o) {
						if (!xs) xr += yr;
						xs |= ys;
					}
				}
				if (lane == 31) {
					s_scan_run[warp] = xr;
					s_scan_reset[warp] = xs;
				}
				__syncthreads();
				if (warp == 0) {
					int vr = s_scan_run[lane], vs = s_scan_reset[lane];
					for (int o = 1; o < 32; o <<= 1) {
						int yr = __shfl_up_sync(0xFFFFFFFFu, vr, o), ys = __shfl_up_sync(0xFFFFFFFFu, vs, o);
						if (lane >= o) {
							if (!vs) vr += yr;
							vs |= ys;
						}
					}
					s_scan_run[lane] = vr;
					s_scan_reset[lane] = vs;
				}
				__syncthreads();
				/* exclusive prefix for this thread: carry (block) . warps before . lanes before */
				int pr = s_carry_run, ps = 0;
				if (warp > 0) {
					if (s_scan_reset[warp - 1]) {
						pr = s_scan_run[warp - 1];
						ps = 1;
					} else pr += s_scan_run[warp - 1];
				}
				{
					int er = __shfl_up_sync(0xFFFFFFFFu, xr, 1), es = __shfl_up_sync(0xFFFFFFFFu, xs, 1);
					if (lane > 0) {
						if (es) {
							pr = er;
							ps = 1;
						} else pr += er;
					}
				}
				(void) ps;
				/* mark selections */
				bool seen_reset = false;
				for (int i = 0; i < 8; i++) {
					const int k = k0 + i;
					if (k >= n || nxt[k] < 0) continue;
					if (rk[k] == mr) {
						const int before = seen_reset ? loc_run[i] : loc_run[i] + pr;
						if ((before & 1) == 0) rk[k] = JTK_RANK_MAX - 1; /* SEL marker (never a real rank: ranks < MAX-1 enforced at registration) */
					} else {
						seen_reset = true;
					}
				}
				__syncthreads();
				if (tid == LNT - 1) {
					/* carry for the next step: inclusive result of the last thread */
					int cr = s_scan_run[LNT / 32 - 1];
					int cs = s_scan_reset[LNT / 32 - 1];
					s_carry_run = cs ? cr : s_carry_run + cr;
				}
				__syncthreads();
			}
			/* merge every selected pair */
			for (int k = tid; k < n; k += LNT) {
				if (rk[k] != JTK_RANK_MAX - 1) continue;
				const int j = nxt[k];
				const int nn = nxt[j];
				tok[k] = mr;
				rk[j] = DEAD;
				nxt[j] = -1;
				nxt[k] = nn;
				if (nn < n) prv[nn] = k;
			}
			__syncthreads();
			/* recompute the ranks around every merged part (:254-257) */
			for (int k = tid; k < n; k += LNT) {
				if (rk[k] != JTK_RANK_MAX - 1) continue;
				const int nn = nxt[k];
				const int32_t r1 = nn < n ? jtk_lookup_pair(T, mr, tok[nn]) : JTK_RANK_MAX;
				/* the right neighbour may itself be a merged part; both sides then compute the same value */
				int32_t r0 = JTK_RANK_MAX;
				const int pv = prv[k];
				if (pv >= 0) r0 = jtk_lookup_pair(T, tok[pv], mr);
				if (r1 < mr || r0 < mr) s_flag = 1; /* a new pair outranks the round: not equal to the sequential order */
				rk[k] = r1;
				if (pv >= 0 && rk[pv] != JTK_RANK_MAX - 1) rk[pv] = r0;
			}
			__syncthreads();
			if (s_flag) {
				strict = true;
				if (tid == 0) atomicAdd(&a.hdr->violations, 1u);
				__syncthreads();
				goto restart;
			}
		}
		/* compact the surviving parts: out[0..count) = tokens in order (written over tok[] front via rk[] as temp) */
		__syncthreads();
		int64_t count = 0;
		{
			/* sequential list walk by one thread would be O(n); do a blocked compaction instead */
			__shared__ int s_base;
			if (tid == 0) s_base = 0;
			__syncthreads();
			for (int base = 0; base < n; base += LNT) {
				const int k = base + tid;
				const bool live = k < n && nxt[k] >= 0;
				const int32_t v = live ? tok[k] : 0;
				const unsigned b = __ballot_sync(0xFFFFFFFFu, live);
				if (lane == 0) s_red[warp] = __popc(b);
				__syncthreads();
				int woff = 0;
				for (int w = 0; w < warp; w++) woff += s_red[w];
				int tot = 0;
				for (int w = 0; w < LNT / 32; w++) tot += s_red[w];
				const int dst = s_base + woff + __popc(b & ((1u << lane) - 1u));
				if (live) rk[dst] = v; /* dst <= k: rk of earlier positions is no longer needed */
				__syncthreads();
				if (tid == 0) s_base += tot;
				__syncthreads();
			}
			count = s_base;
		}
		bool unk = false;
		for (int k = tid; k < (int) count; k += LNT) {
			const int32_t v = rk[k];
			tok[k] = v;
			if (v < JTK_PSEUDO_BASE + 256) unk = true;
		}
		if (unk) flag_doc(a, lp.start, JTK_DOC_UNKNOWN_BYTES);
		if (tid == 0) a.long_list[pi].count = count;
	}
}

/* ids_out = ids_in with the long pieces' tokens spliced in.  long_list is sorted by start; lp.scratch holds the
 * scratch offset and lp.flags is unused; cum[i] = tokens of long pieces 0..i-1. */
__global__ void jtk_long_insert_kernel(const jtk_long_piece *list, const int64_t *cum, unsigned int n_long, const int32_t *scr_tok, const int32_t *ids_in,
                                       int32_t *ids_out, int64_t total_in) {
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; blockIdx.y == 0 && j < total_in; j += stride) {
		/* number of long pieces with insert_at <= j */
		unsigned lo = 0, hi = n_long;
		while (lo < hi) {
			unsigned mid = (lo + hi) >> 1;
			if (list[mid].insert_at <= j) lo = mid + 1;
			else hi = mid;
		}
		ids_out[j + cum[lo]] = ids_in[j];
	}
	for (unsigned i = blockIdx.y; i < n_long; i += gridDim.y) {
		const jtk_long_piece lp = list[i];
		const int32_t *src = scr_tok + lp.scratch;
		int32_t *dst = ids_out + lp.insert_at + cum[i];
		for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; t < lp.count; t += stride) dst[t] = src[t];
	}
}

__global__ void jtk_long_fix_offsets_kernel(const jtk_long_piece *list, const int64_t *cum, unsigned int n_long, const int64_t *doc_off, int64_t ndocs,
                                            int64_t *tok_off) {
	const int64_t d = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (d > ndocs) return;
	const int64_t g = doc_off[d];
	unsigned lo = 0, hi = n_long; /* long pieces with start < g */
	while (lo < hi) {
		unsigned mid = (lo + hi) >> 1;
		if (list[mid].start < g) lo = mid + 1;
		else hi = mid;
	}
	tok_off[d] += cum[lo];
}

} /* namespace */

/* =============================================================================================
 * launch wrappers
 * ============================================================================================= */
cudaError_t jtk_encode_kernel_setup() {
	return cudaFuncSetAttribute(jtk_encode_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, JTK_SMEM_BYTES);
}

cudaError_t jtk_launch_tile_first_doc(const int64_t *doc_off, int64_t ndocs, int64_t ntiles, int32_t *out, cudaStream_t st) {
	if (ntiles <= 0) return cudaSuccess;
	jtk_tile_first_doc_kernel<<<(unsigned) ((ntiles + 255) / 256), 256, 0, st>>>(doc_off, ndocs, ntiles, out);
	return cudaGetLastError();
}

cudaError_t jtk_launch_encode_tiles(const jtk_encode_args &a, int num_sms, cudaStream_t st) {
	if (a.ntiles <= 0) return cudaSuccess;
	int64_t grid = (int64_t) num_sms * 2;
	if (grid > a.ntiles) grid = a.ntiles;
	jtk_encode_tiles_kernel<<<(unsigned) grid, JTK_NT, JTK_SMEM_BYTES, st>>>(a);
	return cudaGetLastError();
}

cudaError_t jtk_launch_scan_gather(const jtk_encode_args &a, int num_sms, cudaStream_t st) {
	jtk_tile_scan_kernel<<<1, 1024, 0, st>>>(a);
	int64_t grid = std::max<int64_t>(std::min<int64_t>(std::max<int64_t>(a.ntiles, (a.ndocs + 256) / 256), (int64_t) num_sms * 16), 1);
	jtk_gather_kernel<<<(unsigned) grid, 256, 0, st>>>(a);
	return cudaGetLastError();
}

cudaError_t jtk_launch_long_bounds(const jtk_encode_args &a, unsigned int n_long, cudaStream_t st) {
	jtk_long_bounds_kernel<<<(n_long + 127) / 128, 128, 0, st>>>(a, n_long);
	return cudaGetLastError();
}

cudaError_t jtk_launch_long_merge(const jtk_encode_args &a, unsigned int n_long, int32_t *scr_tok, int32_t *scr_rk, int32_t *scr_nxt, int32_t *scr_prv,
                                  int num_sms, cudaStream_t st) {
	unsigned grid = n_long < (unsigned) num_sms ? n_long : (unsigned) num_sms;
	jtk_long_merge_kernel<<<grid, LNT, 0, st>>>(a, n_long, scr_tok, scr_rk, scr_nxt, scr_prv);
	return cudaGetLastError();
}

cudaError_t jtk_launch_long_insert(const jtk_long_piece *list, const int64_t *cum, unsigned int n_long, const int32_t *scr_tok, const int32_t *ids_in,
                                   int32_t *ids_out, int64_t total_in, cudaStream_t st) {
	dim3 grid(1024, n_long < 64 ? n_long : 64);
	jtk_long_insert_kernel<<<grid, 256, 0, st>>>(list, cum, n_long, scr_tok, ids_in, ids_out, total_in);
	return cudaGetLastError();
}

cudaError_t jtk_launch_long_fix_offsets(const jtk_long_piece *list, const int64_t *cum, unsigned int n_long, const int64_t *doc_off, int64_t ndocs,
                                        int64_t *tok_off, cudaStream_t st) {
	jtk_long_fix_offsets_kernel<<<(unsigned) ((ndocs + 1 + 255) / 256), 256, 0, st>>>(list, cum, n_long, doc_off, ndocs, tok_off);
	return cudaGetLastError();
}

/* =============================================================================================
 * decode: ids -> bytes (GptBytePairEncoding.decodeBytes / decodeToken, :136-151,302-314)
 *   lengths kernel (id -> token index + byte length), exclusive scan, gather kernel.
 * ============================================================================================= */
namespace {

constexpr int SCAN_NT = 256;
constexpr int SCAN_ITEMS = 16;

__device__ __forceinline__ int decode_find(const jtk_tables &T, int32_t id) {
	uint32_t s = jtk_hash_pair(id, 0) & T.mask_d;
	for (;;) {
		const uint32_t v = T.dec_keys[2 * s + 1];
		if (v == 0) return -1;
		if (T.dec_keys[2 * s] == (uint32_t) id) return (int) v - 1;
		s = (s + 1) & T.mask_d;
	}
}

__global__ void jtk_decode_lengths_kernel(const jtk_decode_args a, int32_t *tok_index, unsigned long long *bad_pos) {
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; j < a.nids; j += stride) {
		const int idx = decode_find(a.T, a.ids[j]);
		tok_index[j] = idx;
		if (idx >= 0) {
			a.id_byte_off[j] = a.T.dec_off[idx + 1] - a.T.dec_off[idx];
		} else {
			a.id_byte_off[j] = 0;
			/* document of token j: last d with tok_off[d] <= j */
			int64_t lo = 0, hi = a.ndocs - 1;
			while (lo < hi) {
				int64_t mid = (lo + hi + 1) >> 1;
				if (a.tok_off[mid] <= j) lo = mid;
				else hi = mid - 1;
			}
			atomicMin(bad_pos + lo, (unsigned long long) j);
		}
	}
}

/* in-place exclusive scan of int64 values, three phases */
__global__ void jtk_scan_block_kernel(int64_t *data, int64_t n, int64_t *block_sums) {
	__shared__ int64_t s_w[SCAN_NT / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int64_t base = (int64_t) blockIdx.x * SCAN_NT * SCAN_ITEMS + (int64_t) tid * SCAN_ITEMS;
	int64_t v[SCAN_ITEMS];
	int64_t sum = 0;
	for (int i = 0; i < SCAN_ITEMS; i++) {
		v[i] = base + i < n ? data[base + i] : 0;
		sum += v[i];
	}
	int64_t x = sum;
	for (int o = 1; o < 32; o <<= 1) {
		int64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
		if (lane >= o) x += y;
	}
	if (lane == 31) s_w[warp] = x;
	__syncthreads();
	if (warp == 0) {
		int64_t w = lane < SCAN_NT / 32 ? s_w[lane] : 0;
		for (int o = 1; o < 32; o <<= 1) {
			int64_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
			if (lane >= o) w += y;
		}
		if (lane < SCAN_NT / 32) s_w[lane] = w;
	}
	__syncthreads();
	int64_t run = x - sum + (warp ? s_w[warp - 1] : 0);
	for (int i = 0; i < SCAN_ITEMS; i++) {
		if (base + i < n) data[base + i] = run;
		run += v[i];
	}
	if (tid == SCAN_NT - 1) block_sums[blockIdx.x] = run;
}

__global__ void jtk_scan_sums_kernel(int64_t *block_sums, int64_t nblocks, int64_t *total_out) {
	/* one block: sequential chunks of blockDim.x with a running carry */
	__shared__ int64_t s_w[32];
	__shared__ int64_t s_carry;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_carry = 0;
	__syncthreads();
	for (int64_t base = 0; base < nblocks; base += blockDim.x) {
		const int64_t i = base + tid;
		const int64_t val = i < nblocks ? block_sums[i] : 0;
		int64_t x = val;
		for (int o = 1; o < 32; o <<= 1) {
			int64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31) s_w[warp] = x;
		__syncthreads();
		if (warp == 0) {
			int64_t w = lane < (int) (blockDim.x / 32) ? s_w[lane] : 0;
			for (int o = 1; o < 32; o <<= 1) {
				int64_t y = __shfl_up_sync(0xFFFFFFFFu, w, o);
				if (lane >= o) w += y;
			}
			s_w[lane] = w;
		}
		__syncthreads();
		const int64_t excl = s_carry + x - val + (warp ? s_w[warp - 1] : 0);
		if (i < nblocks) block_sums[i] = excl;
		__syncthreads();
		if (tid == (int) blockDim.x - 1) s_carry = excl + val;
		__syncthreads();
	}
	if (tid == 0) *total_out = s_carry;
}

__global__ void jtk_scan_add_kernel(int64_t *data, int64_t n, const int64_t *block_sums) {
	const int64_t base = (int64_t) blockIdx.x * SCAN_NT * SCAN_ITEMS;
	const int64_t add = block_sums[blockIdx.x];
	for (int i = threadIdx.x; i < SCAN_NT * SCAN_ITEMS; i += SCAN_NT)
		if (base + i < n) data[base + i] += add;
}

__global__ void jtk_decode_gather_kernel(const jtk_decode_args a, const int32_t *tok_index, const unsigned long long *bad_pos) {
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;
	for (int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; j < a.nids; j += stride) {
		const int idx = tok_index[j];
		if (idx < 0) continue;
		const uint32_t s = a.T.dec_off[idx], e = a.T.dec_off[idx + 1];
		uint8_t *dst = a.out + a.id_byte_off[j];
		for (uint32_t k = s; k < e; k++) dst[k - s] = a.T.dec_bytes[k];
	}
	for (int64_t d = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; d <= a.ndocs; d += stride) {
		a.byte_off[d] = a.id_byte_off[a.tok_off[d]];
		if (d < a.ndocs) {
			const unsigned long long bp = bad_pos[d];
			if (bp != ~0ull) {
				a.doc_status[d] |= JTK_DOC_UNKNOWN_ID;
				a.bad_ids[d] = a.ids[bp];
			} else {
				a.bad_ids[d] = 0;
			}
		}
	}
}

} /* namespace */

int64_t jtk_scan_blocks(int64_t n) { return (n + SCAN_NT * SCAN_ITEMS - 1) / (SCAN_NT * SCAN_ITEMS); }

/* id_byte_off has nids + 1 entries; entry nids must be zero on entry and receives the total. */
cudaError_t jtk_launch_decode_lengths(const jtk_decode_args &a, int32_t *tok_index, unsigned long long *bad_pos, int64_t *block_sums, int64_t *total,
                                      cudaStream_t st) {
	if (a.nids > 0) {
		unsigned grid = (unsigned) std::min<int64_t>((a.nids + 255) / 256, 148 * 16);
		jtk_decode_lengths_kernel<<<grid, 256, 0, st>>>(a, tok_index, bad_pos);
	}
	const int64_t n = a.nids + 1;
	const int64_t nb = jtk_scan_blocks(n);
	jtk_scan_block_kernel<<<(unsigned) nb, SCAN_NT, 0, st>>>(a.id_byte_off, n, block_sums);
	jtk_scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total);
	jtk_scan_add_kernel<<<(unsigned) nb, SCAN_NT, 0, st>>>(a.id_byte_off, n, block_sums);
	return cudaGetLastError();
}

cudaError_t jtk_launch_decode_gather(const jtk_decode_args &a, const int32_t *tok_index, const unsigned long long *bad_pos, cudaStream_t st) {
	const int64_t work = std::max<int64_t>(a.nids, a.ndocs + 1);
	unsigned grid = (unsigned) std::min<int64_t>((work + 255) / 256, 148 * 16);
	jtk_decode_gather_kernel<<<grid, 256, 0, st>>>(a, tok_index, bad_pos);
	return cudaGetLastError();
}
