/*
 * sm_100a kernels of the JTokkit encode path.
 *
 * The batch is processed in sub-batches of tiles (8 KiB of input each) so that the intermediates stay close to
 * the L2; per sub-batch four uniform kernels run back to back on one stream, none of them waits for another CTA:
 *
 *   jtk_split_lookup_kernel  persistent CTAs take tiles by ticket: stage tile + halos in shared memory (16-byte
 *                            loads), classify code points, evaluate the split rules into a piece-start bitmask,
 *                            list the piece starts, then ONE PIECE PER THREAD probe the byte-keyed piece table
 *                            (L2 resident).  Output: one int32 record per piece (the token id, or "needs merging"
 *                            with start/length), the tile's list of pieces that need merging, counts.
 *   jtk_merge_kernel         one CTA per tile: the tile's unresolved pieces, sorted by length, run the exact
 *                            bytePairMerge loop (thread per short piece, warp per medium piece) against the pair table.
 *   jtk_tile_scan_kernel     exclusive scan of the per-tile token counts.
 *   jtk_gather_kernel        one CTA per tile: per-piece token counts -> block scan -> ids written at their final
 *                            position, document token offsets resolved.
 *
 * Reference code this replaces: GptBytePairEncoding.encodeOrdinaryInternal / bytePairMerge / getRank
 * (GptBytePairEncoding.java:71-103,200-300) and the special-token guard of encodeInternal (:52-56).
 */
#include "jtk_kernels.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "jtk_device.cuh"
#include "jtk_regex.h"

namespace {

constexpr int NT = JTK_NT;
constexpr int NWARPS = NT / 32;
constexpr int RECN = JTK_RECN;
constexpr int QCAP = JTK_QCAP;
constexpr int BH = JTK_BACK_HALO;
constexpr int TC = JTK_TILE / 16;  /* 16-byte chunks per tile */
static_assert(TC <= NT && NT % 32 == 0 && NT <= 1024, "one thread per tile chunk at least");

/* piece records: a token id, or (id space is limited to >= JTK_REC_MIN_ID at registration) a payload */
constexpr int32_t REC_BASE = (int32_t) 0x80000000;
constexpr uint32_t REC_LONG = 1u << 28; /* payload = REC_LONG | index into long_list; else (offset << 11) | (count - 1) */
constexpr uint32_t REC_SKIP = 1u << 29; /* general patterns: a gap between matches, no tokens */
constexpr uint32_t REC_MEMO = 1u << 27; /* payload = REC_MEMO | memo entry << 4 | (count - 1): the tokens are read from the memo entry itself */
__device__ __forceinline__ bool rec_is_id(int32_t r) { return r >= JTK_REC_MIN_ID; }
__device__ __forceinline__ int32_t rec_make(int s, int m) { return REC_BASE + (int32_t) (((uint32_t) s << 11) | (uint32_t) (m - 1)); }
__device__ __forceinline__ uint32_t rec_payload(int32_t r) { return (uint32_t) (r - REC_BASE); }
__device__ __forceinline__ int32_t rec_make_memo(uint32_t entry, int m) { return REC_BASE + (int32_t) (REC_MEMO | (entry << 4) | (uint32_t) (m - 1)); }



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

/* block-wide exclusive scan of one int per thread; *total receives the block sum.  wsum: NTHREADS/32 ints of shared memory. */
template <int NTHREADS>
__device__ __forceinline__ int block_exclusive_scan(int v, int *wsum, int *total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int x = v;
	for (int o = 1; o < 32; o <<= 1) {
		int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
		if (lane >= o) x += y;
	}
	if (lane == 31) wsum[warp] = x;
	__syncthreads();
	if (warp == 0) {
		int w = lane < NTHREADS / 32 ? wsum[lane] : 0;
		for (int o = 1; o < 32; o <<= 1) {
			int y = __shfl_up_sync(0xFFFFFFFFu, w, o);
			if (lane >= o) w += y;
		}
		if (lane < NTHREADS / 32) wsum[lane] = w;
	}
	__syncthreads();
	const int excl = x - v + (warp ? wsum[warp - 1] : 0);
	*total = wsum[NTHREADS / 32 - 1];
	__syncthreads(); /* wsum may be reused by the caller */
	return excl;
}

/* ---------------------------------------------------------------------------------------------
 * bytePairMerge for one piece of 33..JTK_LONG_PIECE bytes by a group of G = 2^LOG2G lanes
 * (GptBytePairEncoding.java:200-275).  Position k is owned by group lane k % G (row k / G, at most 32 rows), so
 * rank scans are bank-conflict free; the argmin over adjacent pair ranks is a REDUX min over per-lane minima,
 * neighbours are found with REDUX over per-lane alive masks, and the two rank probes of a merge (:254-257) are
 * issued by two lanes in parallel.  Several groups share a warp (G = 8: four pieces per warp).
 * ------------------------------------------------------------------------------------------- */
/* min / max over the G lanes of a group; full warps use the hardware REDUX, smaller groups a xor-shuffle butterfly */
template <int G>
__device__ __forceinline__ int group_min(unsigned gmask, int v) {
	if (G == 32) return __reduce_min_sync(gmask, v);
#pragma unroll
	for (int o = G / 2; o; o >>= 1) v = min(v, __shfl_xor_sync(gmask, v, o));
	return v;
}
template <int G>
__device__ __forceinline__ int group_max(unsigned gmask, int v) {
	if (G == 32) return __reduce_max_sync(gmask, v);
#pragma unroll
	for (int o = G / 2; o; o >>= 1) v = max(v, __shfl_xor_sync(gmask, v, o));
	return v;
}
/* leftmost minimum in one pass: key = rank (signed, high word) : position (low word) */
template <int G>
__device__ __forceinline__ long long group_min64(unsigned gmask, long long v) {
#pragma unroll
	for (int o = G / 2; o; o >>= 1) {
		const long long y = __shfl_xor_sync(gmask, v, o);
		v = y < v ? y : v;
	}
	return v;
}

template <int LOG2G>
__device__ int merge_group(const jtk_tables &T, const uint8_t *p, int n, int32_t *tok, int32_t *rk, bool *unknown) {
	constexpr int G = 1 << LOG2G;
	const int lane = threadIdx.x & 31;
	const int gl = lane & (G - 1);
	const unsigned gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << (lane & ~(G - 1)));
	for (int k = gl; k < n; k += G) {
		tok[k] = T.byte_id[p[k]];
		rk[k] = (k + 1 < n) ? T.bytepair[((uint32_t) p[k] << 8) | p[k + 1]] : JTK_RANK_MAX;
	}
	__syncwarp(gmask);
	const int rows = (n + G - 1) >> LOG2G;
	uint32_t alive = 0;
	for (int j = 0; j < rows; j++)
		if ((j << LOG2G) + gl < n) alive |= 1u << j;
	int32_t lr = JTK_RANK_MAX;
	int lk = 0x7fffffff;
	auto rescan = [&]() {
		lr = JTK_RANK_MAX;
		lk = 0x7fffffff;
		for (uint32_t m = alive; m;) {
			int j = __ffs((int) m) - 1;
			m &= m - 1;
			int k = (j << LOG2G) + gl;
			int32_t r = rk[k];
			if (r < lr) {
				lr = r;
				lk = k;
			}
		}
	};
	rescan();
	for (;;) {
		const long long best = group_min64<G>(gmask, ((long long) lr << 32) | (unsigned) lk); /* leftmost minimum (:232-240) */
		const int32_t mr = (int32_t) (best >> 32);
		if (mr == JTK_RANK_MAX) break;
		const int mi = (int) (unsigned) best;
		/* next alive position after k / previous alive position before k */
		auto next_alive = [&](int k) {
			int j0 = (k >> LOG2G) + (gl <= (k & (G - 1)) ? 1 : 0);
			uint32_t m = j0 >= 32 ? 0u : (alive & (0xFFFFFFFFu << j0));
			int cand = m ? (((__ffs((int) m) - 1) << LOG2G) + gl) : 0x7fffffff;
			return group_min<G>(gmask, cand);
		};
		auto prev_alive = [&](int k) {
			int j1 = (k >> LOG2G) - (gl < (k & (G - 1)) ? 0 : 1);
			uint32_t m = j1 < 0 ? 0u : (alive & (j1 >= 31 ? 0xFFFFFFFFu : ((2u << j1) - 1u)));
			int cand = m ? (((31 - __clz((int) m)) << LOG2G) + gl) : -1;
			return group_max<G>(gmask, cand);
		};
		const int nx = next_alive(mi);
		const int nn = next_alive(nx);
		const int pv = prev_alive(mi);
		if (gl == (mi & (G - 1))) tok[mi] = mr;
		if (gl == (nx & (G - 1))) {
			alive &= ~(1u << (nx >> LOG2G));
			rk[nx] = JTK_RANK_MAX;
		}
		__syncwarp(gmask);
		if (gl == 0) rk[mi] = (nn != 0x7fffffff) ? jtk_lookup_pair(T, mr, tok[nn]) : JTK_RANK_MAX;
		if (gl == 1 && pv >= 0) rk[pv] = jtk_lookup_pair(T, tok[pv], mr);
		__syncwarp(gmask);
		if (gl == (mi & (G - 1)) || gl == (nx & (G - 1)) || (pv >= 0 && gl == (pv & (G - 1)))) rescan();
	}
	/* compact the surviving parts to the front of tok[] */
	int out = 0;
	bool unk = false;
	for (int j = 0; j < rows; j++) {
		const bool a = (alive >> j) & 1u;
		const int32_t v = a ? tok[(j << LOG2G) + gl] : 0;
		const unsigned b = __ballot_sync(gmask, a) & gmask;
		__syncwarp(gmask);
		if (a) {
			tok[out + __popc(b & ((1u << lane) - 1u))] = v;
			if (v < JTK_PSEUDO_BASE + 256) unk = true;
		}
		out += __popc(b);
		__syncwarp(gmask);
	}
	if (__any_sync(gmask, unk)) *unknown = true;
	return out;
}

/* ---------------------------------------------------------------------------------------------
 * kernel 1: split + whole-piece lookup
 *
 * Persistent CTAs of JTK_NT threads take tiles by ticket.  The region of a tile (back halo + tile + forward halo, 8 192 bytes)
 * is staged by ONE bulk asynchronous copy (cp.async.bulk global -> shared, completion on an mbarrier); the region of the
 * NEXT tile is requested before the current one is processed (two staging buffers), so the copy runs under the compute and
 * no thread spends instructions on staging.  Per tile:
 *   P1  mark document starts (bit mask), wait for the bytes
 *   P2  one 16-byte chunk per thread: code point classes -> bit planes (jtk_classify_chunk); special-token first-byte test
 *   P3  one chunk per thread: split rules over the planes -> piece-start bits
 *   P4a piece starts listed in order (popcount + block scan)
 *   P4b one piece per thread: pieces of up to 8 bytes probe the first half of their table slot (one 16-byte load); longer
 *       pieces, misses and everything unusual go to a short deferred list, which a second, dense pass resolves (full-key
 *       probe, long keys, memo, queues for the merge kernels): the rare paths no longer run inside every warp.
 * ------------------------------------------------------------------------------------------- */
#ifndef JTK_SPLIT_CTAS
#define JTK_SPLIT_CTAS (2048 / JTK_NT) /* resident CTAs per SM of the split+lookup kernel (register budget = 65536 / (JTK_NT * JTK_SPLIT_CTAS)) */
#endif
constexpr int SB_BYTES = JTK_REGION + 32;                 /* one staging buffer: region + pad chunk + 16 zero bytes */
constexpr int COPY_BYTES = JTK_REGION + 16;               /* what the bulk copy covers in the interior of the input */
constexpr int PLANE_BYTES = 16 * (JTK_REGION_CHUNKS + 2); /* four plane words per chunk */
constexpr int PLIST_BYTES = ((2 * JTK_TILE > PLANE_BYTES ? 2 * JTK_TILE : PLANE_BYTES) + 15) / 16 * 16; /* plist reuses the planes' memory after P3 */
/* per-warp lists of the lookup step (what does not fit is resolved in place) */
constexpr int WKEYS = 64;  /* pieces of 9..24 bytes: full-key pass (drained 32 at a time: never more than 63 entries) */
constexpr int WODD = 64;   /* unusual pieces (longer keys, long pieces, the tile's last piece) */
constexpr int WMISS = 96;  /* table misses: memo / queue pass (drained 32 at a time: never more than 31 + 32 + 32 entries) */
constexpr int PLIST_POS = 0x7FFF, PLIST_CUT = 0x8000; /* piece list entry: region index | "a safe cut, not a piece start of the split pattern" */
static_assert(JTK_REGION + 32 <= PLIST_POS, "region indices must fit 15 bits");
constexpr int DEFCAP = WKEYS * (JTK_NT / 32), ODDCAP = WODD * (JTK_NT / 32), MISSCAP = WMISS * (JTK_NT / 32);
static_assert(SB_BYTES % 16 == 0 && COPY_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
static_assert(JTK_REGION_CHUNKS + 1 <= NT, "one thread per region chunk");

/* indices into the small shared "misc" array */
enum { M_NSLOW = 0, M_HITS, M_SLOWTOK, M_CARRY = 4, M_TICKET /* 2 */, M_WSUM = 16, M_HIST = 48 /* .. M_HIST + JTK_SHORT_PIECE */, M_WORDS = 128 };
static_assert(M_HIST + JTK_SHORT_PIECE + 1 <= M_WORDS, "misc too small");

constexpr int MASK_BYTES = (4 * JTK_MASK_WORDS + 15) / 16 * 16; /* one bit mask over the region */
constexpr int SPLIT_SMEM_BYTES = 2 * SB_BYTES + PLIST_BYTES + 3 * MASK_BYTES + 4 * ((TC + 3) / 4 * 4) + 4 * M_WORDS + 1024 + 2048 + 2 * DEFCAP + 2 * ODDCAP + 2 * MISSCAP + 16;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_LOOP:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra WAIT_DONE;\n"
	    "bra WAIT_LOOP;\n"
	    "WAIT_DONE:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
/* one bulk asynchronous copy global -> shared; src, dst and bytes are multiples of 16; completion is counted on `bar` */
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
	             "r"(smem_u32(bar))
	             : "memory");
}

/* the part of tile `tile`'s region that consists of whole 16-byte chunks inside the input: [*lo, *hi) in region bytes */
__device__ __forceinline__ void region_copy_range(int64_t tile, int64_t total, int *lo, int *hi) {
	const int64_t g0 = tile * (int64_t) JTK_TILE - BH;
	*lo = g0 < 0 ? (int) -g0 : 0;
	const int64_t avail = (total & ~(int64_t) 15) - g0; /* end of the last whole chunk of the input, in region bytes */
	*hi = avail >= COPY_BYTES ? COPY_BYTES : (avail > *lo ? (int) avail : *lo);
}

/* dynamic shared memory of the split+lookup kernel: fixed offsets, so that every device function can address it */
extern __shared__ __align__(128) uint8_t jtk_dyn_smem[];
constexpr int OFF_SB = 0;                                  /* two staging buffers */
constexpr int OFF_PLANES = OFF_SB + 2 * SB_BYTES;          /* planes, later the piece list */
constexpr int OFF_BMASK = OFF_PLANES + PLIST_BYTES;
constexpr int OFF_DMASK = OFF_BMASK + MASK_BYTES;
constexpr int OFF_RMASK = OFF_DMASK + MASK_BYTES;  /* starts of the split pattern's pieces alone (bmask also has the safe cuts) */
constexpr int OFF_CPREF = OFF_RMASK + MASK_BYTES;  /* pieces before each 16-byte chunk of the tile */
constexpr int OFF_MISC = OFF_CPREF + 4 * ((TC + 3) / 4 * 4);
constexpr int OFF_LUT = OFF_MISC + 4 * M_WORDS;
constexpr int OFF_CLS2 = OFF_LUT + 1024;
constexpr int OFF_DEF = OFF_CLS2 + 2048;
constexpr int OFF_ODD = OFF_DEF + 2 * DEFCAP;
constexpr int OFF_MISS = OFF_ODD + 2 * ODDCAP;
constexpr int OFF_MBAR = OFF_MISS + 2 * MISSCAP;
static_assert(OFF_MBAR + 16 == SPLIT_SMEM_BYTES && OFF_MBAR % 8 == 0 && OFF_PLANES % 16 == 0 && OFF_BMASK % 16 == 0, "shared memory layout");
struct split_smem {
	__device__ __forceinline__ uint8_t *sb(int i) const { return jtk_dyn_smem + OFF_SB + i * SB_BYTES; }
	__device__ __forceinline__ uint32_t *planes() const { return reinterpret_cast<uint32_t *>(jtk_dyn_smem + OFF_PLANES); }
	__device__ __forceinline__ uint16_t *plist() const { return reinterpret_cast<uint16_t *>(jtk_dyn_smem + OFF_PLANES); }
	__device__ __forceinline__ uint32_t *bmask() const { return reinterpret_cast<uint32_t *>(jtk_dyn_smem + OFF_BMASK); }
	__device__ __forceinline__ uint32_t *dmask() const { return reinterpret_cast<uint32_t *>(jtk_dyn_smem + OFF_DMASK); }
	__device__ __forceinline__ uint32_t *rmask() const { return reinterpret_cast<uint32_t *>(jtk_dyn_smem + OFF_RMASK); }
	__device__ __forceinline__ uint32_t *chunk_pref() const { return reinterpret_cast<uint32_t *>(jtk_dyn_smem + OFF_CPREF); }
	__device__ __forceinline__ uint32_t *misc() const { return reinterpret_cast<uint32_t *>(jtk_dyn_smem + OFF_MISC); }
	__device__ __forceinline__ uint32_t *lut_sp() const { return reinterpret_cast<uint32_t *>(jtk_dyn_smem + OFF_LUT); }
	__device__ __forceinline__ uint8_t *cls2() const { return jtk_dyn_smem + OFF_CLS2; }
	__device__ __forceinline__ uint16_t *deflist() const { return reinterpret_cast<uint16_t *>(jtk_dyn_smem + OFF_DEF); }
	__device__ __forceinline__ uint16_t *oddlist() const { return reinterpret_cast<uint16_t *>(jtk_dyn_smem + OFF_ODD); }
	__device__ __forceinline__ uint16_t *misslist() const { return reinterpret_cast<uint16_t *>(jtk_dyn_smem + OFF_MISS); }
	__device__ __forceinline__ uint64_t *mbar() const { return reinterpret_cast<uint64_t *>(jtk_dyn_smem + OFF_MBAR); }
};

/* A piece of n bytes at region index r that is not a token as a whole (pass 3 of P4b): the memo of this call may hold its tokens,
 * otherwise it is queued for the merge kernels.  Returns the tokens accounted for now (memo hit) and stores the piece's record. */
__device__ __forceinline__ int split_resolve_miss(const jtk_encode_args &a, const uint8_t *sb, int q, int r, int n, long long lt) {
	const split_smem S;
	const int s = r - BH;
	int32_t out = rec_make(s, n);
	int hits = 0;
	bool memo_hit = false;
	if (n <= JTK_MEMO_MAX_PIECE && a.memo) { /* has this call merged the same piece before? */
		uint32_t key[6];
		jtk_build_key4(sb + r, n, key);
		key[4] = key[5] = 0;
		const uint32_t entry = (jtk_hash6(key, (uint32_t) n) * 0x9E3779B1u) >> 8 & a.memo_mask;
		/* entries are written by the merge kernel of EARLIER sub-batches of this call (stream ordered, never while this kernel
		 * runs: the two-lane pipeline runs without a memo) and never change afterwards: the record points at the entry and
		 * the gather kernel reads the tokens from there */
		const uint4 *mp = reinterpret_cast<const uint4 *>(a.memo + entry);
		const uint4 q0 = mp[0];
		const uint2 q1 = *reinterpret_cast<const uint2 *>(mp + 1);
		if ((q1.x >> 8) == a.memo_epoch && (q1.x & 3u) == 2u && (q1.y & 0xFFu) == (uint32_t) n && q0.x == key[0] && q0.y == key[1] && q0.z == key[2] && q0.w == key[3]) {
			const int cnt = (int) (q1.y >> 8);
			out = rec_make_memo(entry, cnt);
			hits = cnt;
			memo_hit = true;
		}
	}
	/* queue for the merge kernels (warp-aggregated slot in the tile's list of short pieces) */
	const bool shortq = !memo_hit && n <= JTK_SHORT_PIECE;
	const unsigned act = __activemask();
	const unsigned sm = __ballot_sync(act, shortq);
	if (sm) {
		const int lane = threadIdx.x & 31, leader = __ffs((int) sm) - 1;
		unsigned slot = 0;
		if (lane == leader) slot = atomicAdd(&S.misc()[M_NSLOW], (unsigned) __popc(sm));
		slot = __shfl_sync(act, slot, leader) + __popc(sm & ((1u << lane) - 1u));
		if (shortq) {
			a.slowq[lt * (long long) QCAP + slot] = (uint16_t) q;
			atomicAdd(&S.misc()[M_HIST + n], 1u);
		}
	}
	if (!memo_hit && n > JTK_SHORT_PIECE) {
		if (n <= JTK_GROUP8_PIECE) a.med8[atomicAdd(&a.sub->n_med8, 1u)] = ((uint32_t) lt << 14) | (uint32_t) q;
		else a.med32[atomicAdd(&a.sub->n_med32, 1u)] = ((uint32_t) lt << 14) | (uint32_t) q;
	}
	(a.rec + lt * (long long) RECN)[q] = out;
	return hits;
}

/* The unusual pieces of a tile: keys longer than 24 bytes, pieces longer than the in-tile limit, the tile's last piece (whose end
 * lies in the halo), and whatever did not fit the shared-memory lists.  Complete on its own: lookup, memo, queues. */
template <bool GENERAL>
__device__ __noinline__ int split_slow_piece(const jtk_encode_args &a, const uint8_t *sb, int q, int npieces, long long lt, int64_t tb) {
	const jtk_tables &T = a.T;
	const split_smem S;
	const int r = S.plist()[q] & PLIST_POS;
	const int s = r - BH;
	int e = (q + 1 < npieces) ? (int) (S.plist()[q + 1] & PLIST_POS) : next_bit(S.bmask(), r + 1, r + JTK_LONG_PIECE);
	if (e >= 0 && e - r > JTK_LONG_PIECE) e = -1;
	if (e < 0) { /* longer than JTK_LONG_PIECE: deferred to the long-piece kernels */
		const unsigned idx = atomicAdd(&a.hdr->n_long, 1u);
		if ((int64_t) idx < a.long_cap) {
			jtk_long_piece lp;
			lp.start = tb + s;
			const int e2 = next_bit(S.bmask(), r + 1, JTK_REGION - 1);
			lp.end = e2 < 0 ? -1 : (tb - BH) + e2;
			lp.insert_at = 0; /* set by the gather kernel */
			lp.count = 0;
			lp.scratch = 0;
			lp.doc = 0;
			lp.flags = 0;
			a.long_list[idx] = lp;
		}
		(a.rec + lt * (long long) RECN)[q] = REC_BASE + (int32_t) (REC_LONG | (idx & 0x0FFFFFFFu));
		return 0;
	}
	const int n = e - r;
	const uint8_t *p = sb + r;
	int32_t out;
	if (!GENERAL && n != 1 && !(((S.rmask()[r >> 5] >> (r & 31)) & (S.rmask()[e >> 5] >> (e & 31))) & 1u)) {
		out = JTK_RANK_MAX; /* a segment between safe cuts, not a piece of the split pattern: no whole-piece shortcut */
	} else if (n <= JTK_INLINE_KEY_MAX) {
		uint32_t key[6];
		jtk_build_key(p, n, key);
		out = jtk_lookup_a(T, key, (uint32_t) n, jtk_hash6(key, (uint32_t) n));
		if (out == JTK_RANK_MAX && n == 1) { /* a byte that is not in the vocabulary (TokenEncoder.java:64-71) */
			flag_doc(a, tb + s, JTK_DOC_UNKNOWN_BYTES);
			out = JTK_REC_MIN_ID; /* the document is in error; keep the record a plain id */
		}
	} else {
		out = n > T.max_token_len ? JTK_RANK_MAX : jtk_lookup_long(T, p, n);
	}
	if (out == JTK_RANK_MAX) return split_resolve_miss(a, sb, q, r, n, lt);
	(a.rec + lt * (long long) RECN)[q] = out;
	return 1;
}

template <bool GENERAL>
__global__ void __launch_bounds__(JTK_NT, JTK_SPLIT_CTAS) jtk_split_lookup_kernel(const __grid_constant__ jtk_encode_args a) {
	const split_smem L;
	struct {
		uint8_t *sb0;
		uint32_t *planes, *bmask, *dmask, *rmask, *chunk_pref, *lut_sp;
		uint16_t *plist, *deflist, *oddlist, *misslist;
		uint8_t *cls2;
		uint64_t *mbar;
	} S;
	S.sb0 = L.sb(0);
	S.planes = L.planes();
	S.plist = L.plist(); /* region index of every piece start, in order (written after the planes' last use) */
	S.bmask = L.bmask();
	S.dmask = L.dmask();
	S.rmask = L.rmask();
	S.chunk_pref = L.chunk_pref();
	S.lut_sp = L.lut_sp();
	S.cls2 = L.cls2();
	S.deflist = L.deflist();
	S.oddlist = L.oddlist();
	S.misslist = L.misslist();
	S.mbar = L.mbar();
	uint32_t *const misc = L.misc();

	const int tid = threadIdx.x;
	const int lane = tid & 31, warp = tid >> 5;
	const jtk_tables &T = a.T;

	/* ---- once per CTA: tables, zeroed masks and pads, barriers, the first two tickets ---- */
	for (int i = tid; i < 256; i += NT) S.lut_sp[i] = T.lut_sp[i];
	for (int i = tid; i < 512; i += NT) reinterpret_cast<uint32_t *>(S.cls2)[i] = reinterpret_cast<const uint32_t *>(T.cls2)[i];
	for (int w = tid; w < 3 * MASK_BYTES / 4; w += NT) S.bmask[w] = 0; /* bmask, dmask and rmask */
	if (tid < M_WORDS) misc[tid] = 0;
	if (tid < 8) { /* the 16 bytes after what the copies cover stay zero for good */
		reinterpret_cast<uint32_t *>(S.sb0 + (tid >> 2) * SB_BYTES + COPY_BYTES)[tid & 3] = 0;
	}
	if (tid == 0) {
		mbar_init(&S.mbar[0], 1);
		mbar_init(&S.mbar[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const int64_t ntiles_sub = a.tile_end - a.tile_begin;
	auto issue = [&](int64_t lt_next, int buf) { /* one thread: request the region of sub-batch tile lt_next into staging buffer buf */
		int lo, hi;
		region_copy_range(a.tile_begin + lt_next, a.total, &lo, &hi);
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the buffer was last touched by ordinary loads / stores */
		mbar_expect_tx(&S.mbar[buf], (uint32_t) (hi - lo));
		if (hi > lo) bulk_load(S.sb0 + buf * SB_BYTES + lo, a.bytes + ((a.tile_begin + lt_next) * (int64_t) JTK_TILE - BH) + lo, (uint32_t) (hi - lo), &S.mbar[buf]);
	};
	if (tid == 0) {
		const unsigned t0 = atomicAdd(&a.sub->ticket, 1u);
		if ((int64_t) t0 < ntiles_sub) issue((int64_t) t0, 0);
		misc[M_TICKET] = t0;
		misc[M_TICKET + 1] = atomicAdd(&a.sub->ticket, 1u);
	}
	__syncthreads();
	int64_t cur = misc[M_TICKET], nxt = misc[M_TICKET + 1];
	__syncthreads(); /* the ticket slots are rewritten inside the loop */

	jtk_tile_ctx c;
	c.bmask = S.bmask;
	c.dmask = S.dmask;
	c.planes = S.planes;
	c.tok = nullptr;
	c.rk = nullptr;
	c.total = a.total;
	c.gbytes = a.bytes;
	c.doc_off = a.doc_off;
	c.ndocs = a.ndocs;
	c.T = &T;
	c.lut_sp = S.lut_sp;
	c.cls2 = S.cls2;
	/* the special-token guard only has to look at bytes that can start a special token */
	const bool check_special = (a.flags & JTK_CHECK_SPECIAL) && T.nspecial > 0;

	for (int it = 0; cur < ntiles_sub; it++) {
		const int buf = it & 1;
		const long long lt = cur; /* index into the per-sub-batch buffers */
		const long long tile = a.tile_begin + lt;
		const int64_t tb = tile * (int64_t) JTK_TILE;
		uint8_t *const sb = S.sb0 + buf * SB_BYTES;
		c.sb = sb;
		c.g0 = tb - BH;
		c.rs = 0;
		c.carry_n = 0;
		/* ---- P0: request the next tile's bytes (its buffer was released by the barrier that ended the previous tile), take a ticket ---- */
		if (tid == 0 && nxt < ntiles_sub) issue(nxt, buf ^ 1);
		if (tid == NT - 32) misc[M_TICKET + buf] = atomicAdd(&a.sub->ticket, 1u);

		/* ---- P1: document starts; the bytes the bulk copy does not cover (start / end of the input); wait for the copy ---- */
		const int64_t first_doc = a.tile_first_doc[tile];
		if (GENERAL) {
			/* general pattern: the piece bits were computed by the jtk_general_* kernels; dmask holds the gap bits */
			const int64_t w0 = c.g0 / 32; /* g0 is a multiple of 32 (negative for the first tile) */
			for (int w = tid; w < JTK_MASK_WORDS; w += NT) {
				const int64_t gw = w0 + w;
				const bool in = gw >= 0 && gw < a.rx_words;
				S.bmask[w] = in ? a.rx_start[gw] : 0u;
				S.dmask[w] = in ? a.rx_skip[gw] : 0u;
			}
		} else {
			jtk_mark_docstarts(c, first_doc, tid, NT);
		}
		{
			int lo, hi;
			region_copy_range(tile, a.total, &lo, &hi);
			if (lo > 0 || hi < COPY_BYTES) { /* first / last tiles only */
				for (int i = tid; i < COPY_BYTES / 4; i += NT) {
					const int r = 4 * i;
					if (r >= lo && r < hi) continue;
					uint32_t v = 0;
					for (int k = 0; k < 4; k++) {
						const int64_t g = c.g0 + r + k;
						if (g >= 0 && g < a.total) v |= (uint32_t) a.bytes[g] << (8 * k);
					}
					reinterpret_cast<uint32_t *>(sb)[i] = v;
				}
			}
		}
		mbar_wait(&S.mbar[buf], (uint32_t) (it >> 1) & 1u);
		__syncthreads();

		if (!GENERAL) {
			c.rs = jtk_region_first(c);
			/* ---- P2: code point classes -> bit planes, one chunk per thread ---- */
			if (tid <= JTK_REGION_CHUNKS) jtk_classify_chunk(c, tid);
		}
		if (check_special && tid >= BH / 16 && tid < BH / 16 + TC) {
			const int r0 = tid * 16;
			bool look = true;
			if (T.special_first_single) { /* one candidate first byte (built-ins: '<'): SWAR "has byte" test on the 16 bytes */
				const uint4 w = *reinterpret_cast<const uint4 *>(sb + r0);
				const uint32_t pat = T.special_first_single * 0x01010101u;
				const uint32_t x0 = w.x ^ pat, x1 = w.y ^ pat, x2 = w.z ^ pat, x3 = w.w ^ pat;
				look = ((((x0 - 0x01010101u) & ~x0) | ((x1 - 0x01010101u) & ~x1) | ((x2 - 0x01010101u) & ~x2) | ((x3 - 0x01010101u) & ~x3)) & 0x80808080u) != 0;
			}
			if (look)
				for (int i = 0; i < 16; i++) {
					const int64_t g = c.g0 + r0 + i;
					if (g >= a.total) break;
					const uint8_t b = sb[r0 + i];
					if ((T.special_first[b >> 5] >> (b & 31)) & 1u) {
						if (jtk_special_at(T, a.bytes, g, jtk_doc_ceil(c, g))) flag_doc(a, g, JTK_DOC_HAS_SPECIAL);
					}
				}
		}
		uint32_t cutbits = 0;
		if (!GENERAL) {
			__syncthreads();
			/* \p{N} carry into the region: needed only when the region starts inside a digit run (the same answer in every thread) */
			if (T.pattern_kind == JTK_PAT_CL100K && !jtk_docstart(c, c.rs) && c.g0 + c.rs > 0 && jtk_clsb(c, c.rs) == JTK_C_N) {
				if (tid == 0) misc[M_CARRY] = (uint32_t) jtk_global_count_n_before(c, c.g0 + c.rs);
				__syncthreads();
				c.carry_n = (int) misc[M_CARRY];
			}
			/* ---- P3: split rules -> piece-start bits ---- */
			{
				const int ch = BH / 16 + tid;
				if (ch < JTK_REGION_CHUNKS) {
					const uint32_t rb = jtk_boundary_chunk(c, ch);
					cutbits = jtk_cut_chunk(c, ch) & ~rb;
					reinterpret_cast<uint16_t *>(S.rmask)[ch] = (uint16_t) rb;
					reinterpret_cast<uint16_t *>(S.bmask)[ch] = (uint16_t) (rb | cutbits);
				}
			}
		}
		/* (does the region have any safe cut?  if not - English text - every listed piece is a piece of the split pattern) */
		const bool any_cuts = __syncthreads_or(cutbits != 0) != 0;

		/* ---- P4a: list the piece starts of the tile in order (thread owns one chunk) ---- */
		uint32_t bits;
		int mycount;
		{
			const int ch = tid;
			const int64_t gbase = tb + ch * 16;
			uint32_t m = ch < TC ? reinterpret_cast<const uint16_t *>(S.bmask)[BH / 16 + ch] : 0u;
			if (gbase >= a.total) m = 0;
			else if (gbase + 16 > a.total) m &= (1u << (int) (a.total - gbase)) - 1u;
			bits = m;
			mycount = __popc(m);
			if (a.piece_flags && ch < TC) {
				/* the split pattern's own pieces: gaps are not pieces (general patterns), safe cuts are not piece starts */
				const uint32_t real = GENERAL ? m & ~(uint32_t) reinterpret_cast<const uint16_t *>(S.dmask)[BH / 16 + ch] : m & reinterpret_cast<const uint16_t *>(S.rmask)[BH / 16 + ch];
				for (int i = 0; i < 16 && gbase + i < a.total; i++) a.piece_flags[gbase + i] = (real >> i) & 1u;
			}
		}
		int npieces, base;
		{
			/* block-wide exclusive scan with one barrier: warp scan, warp totals to shared memory, every warp scans the totals itself */
			int x = mycount;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
				if (lane >= o) x += y;
			}
			if (lane == 31) misc[M_WSUM + warp] = (uint32_t) x;
			__syncthreads();
			int ws = lane < NWARPS ? (int) misc[M_WSUM + lane] : 0;
#pragma unroll
			for (int o = 1; o < NWARPS; o <<= 1) {
				const int y = __shfl_up_sync(0xFFFFFFFFu, ws, o);
				if (lane >= o) ws += y;
			}
			npieces = __shfl_sync(0xFFFFFFFFu, ws, NWARPS - 1);
			const int before = __shfl_sync(0xFFFFFFFFu, ws, warp > 0 ? warp - 1 : 0);
			base = x - mycount + (warp > 0 ? before : 0);
		}
		/* the planes are no longer needed: their memory becomes the piece list; the document-start bits are cleared for the next tile */
		if (!GENERAL)
			for (int w = tid; w < JTK_MASK_WORDS; w += NT) S.dmask[w] = 0;
		if (tid < TC) S.chunk_pref[tid] = (uint32_t) base;
		{
			/* bit 15 of an entry: the start is a safe cut, not a piece start of the split pattern (region indices need 13 bits) */
			const uint32_t cut16 = (any_cuts && tid < TC) ? bits & ~(uint32_t) reinterpret_cast<const uint16_t *>(S.rmask)[BH / 16 + tid] : 0u;
			for (uint32_t m = bits; m;) {
				const int i = __ffs((int) m) - 1;
				m &= m - 1;
				S.plist[base++] = (uint16_t) ((BH + tid * 16 + i) | (((cut16 >> i) & 1u) << 15));
			}
		}
		if (tid == 0) {
			/* first piece start of the tile (the end of the input counts), for the long-piece bounds kernel */
			const int fb = next_bit(S.bmask, BH, BH + JTK_TILE - 1);
			a.tile_first_b[tile] = fb < 0 ? -1 : c.g0 + fb;
			a.npieces[tile] = npieces;
		}
		__syncthreads();

		/* ---- P4b: one piece per thread and round.  Pass 1: keys of up to eight bytes probe the first half of their slot
		 * (whole-piece fast path, GptBytePairEncoding.java:81-83).  Longer keys, misses and unusual pieces go to three short
		 * lists PER WARP (counts live in registers: no atomics, no block barrier), which the warp then works off densely:
		 * pass 2 = full-key probe of the 9..24-byte keys, pass 3 = memo of this call / queues of the merge kernels ---- */
		int32_t *rec = a.rec + lt * (long long) RECN;
		int hits = 0;
		uint16_t *const wkeys = S.deflist + warp * WKEYS, *const wodd = S.oddlist + warp * WODD, *const wmiss = S.misslist + warp * WMISS;
		int nkeys = 0, nodd = 0, nmiss = 0; /* warp-uniform */
		/* the lists are worked off 32 entries at a time as soon as 32 have come together (always from the top), so they stay short */
		auto drain_miss = [&](int cnt) { /* pass 3 on the top cnt <= 32 entries: memo of this call, else the merge kernels' queues */
			__syncwarp();
			if (lane < cnt) {
				const int q = wmiss[nmiss - cnt + lane];
				const int r = S.plist[q] & PLIST_POS;
				hits += split_resolve_miss(a, sb, q, r, (int) (S.plist[q + 1] & PLIST_POS) - r, lt);
			}
			nmiss -= cnt;
			__syncwarp();
		};
		auto drain_keys = [&](int cnt) { /* pass 2 on the top cnt <= 32 entries: full-key probe of the 9..24-byte keys */
			__syncwarp();
			bool miss = false;
			int q = 0;
			if (lane < cnt) {
				q = wkeys[nkeys - cnt + lane];
				const int r = S.plist[q] & PLIST_POS, n = (int) (S.plist[q + 1] & PLIST_POS) - r;
				uint32_t key[6];
				jtk_build_key(sb + r, n, key);
				const int32_t out = jtk_lookup_a(T, key, (uint32_t) n, jtk_hash6(key, (uint32_t) n));
				if (out != JTK_RANK_MAX) {
					rec[q] = out;
					hits++;
				} else {
					miss = true;
				}
			}
			nkeys -= cnt;
			const unsigned mm = __ballot_sync(0xFFFFFFFFu, miss);
			if (miss) wmiss[nmiss + __popc(mm & ((1u << lane) - 1u))] = (uint16_t) q; /* (nmiss < 32 here: room for 32 more) */
			nmiss += __popc(mm);
			if (nmiss >= 32) drain_miss(32);
		};
		for (int q0 = 0; q0 < npieces; q0 += NT) {
			const int q = q0 + tid;
			int kind = 0; /* 1: key of 9..24 bytes, 2: unusual, 3: miss of the short probe */
			if (q < npieces) {
				const int pr = S.plist[q], pe = (q + 1 < npieces) ? (int) S.plist[q + 1] : -1; /* the tile's last piece ends in the halo: unusual */
				const int r = pr & PLIST_POS, e = pe < 0 ? -1 : (pe & PLIST_POS);
				const int n = e - r;
				/* a segment between safe cuts is not a piece of the split pattern: no whole-piece shortcut for it (:81-83 looks the PIECE up);
				 * a single byte is its own token either way */
				const bool part = n != 1 && pe >= 0 && ((pr | pe) & PLIST_CUT);
				if (GENERAL && ((S.dmask[r >> 5] >> (r & 31)) & 1u)) {
					rec[q] = REC_BASE + (int32_t) REC_SKIP;
				} else if (part) {
					kind = n <= JTK_INLINE_KEY_MAX ? 3 : 2;
				} else if ((unsigned) (n - 1) < 8u) {
					const uint32_t *aw = reinterpret_cast<const uint32_t *>(sb + (r & ~3));
					const int sh = (r & 3) * 8;
					const uint32_t a0 = aw[0], a1 = aw[1], a2 = aw[2];
					uint32_t k0 = __funnelshift_r(a0, a1, sh), k1 = __funnelshift_r(a1, a2, sh);
					const uint32_t keep = 0xFFFFFFFFu >> ((32 - 8 * n) & 31); /* n = 4 or 8: all four bytes */
					if (n <= 4) {
						k0 &= keep;
						k1 = 0;
					} else {
						k1 &= keep;
					}
					const int32_t out = jtk_lookup_a8(T, k0, k1, (uint32_t) n, jtk_hash6_short(k0, k1, (uint32_t) n));
					if (out != JTK_RANK_MAX) {
						rec[q] = out;
						hits++;
					} else {
						kind = n == 1 ? 2 : 3; /* (a single byte outside the vocabulary is an error of its document: unusual) */
					}
				} else {
					kind = (unsigned) (n - 9) <= (unsigned) (JTK_INLINE_KEY_MAX - 9) ? 1 : 2;
				}
			}
			const unsigned any = __ballot_sync(0xFFFFFFFFu, kind != 0);
			if (any) {
				const unsigned m1 = __ballot_sync(0xFFFFFFFFu, kind == 1), m3 = __ballot_sync(0xFFFFFFFFu, kind == 3), m2 = any & ~(m1 | m3);
				const unsigned below = (1u << lane) - 1u;
				bool placed = true;
				if (kind == 1) {
					wkeys[nkeys + __popc(m1 & below)] = (uint16_t) q; /* (nkeys, nmiss < 32 before the append: always room) */
				} else if (kind == 3) {
					wmiss[nmiss + __popc(m3 & below)] = (uint16_t) q;
				} else if (kind == 2) {
					const int sl = nodd + __popc(m2 & below);
					placed = sl < WODD;
					if (placed) wodd[sl] = (uint16_t) q;
				}
				nkeys += __popc(m1);
				nmiss += __popc(m3);
				nodd += __popc(m2);
				if (!placed) hits += split_slow_piece<GENERAL>(a, sb, q, npieces, lt, tb); /* list full: resolve in place */
				if (nmiss >= 32) drain_miss(32);
				if (nkeys >= 32) drain_keys(32);
			}
		}
		if (nkeys) drain_keys(nkeys);
		/* the unusual pieces */
		__syncwarp();
		nodd = min(nodd, WODD);
		for (int i = lane; i < nodd; i += 32) hits += split_slow_piece<GENERAL>(a, sb, wodd[i], npieces, lt, tb);
		if (nmiss) drain_miss(nmiss);
		hits = __reduce_add_sync(0xFFFFFFFFu, hits);
		if (lane == 0 && hits) atomicAdd(&misc[M_HITS], (uint32_t) hits);
		/* tile-local piece index of the documents that start in this tile (the end of the input included);
		 * jtk_gather_kernel turns it into a token offset */
		if (a.tok_off) {
			for (int64_t d = first_doc + tid; d <= a.ndocs; d += NT) {
				const int64_t g = a.doc_off[d];
				if (g >= tb + JTK_TILE) break;
				if (g < tb) continue;
				const int s = (int) (g - tb);
				const int ch = s >> 4;
				uint32_t m = reinterpret_cast<const uint16_t *>(S.bmask)[BH / 16 + ch] & ((1u << (s & 15)) - 1u);
				const int64_t gbase = tb + ch * 16;
				if (gbase + 16 > a.total) m &= (1u << (int) (a.total - gbase)) - 1u;
				a.tok_off[d] = (int64_t) S.chunk_pref[ch] + __popc(m);
			}
		}
		__syncthreads();
		/* ---- the tile's counters go out, the next tile begins (its ticket was taken at the top) ---- */
		if (tid <= JTK_SHORT_PIECE && misc[M_HIST + tid]) {
			atomicAdd(&a.sub->short_cnt[tid], misc[M_HIST + tid]);
			misc[M_HIST + tid] = 0;
		}
		if (tid == NT - 1) {
			a.nslow[tile] = (int32_t) misc[M_NSLOW];
			a.tile_count[tile] = (int32_t) misc[M_HITS];
			a.tile_slow_used[tile] = (int32_t) misc[M_SLOWTOK];
			misc[M_NSLOW] = 0;
			misc[M_HITS] = 0;
			misc[M_SLOWTOK] = 0;
		}
		cur = nxt;
		nxt = misc[M_TICKET + buf];
	}
}

/* ---------------------------------------------------------------------------------------------
 * kernels 0 (general patterns only): Matcher.find() over every document with the compiled backtracking program, in the
 * sliced form described in jtk_regex.h: every 512-byte slice is matched speculatively from its first byte (threads take
 * slices by ticket), one thread per document then follows the true chain from slice to slice, re-matching only until it
 * meets the speculative trail, and a last pass turns match start / end bits into piece-start and gap bits.
 * ------------------------------------------------------------------------------------------- */
__device__ __forceinline__ jtk_rx_split_buffers rx_buffers(const jtk_encode_args &a) {
	jtk_rx_split_buffers B;
	B.ms = a.rx_start;
	B.me = a.rx_skip;
	B.s_ms = a.rx_spec;
	B.s_me = a.rx_spec + a.rx_words;
	B.s_from = a.rx_spec + 2 * a.rx_words;
	B.exit_slice = a.rx_rec;
	B.last_ms = a.rx_rec + a.rx_slices;
	B.last_me = a.rx_rec + 2 * a.rx_slices;
	B.join = a.rx_rec + 3 * a.rx_slices;
	B.exit_doc = a.rx_rec + 4 * a.rx_slices;
	B.nwords = a.rx_words;
	B.nslices = a.rx_slices;
	return B;
}

struct rx_atomic_or {
	__device__ void operator()(uint32_t *w, uint32_t m) const { atomicOr(w, m); }
};

/* the DFA's hot entries staged in shared memory: the whole transition table when it fits (the usual case: tens of states x tens of
 * classes), and the class of every code point below U+0100 */
constexpr int RX_DFA_SMEM = 8192;
__device__ __forceinline__ void rx_stage_dfa(jtk_rx_program &P, const jtk_tables &T, uint16_t *s_trans, uint8_t *s_ascii) {
	if (!P.dfa_trans) return;
	const int n = T.rx_dfa_nstates * T.rx_dfa_nsym;
	if (n <= RX_DFA_SMEM) {
		for (int i = threadIdx.x; i < n; i += blockDim.x) s_trans[i] = P.dfa_trans[i];
		P.dfa_trans = s_trans;
	}
	for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ascii[i] = P.dfa_ascii[i];
	P.dfa_ascii = s_ascii;
	__syncthreads();
}

#ifndef JTK_RX_SLICE_CTAS
#define JTK_RX_SLICE_CTAS 8 /* resident CTAs per SM the slice kernel is compiled for (64 registers) */
#endif
__global__ void __launch_bounds__(128, JTK_RX_SLICE_CTAS) jtk_general_slice_kernel(const __grid_constant__ jtk_encode_args a) {
	__shared__ uint16_t s_trans[RX_DFA_SMEM];
	__shared__ uint8_t s_ascii[256];
	jtk_rx_program P = jtk_rx_program_of(a.T);
	rx_stage_dfa(P, a.T, s_trans, s_ascii);
	const jtk_rx_split_buffers B = rx_buffers(a);
	/* eight times the threads of the per-document pass on the same stack memory: what overflows a small stack is redone there */
	jtk_rx_frame *st = static_cast<jtk_rx_frame *>(a.rx_stacks) + (size_t) (blockIdx.x * blockDim.x + threadIdx.x) * JTK_RX_STACK_SMALL;
	/* a warp takes 32 consecutive slices at a time: its lanes start their (equally long) slices together and stay in step in the matching
	 * loop instead of drifting apart over tickets taken one by one; neighbouring lanes write neighbouring words of the bit arrays */
	const int lane = threadIdx.x & 31;
	for (;;) {
		unsigned first = 0;
		if (lane == 0) first = atomicAdd(&a.hdr->rx_ticket, 32u);
		first = __shfl_sync(0xFFFFFFFFu, first, 0);
		if ((int64_t) first >= a.rx_slices) break;
		const int64_t slice = (int64_t) first + lane;
		int64_t bad = -1;
		if (slice < a.rx_slices) jtk_rx_slice_pass(P, a.T, a.bytes, a.total, a.doc_off, a.ndocs, slice, B, st, JTK_RX_STACK_SMALL, &bad, rx_atomic_or());
		__syncwarp();
	}
}

__global__ void __launch_bounds__(128) jtk_general_stitch_kernel(const __grid_constant__ jtk_encode_args a) {
	__shared__ uint16_t s_trans[RX_DFA_SMEM];
	__shared__ uint8_t s_ascii[256];
	jtk_rx_program P = jtk_rx_program_of(a.T);
	rx_stage_dfa(P, a.T, s_trans, s_ascii);
	const jtk_rx_split_buffers B = rx_buffers(a);
	jtk_rx_frame *st = static_cast<jtk_rx_frame *>(a.rx_stacks) + (size_t) (blockIdx.x * blockDim.x + threadIdx.x) * JTK_RX_STACK;
	for (;;) {
		const int64_t d = (int64_t) atomicAdd(&a.hdr->rx_ticket2, 1u);
		if (d >= a.ndocs) break;
		if (a.doc_off[d + 1] == a.doc_off[d]) continue;
		if (!jtk_rx_stitch_doc(P, a.T, a.bytes, a.total, a.doc_off, d, B, st, JTK_RX_STACK, rx_atomic_or()) && a.doc_status) atomicOr(a.doc_status + d, JTK_DOC_PATTERN_STACK);
	}
}

__global__ void jtk_general_finish_kernel(const jtk_encode_args a) {
	const jtk_rx_split_buffers B = rx_buffers(a);
	for (int64_t w = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; w < a.rx_words; w += (int64_t) gridDim.x * blockDim.x) jtk_rx_finish_word(B, w, a.total);
}

/* ---------------------------------------------------------------------------------------------
 * kernels 2a: bytePairMerge for the short pieces (2..32 bytes) the lookup did not resolve
 * (GptBytePairEncoding.java:85, 200-300).  The unresolved short pieces of the whole sub-batch are counting-sorted by
 * length into one global list (offsets kernel + scatter kernel), then persistent warps take 32 consecutive entries
 * at a time, one thread per piece: every warp is full and its threads run merge loops of the same length.
 * Tokens go to the tile's slice of slowtok at the piece's byte position.
 * ------------------------------------------------------------------------------------------- */
__global__ void jtk_short_offsets_kernel(const jtk_encode_args a) {
	if (threadIdx.x != 0) return;
	unsigned run = 0;
	for (int n = 0; n <= JTK_SHORT_PIECE; n++) {
		a.sub->short_base[n] = run;
		a.sub->short_cur[n] = run;
		run += a.sub->short_cnt[n];
	}
	a.sub->short_base[JTK_SHORT_PIECE + 1] = run;
	a.sub->short_next[0] = 0;
	a.sub->short_next[1] = a.sub->short_base[JTK_SHORT_PIECE < 17 ? JTK_SHORT_PIECE + 1 : 17];
	a.sub->short_next[2] = a.sub->short_base[JTK_SHORT_PIECE < 33 ? JTK_SHORT_PIECE + 1 : 33];
}

constexpr int SNT = 128;
constexpr int STW = SNT / 32; /* tiles per CTA: one warp per tile (a tile queues a few dozen pieces: a whole CTA per tile mostly idled) */

__global__ void __launch_bounds__(SNT) jtk_short_scatter_kernel(const __grid_constant__ jtk_encode_args a) {
	__shared__ unsigned s_hist[STW][JTK_SHORT_PIECE + 1], s_cur[STW][JTK_SHORT_PIECE + 1];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const long long tile = a.tile_begin + (long long) blockIdx.x * STW + warp;
	if (tile >= a.tile_end) return;
	const int S = a.nslow[tile];
	if (S == 0) return;
	const long long lt = tile - a.tile_begin;
	const int32_t *rec = a.rec + lt * (long long) RECN;
	const uint16_t *sq = a.slowq + lt * (long long) QCAP;
	unsigned *hist = s_hist[warp], *cur = s_cur[warp];
	for (int n = lane; n <= JTK_SHORT_PIECE; n += 32) hist[n] = 0;
	__syncwarp();
	for (int k = lane; k < S; k += 32) atomicAdd(&hist[(rec_payload(rec[sq[k]]) & 0x7FFu) + 1], 1u);
	__syncwarp();
	for (int n = lane; n <= JTK_SHORT_PIECE; n += 32) cur[n] = hist[n] ? atomicAdd(&a.sub->short_cur[n], hist[n]) : 0u; /* reserve a range per length */
	__syncwarp();
	for (int k = lane; k < S; k += 32) {
		const int q = sq[k];
		const int n = (int) (rec_payload(rec[q]) & 0x7FFu) + 1;
		a.shortlist[atomicAdd(&cur[n], 1u)] = ((uint32_t) lt << 14) | (uint32_t) q;
	}
}

/* NSLOT = 16 / 32 / 64: pieces of 2..16 / 17..32 / 33..64 bytes; NSLOT * NTHREADS is constant (32 KiB of scratch) */
template <int NSLOT, int NTHREADS>
__global__ void __launch_bounds__(NTHREADS, NTHREADS >= 256 ? 4 : 7) jtk_merge_short_kernel(const __grid_constant__ jtk_encode_args a) {
	__shared__ int32_t s_scr[2 * NSLOT * NTHREADS]; /* tok / rk, slot k of thread t at k * NTHREADS + t (bank = thread) */
	const jtk_tables &T = a.T;
	const int tid = threadIdx.x, lane = tid & 31;
	const bool write_tok = !(a.flags & JTK_COUNT_ONLY);
	const unsigned end = a.sub->short_base[(NSLOT < JTK_SHORT_PIECE ? NSLOT : JTK_SHORT_PIECE) + 1];
	unsigned *cursor = &a.sub->short_next[NSLOT == 16 ? 0 : NSLOT == 32 ? 1 : 2];
	int32_t *tk = s_scr + tid, *rk = s_scr + NSLOT * NTHREADS + tid;
	for (;;) {
		unsigned idx = 0;
		if (lane == 0) idx = atomicAdd(cursor, 32u);
		idx = __shfl_sync(0xFFFFFFFFu, idx, 0) + lane;
		if (idx - lane >= end) break;
		if (idx < end) {
			const uint32_t e = a.shortlist[idx];
			const long long lt = e >> 14;
			const int q = (int) (e & 0x3FFFu);
			int32_t *rec = a.rec + lt * (long long) RECN;
			const uint32_t pl = rec_payload(rec[q]);
			const int s = (int) (pl >> 11), n = (int) (pl & 0x7FFu) + 1;
			const int64_t tb = (a.tile_begin + lt) * (int64_t) JTK_TILE;
			bool unk = false;
			const int cnt = NSLOT <= 32 ? jtk_merge_short_t<uint32_t>(T, a.bytes + tb + s, n, tk, rk, NTHREADS, &unk)
			                            : jtk_merge_short_t<uint64_t>(T, a.bytes + tb + s, n, tk, rk, NTHREADS, &unk);
			const int off = atomicAdd(&a.tile_slow_used[a.tile_begin + lt], cnt); /* dense area of the tile's slowtok slice */
			if (write_tok) {
				int32_t *stok = a.slowtok + lt * (long long) RECN + off;
				for (int k = 0; k < cnt; k++) stok[k] = tk[k * NTHREADS];
			}
			rec[q] = rec_make(off, cnt);
			atomicAdd(&a.tile_count[a.tile_begin + lt], cnt);
			if (unk) flag_doc(a, tb + s, JTK_DOC_UNKNOWN_BYTES);
			if (NSLOT == 16 && a.memo && !unk && cnt <= JTK_MEMO_MAX_TOKENS) { /* remember the result for later occurrences in this call */
				uint32_t key[6] = {0, 0, 0, 0, 0, 0};
				const uint8_t *pb = a.bytes + tb + s;
				for (int k = 0; k < n; k++) key[k >> 2] |= (uint32_t) pb[k] << (8 * (k & 3));
				jtk_memo_entry *me = a.memo + ((jtk_hash6(key, (uint32_t) n) * 0x9E3779B1u) >> 8 & a.memo_mask);
				const uint32_t old = me->meta;
				if ((old >> 8) != a.memo_epoch && atomicCAS(&me->meta, old, (a.memo_epoch << 8) | 1u) == old) {
					me->key[0] = key[0];
					me->key[1] = key[1];
					me->key[2] = key[2];
					me->key[3] = key[3];
					me->n_cnt = (uint32_t) n | ((uint32_t) cnt << 8);
					for (int k = 0; k < cnt; k++) me->tok[k] = tk[k * NTHREADS];
					__threadfence();
					me->meta = (a.memo_epoch << 8) | 2u;
				}
			}
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * kernel 2b: bytePairMerge for the medium pieces (33..JTK_LONG_PIECE bytes) of the whole sub-batch, taken from two
 * global lists by persistent lane groups: 8 lanes per piece up to 256 bytes, a full warp above.
 * ------------------------------------------------------------------------------------------- */
constexpr int GNTM = 128;

__global__ void __launch_bounds__(GNTM) jtk_merge_medium_kernel(const __grid_constant__ jtk_encode_args a) {
	__shared__ int32_t s_scr[(GNTM / 32) * 2 * JTK_LONG_PIECE]; /* per warp: tok[1024] rk[1024]; per 8-lane group a quarter of each */
	const jtk_tables &T = a.T;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const bool write_tok = !(a.flags & JTK_COUNT_ONLY);
	int32_t *wtok = s_scr + warp * 2 * JTK_LONG_PIECE, *wrk = wtok + JTK_LONG_PIECE;
	/* 8-lane groups */
	{
		const unsigned n8 = a.sub->n_med8;
		const int g = lane >> 3, gl = lane & 7;
		const unsigned gmask = 0xFFu << (g * 8);
		int32_t *tk = wtok + g * JTK_GROUP8_PIECE, *rk = wrk + g * JTK_GROUP8_PIECE;
		for (;;) {
			unsigned idx = 0;
			if (gl == 0) idx = atomicAdd(&a.sub->cursor8, 1u);
			idx = __shfl_sync(gmask, idx, g * 8);
			if (idx >= n8) break;
			const uint32_t e = a.med8[idx];
			const long long lt = e >> 14;
			const int q = (int) (e & 0x3FFFu);
			int32_t *rec = a.rec + lt * (long long) RECN;
			const uint32_t pl = rec_payload(rec[q]);
			const int s = (int) (pl >> 11), n = (int) (pl & 0x7FFu) + 1;
			const int64_t tb = (a.tile_begin + lt) * (int64_t) JTK_TILE;
			bool unk = false;
			const int cnt = merge_group<3>(T, a.bytes + tb + s, n, tk, rk, &unk);
			__syncwarp(gmask);
			int off = 0;
			if (gl == 0) off = atomicAdd(&a.tile_slow_used[a.tile_begin + lt], cnt);
			off = __shfl_sync(gmask, off, g * 8);
			if (write_tok) {
				int32_t *stok = a.slowtok + lt * (long long) RECN + off;
				for (int k = gl; k < cnt; k += 8) stok[k] = tk[k];
			}
			__syncwarp(gmask);
			if (gl == 0) {
				rec[q] = rec_make(off, cnt);
				atomicAdd(&a.tile_count[a.tile_begin + lt], cnt);
				if (unk) flag_doc(a, tb + s, JTK_DOC_UNKNOWN_BYTES);
			}
		}
	}
	__syncwarp();
	/* whole warps */
	{
		const unsigned n32 = a.sub->n_med32;
		for (;;) {
			unsigned idx = 0;
			if (lane == 0) idx = atomicAdd(&a.sub->cursor32, 1u);
			idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
			if (idx >= n32) break;
			const uint32_t e = a.med32[idx];
			const long long lt = e >> 14;
			const int q = (int) (e & 0x3FFFu);
			int32_t *rec = a.rec + lt * (long long) RECN;
			const uint32_t pl = rec_payload(rec[q]);
			const int s = (int) (pl >> 11), n = (int) (pl & 0x7FFu) + 1;
			const int64_t tb = (a.tile_begin + lt) * (int64_t) JTK_TILE;
			bool unk = false;
			const int cnt = merge_group<5>(T, a.bytes + tb + s, n, wtok, wrk, &unk);
			__syncwarp();
			int off = 0;
			if (lane == 0) off = atomicAdd(&a.tile_slow_used[a.tile_begin + lt], cnt);
			off = __shfl_sync(0xFFFFFFFFu, off, 0);
			if (write_tok) {
				int32_t *stok = a.slowtok + lt * (long long) RECN + off;
				for (int k = lane; k < cnt; k += 32) stok[k] = wtok[k];
			}
			__syncwarp();
			if (lane == 0) {
				rec[q] = rec_make(off, cnt);
				atomicAdd(&a.tile_count[a.tile_begin + lt], cnt);
				if (unk) flag_doc(a, tb + s, JTK_DOC_UNKNOWN_BYTES);
			}
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * kernel 3: exclusive scan of the per-tile token counts of the sub-batch, continuing the batch total
 * ------------------------------------------------------------------------------------------- */
__global__ void __launch_bounds__(1024, 1) jtk_tile_scan_kernel(const jtk_encode_args a) {
	__shared__ long long s_w[32];
	__shared__ long long s_carry;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_carry = (long long) a.hdr->total_tokens;
	__syncthreads();
	constexpr int TSI = 8; /* consecutive tiles per thread and round (one 32-byte sector of counts): eight times fewer rounds of barriers */
	for (int64_t base = a.tile_begin; base < a.tile_end; base += 1024 * TSI) {
		const int64_t i0 = base + (int64_t) tid * TSI;
		int c[TSI];
		long long val = 0;
#pragma unroll
		for (int k = 0; k < TSI; k++) {
			c[k] = i0 + k < a.tile_end ? a.tile_count[i0 + k] : 0;
			val += c[k];
		}
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
		long long run = excl;
#pragma unroll
		for (int k = 0; k < TSI; k++) {
			if (i0 + k < a.tile_end) a.tile_base[i0 + k] = run;
			run += c[k];
		}
		__syncthreads();
		if (tid == 1023) s_carry = excl + val;
		__syncthreads();
	}
	if (tid == 0) {
		a.tile_base[a.tile_end] = s_carry;
		a.hdr->total_tokens = (unsigned long long) s_carry;
		a.sub->ticket = 0; /* the next sub-batch starts its tickets and medium-piece lists at zero */
		a.sub->n_med8 = a.sub->n_med32 = a.sub->cursor8 = a.sub->cursor32 = 0;
		for (int n = 0; n <= JTK_SHORT_PIECE; n++) a.sub->short_cnt[n] = 0;
		if (!(a.flags & JTK_COUNT_ONLY) && a.ids && s_carry > a.ids_cap) a.hdr->overflow = 1;
	}
}

/* ---------------------------------------------------------------------------------------------
 * kernel 4: ids to their final position (one CTA per tile) + document token offsets
 * ------------------------------------------------------------------------------------------- */
#ifndef JTK_GNT
#define JTK_GNT 256
#endif
constexpr int GNT = JTK_GNT;
constexpr int GIPT = 8;     /* consecutive pieces per thread and round: one block scan per 2048 pieces */
constexpr int GCAP = 4096;  /* tokens staged in shared memory at a time for coalesced stores */

__device__ __forceinline__ int rec_count(int32_t r) {
	if (rec_is_id(r)) return 1;
	const uint32_t pl = rec_payload(r);
	if (pl & (REC_LONG | REC_SKIP)) return 0;
	return (pl & REC_MEMO) ? (int) (pl & 15u) + 1 : (int) (pl & 0x7FFu) + 1;
}
/* where the tokens of a merged / memoised piece are: REC_BASE + index into the tile's slowtok slice, or REC_BASE + REC_MEMO + entry << 4 */
__device__ __forceinline__ int32_t rec_token_source(int32_t r) {
	const uint32_t pl = rec_payload(r);
	return REC_BASE + (int32_t) ((pl & REC_MEMO) ? (pl & ~15u) : ((pl >> 11) & 0x3FFFu));
}

__global__ void __launch_bounds__(GNT, 2048 / GNT) jtk_gather_kernel(const __grid_constant__ jtk_encode_args a) {
	__shared__ __align__(16) int32_t s_tok[GCAP];          /* the round's tokens in order */
	__shared__ uint16_t s_gpref[RECN / GIPT + 2];            /* tokens of the tile before each group of GIPT pieces */
	__shared__ int s_w[GNT / 32];
	const int tid = threadIdx.x;
	const long long tile = a.tile_begin + blockIdx.x;
	if (tile >= a.tile_end) return;
	const long long lt = tile - a.tile_begin;
	const int64_t tb = tile * (int64_t) JTK_TILE;
	const int P = a.npieces[tile];
	const long long base = a.tile_base[tile];
	const int32_t *rec = a.rec + lt * (long long) RECN;
	const int32_t *stok = a.slowtok + lt * (long long) RECN;
	const bool write_ids = !(a.flags & JTK_COUNT_ONLY) && a.ids != nullptr && !a.hdr->overflow;
	int32_t *dst = a.ids + base;
	int carry = 0;
	for (int q0 = 0; q0 < P; q0 += GNT * GIPT) {
		const int qb = q0 + tid * GIPT;
		int32_t r[GIPT];
		int cnt[GIPT];
		if (qb + GIPT <= P) { /* the tile's slice of rec is 16-byte aligned and qb a multiple of 8 */
			const int4 v0 = *reinterpret_cast<const int4 *>(rec + qb), v1 = *reinterpret_cast<const int4 *>(rec + qb + 4);
			r[0] = v0.x, r[1] = v0.y, r[2] = v0.z, r[3] = v0.w, r[4] = v1.x, r[5] = v1.y, r[6] = v1.z, r[7] = v1.w;
		} else {
#pragma unroll
			for (int j = 0; j < GIPT; j++) r[j] = qb + j < P ? rec[qb + j] : (int32_t) (REC_BASE + (int32_t) REC_SKIP);
		}
		int mine = 0;
		bool any_empty = false; /* a record without tokens: a long piece (placed later), a gap, or the padding after the tile's last piece */
#pragma unroll
		for (int j = 0; j < GIPT; j++) {
			cnt[j] = rec_count(r[j]);
			mine += cnt[j];
			any_empty |= cnt[j] == 0;
		}
		int round_total;
		int excl = carry + block_exclusive_scan<GNT>(mine, s_w, &round_total);
		if (qb < P) s_gpref[qb / GIPT] = (uint16_t) excl;
		if (any_empty)
#pragma unroll
		for (int j = 0; j < GIPT; j++)
			if (!cnt[j] && qb + j < P && !rec_is_id(r[j]) && (rec_payload(r[j]) & REC_LONG)) { /* (no tokens here: rare, tested first) */
				int before = excl - carry;
				for (int i = 0; i < j; i++) before += cnt[i];
				a.long_list[rec_payload(r[j]) & 0x0FFFFFFFu].insert_at = base + carry + before;
			}
		if (write_ids) {
			/* stage the round's tokens in windows of GCAP (one window unless the tile is unusually token-dense), store coalesced */
			for (int w0 = 0; w0 < round_total; w0 += GCAP) {
				int pos = excl - carry - w0; /* window-relative position of this thread's first token */
				if (round_total <= GCAP) { /* the usual case, one window: nothing to clip */
#pragma unroll
					for (int j = 0; j < GIPT; j++) {
						if (rec_is_id(r[j])) {
							s_tok[pos] = r[j];
						} else if (cnt[j]) {
							/* merged / memoised piece: two to four tokens as a rule - straight-line stores for those, a loop for the rest */
							int32_t *o = s_tok + pos;
							const int32_t src = rec_token_source(r[j]);
							const int n = cnt[j];
							o[0] = src;
							if (n > 1) o[1] = src + 1;
							if (n > 2) {
								o[2] = src + 2;
								if (n > 3) o[3] = src + 3;
								for (int k = 4; k < n; k++) o[k] = src + k;
							}
						}
						pos += cnt[j];
					}
				} else
#pragma unroll
				for (int j = 0; j < GIPT; j++) {
					if (cnt[j] && pos + cnt[j] > 0 && pos < GCAP) {
						if (rec_is_id(r[j])) {
							s_tok[pos] = r[j];
						} else {
							/* merged / memoised piece: only note WHERE its tokens are (REC_BASE + index into the tile's slowtok slice);
							 * the loads happen in the coalesced copy loop below, all in flight together, instead of one dependent
							 * load per iteration of this divergent loop */
							const int k0 = pos < 0 ? -pos : 0, k1 = min(cnt[j], GCAP - pos); /* the part of the piece inside the window */
							int32_t *o = s_tok + pos;
							const int32_t src = rec_token_source(r[j]);
							for (int k = k0; k < k1; k++) o[k] = src + k;
						}
					}
					pos += cnt[j];
				}
				__syncthreads();
				const int nw = min(GCAP, round_total - w0);
				for (int k = tid; k < nw; k += 4 * GNT) { /* four tokens per thread and step: the dependent loads of merged tokens overlap */
					int32_t v[4];
#pragma unroll
					for (int i = 0; i < 4; i++) v[i] = k + i * GNT < nw ? s_tok[k + i * GNT] : 0;
#pragma unroll
					for (int i = 0; i < 4; i++)
						if (!rec_is_id(v[i])) {
							const uint32_t pl = rec_payload(v[i]);
							v[i] = (pl & REC_MEMO) ? a.memo[(pl & ~REC_MEMO) >> 4].tok[pl & 15u] : stok[pl];
						}
#pragma unroll
					for (int i = 0; i < 4; i++)
						if (k + i * GNT < nw) dst[carry + w0 + k + i * GNT] = v[i];
				}
				if (w0 + GCAP < round_total) __syncthreads();
			}
		}
		carry += round_total;
		__syncthreads(); /* s_tok / s_w are reused by the next round */
	}
	if (a.tok_off) {
		__syncthreads();
		for (int64_t d = a.tile_first_doc[tile] + tid; d <= a.ndocs; d += GNT) {
			const int64_t g = a.doc_off[d];
			if (g >= tb + JTK_TILE) break;
			if (g < tb) continue;
			const int pi = (int) a.tok_off[d]; /* tile-local index of the document's first piece (jtk_split_lookup_kernel) */
			int before = carry;
			if (pi < P) {
				before = s_gpref[pi / GIPT];
				for (int q = pi - pi % GIPT; q < pi; q++) before += rec_count(rec[q]);
			}
			a.tok_off[d] = base + before;
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

/* token offsets of the (empty) documents that start at the very end of the input when it ends on a tile boundary */
__global__ void jtk_finalize_kernel(const jtk_encode_args a) {
	if (!a.tok_off) return;
	const long long total_tokens = (long long) a.hdr->total_tokens;
	const int64_t covered = a.ntiles * (int64_t) JTK_TILE;
	int64_t lo = 0, hi = a.ndocs + 1; /* first d with doc_off[d] >= covered */
	while (lo < hi) {
		int64_t mid = (lo + hi) >> 1;
		if (a.doc_off[mid] >= covered) hi = mid;
		else lo = mid + 1;
	}
	for (int64_t d = lo + blockIdx.x * (int64_t) blockDim.x + threadIdx.x; d <= a.ndocs; d += (int64_t) gridDim.x * blockDim.x) a.tok_off[d] = total_tokens;
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

/* Scratch of the long-piece kernel: eight int32 arrays of `stride` elements (one slot per input byte of all long pieces). */
struct jtk_long_scratch {
	int32_t *base;
	int64_t stride;
};

enum { LST_NONE = 0, LST_UNK = 1, LST_SEL = 2, LST_REJ = 3 }; /* per-part status of the round; bit 2: the pair to the left is selected too */

__global__ void __launch_bounds__(LNT, 1) jtk_long_merge_kernel(const jtk_encode_args a, unsigned int n_long, const jtk_long_scratch scr) {
	__shared__ int32_t s_red[LNT / 32];
	__shared__ int s_scan_run[LNT / 32];
	__shared__ int s_scan_reset[LNT / 32];
	__shared__ int32_t s_min;
	__shared__ int s_flag, s_piece, s_carry_run, s_first, s_m, s_unk;
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
		int32_t *tok = scr.base + lp.scratch, *rk = tok + scr.stride, *nxt = rk + scr.stride, *prv = nxt + scr.stride, *st = prv + scr.stride,
		        *aux0 = st + scr.stride, *aux1 = aux0 + scr.stride, *cand = aux1 + scr.stride;
		bool strict = false;
	restart:
		for (int k = tid; k < n; k += LNT) {
			tok[k] = T.byte_id[p[k]];
			rk[k] = (k + 1 < n) ? T.bytepair[((uint32_t) p[k] << 8) | p[k + 1]] : JTK_RANK_MAX;
			nxt[k] = k + 1;
			prv[k] = k - 1;
			st[k] = LST_NONE;
		}
		__syncthreads();
		/* Rounds.  A round merges, in one go, the pairs the reference loop would merge next: every pair whose rank is at most
		 * T = (current minimum) + window and that survives the (rank, position) priority among adjacent candidates.  That is
		 * exactly the sequential order as long as no pair CREATED by these merges - the final ones and the transient one
		 * between two merges that are one part apart - ranks at or below T; this is checked before anything is changed.
		 * On a failed check the round is redone with window 0 (only the minimum rank: adjacent candidates then form runs of
		 * equal rank, resolved by a parity scan), and if that fails too (a vocabulary in which a concatenation ranks below
		 * its parts) the piece restarts in strict mode: one merge per round, the reference loop verbatim. */
		int wcur = 0, cool = 0, cool_len = 2; /* window; rounds to stay at window 0 after a failed windowed round (doubles per failure) */
		bool force_exact = false;
		for (;;) {
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
					s_m = 0;
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
			const int win = (force_exact || cool > 0) ? 0 : wcur;
			const int32_t thr = (win > 0 && mr < JTK_RANK_MAX - 2 - win) ? mr + win : mr;
			const bool windowed = thr > mr;
			/* candidates: parts whose pair ranks at most thr (removed parts and the last part have rk = JTK_RANK_MAX) */
			for (int k = tid; k < n; k += LNT)
				if (rk[k] <= thr) {
					cand[atomicAdd(&s_m, 1)] = k;
					st[k] = LST_UNK;
				}
			__syncthreads();
			const int m = s_m;
			bool failed = false, chains = false;
			if (windowed) {
				/* priority among adjacent candidates: a candidate is selected iff no adjacent candidate with a smaller (rank, position)
				 * key is selected.  Relaxation; chains of decreasing keys are short unless the ranks are equal (window 0 handles those). */
				for (int it = 0;; it++) {
					if (tid == 0) s_unk = 0;
					__syncthreads();
					int undecided = 0;
					for (int i = tid; i < m; i += LNT) {
						const int k = cand[i];
						if (st[k] != LST_UNK) continue;
						const int32_t r = rk[k];
						const int l = prv[k], rt = nxt[k];
						const int sl = l >= 0 ? (st[l] & 3) : LST_NONE, sr = rt < n ? (st[rt] & 3) : LST_NONE;
						const bool lsm = sl != LST_NONE && rk[l] <= r; /* left neighbour: smaller position, so smaller key on equal ranks */
						const bool rsm = sr != LST_NONE && rk[rt] < r;
						if ((lsm && sl == LST_SEL) || (rsm && sr == LST_SEL)) st[k] = LST_REJ;
						else if ((!lsm || sl == LST_REJ) && (!rsm || sr == LST_REJ)) st[k] = LST_SEL;
						else undecided++;
					}
					undecided = __reduce_add_sync(0xFFFFFFFFu, undecided);
					if (lane == 0 && undecided) atomicAdd(&s_unk, undecided);
					__syncthreads();
					const int unk = s_unk;
					__syncthreads();
					if (!unk) break;
					/* chains of decreasing keys shrink geometrically; a long run of equal ranks (repeated text) loses one part per
					 * iteration: leave those to window 0, which resolves them with a scan */
					if (it >= 40 || (it >= 6 && unk > m / 2)) {
						failed = chains = true;
						break;
					}
				}
			} else {
				/* window 0: in list order, a flagged part (rk == mr) is selected iff an even number of flagged parts directly precede
				 * it.  Blocked scan over positions, LNT * 8 positions per step, removed parts are transparent. */
				for (int base = 0; base < n; base += LNT * 8) {
					const int k0 = base + tid * 8;
					int run = 0, reset = 0; /* summary of this thread's 8 positions */
					int loc_run[8] = {0, 0, 0, 0, 0, 0, 0, 0};
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
					/* scan of (run, reset) across threads */
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
					int pr = s_carry_run;
					if (warp > 0) {
						if (s_scan_reset[warp - 1]) pr = s_scan_run[warp - 1];
						else pr += s_scan_run[warp - 1];
					}
					{
						int er = __shfl_up_sync(0xFFFFFFFFu, xr, 1), es = __shfl_up_sync(0xFFFFFFFFu, xs, 1);
						if (lane > 0) {
							if (es) pr = er;
							else pr += er;
						}
					}
					bool seen_reset = false;
					for (int i = 0; i < 8; i++) {
						const int k = k0 + i;
						if (k >= n || nxt[k] < 0) continue;
						if (rk[k] == mr) {
							const int before = seen_reset ? loc_run[i] : loc_run[i] + pr;
							st[k] = (before & 1) == 0 ? LST_SEL : LST_REJ;
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
			}
			/* check, without changing anything: the ranks of the pairs these merges create */
			if (!failed) {
				for (int i = tid; i < m; i += LNT) {
					const int k = cand[i];
					if ((st[k] & 3) != LST_SEL) continue;
					const int32_t mk = rk[k]; /* rank of the pair == id of the merged token */
					const int j = nxt[k], nn = nxt[j], pv = prv[k];
					int32_t r1 = JTK_RANK_MAX, r0 = JTK_RANK_MAX;
					bool bad = false;
					if (nn < n) {
						int32_t right = tok[nn];
						if ((st[nn] & 3) == LST_SEL) { /* the next pair merges too: between the two merges the pair (merged, its first part) exists */
							if (jtk_lookup_pair(T, mk, right) <= thr) bad = true;
							right = rk[nn];
						}
						r1 = jtk_lookup_pair(T, mk, right);
						if (r1 <= thr) bad = true;
					}
					if (pv >= 0) {
						const int ppv = prv[pv];
						int32_t left = tok[pv];
						if (ppv >= 0 && (st[ppv] & 3) == LST_SEL) { /* pv is the second part of a selected pair */
							if (jtk_lookup_pair(T, left, mk) <= thr) bad = true;
							left = rk[ppv];
							st[k] = LST_SEL | 4;
						}
						r0 = jtk_lookup_pair(T, left, mk);
						if (r0 <= thr) bad = true;
					}
					aux0[k] = r0;
					aux1[k] = r1;
					if (bad) s_flag = 1;
				}
				__syncthreads();
				failed = s_flag != 0;
			}
			if (!failed) {
				/* apply */
				for (int i = tid; i < m; i += LNT) {
					const int k = cand[i];
					const int s0 = st[k];
					if ((s0 & 3) != LST_SEL) continue;
					const int j = nxt[k], nn = nxt[j];
					tok[k] = rk[k];
					rk[j] = DEAD;
					nxt[j] = -1;
					nxt[k] = nn;
					if (nn < n) prv[nn] = k;
					rk[k] = aux1[k];
					if (!(s0 & 4)) { /* otherwise the pair to the left was selected: its own new rank covers this one */
						const int pv = prv[k];
						if (pv >= 0) rk[pv] = aux0[k];
					}
				}
			}
			__syncthreads();
			for (int i = tid; i < m; i += LNT) st[cand[i]] = LST_NONE;
			__syncthreads();
			if (failed) {
				if (windowed) { /* redo the round with the minimum rank only, and stay there for a while */
					wcur >>= 2;
					force_exact = true;
					if (chains) { /* repeated text: windows will keep failing for a while */
						cool_len = min(cool_len * 2, 1 << 20);
						cool = cool_len;
					}
					continue;
				}
				strict = true;
				if (tid == 0) atomicAdd(&a.hdr->violations, 1u);
				__syncthreads();
				goto restart;
			}
			if (windowed) {
				wcur = min(wcur * 2, 1 << 24);
				cool_len = 2;
			} else if (cool > 0) {
				cool--;
			}
			if (wcur < 16) wcur = 16;
			force_exact = false;
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
	cudaError_t e = cudaFuncSetAttribute(jtk_split_lookup_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SPLIT_SMEM_BYTES);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(jtk_split_lookup_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SPLIT_SMEM_BYTES);
	if (e != cudaSuccess) return e;
	/* merge kernels: 75 % shared-memory carve-out = five CTAs of the 16-slot kernel per SM and ~85 KB of L1 for the pair table
	 * (measured: mixed corpus 8.12 -> 7.65 ms per 512 MiB against the default carve-out; 50 % and 25 % are slower) */
	int pct = 75;
	if (const char *env = getenv("JTK_MERGE_CARVEOUT")) pct = atoi(env);
	if (pct > 0) {
		cudaFuncSetAttribute(jtk_merge_short_kernel<16, 256>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(jtk_merge_short_kernel<32, 128>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(jtk_merge_short_kernel<64, 64>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
		cudaFuncSetAttribute(jtk_merge_medium_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
	}
	return cudaSuccess;
}

cudaError_t jtk_launch_tile_first_doc(const int64_t *doc_off, int64_t ndocs, int64_t ntiles, int32_t *out, cudaStream_t st) {
	if (ntiles <= 0) return cudaSuccess;
	jtk_tile_first_doc_kernel<<<(unsigned) ((ntiles + 255) / 256), 256, 0, st>>>(doc_off, ndocs, ntiles, out);
	return cudaGetLastError();
}

static unsigned l2_window_attr(const jtk_encode_args &a, cudaLaunchAttribute *attr) {
	/* the table-probing kernels run with the hot tables pinned in L2 (persisting hits, streaming misses) */
	static const bool off = getenv("JTK_L2_WINDOW") && getenv("JTK_L2_WINDOW")[0] == '0'; /* development switch */
	if (a.l2_bytes == 0 || off) return 0;
	attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
	attr[0].val.accessPolicyWindow.base_ptr = const_cast<void *>(a.l2_base);
	attr[0].val.accessPolicyWindow.num_bytes = a.l2_bytes;
	attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
	attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
	attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	return 1;
}

/* split + whole-piece lookup of the sub-batch [a.tile_begin, a.tile_end).  ctas_per_sm > 0: persistent CTAs taking tiles by
 * ticket; 0: one CTA per tile, so that kernels of the previous sub-batch (higher-priority streams) get SM slots as CTAs retire. */
cudaError_t jtk_launch_split(const jtk_encode_args &a, int num_sms, int ctas_per_sm, cudaEvent_t k0, cudaEvent_t k1, cudaStream_t st) {
	const int64_t nt = a.tile_end - a.tile_begin;
	if (nt <= 0) return cudaSuccess;
	int64_t grid = ctas_per_sm > 0 ? (int64_t) num_sms * JTK_SPLIT_CTAS : nt;
	if (grid > nt) grid = nt;
	cudaLaunchAttribute attr[1];
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.stream = st;
	cfg.attrs = attr;
	cfg.numAttrs = l2_window_attr(a, attr);
	cfg.gridDim = dim3((unsigned) grid);
	cfg.blockDim = dim3(JTK_NT);
	cfg.dynamicSmemBytes = SPLIT_SMEM_BYTES;
	if (k0) cudaEventRecord(k0, st);
	if (a.T.pattern_kind == JTK_PAT_GENERAL) cudaLaunchKernelEx(&cfg, jtk_split_lookup_kernel<true>, a);
	else cudaLaunchKernelEx(&cfg, jtk_split_lookup_kernel<false>, a);
	if (k1) cudaEventRecord(k1, st);
	return cudaGetLastError();
}

/* everything after it: sort the unresolved short pieces by length, merge, scan, gather */
cudaError_t jtk_launch_post(const jtk_encode_args &a, int num_sms, cudaStream_t st, const jtk_side_streams *side, cudaEvent_t *marks) {
	const int64_t nt = a.tile_end - a.tile_begin;
	if (nt <= 0) return cudaSuccess;
	cudaLaunchAttribute attr[1];
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	cfg.attrs = attr;
	cfg.numAttrs = l2_window_attr(a, attr);
	jtk_short_offsets_kernel<<<1, 32, 0, st>>>(a);
	jtk_short_scatter_kernel<<<(unsigned) ((nt + STW - 1) / STW), SNT, 0, st>>>(a);
	if (marks) cudaEventRecord(marks[0], st);
	/* the four merge kernels are independent of each other (own lists, own pieces): longest chains first, side by side */
	if (side) {
		cudaEventRecord(side->fork, st);
		for (int i = 0; i < 3; i++) cudaStreamWaitEvent(side->s[i], side->fork, 0);
	}
	static const int merge_grid = getenv("JTK_MERGE_GRID") ? atoi(getenv("JTK_MERGE_GRID")) : 6; /* persistent CTAs per SM of each merge kernel */
	cfg.gridDim = dim3((unsigned) (num_sms * merge_grid));
	cfg.blockDim = dim3(GNTM);
	cfg.stream = side ? side->s[0] : st;
	cudaLaunchKernelEx(&cfg, jtk_merge_medium_kernel, a);
	cfg.blockDim = dim3(64);
	cfg.stream = side ? side->s[1] : st;
	cudaLaunchKernelEx(&cfg, jtk_merge_short_kernel<64, 64>, a);
	cfg.blockDim = dim3(128);
	cfg.stream = side ? side->s[2] : st;
	cudaLaunchKernelEx(&cfg, jtk_merge_short_kernel<32, 128>, a);
	cfg.blockDim = dim3(256);
	cfg.stream = st;
	cudaLaunchKernelEx(&cfg, jtk_merge_short_kernel<16, 256>, a);
	if (side)
		for (int i = 0; i < 3; i++) {
			cudaEventRecord(side->join[i], side->s[i]);
			cudaStreamWaitEvent(st, side->join[i], 0);
		}
	if (marks) cudaEventRecord(marks[1], st);
	jtk_tile_scan_kernel<<<1, 1024, 0, st>>>(a);
	if (marks) cudaEventRecord(marks[2], st);
	jtk_gather_kernel<<<(unsigned) nt, GNT, 0, st>>>(a);
	if (marks) cudaEventRecord(marks[3], st);
	return cudaGetLastError();
}

cudaError_t jtk_launch_general_split(const jtk_encode_args &a, cudaStream_t st) {
	int64_t blocks = (a.rx_slices + 127) / 128;
	if (blocks > JTK_RX_THREADS / 128 * (JTK_RX_STACK / JTK_RX_STACK_SMALL)) blocks = JTK_RX_THREADS / 128 * (JTK_RX_STACK / JTK_RX_STACK_SMALL);
	if (blocks < 1) blocks = 1;
	jtk_general_slice_kernel<<<(unsigned) blocks, 128, 0, st>>>(a);
	blocks = (a.ndocs + 127) / 128;
	if (blocks > JTK_RX_THREADS / 128) blocks = JTK_RX_THREADS / 128;
	if (blocks < 1) blocks = 1;
	jtk_general_stitch_kernel<<<(unsigned) blocks, 128, 0, st>>>(a);
	blocks = (a.rx_words + 255) / 256;
	if (blocks > 148 * 16) blocks = 148 * 16;
	jtk_general_finish_kernel<<<(unsigned) blocks, 256, 0, st>>>(a);
	return cudaGetLastError();
}

cudaError_t jtk_launch_finalize(const jtk_encode_args &a, cudaStream_t st) {
	jtk_finalize_kernel<<<8, 256, 0, st>>>(a);
	return cudaGetLastError();
}

cudaError_t jtk_launch_long_bounds(const jtk_encode_args &a, unsigned int n_long, cudaStream_t st) {
	jtk_long_bounds_kernel<<<(n_long + 127) / 128, 128, 0, st>>>(a, n_long);
	return cudaGetLastError();
}

cudaError_t jtk_launch_long_merge(const jtk_encode_args &a, unsigned int n_long, int32_t *scratch, int64_t stride, int num_sms, cudaStream_t st) {
	unsigned grid = n_long < (unsigned) num_sms ? n_long : (unsigned) num_sms;
	jtk_long_scratch scr;
	scr.base = scratch;
	scr.stride = stride;
	jtk_long_merge_kernel<<<grid, LNT, 0, st>>>(a, n_long, scr);
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

/* token id -> (offset into dec_bytes, length, first twelve bytes); false for an id that neither map knows (GptBytePairEncoding.java:302-314) */
__device__ __forceinline__ bool decode_entry(const jtk_tables &T, int32_t id, uint32_t *off, uint32_t *len, uint32_t *b0, uint32_t *b1, uint32_t *b2) {
	if (T.dec_direct) { /* ids are small non-negative numbers (all predefined encodings): one 16-byte read answers tokens of up to twelve bytes */
		if ((uint32_t) id >= T.dec_direct_size) return false;
		const uint4 e = __ldg(T.dec_direct + id);
		*off = e.x >> 8;
		*len = e.x & 0xFFu;
		*b0 = e.y, *b1 = e.z, *b2 = e.w;
		return e.x != 0xFFFFFFFFu;
	}
	uint32_t s = jtk_hash_pair(id, 0) & T.mask_d;
	for (;;) {
		const uint32_t v = T.dec_keys[2 * s + 1];
		if (v == 0) return false;
		if (T.dec_keys[2 * s] == (uint32_t) id) {
			*off = T.dec_off[v - 1];
			*len = T.dec_off[v] - *off;
			uint32_t f[3] = {0, 0, 0};
			for (uint32_t k = 0; k < 12 && k < *len; k++) f[k >> 2] |= (uint32_t) T.dec_bytes[*off + k] << (8 * (k & 3));
			*b0 = f[0], *b1 = f[1], *b2 = f[2];
			return true;
		}
		s = (s + 1) & T.mask_d;
	}
}
__device__ __forceinline__ bool decode_len(const jtk_tables &T, int32_t id, uint32_t *len) {
	if (T.dec_direct) {
		if ((uint32_t) id >= T.dec_direct_size) return false;
		const uint32_t x = __ldg(reinterpret_cast<const uint32_t *>(T.dec_direct + id));
		*len = x & 0xFFu;
		return x != 0xFFFFFFFFu;
	}
	uint32_t off, b0, b1, b2;
	return decode_entry(T, id, &off, len, &b0, &b1, &b2);
}

#ifndef JTK_DECODE_CTAS
#define JTK_DECODE_CTAS 6 /* resident CTAs per SM the single-pass decode kernel is compiled for (register budget: 42) */
#endif
constexpr int DNT = 256;           /* threads per CTA of the decode kernels */
#ifndef JTK_DECODE_DPT
#define JTK_DECODE_DPT 8
#endif
constexpr int DPT = JTK_DECODE_DPT; /* consecutive tokens per thread */
constexpr int DTILE = DNT * DPT;   /* tokens per tile */
/* how much of a token's 16-byte table entry stays in registers between the length pass and the copy (JTK_DECODE_KEEP = 4: the length
 * word only, every token is looked up again; 8: + the first four bytes; 16: the whole entry, tokens of up to twelve bytes need no second
 * read): registers (occupancy) against gathers through L1 */
#ifndef JTK_DECODE_KEEP
#define JTK_DECODE_KEEP 4 /* measured on B200, 404 M tokens: 4 at six CTAs per SM 3.90 ms, 8 at eight 4.38 ms, 16 at four 4.26 ms */
#endif
constexpr int DKEEP = JTK_DECODE_KEEP;
#if JTK_DECODE_KEEP == 16
typedef uint4 dec_ent;
__device__ __forceinline__ dec_ent dec_ent_load(const uint4 *p) { return __ldg(p); }
__device__ __forceinline__ dec_ent dec_ent_none(uint32_t x) { return make_uint4(x, 0u, 0u, 0u); }
__device__ __forceinline__ void dec_ent_bytes(const dec_ent &e, uint32_t *f0, uint32_t *f1, uint32_t *f2) { *f0 = e.y, *f1 = e.z, *f2 = e.w; }
#elif JTK_DECODE_KEEP == 8
typedef uint2 dec_ent;
__device__ __forceinline__ dec_ent dec_ent_load(const uint4 *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ dec_ent dec_ent_none(uint32_t x) { return make_uint2(x, 0u); }
__device__ __forceinline__ void dec_ent_bytes(const dec_ent &e, uint32_t *f0, uint32_t *f1, uint32_t *f2) { *f0 = e.y, *f1 = 0, *f2 = 0; }
#else
struct dec_ent {
	uint32_t x;
};
__device__ __forceinline__ dec_ent dec_ent_load(const uint4 *p) { return dec_ent{__ldg(reinterpret_cast<const uint32_t *>(p))}; }
__device__ __forceinline__ dec_ent dec_ent_none(uint32_t x) { return dec_ent{x}; }
__device__ __forceinline__ void dec_ent_bytes(const dec_ent &, uint32_t *f0, uint32_t *f1, uint32_t *f2) { *f0 = *f1 = *f2 = 0; }
#endif
constexpr int SPAD = 16;           /* front pad of the staging window (a flushed 16-byte group may begin up to 15 bytes before the window) */
constexpr int DWIN = 12 * 1024;    /* bytes of a tile staged in shared memory at a time (a tile of typical text is ~5.5 KB) */

/* pass 1: bytes per tile of DTILE tokens; unknown ids are reported per document (the smallest position) */
__global__ void __launch_bounds__(DNT) jtk_decode_count_kernel(const __grid_constant__ jtk_decode_args a) {
	__shared__ int s_w[DNT / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int64_t t0 = (int64_t) blockIdx.x * DTILE + (int64_t) tid * DPT;
	int32_t ids[DPT];
	if (t0 + DPT <= a.nids) { /* tiles start at multiples of 4096 tokens: 16-byte aligned */
#pragma unroll
		for (int i = 0; i < DPT / 4; i++) {
			const int4 v = __ldg(reinterpret_cast<const int4 *>(a.ids + t0) + i);
			ids[4 * i] = v.x, ids[4 * i + 1] = v.y, ids[4 * i + 2] = v.z, ids[4 * i + 3] = v.w;
		}
	} else {
#pragma unroll
		for (int i = 0; i < DPT; i++) ids[i] = t0 + i < a.nids ? a.ids[t0 + i] : 0;
	}
	int sum = 0;
#pragma unroll
	for (int i = 0; i < DPT; i++) {
		if (t0 + i >= a.nids) break;
		uint32_t len;
		if (decode_len(a.T, ids[i], &len)) {
			sum += (int) len;
		} else {
			/* document of token j: the last d with tok_off[d] <= j */
			const int64_t j = t0 + i;
			int64_t lo = 0, hi = a.ndocs - 1;
			while (lo < hi) {
				const int64_t mid = (lo + hi + 1) >> 1;
				if (a.tok_off[mid] <= j) lo = mid;
				else hi = mid - 1;
			}
			atomicMin(a.bad_pos + lo, (unsigned long long) j);
		}
	}
	sum = __reduce_add_sync(0xFFFFFFFFu, sum);
	if (lane == 0) s_w[warp] = sum;
	__syncthreads();
	if (tid == 0) {
		int t = 0;
		for (int w = 0; w < DNT / 32; w++) t += s_w[w];
		a.tile_bytes[blockIdx.x] = t;
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

/* Single pass: tiles of DTILE tokens are taken by ticket (so every predecessor of a tile has started), the tile's byte count is
 * published, the exclusive prefix over all earlier tiles comes from a decoupled look-back over the published words (a warp
 * inspects 32 predecessors at a time), while the token bytes are already being copied into a shared-memory window at their tile-relative
 * position; the flush shifts them to the output's 16-byte grid and writes 16-byte stores.  Algorithmic bytes: 4 per token read, the bytes written.
 * status word of a tile: bits 62..63 = 1 aggregate (bytes of the tile) / 2 inclusive prefix, bits 0..61 the value. */
constexpr unsigned long long DST_AGG = 1ull << 62, DST_PREFIX = 2ull << 62, DST_MASK = (1ull << 62) - 1;

__global__ void __launch_bounds__(DNT, JTK_DECODE_CTAS) jtk_decode_fused_kernel(const __grid_constant__ jtk_decode_args a) {
	__shared__ __align__(16) uint8_t s_out[SPAD + DWIN + 32];
	__shared__ int s_pref[DNT + 1];
	__shared__ int s_w[DNT / 32];
	__shared__ long long s_base;
	__shared__ unsigned s_tile;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
	__syncthreads();
	const int64_t tile = s_tile;
	const int64_t tile0 = tile * DTILE;
	const int64_t t0 = tile0 + (int64_t) tid * DPT;
	int32_t ids[DPT];
	if (t0 + DPT <= a.nids) {
#pragma unroll
		for (int i = 0; i < DPT / 4; i++) {
			const int4 v = __ldg(reinterpret_cast<const int4 *>(a.ids + t0) + i);
			ids[4 * i] = v.x, ids[4 * i + 1] = v.y, ids[4 * i + 2] = v.z, ids[4 * i + 3] = v.w;
		}
	} else {
#pragma unroll
		for (int i = 0; i < DPT; i++) ids[i] = t0 + i < a.nids ? a.ids[t0 + i] : 0;
	}
	/* ONE table read per token: with the direct table the 16-byte entry (length, offset, first twelve bytes) stays in registers across
	 * the scan; without it (ids that are not small non-negative numbers) only the length is kept and the bytes are looked up again */
	const bool direct = a.T.dec_direct != nullptr;
	dec_ent ent[DPT];
	int sum = 0;
#pragma unroll
	for (int i = 0; i < DPT; i++) {
		uint32_t l = 0;
		bool known;
		if (direct) {
			const bool in = t0 + i < a.nids;
			ent[i] = (in && (uint32_t) ids[i] < a.T.dec_direct_size) ? dec_ent_load(a.T.dec_direct + ids[i]) : dec_ent_none(0xFFFFFFFFu);
			known = !in || ent[i].x != 0xFFFFFFFFu;
			l = known && in ? (ent[i].x & 0xFFu) : 0u;
		} else {
			known = !(t0 + i < a.nids) || decode_len(a.T, ids[i], &l);
			if (!known) l = 0;
			ent[i] = dec_ent_none(0u);
		}
		if (!known) {
			/* unknown id: flag its document (the last d with tok_off[d] <= j) and remember the first one */
			const int64_t j = t0 + i;
			int64_t lo = 0, hi = a.ndocs - 1;
			while (lo < hi) {
				const int64_t mid = (lo + hi + 1) >> 1;
				if (a.tok_off[mid] <= j) lo = mid;
				else hi = mid - 1;
			}
			atomicMin(a.bad_pos + lo, (unsigned long long) j);
			atomicOr(a.doc_status + lo, JTK_DOC_UNKNOWN_ID);
		}
		ent[i].x = (ent[i].x & ~0xFFu) | l; /* the length in the low byte either way */
		sum += (int) l;
	}
	int x = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const int y = __shfl_up_sync(0xFFFFFFFFu, x, o);
		if (lane >= o) x += y;
	}
	if (lane == 31) s_w[warp] = x;
	__syncthreads();
	int wpre = 0, total = 0;
#pragma unroll
	for (int w = 0; w < DNT / 32; w++) {
		if (w < warp) wpre += s_w[w];
		total += s_w[w];
	}
	const int mine = wpre + x - sum; /* tile-relative byte offset of this thread's first token */
	s_pref[tid] = mine;
	if (tid == DNT - 1) s_pref[DNT] = total;
	/* ---- publish the aggregate, look back for the exclusive prefix (warp 0) ---- */
	if (warp == 0) {
		volatile unsigned long long *st = a.tile_state;
		if (lane == 0) {
			st[tile] = (tile == 0 ? DST_PREFIX : DST_AGG) | (unsigned long long) total;
			__threadfence();
		}
		long long excl = 0;
		int64_t p = tile - 1; /* predecessor inspected by lane 0 of this round */
		while (p >= 0) {
			const int64_t mine_p = p - lane;
			unsigned long long w = DST_PREFIX; /* lanes before the first tile contribute nothing */
			if (mine_p >= 0) {
				do {
					w = st[mine_p];
				} while ((w >> 62) == 0);
			}
			const unsigned pm = __ballot_sync(0xFFFFFFFFu, (w >> 62) == 2);
			/* everything up to and including the nearest inclusive prefix counts */
			const int stop = pm ? __ffs((int) pm) - 1 : 31;
			long long v = lane <= stop && mine_p >= 0 ? (long long) (w & DST_MASK) : 0;
#pragma unroll
			for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
			excl += v;
			if (pm) break;
			p -= 32;
		}
		if (lane == 0) {
			if (tile > 0) {
				st[tile] = DST_PREFIX | (unsigned long long) (excl + total);
				__threadfence();
			}
			s_base = excl;
			if (tile == a.ntiles - 1) *a.total_out = excl + total;
		}
	}
	/* ---- the token bytes go to the staging window at their TILE-RELATIVE position right away: this needs no global prefix, so
	 * the other warps copy while warp 0 looks back (and warp 0 copies its own tokens afterwards); only the flush waits for `base`.
	 * Staging byte SPAD + k <-> output byte base + w0 + k. ---- */
	long long base = 0;
	bool fits = true;
	for (int w0 = 0; w0 < total; w0 += DWIN) {
		const int wn = min(DWIN, total - w0); /* bytes of this window */
		if (w0) __syncthreads();
		int pos = mine - w0;
#pragma unroll
		for (int i = 0; i < DPT; i++) {
			const int l = (int) (ent[i].x & 0xFFu);
			if (l && pos + l > 0 && pos < wn) {
				uint32_t o2 = ent[i].x >> 8, l2, f0 = 0, f1 = 0, f2 = 0;
				if (direct && l <= DKEEP - 4) dec_ent_bytes(ent[i], &f0, &f1, &f2); /* the bytes came with the length */
				else decode_entry(a.T, ids[i], &o2, &l2, &f0, &f1, &f2);
				if (l <= 12 && pos >= 0 && pos + l <= wn) { /* the table entry holds the bytes: straight-line predicated byte stores */
					uint8_t *d = s_out + SPAD + pos;
					d[0] = (uint8_t) f0;
					if (l > 1) d[1] = (uint8_t) (f0 >> 8);
					if (l > 2) d[2] = (uint8_t) (f0 >> 16);
					if (l > 3) d[3] = (uint8_t) (f0 >> 24);
					if (l > 4) {
						d[4] = (uint8_t) f1;
						if (l > 5) d[5] = (uint8_t) (f1 >> 8);
						if (l > 6) d[6] = (uint8_t) (f1 >> 16);
						if (l > 7) d[7] = (uint8_t) (f1 >> 24);
						if (l > 8) {
							d[8] = (uint8_t) f2;
							if (l > 9) d[9] = (uint8_t) (f2 >> 8);
							if (l > 10) d[10] = (uint8_t) (f2 >> 16);
							if (l > 11) d[11] = (uint8_t) (f2 >> 24);
						}
					}
				} else {
					const uint8_t *src = a.T.dec_bytes + o2;
					const int k0 = pos < 0 ? -pos : 0, k1 = min(l, wn - pos);
					for (int k = k0; k < k1; k++) s_out[SPAD + pos + k] = src[k];
				}
			}
			pos += l;
		}
		__syncthreads(); /* the window is complete; in the first round this also publishes s_base */
		if (w0 == 0) {
			base = s_base;
			fits = base + total <= a.out_capacity;
			if (!fits && tid == 0) *a.overflow = 1;
		}
		if (!fits) break;
		/* flush: whole 16-byte groups of the OUTPUT with one store each - the staged bytes of a group start at any byte offset, so a
		 * group is five aligned shared-memory words funnel-shifted into four; the ragged ends bytewise */
		const int mis = (int) ((base + w0) & 15);
		uint8_t *const gout = a.out + (base + w0 - mis);
		const int nchunk = (mis + wn + 15) >> 4;
		const int sh = ((SPAD - mis) & 3) * 8;
		for (int c = tid; c < nchunk; c += DNT) {
			const int lo = c == 0 ? mis : 0, hi = min(16, mis + wn - 16 * c);
			const int sa = SPAD + 16 * c - mis; /* staging index of the group's first byte (>= 1) */
			if (lo == 0 && hi == 16) {
				const uint32_t *wp = reinterpret_cast<const uint32_t *>(s_out) + (sa >> 2);
				const uint32_t x0 = wp[0], x1 = wp[1], x2 = wp[2], x3 = wp[3], x4 = wp[4];
				uint4 v;
				v.x = __funnelshift_r(x0, x1, sh);
				v.y = __funnelshift_r(x1, x2, sh);
				v.z = __funnelshift_r(x2, x3, sh);
				v.w = __funnelshift_r(x3, x4, sh);
				*reinterpret_cast<uint4 *>(gout + 16 * c) = v;
			} else {
				for (int k = lo; k < hi; k++) gout[16 * c + k] = s_out[sa + k];
			}
		}
	}
	if (total == 0) { /* (no window round ran: the barrier that publishes s_base) */
		__syncthreads();
		base = s_base;
	}
	/* documents whose first token lies in this tile (the last tile also takes the documents that start at the very end) */
	const bool last = tile == a.ntiles - 1;
	const int64_t tile_end = last ? a.nids + 1 : tile0 + DTILE;
	const int64_t lo = a.tile_first_doc[tile]; /* first d in [0, ndocs] with tok_off[d] >= tile0 (jtk_decode_first_doc_kernel) */
	for (int64_t d = lo + tid; d <= a.ndocs; d += DNT) {
		const int64_t j = a.tok_off[d];
		if (j >= tile_end) break;
		const int jl = (int) (j - tile0);
		int before = jl >= DTILE ? total : s_pref[jl / DPT];
		if (jl < DTILE)
			for (int i = 0; i < jl % DPT; i++) {
				uint32_t l2;
				const int64_t jj = tile0 + jl - jl % DPT + i;
				if (jj < a.nids && decode_len(a.T, a.ids[jj], &l2)) before += (int) l2;
			}
		a.byte_off[d] = base + before;
	}
}

/* first document whose first token is not before the tile (one thread per tile, before the main pass) */
__global__ void jtk_decode_first_doc_kernel(const jtk_decode_args a) {
	const int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (t >= a.ntiles) return;
	const int64_t tile0 = t * DTILE;
	int64_t lo = 0, hi = a.ndocs + 1;
	while (lo < hi) {
		const int64_t mid = (lo + hi) >> 1;
		if (a.tok_off[mid] >= tile0) hi = mid;
		else lo = mid + 1;
	}
	a.tile_first_doc[t] = lo;
}

/* the first unknown id of every document (positions collected by the pass above) */
__global__ void jtk_decode_bad_ids_kernel(const jtk_decode_args a) {
	const int64_t d = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (d >= a.ndocs) return;
	const unsigned long long bp = a.bad_pos[d];
	a.bad_ids[d] = bp != ~0ull ? a.ids[bp] : 0;
}

} /* namespace */

/* ---------------------------------------------------------------------------------------------
 * special-token encoding (see jtk_special_args)
 * ------------------------------------------------------------------------------------------- */
/* One thread per document walks its bytes (leftmost-first, non-overlapping matches are inherently sequential; candidates
 * are found eight bytes at a time when all special tokens start with the same byte).  FILL = false: count; true: write the segments. */
template <bool FILL>
__global__ void jtk_special_scan_kernel(const jtk_special_args a) {
	const jtk_tables &T = a.T;
	for (int64_t d = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; d < a.ndocs; d += (int64_t) gridDim.x * blockDim.x) {
		const int64_t lo = a.doc_off[d], hi = a.doc_off[d + 1];
		int64_t seg = 0, cnt = 0;
		if (FILL) {
			seg = d + 2 * a.match_base[d];
			a.seg_off[seg] = lo;
			a.seg_special[seg] = 0;
		}
		const uint64_t pat = (uint64_t) T.special_first_single * 0x0101010101010101ull;
		for (int64_t i = lo; i < hi;) {
			if (T.special_first_single && (i & 7) == 0 && i + 8 <= hi) { /* SWAR "has byte": skip eight bytes without a candidate */
				const uint64_t x = *reinterpret_cast<const uint64_t *>(a.bytes + i) ^ pat;
				if (!((x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull)) {
					i += 8;
					continue;
				}
			}
			const uint8_t b = a.bytes[i];
			int len = 0;
			if (((T.special_first[b >> 5] >> (b & 31)) & 1u) && jtk_special_match(T, a.bytes, i, hi, &len) >= 0) {
				if (FILL) {
					const int idx = jtk_special_match(T, a.bytes, i, hi, &len);
					a.seg_off[seg + 1] = i;
					a.seg_special[seg + 1] = idx + 1;
					a.seg_off[seg + 2] = i + len;
					a.seg_special[seg + 2] = 0;
					seg += 2;
				}
				cnt++;
				i += len;
			} else {
				i++;
			}
		}
		if (!FILL) a.match_base[d] = cnt;
	}
	if (FILL && blockIdx.x == 0 && threadIdx.x == 0) a.seg_off[a.nseg] = a.total;
	if (!FILL && blockIdx.x == 0 && threadIdx.x == 0) a.match_base[a.ndocs] = 0;
}

__global__ void jtk_special_shift_kernel(const jtk_special_args a) {
	for (int64_t s = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; s <= a.nseg; s += (int64_t) gridDim.x * blockDim.x)
		a.shift[s] = (s < a.nseg && a.seg_special[s]) ? 1 - (a.seg_tok_off[s + 1] - a.seg_tok_off[s]) : 0;
}

__global__ void jtk_special_gather_kernel(const jtk_special_args a, int64_t ntok) {
	const int64_t stride = (int64_t) gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	/* tokens of text segments move by the shift of their segment */
	for (int64_t t = t0; t < ntok; t += stride) {
		int64_t lo = 0, hi = a.nseg - 1; /* last segment with seg_tok_off[s] <= t */
		while (lo < hi) {
			const int64_t mid = (lo + hi + 1) >> 1;
			if (a.seg_tok_off[mid] <= t) lo = mid;
			else hi = mid - 1;
		}
		if (!a.seg_special[lo]) a.ids[t + a.shift[lo]] = a.seg_ids[t];
	}
	/* a special segment becomes the token's id */
	for (int64_t s = t0; s < a.nseg; s += stride)
		if (a.seg_special[s]) a.ids[a.seg_tok_off[s] + a.shift[s]] = a.T.special_ids[a.seg_special[s] - 1];
	/* documents: token offsets and the status bits of their text segments */
	for (int64_t d = t0; d <= a.ndocs; d += stride) {
		const int64_t s0 = d + 2 * a.match_base[d];
		a.tok_off[d] = a.seg_tok_off[s0] + a.shift[s0];
		if (d < a.ndocs) {
			const int64_t s1 = d + 1 + 2 * a.match_base[d + 1];
			int32_t st = 0;
			for (int64_t s = s0; s < s1; s += 2) st |= a.seg_status[s];
			a.doc_status[d] = st;
		}
	}
}

cudaError_t jtk_launch_special_count(const jtk_special_args &a, cudaStream_t st) {
	const unsigned grid = (unsigned) std::min<int64_t>((a.ndocs + 127) / 128 + 1, 148 * 16);
	jtk_special_scan_kernel<false><<<grid, 128, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t jtk_launch_special_fill(const jtk_special_args &a, cudaStream_t st) {
	const unsigned grid = (unsigned) std::min<int64_t>((a.ndocs + 127) / 128 + 1, 148 * 16);
	jtk_special_scan_kernel<true><<<grid, 128, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t jtk_launch_special_shift(const jtk_special_args &a, cudaStream_t st) {
	const unsigned grid = (unsigned) std::min<int64_t>((a.nseg + 256) / 256 + 1, 148 * 16);
	jtk_special_shift_kernel<<<grid, 256, 0, st>>>(a);
	return cudaGetLastError();
}
cudaError_t jtk_launch_special_gather(const jtk_special_args &a, int64_t ntok, cudaStream_t st) {
	const int64_t work = std::max<int64_t>(ntok, a.nseg + 1);
	const unsigned grid = (unsigned) std::min<int64_t>((work + 255) / 256 + 1, 148 * 16);
	jtk_special_gather_kernel<<<grid, 256, 0, st>>>(a, ntok);
	return cudaGetLastError();
}

int64_t jtk_scan_blocks(int64_t n) { return (n + SCAN_NT * SCAN_ITEMS - 1) / (SCAN_NT * SCAN_ITEMS); }

int64_t jtk_decode_tiles(int64_t nids) { return nids > 0 ? (nids + DTILE - 1) / DTILE : 0; }

/* pass 1 + scan: a.tile_bytes (jtk_decode_tiles(nids) + 1 entries) ends up holding the exclusive scan of the bytes per tile, *total
 * (device) the byte count.  a.bad_pos: ndocs entries preset to ~0; block_sums: jtk_scan_blocks(tiles + 1) scratch. */
cudaError_t jtk_launch_decode_count(const jtk_decode_args &a, int64_t *block_sums, int64_t *total, cudaStream_t st) {
	const int64_t nt = jtk_decode_tiles(a.nids);
	if (nt > 0) jtk_decode_count_kernel<<<(unsigned) nt, DNT, 0, st>>>(a);
	cudaMemsetAsync(a.tile_bytes + nt, 0, sizeof(int64_t), st);
	const int64_t n = nt + 1;
	const int64_t nb = jtk_scan_blocks(n);
	jtk_scan_block_kernel<<<(unsigned) nb, SCAN_NT, 0, st>>>(a.tile_bytes, n, block_sums);
	jtk_scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total);
	jtk_scan_add_kernel<<<(unsigned) nb, SCAN_NT, 0, st>>>(a.tile_bytes, n, block_sums);
	return cudaGetLastError();
}

cudaError_t jtk_launch_exclusive_scan(int64_t *data, int64_t n, int64_t *block_sums, int64_t *total, cudaStream_t st) {
	const int64_t nb = jtk_scan_blocks(n);
	jtk_scan_block_kernel<<<(unsigned) nb, SCAN_NT, 0, st>>>(data, n, block_sums);
	jtk_scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total);
	jtk_scan_add_kernel<<<(unsigned) nb, SCAN_NT, 0, st>>>(data, n, block_sums);
	return cudaGetLastError();
}

/* the single-pass decode; a.ticket / a.tile_state (ntiles words) / a.overflow must be zero, a.bad_pos preset to ~0 */
cudaError_t jtk_launch_decode_fused(const jtk_decode_args &a, cudaStream_t st) {
	jtk_decode_first_doc_kernel<<<(unsigned) ((a.ntiles + 255) / 256), 256, 0, st>>>(a);
	jtk_decode_fused_kernel<<<(unsigned) a.ntiles, DNT, 0, st>>>(a);
	if (a.ndocs > 0) jtk_decode_bad_ids_kernel<<<(unsigned) ((a.ndocs + 255) / 256), 256, 0, st>>>(a);
	return cudaGetLastError();
}
