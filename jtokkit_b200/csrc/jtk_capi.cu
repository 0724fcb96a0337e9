/*
 * C ABI implementation (include/jtokkit_b200.h): registration, the device-resident batch call, the
 * host-buffer batch call with chunked H2D / kernel / D2H pipelining and document sharding over devices.
 * There is no CPU encode path in this file: every entry point ends in kernel launches.
 */
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "jtk_kernels.cuh"
#include "jtk_regex.h"
#include "jtk_tables.h"

/* ------------------------------------------------------------------ error plumbing */
static thread_local std::string g_last_error;
static int set_error(int code, const std::string &msg) {
	g_last_error = msg;
	return code;
}
#define CUDA_TRY(expr)                                                                                             \
	do {                                                                                                           \
		cudaError_t _e = (expr);                                                                                   \
		if (_e != cudaSuccess) return set_error(JTK_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
	} while (0)

extern "C" const char *jtk_last_error(void) { return g_last_error.c_str(); }
extern "C" const char *jtk_version(void) {
	static const std::string v = std::string("jtokkit_b200 0.1 sm_100a unicode ") + jtk_unicode_version();
	return v.c_str();
}

/* ------------------------------------------------------------------ per-device state */
struct jtk_pinned_buf {
	void *p = nullptr;
	int64_t cap = 0;
};
/* grow-only pinned staging buffer owned by a workspace / input buffer (never returned to the shared pool: a batch over eight devices
 * uses ~90 of them per call, the pool keeps twelve) */
static int pinned_grow(jtk_pinned_buf *b, int64_t bytes);

struct jtk_workspace {
	/* sized for ntiles_cap tiles / long_cap long pieces */
	int64_t ntiles_cap = 0, long_cap = 0;
	/* general split patterns: start + gap bit arrays (2 * rx_words_cap words) and the backtrack stacks */
	uint32_t *rx_bits = nullptr;
	int64_t rx_words_cap = 0;
	int64_t *rx_rec = nullptr; /* per-slice / per-document records of the sliced matcher */
	int64_t rx_rec_cap = 0;
	void *rx_stacks = nullptr;
	/* side streams for the merge kernels (JTK_SIDE_STREAMS=0 keeps everything on one stream) */
	jtk_side_streams side;
	bool side_ok = false;
	int32_t *tile_first_doc = nullptr;
	int32_t *tile_count = nullptr, *npieces = nullptr, *nslow = nullptr, *tile_slow_used = nullptr;
	int64_t *tile_base = nullptr;
	/* per sub-batch: piece records, merged-token staging, unresolved-piece lists.  Two lanes: the split+lookup kernel of sub-batch
	 * k + 1 runs on the caller's stream while the merge / gather kernels of sub-batch k run on post_stream (JTK_PIPELINE=0: one lane) */
	int64_t sub_tiles = 0;
	struct lane_t {
		int32_t *rec = nullptr, *slowtok = nullptr;
		uint16_t *slowq = nullptr;
		uint32_t *med8 = nullptr, *med32 = nullptr, *shortlist = nullptr;
		jtk_sub_header *sub = nullptr;
		cudaEvent_t split_done = nullptr, post_done = nullptr;
	} lane[2];
	int nlanes = 0;
	cudaStream_t post_stream = nullptr;
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
	std::vector<cudaEvent_t> kev; /* event pairs around the split+lookup kernel of every sub-batch (JTK_TIME_KERNEL) */
	int64_t *tile_first_b = nullptr;
	jtk_long_piece *long_list = nullptr;
	jtk_batch_header *hdr = nullptr;      /* device */
	jtk_batch_header *hdr_host = nullptr; /* pinned */
	/* buffers of the host-buffer path */
	int64_t in_cap = 0, docs_cap = 0;
	int32_t *d_ids = nullptr;
	int64_t *d_tok_off = nullptr;
	int32_t *d_status = nullptr;
	cudaStream_t stream = nullptr;
	/* scratch of the decode path */
	int64_t *dec_tile = nullptr, *dec_sums = nullptr, *dec_total = nullptr;
	unsigned long long *dec_badpos = nullptr;
	int64_t dec_tiles_cap = 0, dec_docs_cap = 0;
	int64_t *dec_total_host = nullptr; /* pinned */
	jtk_pinned_buf h_stage; /* host-buffer path: the chunk's token offsets and statuses on their way to the result arrays */
};

/* per-call piece memo buffer (see jtk_memo_entry); pooled per device, an epoch per buffer makes old entries invisible */
struct jtk_memo_buf {
	jtk_memo_entry *p = nullptr;
	uint32_t epoch = 0;
};
/* entries of the per-call piece memo (64 bytes each): 2^20 = 64 MiB by default (measured on the mixed corpus: 2^18 7.62 ms,
 * 2^19 7.36, 2^20 7.32, 2^21 7.19 per 512 MiB; without a memo 9.51), JTK_MEMO_LOG2 overrides (16..22) */
static uint32_t memo_entries() {
	static const uint32_t v = [] {
		const char *env = getenv("JTK_MEMO_LOG2");
		const int l = env ? atoi(env) : 20;
		return 1u << (l < 16 ? 16 : l > 22 ? 22 : l); /* a piece record holds the entry index in 22 bits */
	}();
	return v;
}
#define JTK_MEMO_ENTRIES memo_entries()

struct jtk_device_state {
	int device = 0;
	int num_sms = 0;
	jtk_tables T;
	std::vector<void *> allocs;
	const void *l2_base = nullptr; /* hot tables: L2 access-policy window */
	size_t l2_bytes = 0;
	std::mutex mu;
	std::vector<jtk_workspace *> free_ws;
	std::vector<jtk_memo_buf *> free_memo;
	std::vector<struct jtk_in_buf *> free_in;
};

/* Device-side input buffer of the host-buffer path (one chunk of documents + their offsets).  The copy-in runs on a ring of
 * these, several chunks ahead of the kernels, so that it finishes early and leaves the PCIe link to the copy-out. */
struct jtk_in_buf {
	uint8_t *d_in = nullptr;
	int64_t cap = 0;
	int64_t *d_doc = nullptr;
	int64_t doc_cap = 0;
	jtk_pinned_buf h_doc; /* chunk-relative document offsets staged for the copy-in */
};

struct jtk_encoding {
	std::string name;
	jtk_host_tables host;
	std::vector<jtk_device_state *> devs;
	std::mutex pool_mu;
	std::vector<jtk_pinned_buf> pinned_pool; /* result buffers returned by jtk_result_free */
	int64_t chunk_bytes = 64ll << 20;
	std::atomic<int64_t> tokens_per_kib{440}; /* running maximum of tokens per KiB seen, sizes the pinned result buffers */
};

struct jtk_result {
	jtk_encoding *enc = nullptr;
	int64_t ndocs = 0, ntokens = 0, nbytes = 0;
	jtk_pinned_buf ids, tok_off, status, bytes, byte_off, bad_ids;
	double device_ms = 0;
	int64_t launches = 0;
};

static int device_index(const jtk_encoding *e, int device) {
	for (size_t i = 0; i < e->devs.size(); i++)
		if (e->devs[i]->device == device) return (int) i;
	return -1;
}

/* All tables of an encoding live in ONE device allocation so that a single L2 access-policy window can pin them
 * (persisting hits, streaming misses for everything else): the piece table, pair table and byte-pair table are probed
 * once or twice per piece and must not be evicted by the input / id streams. */
struct table_arena {
	std::vector<std::pair<const void *, size_t>> parts;
	std::vector<size_t> offsets;
	size_t total = 0;
	template <typename Tv>
	size_t add(const std::vector<Tv> &v) {
		size_t bytes = std::max<size_t>(v.size() * sizeof(Tv), 16);
		size_t off = total;
		parts.emplace_back(v.data(), v.size() * sizeof(Tv));
		offsets.push_back(off);
		total = (off + bytes + 255) & ~(size_t) 255;
		return off;
	}
};

static int init_device(jtk_encoding *e, int device) {
	jtk_device_state *ds = new jtk_device_state();
	e->devs.push_back(ds);
	ds->device = device;
	CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) return set_error(JTK_E_CUDA, "device is not sm_100-class (this library only carries sm_100a code)");
	ds->num_sms = prop.multiProcessorCount;
	const jtk_host_tables &h = e->host;
	jtk_tables &T = ds->T;
	memset(&T, 0, sizeof(T));
	T.pattern_kind = h.pattern_kind;
	T.max_token_len = h.max_token_len;
	table_arena ar;
	/* hot tables first: they form the persisting window */
	const size_t o_tab_a = ar.add(h.tab_a), o_pair = ar.add(h.pair), o_bytepair = ar.add(h.bytepair), o_byte_id = ar.add(h.byte_id);
	const size_t o_ascii = ar.add(h.ascii_cls), o_st1 = ar.add(h.cp_stage1), o_st2 = ar.add(h.cp_stage2), o_tab_b = ar.add(h.tab_b), o_filt = ar.add(h.long_filter);
	const size_t o_lsp = ar.add(h.lut_sp), o_cls2 = ar.add(h.cls2), o_nib = ar.add(h.bmp_nib), o_big = ar.add(h.bigram_bits), o_tri = ar.add(h.trigram_bits);
	const size_t hot_bytes = ar.total;
	const size_t o_tokb = ar.add(h.tok_bytes), o_toko = ar.add(h.tok_off), o_spb = ar.add(h.special_bytes), o_spo = ar.add(h.special_off);
	const size_t o_deck = ar.add(h.dec_keys), o_decb = ar.add(h.dec_bytes), o_deco = ar.add(h.dec_off);
	const size_t o_decd = ar.add(h.dec_direct);
	const size_t o_rxi = ar.add(h.rx_inst), o_rxs = ar.add(h.rx_sets), o_rxr = ar.add(h.rx_ranges), o_spi = ar.add(h.special_ids);
	const size_t o_dft = ar.add(h.rx_dfa_trans), o_df1 = ar.add(h.rx_dfa_stage1), o_df2 = ar.add(h.rx_dfa_stage2), o_dfs = ar.add(h.rx_dfa_stay);
	uint8_t *base = nullptr;
	CUDA_TRY(cudaMalloc(&base, ar.total));
	ds->allocs.push_back(base);
	for (size_t i = 0; i < ar.parts.size(); i++)
		if (ar.parts[i].second) CUDA_TRY(cudaMemcpy(base + ar.offsets[i], ar.parts[i].first, ar.parts[i].second, cudaMemcpyHostToDevice));
	T.tab_a = reinterpret_cast<const jtk_slot_a *>(base + o_tab_a);
	T.pair = reinterpret_cast<const jtk_slot *>(base + o_pair);
	T.bytepair = reinterpret_cast<const int32_t *>(base + o_bytepair);
	T.byte_id = reinterpret_cast<const int32_t *>(base + o_byte_id);
	T.ascii_cls = base + o_ascii;
	T.cp_stage1 = reinterpret_cast<const uint16_t *>(base + o_st1);
	T.cp_stage2 = base + o_st2;
	T.lut_sp = reinterpret_cast<const uint32_t *>(base + o_lsp);
	T.cls2 = base + o_cls2;
	T.bmp_nib = base + o_nib;
	T.bigram_bits = reinterpret_cast<const uint32_t *>(base + o_big);
	T.trigram_bits = reinterpret_cast<const uint32_t *>(base + o_tri);
	T.tab_b = reinterpret_cast<const jtk_slot *>(base + o_tab_b);
	T.long_filter = reinterpret_cast<const uint32_t *>(base + o_filt);
	T.tok_bytes = base + o_tokb;
	T.tok_off = reinterpret_cast<const uint32_t *>(base + o_toko);
	T.special_bytes = base + o_spb;
	T.special_off = reinterpret_cast<const uint32_t *>(base + o_spo);
	T.special_ids = reinterpret_cast<const int32_t *>(base + o_spi);
	T.dec_keys = reinterpret_cast<const uint32_t *>(base + o_deck);
	T.dec_bytes = base + o_decb;
	T.dec_off = reinterpret_cast<const uint32_t *>(base + o_deco);
	T.dec_direct = h.dec_direct.empty() ? nullptr : reinterpret_cast<const uint4 *>(base + o_decd);
	T.dec_direct_size = (uint32_t) (h.dec_direct.size() / 4);
	T.rx_inst = base + o_rxi;
	T.rx_sets = base + o_rxs;
	T.rx_ranges = reinterpret_cast<const uint32_t *>(base + o_rxr);
	T.rx_ninst = h.rx_ninst;
	memcpy(T.rx_first, h.rx_first, sizeof(T.rx_first));
	T.rx_dfa_trans = h.rx_dfa_trans.empty() ? nullptr : reinterpret_cast<const uint16_t *>(base + o_dft);
	T.rx_dfa_stage1 = reinterpret_cast<const uint16_t *>(base + o_df1);
	T.rx_dfa_stage2 = base + o_df2;
	T.rx_dfa_stay = h.rx_dfa_stay.empty() ? nullptr : base + o_dfs;
	T.rx_dfa_nsym = h.rx_dfa_nsym;
	T.rx_dfa_nstates = h.rx_dfa_nstates;
	T.rx_dfa_start = h.rx_dfa_start;
	T.rx_dfa_start_bol = h.rx_dfa_start_bol;
	T.rx_dfa_acc_lo = h.rx_dfa_acc_lo;
	T.mask_a = h.mask_a;
	T.mask_b = h.mask_b;
	T.mask_p = h.mask_p;
	T.mask_d = h.mask_d;
	T.nspecial = h.nspecial;
	T.special_has_empty = h.special_has_empty;
	memcpy(T.special_first, h.special_first, sizeof(T.special_first));
	{
		int nfirst = 0, first = 0;
		for (int b = 0; b < 256; b++)
			if ((h.special_first[b >> 5] >> (b & 31)) & 1u) {
				nfirst++;
				first = b;
			}
		T.special_first_single = (nfirst == 1 && first != 0) ? (uint32_t) first : 0u;
	}
	/* L2 persisting window over the hot tables (bounded by what the device allows) */
	ds->l2_base = base;
	ds->l2_bytes = std::min<size_t>(hot_bytes, (size_t) std::max(prop.accessPolicyMaxWindowSize, 0));
	if (prop.persistingL2CacheMaxSize > 0 && ds->l2_bytes > 0) {
		size_t want = std::min<size_t>((size_t) prop.persistingL2CacheMaxSize, ds->l2_bytes + (ds->l2_bytes >> 2));
		if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) {
			cudaGetLastError();
			ds->l2_bytes = 0;
		}
	} else {
		ds->l2_bytes = 0;
	}
	CUDA_TRY(jtk_encode_kernel_setup());
	return JTK_OK;
}

/* ------------------------------------------------------------------ registration */
extern "C" int jtk_encoding_create(const jtk_params *params, const int *devices, int ndev, jtk_encoding **out) {
	if (!params || !out) return set_error(JTK_E_ARG, "params / out is null");
	*out = nullptr;
	jtk_encoding *e = new jtk_encoding();
	std::string err;
	int rc = jtk_build_host_tables(params, &e->host, &err);
	if (rc != JTK_OK) {
		delete e;
		return set_error(rc, err);
	}
	e->name = e->host.name;
	if (const char *env = getenv("JTK_CHUNK_MB")) {
		long mb = atol(env);
		if (mb >= 1) e->chunk_bytes = (int64_t) mb << 20;
	}
	int count = 0;
	cudaError_t ce = cudaGetDeviceCount(&count);
	if (ce != cudaSuccess || count == 0) {
		delete e;
		return set_error(JTK_E_CUDA, std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0") +
		                                 " (jtokkit_b200 has no CPU fallback)");
	}
	std::vector<int> devs;
	if (!devices || ndev <= 0) devs.push_back(0);
	else devs.assign(devices, devices + ndev);
	for (int d : devs) {
		if (d < 0 || d >= count) {
			jtk_encoding_destroy(e);
			return set_error(JTK_E_ARG, "device index out of range");
		}
		rc = init_device(e, d);
		if (rc != JTK_OK) {
			std::string keep = g_last_error;
			jtk_encoding_destroy(e);
			return set_error(rc, keep);
		}
	}
	*out = e;
	return JTK_OK;
}

extern "C" int jtk_encoding_create_builtin(const char *name, const char *tiktoken_path, const int *devices, int ndev, jtk_encoding **out) {
	if (!name || !tiktoken_path || !out) return set_error(JTK_E_ARG, "name / path / out is null");
	const jtk_builtin_def *def = jtk_find_builtin(name);
	if (!def) return set_error(JTK_E_ARG, std::string("unknown predefined encoding: ") + name);
	std::vector<uint8_t> bytes;
	std::vector<int64_t> off;
	std::vector<int32_t> ranks;
	std::string err;
	int rc = jtk_load_tiktoken_file(tiktoken_path, &bytes, &off, &ranks, &err);
	if (rc != JTK_OK) return set_error(rc, err);
	std::vector<uint8_t> sb;
	std::vector<int64_t> so(1, 0);
	std::vector<int32_t> sid;
	for (int i = 0; i < def->nspecial; i++) {
		sb.insert(sb.end(), def->special[i], def->special[i] + strlen(def->special[i]));
		so.push_back((int64_t) sb.size());
		sid.push_back(def->special_ids[i]);
	}
	jtk_params p;
	p.name = def->name;
	p.pattern = def->pattern;
	p.pattern_flags = JTK_RE_UNICODE_CHARACTER_CLASS; /* EncodingFactory.java:129 */
	p.vocab_bytes = bytes.data();
	p.vocab_off = off.data();
	p.vocab_ranks = ranks.data();
	p.vocab_size = (int64_t) ranks.size();
	p.special_bytes = sb.data();
	p.special_off = so.data();
	p.special_ids = sid.data();
	p.special_size = def->nspecial;
	return jtk_encoding_create(&p, devices, ndev, out);
}

static void free_workspace(jtk_workspace *w) {
	if (!w) return;
	cudaFree(w->tile_first_doc);
	cudaFree(w->tile_count);
	cudaFree(w->npieces);
	cudaFree(w->nslow);
	cudaFree(w->tile_slow_used);
	cudaFree(w->tile_base);
	for (auto &ln : w->lane) {
		cudaFree(ln.rec);
		cudaFree(ln.slowtok);
		cudaFree(ln.slowq);
		cudaFree(ln.med8);
		cudaFree(ln.med32);
		cudaFree(ln.shortlist);
		cudaFree(ln.sub);
		if (ln.split_done) cudaEventDestroy(ln.split_done);
		if (ln.post_done) cudaEventDestroy(ln.post_done);
	}
	if (w->post_stream) cudaStreamDestroy(w->post_stream);
	if (w->ev_begin) cudaEventDestroy(w->ev_begin);
	if (w->ev_end) cudaEventDestroy(w->ev_end);
	for (cudaEvent_t ev : w->kev) cudaEventDestroy(ev);
	cudaFree(w->tile_first_b);
	cudaFree(w->long_list);
	cudaFree(w->rx_bits);
	cudaFree(w->rx_rec);
	cudaFree(w->rx_stacks);
	cudaFree(w->hdr);
	cudaFreeHost(w->hdr_host);
	cudaFree(w->d_ids);
	cudaFree(w->d_tok_off);
	cudaFree(w->d_status);
	cudaFree(w->dec_tile);
	cudaFree(w->dec_sums);
	cudaFree(w->dec_total);
	cudaFree(w->dec_badpos);
	cudaFreeHost(w->dec_total_host);
	cudaFreeHost(w->h_stage.p);
	if (w->side_ok) {
		for (int i = 0; i < 3; i++) {
			cudaStreamDestroy(w->side.s[i]);
			cudaEventDestroy(w->side.join[i]);
		}
		cudaEventDestroy(w->side.fork);
	}
	if (w->stream) cudaStreamDestroy(w->stream);
	delete w;
}

extern "C" void jtk_encoding_destroy(jtk_encoding *e) {
	if (!e) return;
	for (jtk_device_state *ds : e->devs) {
		cudaSetDevice(ds->device);
		for (jtk_in_buf *b : ds->free_in) {
			cudaFree(b->d_in);
			cudaFree(b->d_doc);
			cudaFreeHost(b->h_doc.p);
			delete b;
		}
		for (jtk_workspace *w : ds->free_ws) free_workspace(w);
		for (jtk_memo_buf *m : ds->free_memo) {
			cudaFree(m->p);
			delete m;
		}
		for (void *p : ds->allocs) cudaFree(p);
		delete ds;
	}
	for (jtk_pinned_buf &b : e->pinned_pool) cudaFreeHost(b.p);
	delete e;
}

extern "C" const char *jtk_encoding_name(const jtk_encoding *e) { return e ? e->name.c_str() : ""; }
extern "C" int jtk_encoding_num_devices(const jtk_encoding *e) { return e ? (int) e->devs.size() : 0; }

/* ------------------------------------------------------------------ workspaces */
static jtk_workspace *acquire_ws(jtk_device_state *ds) {
	std::lock_guard<std::mutex> lk(ds->mu);
	if (!ds->free_ws.empty()) {
		jtk_workspace *w = ds->free_ws.back();
		ds->free_ws.pop_back();
		return w;
	}
	return new jtk_workspace();
}
static void release_ws(jtk_device_state *ds, jtk_workspace *w) {
	std::lock_guard<std::mutex> lk(ds->mu);
	ds->free_ws.push_back(w);
}

static int64_t sub_batch_tiles() {
	static const int64_t v = [] {
		const char *env = getenv("JTK_SUB_TILES");
		long n = env ? atol(env) : 0;
		return (int64_t) (n >= 64 ? std::min<long>(n, 1l << 18) : JTK_DEFAULT_SUB_TILES); /* list entries carry the tile index in 18 bits */
	}();
	return v;
}

static bool side_streams_enabled() {
	static const bool v = [] {
		const char *env = getenv("JTK_SIDE_STREAMS");
		return !(env && env[0] == '0');
	}();
	return v;
}

static bool pipeline_enabled() {
	static const bool v = [] {
		const char *env = getenv("JTK_PIPELINE"); /* measured: no gain on B200 (profiles/r1_history.md), so off unless asked for */
		return env && env[0] == '1';
	}();
	return v;
}

/* CTAs per SM of the split+lookup kernel; 0 = one CTA per tile (lets the other lane's kernels in as CTAs retire) */
static int split_ctas_per_sm() {
	static const int v = [] {
		const char *env = getenv("JTK_SPLIT_CTAS");
		return env ? atoi(env) : (pipeline_enabled() ? 0 : 4);
	}();
	return v;
}

static bool memo_enabled() {
	static const bool v = [] {
		const char *env = getenv("JTK_MEMO");
		return !(env && env[0] == '0');
	}();
	return v;
}

/* A memo buffer for one call; nullptr (memo off) when disabled or out of memory. */
static jtk_memo_buf *acquire_memo(jtk_device_state *ds) {
	if (!memo_enabled()) return nullptr;
	jtk_memo_buf *m = nullptr;
	{
		std::lock_guard<std::mutex> lk(ds->mu);
		if (!ds->free_memo.empty()) {
			m = ds->free_memo.back();
			ds->free_memo.pop_back();
		}
	}
	if (!m) {
		m = new jtk_memo_buf();
		if (cudaMalloc(&m->p, sizeof(jtk_memo_entry) * JTK_MEMO_ENTRIES) != cudaSuccess || cudaMemset(m->p, 0, sizeof(jtk_memo_entry) * JTK_MEMO_ENTRIES) != cudaSuccess) {
			cudaGetLastError();
			cudaFree(m->p);
			delete m;
			return nullptr;
		}
	}
	if (++m->epoch >= 0xFFFFFFu) {
		cudaMemset(m->p, 0, sizeof(jtk_memo_entry) * JTK_MEMO_ENTRIES);
		m->epoch = 1;
	}
	return m;
}
static void release_memo(jtk_device_state *ds, jtk_memo_buf *m) {
	if (!m) return;
	std::lock_guard<std::mutex> lk(ds->mu);
	ds->free_memo.push_back(m);
}

static int ensure_ws_tiles(jtk_workspace *w, int64_t ntiles, int64_t long_cap) {
	int prio_lo = 0, prio_hi = 0;
	CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi)); /* numerically lower = higher priority */
	if (!w->side_ok && side_streams_enabled()) {
		for (int i = 0; i < 3; i++) {
			CUDA_TRY(cudaStreamCreateWithPriority(&w->side.s[i], cudaStreamNonBlocking, prio_hi));
			CUDA_TRY(cudaEventCreateWithFlags(&w->side.join[i], cudaEventDisableTiming));
		}
		CUDA_TRY(cudaEventCreateWithFlags(&w->side.fork, cudaEventDisableTiming));
		w->side_ok = true;
	}
	if (w->nlanes == 0) {
		w->nlanes = pipeline_enabled() ? 2 : 1;
		for (int l = 0; l < w->nlanes; l++) {
			CUDA_TRY(cudaMalloc(&w->lane[l].sub, sizeof(jtk_sub_header)));
			CUDA_TRY(cudaMemset(w->lane[l].sub, 0, sizeof(jtk_sub_header)));
			CUDA_TRY(cudaEventCreateWithFlags(&w->lane[l].split_done, cudaEventDisableTiming));
			CUDA_TRY(cudaEventCreateWithFlags(&w->lane[l].post_done, cudaEventDisableTiming));
		}
		if (w->nlanes > 1) {
			CUDA_TRY(cudaStreamCreateWithPriority(&w->post_stream, cudaStreamNonBlocking, prio_hi));
			CUDA_TRY(cudaEventCreateWithFlags(&w->ev_begin, cudaEventDisableTiming));
			CUDA_TRY(cudaEventCreateWithFlags(&w->ev_end, cudaEventDisableTiming));
		}
	}
	if (!w->hdr) {
		CUDA_TRY(cudaMalloc(&w->hdr, sizeof(jtk_batch_header)));
		CUDA_TRY(cudaHostAlloc(&w->hdr_host, sizeof(jtk_batch_header), cudaHostAllocDefault));
	}
	if (ntiles > w->ntiles_cap) {
		cudaFree(w->tile_first_doc);
		cudaFree(w->tile_count);
		cudaFree(w->npieces);
		cudaFree(w->nslow);
		cudaFree(w->tile_slow_used);
		cudaFree(w->tile_base);
		cudaFree(w->tile_first_b);
		w->tile_first_doc = w->tile_count = w->npieces = w->nslow = w->tile_slow_used = nullptr;
		w->tile_base = nullptr;
		w->tile_first_b = nullptr;
		w->ntiles_cap = 0;
		int64_t cap = ntiles + ntiles / 4 + 16;
		CUDA_TRY(cudaMalloc(&w->tile_first_doc, sizeof(int32_t) * cap));
		CUDA_TRY(cudaMalloc(&w->tile_count, sizeof(int32_t) * cap));
		CUDA_TRY(cudaMalloc(&w->npieces, sizeof(int32_t) * cap));
		CUDA_TRY(cudaMalloc(&w->nslow, sizeof(int32_t) * cap));
		CUDA_TRY(cudaMalloc(&w->tile_slow_used, sizeof(int32_t) * cap));
		CUDA_TRY(cudaMalloc(&w->tile_base, sizeof(int64_t) * (cap + 1)));
		CUDA_TRY(cudaMalloc(&w->tile_first_b, sizeof(int64_t) * cap));
		w->ntiles_cap = cap;
	}
	const int64_t sub = std::min<int64_t>(std::max<int64_t>(ntiles, 1), sub_batch_tiles());
	if (sub > w->sub_tiles) {
		w->sub_tiles = 0;
		for (int l = 0; l < w->nlanes; l++) {
			jtk_workspace::lane_t &ln = w->lane[l];
			cudaFree(ln.rec);
			cudaFree(ln.slowtok);
			cudaFree(ln.slowq);
			cudaFree(ln.med8);
			cudaFree(ln.med32);
			cudaFree(ln.shortlist);
			ln.rec = ln.slowtok = nullptr;
			ln.slowq = nullptr;
			ln.med8 = ln.med32 = ln.shortlist = nullptr;
			CUDA_TRY(cudaMalloc(&ln.rec, sizeof(int32_t) * (size_t) sub * JTK_RECN));
			CUDA_TRY(cudaMalloc(&ln.slowtok, sizeof(int32_t) * (size_t) sub * JTK_RECN));
			CUDA_TRY(cudaMalloc(&ln.slowq, sizeof(uint16_t) * (size_t) sub * JTK_QCAP));
			CUDA_TRY(cudaMalloc(&ln.med8, sizeof(uint32_t) * (size_t) sub * JTK_MED8_PER_TILE));
			CUDA_TRY(cudaMalloc(&ln.med32, sizeof(uint32_t) * (size_t) sub * JTK_MED32_PER_TILE));
			CUDA_TRY(cudaMalloc(&ln.shortlist, sizeof(uint32_t) * (size_t) sub * JTK_QCAP));
		}
		w->sub_tiles = sub;
	}
	if (long_cap > w->long_cap) {
		cudaFree(w->long_list);
		w->long_list = nullptr;
		w->long_cap = 0;
		int64_t cap = long_cap + long_cap / 4 + 16;
		CUDA_TRY(cudaMalloc(&w->long_list, sizeof(jtk_long_piece) * cap));
		w->long_cap = cap;
	}
	return JTK_OK;
}

static void set_lane(jtk_encode_args &a, const jtk_workspace *w, int l) {
	const jtk_workspace::lane_t &ln = w->lane[l];
	a.rec = ln.rec;
	a.slowtok = ln.slowtok;
	a.slowq = ln.slowq;
	a.med8 = ln.med8;
	a.med32 = ln.med32;
	a.shortlist = ln.shortlist;
	a.sub = ln.sub;
}

static void fill_args(jtk_encode_args &a, const jtk_device_state *ds, const jtk_workspace *w) {
	a.T = ds->T;
	a.tile_first_doc = w->tile_first_doc;
	a.npieces = w->npieces;
	a.nslow = w->nslow;
	a.tile_slow_used = w->tile_slow_used;
	a.tile_count = w->tile_count;
	a.tile_base = w->tile_base;
	a.tile_first_b = w->tile_first_b;
	set_lane(a, w, 0);
	a.hdr = w->hdr;
	a.long_list = w->long_list;
	a.long_cap = w->long_cap;
	a.l2_base = ds->l2_base;
	a.l2_bytes = ds->l2_bytes;
}

/* ------------------------------------------------------------------ the device-resident batch */
static int run_long_pieces(jtk_device_state *ds, jtk_workspace *w, jtk_encode_args &a, unsigned n_long, cudaStream_t st, jtk_device_info *info) {
	CUDA_TRY(jtk_launch_long_bounds(a, n_long, st));
	std::vector<jtk_long_piece> list(n_long);
	CUDA_TRY(cudaMemcpyAsync(list.data(), a.long_list, sizeof(jtk_long_piece) * n_long, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	std::sort(list.begin(), list.end(), [](const jtk_long_piece &x, const jtk_long_piece &y) { return x.start < y.start; });
	int64_t parts = 0;
	for (jtk_long_piece &lp : list) {
		lp.scratch = parts;
		parts += lp.end - lp.start;
		if (lp.end - lp.start > 0x7ffffff0ll) return set_error(JTK_E_ARG, "a single piece exceeds 2 GiB");
	}
	int32_t *scr = nullptr;
	CUDA_TRY(cudaMalloc(&scr, sizeof(int32_t) * JTK_LONG_SCRATCH_ARRAYS * (size_t) parts));
	int32_t *scr_tok = scr;
	int rc = JTK_OK;
	int64_t *d_cum = nullptr;
	int32_t *ids2 = nullptr;
	do {
		cudaError_t ce;
#define LTRY(expr)                                                                              \
	if ((ce = (expr)) != cudaSuccess) {                                                         \
		rc = set_error(JTK_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(ce));        \
		break;                                                                                  \
	}
		LTRY(cudaMemcpyAsync(a.long_list, list.data(), sizeof(jtk_long_piece) * n_long, cudaMemcpyHostToDevice, st));
		LTRY(jtk_launch_long_merge(a, n_long, scr, parts, ds->num_sms, st));
		LTRY(cudaMemcpyAsync(list.data(), a.long_list, sizeof(jtk_long_piece) * n_long, cudaMemcpyDeviceToHost, st));
		LTRY(cudaMemcpyAsync(w->hdr_host, a.hdr, sizeof(jtk_batch_header), cudaMemcpyDeviceToHost, st));
		LTRY(cudaStreamSynchronize(st));
		std::vector<int64_t> cum(n_long + 1, 0);
		for (unsigned i = 0; i < n_long; i++) cum[i + 1] = cum[i] + list[i].count;
		const int64_t total_in = info->num_tokens, total_out = total_in + cum[n_long];
		LTRY(cudaMalloc(&d_cum, sizeof(int64_t) * (n_long + 1)));
		LTRY(cudaMemcpyAsync(d_cum, cum.data(), sizeof(int64_t) * (n_long + 1), cudaMemcpyHostToDevice, st));
		if (!(a.flags & JTK_COUNT_ONLY) && a.ids) {
			if (total_out > a.ids_cap) {
				rc = set_error(JTK_E_CAPACITY, "ids buffer too small");
				break;
			}
			LTRY(cudaMalloc(&ids2, sizeof(int32_t) * (size_t) std::max<int64_t>(total_out, 1)));
			LTRY(jtk_launch_long_insert(a.long_list, d_cum, n_long, scr_tok, a.ids, ids2, total_in, st));
			LTRY(cudaMemcpyAsync(a.ids, ids2, sizeof(int32_t) * (size_t) total_out, cudaMemcpyDeviceToDevice, st));
			info->gpu_launches += 1;
		}
		if (a.tok_off) {
			LTRY(jtk_launch_long_fix_offsets(a.long_list, d_cum, n_long, a.doc_off, a.ndocs, a.tok_off, st));
			info->gpu_launches += 1;
		}
		LTRY(cudaStreamSynchronize(st));
		info->num_tokens = total_out;
		info->gpu_launches += 2;
		info->reserved = (int32_t) w->hdr_host->violations;
#undef LTRY
	} while (0);
	cudaFree(scr);
	cudaFree(d_cum);
	cudaFree(ids2);
	return rc;
}

static int encode_device_impl(jtk_encoding *e, jtk_device_state *ds, jtk_workspace *w, const uint8_t *d_utf8, int64_t nbytes, const int64_t *d_doc_off,
                              int64_t ndocs, uint32_t flags, int32_t *d_ids, int64_t ids_capacity, int64_t *d_tok_off, int32_t *d_doc_status,
                              uint8_t *d_piece_flags, cudaStream_t st, jtk_device_info *info, bool sync_and_long, const jtk_memo_buf *memo) {
	(void) e;
	if (nbytes < 0 || ndocs < 0) return set_error(JTK_E_ARG, "negative size");
	if ((reinterpret_cast<uintptr_t>(d_utf8) & 15) != 0) return set_error(JTK_E_ARG, "d_utf8 must be 16-byte aligned");
	const int64_t ntiles = (nbytes + JTK_TILE - 1) / JTK_TILE;
	const int64_t long_cap = nbytes / (JTK_LONG_PIECE + 1) + 1;
	int rc = ensure_ws_tiles(w, ntiles, long_cap);
	if (rc != JTK_OK) return rc;
	jtk_encode_args a;
	memset(&a, 0, sizeof(a));
	fill_args(a, ds, w);
	a.bytes = d_utf8;
	a.total = nbytes;
	a.doc_off = d_doc_off;
	a.ndocs = ndocs;
	a.ntiles = ntiles;
	a.ids = d_ids;
	a.ids_cap = ids_capacity;
	a.tok_off = d_tok_off;
	a.doc_status = d_doc_status;
	a.flags = flags;
	a.piece_flags = d_piece_flags;
	if (memo && w->nlanes == 1) { /* the two-lane pipeline would read the memo while the other lane's merge kernel fills it */
		a.memo = memo->p;
		a.memo_mask = JTK_MEMO_ENTRIES - 1;
		a.memo_epoch = memo->epoch;
	}
	CUDA_TRY(cudaMemsetAsync(w->hdr, 0, sizeof(jtk_batch_header), st));
	for (int l = 0; l < w->nlanes; l++) CUDA_TRY(cudaMemsetAsync(w->lane[l].sub, 0, sizeof(jtk_sub_header), st));
	CUDA_TRY(jtk_launch_tile_first_doc(d_doc_off, ndocs, ntiles, w->tile_first_doc, st));
	int general_launches = 0;
	if (ds->T.pattern_kind == JTK_PAT_GENERAL && ntiles > 0) {
		/* general pattern: piece / gap bits first (jtk_general_slice / stitch / finish kernels), the tile kernels read them */
		const int64_t words = (ntiles * (int64_t) JTK_TILE + JTK_REGION) / 32 + 2;
		const int64_t slices = (nbytes + JTK_RX_SLICE - 1) / JTK_RX_SLICE, recs = 4 * slices + ndocs + 1;
		if (words > w->rx_words_cap) {
			cudaFree(w->rx_bits);
			w->rx_bits = nullptr;
			w->rx_words_cap = 0;
			CUDA_TRY(cudaMalloc(&w->rx_bits, sizeof(uint32_t) * 5 * (size_t) (words + words / 4)));
			w->rx_words_cap = words + words / 4;
		}
		if (recs > w->rx_rec_cap) {
			cudaFree(w->rx_rec);
			w->rx_rec = nullptr;
			w->rx_rec_cap = 0;
			CUDA_TRY(cudaMalloc(&w->rx_rec, sizeof(int64_t) * (size_t) (recs + recs / 4)));
			w->rx_rec_cap = recs + recs / 4;
		}
		if (!w->rx_stacks) CUDA_TRY(cudaMalloc(&w->rx_stacks, sizeof(jtk_rx_frame) * (size_t) JTK_RX_STACK * JTK_RX_THREADS));
		CUDA_TRY(cudaMemsetAsync(w->rx_bits, 0, sizeof(uint32_t) * 5 * (size_t) words, st));
		a.rx_start = w->rx_bits;
		a.rx_skip = w->rx_bits + words;
		a.rx_spec = w->rx_bits + 2 * words;
		a.rx_words = words;
		a.rx_rec = w->rx_rec;
		a.rx_slices = slices;
		a.rx_stacks = w->rx_stacks;
		CUDA_TRY(jtk_launch_general_split(a, st));
		general_launches = 3;
	}
	const bool time_kernel = (flags & JTK_TIME_KERNEL) && sync_and_long;
	/* sub-batches: a small first one (16 MiB) so that the piece memo is warm early, then full-size ones */
	const int64_t sub = w->sub_tiles;
	std::vector<int64_t> cuts(1, 0);
	{
		/* with a memo the sub-batches grow geometrically (16 MiB, 64 MiB, 256 MiB, ...): pieces merged in one sub-batch are memo
		 * hits in all later ones, so the early ones are kept short; the later ones are long to amortise launches and tails */
		int64_t step = (a.memo && ntiles > 4 * JTK_FIRST_SUB_TILES) ? JTK_FIRST_SUB_TILES : sub;
		while (cuts.back() < ntiles) {
			cuts.push_back(std::min<int64_t>(ntiles, cuts.back() + std::min<int64_t>(step, sub)));
			step *= 4;
		}
	}
	const int64_t nsub = (int64_t) cuts.size() - 1;
	if (time_kernel)
		while ((int64_t) w->kev.size() < 2 * nsub) {
			cudaEvent_t ev;
			CUDA_TRY(cudaEventCreate(&ev));
			w->kev.push_back(ev);
		}
	/* Two sub-batches in flight: split+lookup of sub-batch i on the caller's stream, everything after it on post_stream with
	 * the other lane of buffers; the split of sub-batch i + 2 waits for the lane to be free again. */
	const bool piped = w->nlanes > 1 && nsub > 1;
	const jtk_side_streams *side = w->side_ok ? &w->side : nullptr;
	/* JTK_PHASE_TRACE=1 (development aid, device-resident call only): per sub-batch the time of split / scatter / merges / scan / gather on stderr */
	static const bool phase_trace = getenv("JTK_PHASE_TRACE") != nullptr;
	std::vector<cudaEvent_t> pm;
	if (phase_trace && sync_and_long && !piped)
		for (int64_t i = 0; i < 6 * nsub; i++) {
			cudaEvent_t ev;
			cudaEventCreate(&ev);
			pm.push_back(ev);
		}
	if (piped) {
		CUDA_TRY(cudaEventRecord(w->ev_begin, st));
		CUDA_TRY(cudaStreamWaitEvent(w->post_stream, w->ev_begin, 0));
	}
	for (int64_t i = 0; i < nsub; i++) {
		const int l = piped ? (int) (i & 1) : 0;
		set_lane(a, w, l);
		a.tile_begin = cuts[(size_t) i];
		a.tile_end = cuts[(size_t) i + 1];
		if (piped && i >= 2) CUDA_TRY(cudaStreamWaitEvent(st, w->lane[l].post_done, 0));
		if (!pm.empty()) cudaEventRecord(pm[(size_t) (6 * i)], st);
		CUDA_TRY(jtk_launch_split(a, ds->num_sms, split_ctas_per_sm(), time_kernel ? w->kev[(size_t) (2 * i)] : nullptr, time_kernel ? w->kev[(size_t) (2 * i + 1)] : nullptr, st));
		if (piped) {
			CUDA_TRY(cudaEventRecord(w->lane[l].split_done, st));
			CUDA_TRY(cudaStreamWaitEvent(w->post_stream, w->lane[l].split_done, 0));
			CUDA_TRY(jtk_launch_post(a, ds->num_sms, w->post_stream, side));
			CUDA_TRY(cudaEventRecord(w->lane[l].post_done, w->post_stream));
		} else if (!pm.empty()) {
			cudaEventRecord(pm[(size_t) (6 * i + 1)], st);
			CUDA_TRY(jtk_launch_post(a, ds->num_sms, st, side, &pm[(size_t) (6 * i + 2)]));
		} else {
			CUDA_TRY(jtk_launch_post(a, ds->num_sms, st, side));
		}
	}
	if (piped) {
		CUDA_TRY(cudaEventRecord(w->ev_end, w->post_stream));
		CUDA_TRY(cudaStreamWaitEvent(st, w->ev_end, 0));
	}
	CUDA_TRY(jtk_launch_finalize(a, st));
	info->gpu_launches = (ntiles > 0 ? 1 : 0) + general_launches + 9 * nsub + 1;
	CUDA_TRY(cudaMemcpyAsync(w->hdr_host, w->hdr, sizeof(jtk_batch_header), cudaMemcpyDeviceToHost, st));
	if (!sync_and_long) return JTK_OK;
	CUDA_TRY(cudaStreamSynchronize(st));
	if (!pm.empty()) {
		float tot[5] = {0, 0, 0, 0, 0};
		for (int64_t i = 0; i < nsub; i++) {
			std::string line = "jtk phases sub-batch " + std::to_string(i) + " (" + std::to_string(cuts[(size_t) i + 1] - cuts[(size_t) i]) + " tiles):";
			static const char *const names[5] = {"split", "sort", "merge", "scan", "gather"};
			for (int k = 0; k < 5; k++) {
				float ms = 0;
				cudaEventElapsedTime(&ms, pm[(size_t) (6 * i + k)], pm[(size_t) (6 * i + k + 1)]);
				tot[k] += ms;
				char buf[48];
				snprintf(buf, sizeof(buf), " %s %.3f", names[k], ms);
				line += buf;
			}
			fprintf(stderr, "%s\n", line.c_str());
		}
		fprintf(stderr, "jtk phases total: split %.3f sort %.3f merge %.3f scan %.3f gather %.3f ms\n", tot[0], tot[1], tot[2], tot[3], tot[4]);
		for (cudaEvent_t ev : pm) cudaEventDestroy(ev);
	}
	if (time_kernel) {
		info->tile_kernel_ms = 0;
		for (int64_t i = 0; i < nsub; i++) {
			float ms = 0;
			cudaEventElapsedTime(&ms, w->kev[(size_t) (2 * i)], w->kev[(size_t) (2 * i + 1)]);
			info->tile_kernel_ms += ms;
		}
	}
	info->num_tokens = (int64_t) w->hdr_host->total_tokens;
	info->num_long_pieces = w->hdr_host->n_long;
	info->reserved = 0;
	if (w->hdr_host->overflow) return set_error(JTK_E_CAPACITY, "ids buffer too small");
	if (w->hdr_host->n_long > 0) {
		if ((int64_t) w->hdr_host->n_long > w->long_cap) return set_error(JTK_E_NOMEM, "long piece list overflow");
		rc = run_long_pieces(ds, w, a, w->hdr_host->n_long, st, info);
		if (rc != JTK_OK) return rc;
	}
	if ((flags & JTK_CHECK_SPECIAL) && ds->T.special_has_empty && d_doc_status && ndocs > 0) {
		/* "".contains(...) is true for every text: flag every document (pathological registration) */
		std::vector<int32_t> ones((size_t) ndocs, JTK_DOC_HAS_SPECIAL);
		CUDA_TRY(cudaMemcpyAsync(d_doc_status, ones.data(), sizeof(int32_t) * ndocs, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaStreamSynchronize(st));
	}
	return JTK_OK;
}

extern "C" int jtk_encode_batch_device(jtk_encoding *e, int device, const uint8_t *d_utf8, int64_t nbytes, const int64_t *d_doc_off, int64_t ndocs,
                                       uint32_t flags, int32_t *d_ids, int64_t ids_capacity, int64_t *d_tok_off, int32_t *d_doc_status,
                                       void *cuda_stream, jtk_device_info *info) {
	if (!e || !info || !d_doc_off || (nbytes > 0 && !d_utf8)) return set_error(JTK_E_ARG, "null argument");
	int di = device_index(e, device);
	if (di < 0) return set_error(JTK_E_ARG, "device is not one of the encoding's devices");
	jtk_device_state *ds = e->devs[(size_t) di];
	CUDA_TRY(cudaSetDevice(device));
	jtk_workspace *w = acquire_ws(ds);
	jtk_memo_buf *memo = acquire_memo(ds);
	memset(info, 0, sizeof(*info));
	int rc = encode_device_impl(e, ds, w, d_utf8, nbytes, d_doc_off, ndocs, flags, d_ids, ids_capacity, d_tok_off, d_doc_status, nullptr,
	                            reinterpret_cast<cudaStream_t>(cuda_stream), info, true, memo);
	release_memo(ds, memo);
	release_ws(ds, w);
	return rc;
}

extern "C" int jtk_split_batch_device(jtk_encoding *e, int device, const uint8_t *d_utf8, int64_t nbytes, const int64_t *d_doc_off, int64_t ndocs,
                                      uint8_t *d_piece_flags, void *cuda_stream) {
	if (!e || !d_doc_off || !d_piece_flags || (nbytes > 0 && !d_utf8)) return set_error(JTK_E_ARG, "null argument");
	int di = device_index(e, device);
	if (di < 0) return set_error(JTK_E_ARG, "device is not one of the encoding's devices");
	jtk_device_state *ds = e->devs[(size_t) di];
	CUDA_TRY(cudaSetDevice(device));
	jtk_workspace *w = acquire_ws(ds);
	jtk_device_info info;
	memset(&info, 0, sizeof(info));
	int rc = encode_device_impl(e, ds, w, d_utf8, nbytes, d_doc_off, ndocs, JTK_COUNT_ONLY, nullptr, 0, nullptr, nullptr, d_piece_flags,
	                            reinterpret_cast<cudaStream_t>(cuda_stream), &info, true, nullptr);
	release_ws(ds, w);
	return rc;
}

/* ------------------------------------------------------------------ pinned buffer pool */
static int pinned_get(jtk_encoding *e, int64_t bytes, jtk_pinned_buf *out) {
	bytes = std::max<int64_t>(bytes, 64);
	{
		std::lock_guard<std::mutex> lk(e->pool_mu);
		int best = -1;
		for (size_t i = 0; i < e->pinned_pool.size(); i++)
			if (e->pinned_pool[i].cap >= bytes && (best < 0 || e->pinned_pool[i].cap < e->pinned_pool[(size_t) best].cap)) best = (int) i;
		if (best >= 0) {
			*out = e->pinned_pool[(size_t) best];
			e->pinned_pool.erase(e->pinned_pool.begin() + best);
			return JTK_OK;
		}
	}
	out->p = nullptr;
	out->cap = bytes;
	CUDA_TRY(cudaHostAlloc(&out->p, (size_t) bytes, cudaHostAllocPortable));
	return JTK_OK;
}
static void pinned_put(jtk_encoding *e, jtk_pinned_buf &b) {
	if (!b.p) return;
	std::lock_guard<std::mutex> lk(e->pool_mu);
	e->pinned_pool.push_back(b);
	b.p = nullptr;
	b.cap = 0;
	/* keep the pool bounded: drop the smallest buffers beyond 12 entries */
	while (e->pinned_pool.size() > 12) {
		size_t m = 0;
		for (size_t i = 1; i < e->pinned_pool.size(); i++)
			if (e->pinned_pool[i].cap < e->pinned_pool[m].cap) m = i;
		cudaFreeHost(e->pinned_pool[m].p);
		e->pinned_pool.erase(e->pinned_pool.begin() + (long) m);
	}
}

static int pinned_grow(jtk_pinned_buf *b, int64_t bytes) {
	if (b->cap >= bytes) return JTK_OK;
	if (b->p) cudaFreeHost(b->p);
	b->p = nullptr;
	b->cap = 0;
	const int64_t cap = bytes + bytes / 4 + 4096;
	CUDA_TRY(cudaHostAlloc(&b->p, (size_t) cap, cudaHostAllocPortable));
	b->cap = cap;
	return JTK_OK;
}

extern "C" void *jtk_host_alloc(int64_t nbytes) {
	void *p = nullptr;
	if (cudaHostAlloc(&p, (size_t) std::max<int64_t>(nbytes, 64), cudaHostAllocPortable) != cudaSuccess) {
		set_error(JTK_E_CUDA, "cudaHostAlloc failed");
		return nullptr;
	}
	return p;
}
extern "C" void jtk_host_free(void *p) {
	if (p) cudaFreeHost(p);
}
extern "C" void jtk_free(void *p) { free(p); }

/* ------------------------------------------------------------------ the host-buffer batch
 * One batch = one global list of CHUNKS (contiguous ranges of whole documents, byte balanced).  Chunk c runs on device c % G:
 * every device takes every G-th chunk, so all devices walk through the batch side by side (the reference's
 * AbstractMultiThreadedBenchmark.java:34-45 hands one document per task to a thread pool; here a task is a chunk of documents
 * and a worker is a GPU).  There is no data-path collective and no host-side concatenation of ids: the token count of a
 * chunk is published as soon as its kernels have run, the exclusive prefix over the chunks before it is the chunk's position
 * in the ONE pinned result buffer, and the chunk's device-to-host copy lands there directly. */
struct batch_ctx {
	jtk_encoding *e = nullptr;
	const uint8_t *utf8 = nullptr;
	const int64_t *doc_off = nullptr;
	uint32_t flags = 0;
	int G = 1;
	std::vector<int64_t> cb; /* chunk c = documents [cb[c], cb[c + 1]) */
	/* result arrays (pinned) */
	int32_t *ids = nullptr;
	int64_t ids_cap = 0; /* tokens */
	int64_t *tok_off = nullptr;
	int32_t *status = nullptr;
	/* token counts and result positions of the chunks */
	std::mutex mu;
	std::condition_variable cv;
	std::vector<int64_t> ntok;  /* -1 = not known yet */
	std::vector<int64_t> base;  /* nchunks + 1; valid up to index known */
	size_t known = 0;           /* base[0 .. known] are final */
	bool failed = false;
	/* chunks that did not fit the result buffer (its size is an estimate): kept in buffers of their own, merged at the end */
	std::vector<jtk_pinned_buf> spill;
};

struct device_job {
	batch_ctx *B = nullptr;
	jtk_device_state *ds = nullptr;
	int g = 0;
	double device_ms = 0;
	int64_t launches = 0;
	int rc = JTK_OK;
	std::string err;
};

static void publish_chunk(batch_ctx *B, size_t c, int64_t ntok) {
	std::lock_guard<std::mutex> lk(B->mu);
	B->ntok[c] = ntok;
	while (B->known < B->ntok.size() && B->ntok[B->known] >= 0) {
		B->base[B->known + 1] = B->base[B->known] + B->ntok[B->known];
		B->known++;
	}
	B->cv.notify_all();
}
static void fail_batch(batch_ctx *B) {
	std::lock_guard<std::mutex> lk(B->mu);
	B->failed = true;
	B->cv.notify_all();
}
/* position of chunk c in the result: waits until every earlier chunk (on whichever device) has published its token count */
static bool wait_base(batch_ctx *B, size_t c, int64_t *base) {
	std::unique_lock<std::mutex> lk(B->mu);
	B->cv.wait(lk, [&] { return B->failed || B->known >= c; });
	if (B->known < c) return false;
	*base = B->base[c];
	return true;
}

/* output-side buffers of a workspace for a chunk of nbytes / ndocs (ids: at most one token per byte) */
static int ensure_ws_io(jtk_workspace *w, int64_t nbytes, int64_t ndocs, bool want_ids) {
	if (nbytes > w->in_cap) {
		cudaFree(w->d_ids);
		w->d_ids = nullptr;
		w->in_cap = nbytes + nbytes / 8 + 4096;
	}
	if (want_ids && !w->d_ids) CUDA_TRY(cudaMalloc(&w->d_ids, sizeof(int32_t) * (size_t) (w->in_cap + 16)));
	if (ndocs > w->docs_cap) {
		cudaFree(w->d_tok_off);
		cudaFree(w->d_status);
		w->d_tok_off = nullptr;
		w->d_status = nullptr;
		w->docs_cap = 0;
		int64_t cap = ndocs + ndocs / 8 + 1024;
		CUDA_TRY(cudaMalloc(&w->d_tok_off, sizeof(int64_t) * (size_t) (cap + 1)));
		CUDA_TRY(cudaMalloc(&w->d_status, sizeof(int32_t) * (size_t) (cap + 1)));
		w->docs_cap = cap;
	}
	return JTK_OK;
}

static int ensure_in_buf(jtk_in_buf *b, int64_t nbytes, int64_t ndocs) {
	if (nbytes > b->cap) {
		cudaFree(b->d_in);
		b->d_in = nullptr;
		b->cap = 0;
		const int64_t cap = nbytes + nbytes / 8 + 4096;
		CUDA_TRY(cudaMalloc(&b->d_in, (size_t) cap + 64));
		b->cap = cap;
	}
	if (ndocs > b->doc_cap) {
		cudaFree(b->d_doc);
		b->d_doc = nullptr;
		b->doc_cap = 0;
		const int64_t cap = ndocs + ndocs / 8 + 1024;
		CUDA_TRY(cudaMalloc(&b->d_doc, sizeof(int64_t) * (size_t) (cap + 1)));
		b->doc_cap = cap;
	}
	return JTK_OK;
}

/* Chunk boundaries of a batch for G devices: whole documents; the chunk size ramps up from 8 MiB to chunk_bytes wave by wave
 * (a wave = one chunk per device) and back down towards the end of the batch, so that the exposed head (first copy-in + first
 * compute) and tail (last compute + last copy-out) of each device's copy-in / compute / copy-out pipeline are short while the
 * steady state moves large chunks. */
static void plan_chunks(const int64_t *off, int64_t ndocs, int G, int64_t chunk_bytes, std::vector<int64_t> *cb) {
	cb->assign(1, 0);
	const int64_t small = std::min<int64_t>(chunk_bytes, 8ll << 20), end_byte = off[ndocs];
	int64_t ramp = small, want = small;
	while (cb->back() < ndocs) {
		const int64_t d = cb->back();
		if ((cb->size() - 1) % (size_t) G == 0) { /* a new wave: one size for all its chunks */
			const int64_t remaining = end_byte - off[d];
			want = std::max<int64_t>(small, std::min<int64_t>(std::min<int64_t>(ramp, chunk_bytes), remaining / (2 * (int64_t) G)));
			ramp = std::min<int64_t>(ramp * 2, chunk_bytes);
		}
		const int64_t lim = off[d] + want;
		int64_t hi = std::upper_bound(off + d + 1, off + ndocs + 1, lim) - off - 1; /* last doc end <= lim */
		if (hi <= d) hi = d + 1;                                                     /* a single oversized document */
		cb->push_back(hi);
	}
}

extern "C" int64_t jtk_plan_chunks(const int64_t *doc_off, int64_t ndocs, int ndev, int64_t chunk_bytes, int64_t *cuts, int64_t cuts_capacity) {
	if (!doc_off || ndocs < 0 || ndev < 1 || doc_off[0] != 0) return set_error(JTK_E_ARG, "jtk_plan_chunks: bad argument");
	for (int64_t d = 0; d < ndocs; d++)
		if (doc_off[d + 1] < doc_off[d]) return set_error(JTK_E_ARG, "doc_off is not monotone");
	if (chunk_bytes <= 0) {
		chunk_bytes = 64ll << 20;
		if (const char *env = getenv("JTK_CHUNK_MB")) {
			long mb = atol(env);
			if (mb >= 1) chunk_bytes = (int64_t) mb << 20;
		}
	}
	std::vector<int64_t> cb;
	plan_chunks(doc_off, ndocs, ndev, chunk_bytes, &cb);
	const int64_t n = (int64_t) cb.size() - 1;
	if (cuts && cuts_capacity >= n + 1) memcpy(cuts, cb.data(), sizeof(int64_t) * cb.size());
	return n;
}

/* One worker = one device.  Its chunks run through a three-stage pipeline on three streams (copy-in, compute, copy-out) over
 * NS workspaces: while chunk k is encoded, later chunks are copied in and the ids of earlier chunks are copied out, so PCIe
 * in both directions and the SMs are busy at the same time.  Stages are ordered by events between the streams; the host only
 * waits where it has to read a result (token counts, staged offsets). */
static void run_device(device_job *job) {
	batch_ctx *B = job->B;
	jtk_encoding *e = B->e;
	jtk_device_state *ds = job->ds;
	auto fail = [&](int rc) {
		job->rc = rc;
		job->err = g_last_error;
		fail_batch(B);
	};
	if (cudaSetDevice(ds->device) != cudaSuccess) return fail(set_error(JTK_E_CUDA, "cudaSetDevice failed"));
	const int64_t *off = B->doc_off;
	const bool want_ids = !(B->flags & JTK_COUNT_ONLY);
	/* this device's chunks: local index k <-> global chunk g + k * G */
	const size_t nglobal = B->cb.size() - 1;
	const size_t nchunks = nglobal > (size_t) job->g ? (nglobal - (size_t) job->g + (size_t) B->G - 1) / (size_t) B->G : 0;
	if (nchunks == 0) return;
	auto gc = [&](size_t k) { return (size_t) job->g + k * (size_t) B->G; };
	const std::vector<int64_t> &cb = B->cb;
	constexpr int NS = 3;
	/* JTK_TRACE=1: per-chunk begin / end times of the three stages on stderr (development aid) */
	static const bool trace = getenv("JTK_TRACE") != nullptr;
	std::vector<cudaEvent_t> tr;
	cudaEvent_t tr0 = nullptr;
	auto mark = [&](cudaStream_t st) {
		if (!trace) return;
		cudaEvent_t ev;
		cudaEventCreate(&ev);
		cudaEventRecord(ev, st);
		tr.push_back(ev);
	};
	jtk_memo_buf *memo = acquire_memo(ds); /* shared by all chunks of this call on this device */
	constexpr int NR = 8; /* input ring: the copy-in may run this many chunks ahead of the kernels */
	jtk_workspace *ws[NS];
	jtk_in_buf *ring[NR];
	cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr, s_meta = nullptr;
	cudaEvent_t ev_in[NR], ev_free[NR], ev_k0[NS], ev_k1[NS], ev_out[NS], ev_meta[NS];
	jtk_device_info infos[NS];
	int64_t out_chunk[NS]; /* local chunk whose copy-out is in flight on the slot, -1 = none */
	int64_t out_base[NS];
	memset(infos, 0, sizeof(infos));
	int rc = JTK_OK;
	for (int i = 0; i < NS; i++) {
		ws[i] = acquire_ws(ds);
		ev_k0[i] = ev_k1[i] = ev_out[i] = ev_meta[i] = nullptr;
		out_chunk[i] = -1;
		out_base[i] = 0;
	}
	for (int i = 0; i < NR; i++) {
		ev_in[i] = ev_free[i] = nullptr;
		ring[i] = nullptr;
		std::lock_guard<std::mutex> lk(ds->mu);
		if (!ds->free_in.empty()) {
			ring[i] = ds->free_in.back();
			ds->free_in.pop_back();
		} else {
			ring[i] = new jtk_in_buf();
		}
	}
	if (cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&s_meta, cudaStreamNonBlocking) != cudaSuccess)
		rc = set_error(JTK_E_CUDA, "cudaStreamCreate failed");
	for (int i = 0; i < NS && rc == JTK_OK; i++)
		if (cudaEventCreate(&ev_k0[i]) != cudaSuccess || cudaEventCreate(&ev_k1[i]) != cudaSuccess ||
		    cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ev_meta[i], cudaEventDisableTiming) != cudaSuccess)
			rc = set_error(JTK_E_CUDA, "cudaEventCreate failed");
	for (int i = 0; i < NR && rc == JTK_OK; i++)
		if (cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming) != cudaSuccess)
			rc = set_error(JTK_E_CUDA, "cudaEventCreate failed");

	/* the copy-out of a slot has landed: rebase its token offsets / statuses into the result's arrays */
	auto retire = [&](int slot) -> int {
		if (out_chunk[slot] < 0) return JTK_OK;
		CUDA_TRY(cudaEventSynchronize(ev_out[slot]));
		const size_t c = gc((size_t) out_chunk[slot]);
		const int64_t d0 = cb[c], nd = cb[c + 1] - d0;
		const int64_t *h_tok = static_cast<const int64_t *>(ws[slot]->h_stage.p);
		const int32_t *h_st = reinterpret_cast<const int32_t *>(h_tok + nd + 1);
		const int64_t base = out_base[slot];
		for (int64_t i = 0; i < nd; i++) {
			B->tok_off[d0 + i] = h_tok[i] + base;
			B->status[d0 + i] = h_st[i];
		}
		out_chunk[slot] = -1;
		return JTK_OK;
	};
	/* local chunk k has been encoded on its slot: read the totals, run the long-piece path if needed, publish the token count,
	 * wait for the chunk's position in the result, start the copy-out */
	auto copy_out = [&](size_t k) -> int {
		const int slot = (int) (k % NS);
		const size_t c = gc(k);
		jtk_workspace *w = ws[slot];
		const int64_t d0 = cb[c], d1 = cb[c + 1], nd = d1 - d0;
		const int64_t cbytes = off[d1] - off[d0];
		CUDA_TRY(cudaEventSynchronize(ev_k1[slot]));
		float ms = 0;
		cudaEventElapsedTime(&ms, ev_k0[slot], ev_k1[slot]);
		job->device_ms += ms;
		jtk_device_info &info = infos[slot];
		info.num_tokens = (int64_t) w->hdr_host->total_tokens;
		info.num_long_pieces = w->hdr_host->n_long;
		if (w->hdr_host->overflow) return set_error(JTK_E_CAPACITY, "internal ids buffer too small");
		if (w->hdr_host->n_long > 0) {
			jtk_encode_args a;
			memset(&a, 0, sizeof(a));
			fill_args(a, ds, w);
			a.bytes = ring[k % NR]->d_in;
			a.total = cbytes;
			a.doc_off = ring[k % NR]->d_doc;
			a.ndocs = nd;
			a.ntiles = (cbytes + JTK_TILE - 1) / JTK_TILE;
			a.ids = want_ids ? w->d_ids : nullptr;
			a.ids_cap = w->in_cap;
			a.tok_off = w->d_tok_off;
			a.doc_status = w->d_status;
			a.flags = B->flags;
			cudaEvent_t l0, l1;
			cudaEventCreate(&l0);
			cudaEventCreate(&l1);
			cudaEventRecord(l0, s_comp);
			int r2 = run_long_pieces(ds, w, a, w->hdr_host->n_long, s_comp, &info);
			cudaEventRecord(l1, s_comp);
			cudaEventSynchronize(l1);
			float lms = 0;
			cudaEventElapsedTime(&lms, l0, l1);
			job->device_ms += lms;
			cudaEventDestroy(l0);
			cudaEventDestroy(l1);
			if (r2 != JTK_OK) return r2;
		}
		CUDA_TRY(cudaEventRecord(ev_free[k % NR], s_comp)); /* the chunk's input buffer may be overwritten from here on */
		job->launches += info.gpu_launches;
		const int64_t ntok = info.num_tokens;
		publish_chunk(B, c, ntok);
		{
			const int64_t per_kib = ntok * 1024 / std::max<int64_t>(cbytes, 1024) + 8;
			int64_t cur = e->tokens_per_kib.load();
			while (per_kib > cur && !e->tokens_per_kib.compare_exchange_weak(cur, per_kib)) {
			}
		}
		int64_t base = 0;
		if (!wait_base(B, c, &base)) return set_error(JTK_E_CUDA, "another device of the batch failed");
		if (want_ids && ntok > 0) {
			int32_t *dst = B->ids + base;
			if (base + ntok > B->ids_cap) {
				/* rare: the batch is denser than anything seen before and the result buffer (sized from the densest chunk so far) is
				 * too small; this chunk and all later ones go to buffers of their own and are merged when the batch is done */
				jtk_pinned_buf sp;
				int r2 = pinned_get(e, sizeof(int32_t) * ntok, &sp);
				if (r2 != JTK_OK) return r2;
				{
					std::lock_guard<std::mutex> lk(B->mu);
					B->spill[c] = sp;
				}
				dst = static_cast<int32_t *>(sp.p);
			}
			mark(s_out);
			CUDA_TRY(cudaMemcpyAsync(dst, w->d_ids, sizeof(int32_t) * (size_t) ntok, cudaMemcpyDeviceToHost, s_out));
			mark(s_out);
		}
		{
			int r2 = retire(slot); /* the slot's previous copy-out must have been consumed before its staging buffer is reused */
			if (r2 != JTK_OK) return r2;
		}
		jtk_pinned_buf &so = w->h_stage;
		{
			int r2 = pinned_grow(&so, (int64_t) sizeof(int64_t) * (nd + 1) + (int64_t) sizeof(int32_t) * (nd + 1));
			if (r2 != JTK_OK) return r2;
		}
		int64_t *h_tok = static_cast<int64_t *>(so.p);
		int32_t *h_st = reinterpret_cast<int32_t *>(h_tok + nd + 1);
		/* the small per-document arrays go on their own stream so that the id copies stay back to back on s_out */
		CUDA_TRY(cudaMemcpyAsync(h_tok, w->d_tok_off, sizeof(int64_t) * (size_t) (nd + 1), cudaMemcpyDeviceToHost, s_meta));
		if (nd > 0) CUDA_TRY(cudaMemcpyAsync(h_st, w->d_status, sizeof(int32_t) * (size_t) nd, cudaMemcpyDeviceToHost, s_meta));
		CUDA_TRY(cudaEventRecord(ev_meta[slot], s_meta));
		CUDA_TRY(cudaStreamWaitEvent(s_out, ev_meta[slot], 0));
		CUDA_TRY(cudaEventRecord(ev_out[slot], s_out));
		out_chunk[slot] = (int64_t) k;
		out_base[slot] = base;
		return JTK_OK;
	};

	/* Stream-side ordering, the host does not wait for copies.  The copy-in of chunk j overwrites ring buffer j % NR, free once
	 * chunk j - NR has been encoded (ev_free, recorded by copy_out); the kernels of chunk k reuse the slot's id / offset
	 * buffers, free once the copy-out of chunk k - NS has landed (ev_out).  Copy-ins are queued as far ahead as the ring
	 * allows, so the copy-in finishes early and leaves the PCIe link to the copy-out (measured, profiles/r1_pcie*.txt:
	 * ~46 GB/s per direction while both are busy, ~56 GB/s alone). */
	auto issue_in = [&](size_t j) -> int {
		const size_t c = gc(j);
		const int64_t d0 = cb[c], d1 = cb[c + 1], nd = d1 - d0;
		const int64_t b0 = off[d0], cbytes = off[d1] - b0;
		jtk_in_buf *in = ring[j % NR];
		if (j >= (size_t) NR) {
			if (cbytes <= in->cap && nd <= in->doc_cap) CUDA_TRY(cudaStreamWaitEvent(s_in, ev_free[j % NR], 0));
			else CUDA_TRY(cudaEventSynchronize(ev_free[j % NR])); /* about to be reallocated */
		}
		int r2 = ensure_in_buf(in, cbytes, nd);
		if (r2 != JTK_OK) return r2;
		/* chunk-relative document offsets (the kernels require doc_off[0] == 0), staged in pinned memory */
		jtk_pinned_buf &sd = in->h_doc;
		if (j >= (size_t) NR) CUDA_TRY(cudaEventSynchronize(ev_in[j % NR])); /* the copy-in of chunk j - NR has read the staging buffer (long ago) */
		r2 = pinned_grow(&sd, (int64_t) sizeof(int64_t) * (nd + 1));
		if (r2 != JTK_OK) return r2;
		int64_t *rebase = static_cast<int64_t *>(sd.p);
		for (int64_t i = 0; i <= nd; i++) rebase[i] = off[d0 + i] - b0;
		if (trace && !tr0) {
			cudaEventCreate(&tr0);
			cudaEventRecord(tr0, s_in);
		}
		mark(s_in);
		CUDA_TRY(cudaMemcpyAsync(in->d_doc, rebase, sizeof(int64_t) * (size_t) (nd + 1), cudaMemcpyHostToDevice, s_in));
		if (cbytes > 0) CUDA_TRY(cudaMemcpyAsync(in->d_in, B->utf8 + b0, (size_t) cbytes, cudaMemcpyHostToDevice, s_in));
		mark(s_in);
		CUDA_TRY(cudaEventRecord(ev_in[j % NR], s_in));
		return JTK_OK;
	};
	size_t next_in = 0;
	for (size_t k = 0; k < nchunks && rc == JTK_OK; k++) {
		const int slot = (int) (k % NS);
		const size_t c = gc(k);
		jtk_workspace *w = ws[slot];
		const int64_t d0 = cb[c], d1 = cb[c + 1], nd = d1 - d0;
		const int64_t cbytes = off[d1] - off[d0];
		/* ev_free of chunk j - NR is recorded by copy_out(j - NR); at this point copy_out has been called for the chunks up to k - 2 */
		while (rc == JTK_OK && next_in < nchunks && (next_in < (size_t) NR || next_in + 2 <= k + (size_t) NR)) rc = issue_in(next_in++);
		if (rc != JTK_OK) break;
		jtk_in_buf *in = ring[k % NR];
		if (k >= (size_t) NS) {
			if (cbytes <= w->in_cap && nd <= w->docs_cap) {
				cudaStreamWaitEvent(s_comp, ev_out[slot], 0);
			} else {
				rc = retire(slot); /* the slot's buffers are about to be reallocated: its copy-out must have landed */
				if (rc != JTK_OK) break;
			}
		}
		rc = ensure_ws_io(w, cbytes, nd, want_ids);
		if (rc != JTK_OK) break;
		cudaError_t ce = cudaStreamWaitEvent(s_comp, ev_in[k % NR], 0);
		if (ce == cudaSuccess && nd > 0) ce = cudaMemsetAsync(w->d_status, 0, sizeof(int32_t) * (size_t) nd, s_comp); /* after the slot's previous copy-out */
		if (ce != cudaSuccess) {
			rc = set_error(JTK_E_CUDA, std::string("chunk set-up: ") + cudaGetErrorString(ce));
			break;
		}
		cudaEventRecord(ev_k0[slot], s_comp);
		rc = encode_device_impl(e, ds, w, in->d_in, cbytes, in->d_doc, nd, B->flags, want_ids ? w->d_ids : nullptr, w->in_cap, w->d_tok_off, w->d_status,
		                        nullptr, s_comp, &infos[slot], false, memo);
		cudaEventRecord(ev_k1[slot], s_comp);
		if (rc != JTK_OK) break;
		if (k >= 1) rc = copy_out(k - 1);
	}
	if (rc == JTK_OK) rc = copy_out(nchunks - 1);
	for (int i = 0; i < NS && rc == JTK_OK; i++) rc = retire(i);
	if (s_in) cudaStreamSynchronize(s_in);
	if (s_comp) cudaStreamSynchronize(s_comp);
	if (s_out) cudaStreamSynchronize(s_out);
	if (s_meta) cudaStreamSynchronize(s_meta);
	if (trace && tr0) {
		/* marks come in pairs: per chunk one copy-in pair (in loop order) and, when ids are wanted, one copy-out pair */
		char head[64];
		snprintf(head, sizeof(head), "jtk trace device %d (ms since first copy-in):", ds->device);
		std::string line = head;
		for (size_t i = 0; i + 1 < tr.size(); i += 2) {
			float t0 = 0, t1 = 0;
			cudaEventElapsedTime(&t0, tr0, tr[i]);
			cudaEventElapsedTime(&t1, tr0, tr[i + 1]);
			char buf[64];
			snprintf(buf, sizeof(buf), " [%.2f %.2f]", t0, t1);
			line += buf;
		}
		fprintf(stderr, "%s\n", line.c_str());
		for (cudaEvent_t ev : tr) cudaEventDestroy(ev);
		cudaEventDestroy(tr0);
	}
	for (int i = 0; i < NS; i++) {
		if (ev_k0[i]) cudaEventDestroy(ev_k0[i]);
		if (ev_k1[i]) cudaEventDestroy(ev_k1[i]);
		if (ev_out[i]) cudaEventDestroy(ev_out[i]);
		if (ev_meta[i]) cudaEventDestroy(ev_meta[i]);
		release_ws(ds, ws[i]);
	}
	for (int i = 0; i < NR; i++) {
		if (ev_in[i]) cudaEventDestroy(ev_in[i]);
		if (ev_free[i]) cudaEventDestroy(ev_free[i]);
		std::lock_guard<std::mutex> lk(ds->mu);
		ds->free_in.push_back(ring[i]);
	}
	if (s_in) cudaStreamDestroy(s_in);
	if (s_comp) cudaStreamDestroy(s_comp);
	if (s_out) cudaStreamDestroy(s_out);
	if (s_meta) cudaStreamDestroy(s_meta);
	release_memo(ds, memo);
	if (rc != JTK_OK) fail(rc);
}

extern "C" int jtk_encode_batch(jtk_encoding *e, const uint8_t *utf8, const int64_t *doc_off, int64_t ndocs, uint32_t flags, jtk_result **out) {
	if (!e || !doc_off || !out || ndocs < 0) return set_error(JTK_E_ARG, "null argument");
	*out = nullptr;
	if (doc_off[0] != 0) return set_error(JTK_E_ARG, "doc_off[0] must be 0");
	for (int64_t d = 0; d < ndocs; d++)
		if (doc_off[d + 1] < doc_off[d]) return set_error(JTK_E_ARG, "doc_off is not monotone");
	const int64_t total = doc_off[ndocs];
	if (total > 0 && !utf8) return set_error(JTK_E_ARG, "utf8 is null");
	const bool want_ids = !(flags & JTK_COUNT_ONLY);
	jtk_result *r = new jtk_result();
	r->enc = e;
	r->ndocs = ndocs;
	int rc = pinned_get(e, sizeof(int64_t) * (ndocs + 1), &r->tok_off);
	if (rc == JTK_OK) rc = pinned_get(e, sizeof(int32_t) * (ndocs + 1), &r->status);
	/* the ids of all chunks land in one buffer, sized from the densest chunk seen so far (chunks that do not fit are kept aside) */
	if (rc == JTK_OK && want_ids) rc = pinned_get(e, sizeof(int32_t) * (total / 1024 * e->tokens_per_kib.load() + 4096), &r->ids);
	if (rc != JTK_OK) {
		jtk_result_free(r);
		return rc;
	}
	batch_ctx B;
	B.e = e;
	B.utf8 = utf8;
	B.doc_off = doc_off;
	B.flags = flags;
	B.G = (int) e->devs.size();
	plan_chunks(doc_off, ndocs, B.G, e->chunk_bytes, &B.cb);
	const size_t nchunks = B.cb.size() - 1;
	B.ids = static_cast<int32_t *>(r->ids.p);
	B.ids_cap = want_ids ? r->ids.cap / (int64_t) sizeof(int32_t) : 0;
	B.tok_off = static_cast<int64_t *>(r->tok_off.p);
	B.status = static_cast<int32_t *>(r->status.p);
	B.ntok.assign(nchunks, -1);
	B.base.assign(nchunks + 1, 0);
	B.spill.assign(nchunks, jtk_pinned_buf());
	std::vector<device_job> jobs((size_t) B.G);
	for (int g = 0; g < B.G; g++) {
		jobs[(size_t) g].B = &B;
		jobs[(size_t) g].ds = e->devs[(size_t) g];
		jobs[(size_t) g].g = g;
	}
	if (B.G == 1 || nchunks <= 1) {
		run_device(&jobs[0]);
	} else {
		std::vector<std::thread> th;
		for (int g = 0; g < B.G; g++) th.emplace_back(run_device, &jobs[(size_t) g]);
		for (auto &t : th) t.join();
	}
	auto drop_spill = [&] {
		for (jtk_pinned_buf &sp : B.spill) pinned_put(e, sp);
	};
	for (int g = 0; g < B.G; g++)
		if (jobs[(size_t) g].rc != JTK_OK) {
			rc = jobs[(size_t) g].rc;
			std::string msg = jobs[(size_t) g].err;
			drop_spill();
			jtk_result_free(r);
			return set_error(rc, msg);
		}
	const int64_t total_tokens = B.base[nchunks];
	r->ntokens = total_tokens;
	B.tok_off[ndocs] = total_tokens;
	for (int g = 0; g < B.G; g++) {
		r->device_ms = std::max(r->device_ms, jobs[(size_t) g].device_ms);
		r->launches += jobs[(size_t) g].launches;
	}
	if ((flags & JTK_CHECK_SPECIAL) && e->host.special_has_empty) /* "".contains(...) is true for every text */
		for (int64_t d = 0; d < ndocs; d++) B.status[d] |= JTK_DOC_HAS_SPECIAL;
	if (want_ids && total_tokens > B.ids_cap) {
		/* the size estimate was too low: one buffer of the right size, the part that did land plus the chunks kept aside */
		jtk_pinned_buf full;
		rc = pinned_get(e, sizeof(int32_t) * total_tokens, &full);
		if (rc != JTK_OK) {
			drop_spill();
			jtk_result_free(r);
			return rc;
		}
		int32_t *dst = static_cast<int32_t *>(full.p);
		for (size_t c = 0; c < nchunks; c++) {
			const int32_t *src = B.spill[c].p ? static_cast<const int32_t *>(B.spill[c].p) : B.ids + B.base[c];
			memcpy(dst + B.base[c], src, sizeof(int32_t) * (size_t) B.ntok[c]);
		}
		pinned_put(e, r->ids);
		r->ids = full;
	}
	drop_spill();
	*out = r;
	return JTK_OK;
}


extern "C" int64_t jtk_result_num_docs(const jtk_result *r) { return r ? r->ndocs : 0; }
extern "C" int64_t jtk_result_num_tokens(const jtk_result *r) { return r ? r->ntokens : 0; }
extern "C" const int32_t *jtk_result_ids(const jtk_result *r) { return r ? static_cast<const int32_t *>(r->ids.p) : nullptr; }
extern "C" const int64_t *jtk_result_token_offsets(const jtk_result *r) { return r ? static_cast<const int64_t *>(r->tok_off.p) : nullptr; }
extern "C" const int32_t *jtk_result_doc_status(const jtk_result *r) { return r ? static_cast<const int32_t *>(r->status.p) : nullptr; }
extern "C" double jtk_result_device_ms(const jtk_result *r) { return r ? r->device_ms : 0; }
extern "C" int64_t jtk_result_gpu_launches(const jtk_result *r) { return r ? r->launches : 0; }
extern "C" const uint8_t *jtk_result_bytes(const jtk_result *r) { return r ? static_cast<const uint8_t *>(r->bytes.p) : nullptr; }
extern "C" const int64_t *jtk_result_byte_offsets(const jtk_result *r) { return r ? static_cast<const int64_t *>(r->byte_off.p) : nullptr; }
extern "C" const int32_t *jtk_result_bad_ids(const jtk_result *r) { return r ? static_cast<const int32_t *>(r->bad_ids.p) : nullptr; }

extern "C" void jtk_result_free(jtk_result *r) {
	if (!r) return;
	jtk_encoding *e = r->enc;
	pinned_put(e, r->ids);
	pinned_put(e, r->tok_off);
	pinned_put(e, r->status);
	pinned_put(e, r->bytes);
	pinned_put(e, r->byte_off);
	pinned_put(e, r->bad_ids);
	delete r;
}

/* ------------------------------------------------------------------ decode */
/* Runs the decode kernels on the encoding's first device.  On success *h_bytes (pinned) holds the bytes,
 * h_id_off (optional, nids + 1) the byte offset of every token, r->byte_off / status / bad_ids per document. */
/* ------------------------------------------------------------------ special-token encoding
 * Not in the reference (README.md:46 "not started"; encodeInternal throws, GptBytePairEncoding.java:52-56): semantics of
 * tiktoken's encode(text, allowed_special="all").  One device, whole batch resident (a convenience path, not the pipelined
 * hot path): find the occurrences per document, encode the segments between them as documents of their own with the
 * ordinary kernels, replace each special segment by the token's id. */
extern "C" int jtk_encode_batch_special(jtk_encoding *e, const uint8_t *utf8, const int64_t *doc_off, int64_t ndocs, uint32_t flags, jtk_result **out) {
	if (!e || !doc_off || !out || ndocs < 0) return set_error(JTK_E_ARG, "null argument");
	*out = nullptr;
	if (doc_off[0] != 0) return set_error(JTK_E_ARG, "doc_off[0] must be 0");
	for (int64_t d = 0; d < ndocs; d++)
		if (doc_off[d + 1] < doc_off[d]) return set_error(JTK_E_ARG, "doc_off is not monotone");
	const int64_t total = doc_off[ndocs];
	if (total > 0 && !utf8) return set_error(JTK_E_ARG, "utf8 is null");
	jtk_device_state *ds = e->devs[0];
	CUDA_TRY(cudaSetDevice(ds->device));
	jtk_workspace *w = acquire_ws(ds);
	jtk_result *r = new jtk_result();
	r->enc = e;
	r->ndocs = ndocs;
	cudaStream_t st = nullptr;
	std::vector<void *> dev; /* device buffers of this call */
	auto dalloc = [&](size_t bytes) -> void * {
		void *p = nullptr;
		if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) {
			cudaGetLastError();
			return nullptr;
		}
		dev.push_back(p);
		return p;
	};
	int rc = JTK_OK;
	do {
		cudaError_t ce;
#define STRY(expr)                                                                              \
	if ((ce = (expr)) != cudaSuccess) {                                                         \
		rc = set_error(JTK_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(ce));        \
		break;                                                                                  \
	}
#define SALLOC(var, type, count)                                           \
	type *var = static_cast<type *>(dalloc(sizeof(type) * (size_t) (count))); \
	if (!var) {                                                            \
		rc = set_error(JTK_E_NOMEM, "device allocation failed");          \
		break;                                                             \
	}
		STRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
		SALLOC(d_in, uint8_t, total + 64);
		SALLOC(d_doc, int64_t, ndocs + 1);
		SALLOC(d_base, int64_t, ndocs + 1);
		SALLOC(d_sums, int64_t, jtk_scan_blocks(ndocs + 1) + 1);
		SALLOC(d_total, int64_t, 1);
		STRY(cudaMemsetAsync(d_in + total, 0, 64, st));
		if (total > 0) STRY(cudaMemcpyAsync(d_in, utf8, (size_t) total, cudaMemcpyHostToDevice, st));
		STRY(cudaMemcpyAsync(d_doc, doc_off, sizeof(int64_t) * (size_t) (ndocs + 1), cudaMemcpyHostToDevice, st));
		jtk_special_args a;
		memset(&a, 0, sizeof(a));
		a.T = ds->T;
		a.bytes = d_in;
		a.total = total;
		a.doc_off = d_doc;
		a.ndocs = ndocs;
		a.match_base = d_base;
		STRY(jtk_launch_special_count(a, st));
		STRY(jtk_launch_exclusive_scan(d_base, ndocs + 1, d_sums, d_total, st));
		int64_t nmatch = 0;
		STRY(cudaMemcpyAsync(&nmatch, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
		STRY(cudaStreamSynchronize(st));
		const int64_t nseg = ndocs + 2 * nmatch;
		a.nseg = nseg;
		SALLOC(d_seg_off, int64_t, nseg + 1);
		SALLOC(d_seg_special, int32_t, nseg + 1);
		SALLOC(d_seg_ids, int32_t, total + 16);
		SALLOC(d_seg_tok, int64_t, nseg + 1);
		SALLOC(d_seg_st, int32_t, nseg + 1);
		SALLOC(d_shift, int64_t, nseg + 1);
		SALLOC(d_sums2, int64_t, jtk_scan_blocks(nseg + 1) + 1);
		a.seg_off = d_seg_off;
		a.seg_special = d_seg_special;
		STRY(jtk_launch_special_fill(a, st));
		STRY(cudaMemsetAsync(d_seg_st, 0, sizeof(int32_t) * (size_t) (nseg + 1), st));
		jtk_device_info info;
		memset(&info, 0, sizeof(info));
		jtk_memo_buf *memo = acquire_memo(ds);
		rc = encode_device_impl(e, ds, w, d_in, total, d_seg_off, nseg, flags & ~(uint32_t) (JTK_CHECK_SPECIAL | JTK_COUNT_ONLY | JTK_TIME_KERNEL), d_seg_ids, total + 16,
		                        d_seg_tok, d_seg_st, nullptr, st, &info, true, memo);
		release_memo(ds, memo);
		if (rc != JTK_OK) break;
		const int64_t nseg_tok = info.num_tokens;
		a.seg_ids = d_seg_ids;
		a.seg_tok_off = d_seg_tok;
		a.seg_status = d_seg_st;
		a.shift = d_shift;
		STRY(jtk_launch_special_shift(a, st));
		STRY(jtk_launch_exclusive_scan(d_shift, nseg + 1, d_sums2, d_total, st));
		int64_t shift_total = 0;
		STRY(cudaMemcpyAsync(&shift_total, d_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
		STRY(cudaStreamSynchronize(st));
		const int64_t ntok = nseg_tok + shift_total;
		SALLOC(d_ids, int32_t, ntok + 1);
		SALLOC(d_tok, int64_t, ndocs + 1);
		SALLOC(d_st, int32_t, ndocs + 1);
		a.ids = d_ids;
		a.tok_off = d_tok;
		a.doc_status = d_st;
		STRY(jtk_launch_special_gather(a, nseg_tok, st));
		if ((rc = pinned_get(e, sizeof(int32_t) * std::max<int64_t>(ntok, 1), &r->ids)) != JTK_OK) break;
		if ((rc = pinned_get(e, sizeof(int64_t) * (ndocs + 1), &r->tok_off)) != JTK_OK) break;
		if ((rc = pinned_get(e, sizeof(int32_t) * (ndocs + 1), &r->status)) != JTK_OK) break;
		if (ntok > 0) STRY(cudaMemcpyAsync(r->ids.p, d_ids, sizeof(int32_t) * (size_t) ntok, cudaMemcpyDeviceToHost, st));
		STRY(cudaMemcpyAsync(r->tok_off.p, d_tok, sizeof(int64_t) * (size_t) (ndocs + 1), cudaMemcpyDeviceToHost, st));
		if (ndocs > 0) STRY(cudaMemcpyAsync(r->status.p, d_st, sizeof(int32_t) * (size_t) ndocs, cudaMemcpyDeviceToHost, st));
		STRY(cudaStreamSynchronize(st));
		r->ntokens = ntok;
		r->launches = info.gpu_launches + 10;
#undef STRY
#undef SALLOC
	} while (0);
	for (void *p : dev) cudaFree(p);
	if (st) cudaStreamDestroy(st);
	release_ws(ds, w);
	if (rc != JTK_OK) {
		jtk_result_free(r);
		return rc;
	}
	*out = r;
	return JTK_OK;
}

/* Decode of a device-resident batch.  d_out == nullptr: size query (count pass + scan, *total_bytes).  Otherwise the single-pass
 * kernel (ticketed tiles, decoupled look-back) writes the bytes; JTK_E_CAPACITY with *total_bytes set when they do not fit.
 * Scratch lives in the workspace. */
static int decode_device_impl(jtk_device_state *ds, jtk_workspace *w, const int32_t *d_ids, int64_t nids, const int64_t *d_tok_off, int64_t ndocs, uint8_t *d_out,
                              int64_t out_capacity, int64_t *d_byte_off, int32_t *d_doc_status, int32_t *d_bad_ids, cudaStream_t st, int64_t *total_bytes,
                              int64_t *launches) {
	if ((reinterpret_cast<uintptr_t>(d_ids) & 15) != 0) return set_error(JTK_E_ARG, "d_ids must be 16-byte aligned");
	const int64_t nt = jtk_decode_tiles(nids);
	if (nt + 1 > w->dec_tiles_cap) {
		cudaFree(w->dec_tile);
		cudaFree(w->dec_sums);
		w->dec_tile = w->dec_sums = nullptr;
		w->dec_tiles_cap = 0;
		const int64_t cap = nt + nt / 4 + 16;
		CUDA_TRY(cudaMalloc(&w->dec_tile, sizeof(int64_t) * (size_t) (2 * cap))); /* tile states / byte counts, then the tiles' first documents */
		CUDA_TRY(cudaMalloc(&w->dec_sums, sizeof(int64_t) * (size_t) (jtk_scan_blocks(cap) + 1)));
		w->dec_tiles_cap = cap;
	}
	if (!w->dec_total) {
		CUDA_TRY(cudaMalloc(&w->dec_total, 4 * sizeof(int64_t))); /* byte count, ticket, overflow flag */
		CUDA_TRY(cudaHostAlloc(&w->dec_total_host, 4 * sizeof(int64_t), cudaHostAllocDefault));
	}
	if (ndocs + 1 > w->dec_docs_cap) {
		cudaFree(w->dec_badpos);
		w->dec_badpos = nullptr;
		w->dec_docs_cap = 0;
		const int64_t cap = ndocs + ndocs / 4 + 16;
		CUDA_TRY(cudaMalloc(&w->dec_badpos, sizeof(unsigned long long) * (size_t) cap));
		w->dec_docs_cap = cap;
	}
	jtk_decode_args a;
	memset(&a, 0, sizeof(a));
	a.T = ds->T;
	a.ids = d_ids;
	a.nids = nids;
	a.tok_off = d_tok_off;
	a.ndocs = ndocs;
	a.tile_bytes = w->dec_tile;
	a.ntiles = std::max<int64_t>(nt, 1);
	a.tile_state = reinterpret_cast<unsigned long long *>(w->dec_tile);
	a.tile_first_doc = w->dec_tile + w->dec_tiles_cap;
	a.total_out = reinterpret_cast<long long *>(w->dec_total);
	a.ticket = reinterpret_cast<unsigned int *>(w->dec_total + 1);
	a.overflow = reinterpret_cast<unsigned int *>(w->dec_total + 2);
	a.out_capacity = out_capacity;
	a.bad_pos = w->dec_badpos;
	a.out = d_out;
	a.byte_off = d_byte_off;
	a.doc_status = d_doc_status;
	a.bad_ids = d_bad_ids;
	CUDA_TRY(cudaMemsetAsync(w->dec_badpos, 0xFF, sizeof(unsigned long long) * (size_t) (ndocs + 1), st));
	if (!d_out) {
		CUDA_TRY(jtk_launch_decode_count(a, w->dec_sums, w->dec_total, st));
		CUDA_TRY(cudaMemcpyAsync(w->dec_total_host, w->dec_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(cudaStreamSynchronize(st));
		*launches = (nt > 0 ? 1 : 0) + 3;
		*total_bytes = *w->dec_total_host;
		return JTK_OK;
	}
	CUDA_TRY(cudaMemsetAsync(w->dec_tile, 0, sizeof(int64_t) * (size_t) a.ntiles, st));
	CUDA_TRY(cudaMemsetAsync(w->dec_total, 0, 4 * sizeof(int64_t), st));
	CUDA_TRY(jtk_launch_decode_fused(a, st));
	CUDA_TRY(cudaMemcpyAsync(w->dec_total_host, w->dec_total, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(cudaStreamSynchronize(st));
	*launches = 2 + (ndocs > 0 ? 1 : 0);
	*total_bytes = w->dec_total_host[0];
	if (w->dec_total_host[2] != 0) return set_error(JTK_E_CAPACITY, "output buffer too small for the decoded bytes");
	return JTK_OK;
}

extern "C" int jtk_decode_batch_device(jtk_encoding *e, int device, const int32_t *d_ids, int64_t nids, const int64_t *d_tok_off, int64_t ndocs, uint8_t *d_out,
                                       int64_t out_capacity, int64_t *d_byte_off, int32_t *d_doc_status, int32_t *d_bad_ids, void *cuda_stream,
                                       int64_t *total_bytes, int64_t *gpu_launches) {
	if (!e || !d_tok_off || !total_bytes || nids < 0 || ndocs < 0 || (nids > 0 && !d_ids) || (d_out && (!d_byte_off || !d_doc_status || !d_bad_ids)))
		return set_error(JTK_E_ARG, "null argument");
	int di = device_index(e, device);
	if (di < 0) return set_error(JTK_E_ARG, "device is not one of the encoding's devices");
	jtk_device_state *ds = e->devs[(size_t) di];
	CUDA_TRY(cudaSetDevice(device));
	jtk_workspace *w = acquire_ws(ds);
	int64_t launches = 0;
	int rc = decode_device_impl(ds, w, d_ids, nids, d_tok_off, ndocs, d_out, out_capacity, d_byte_off, d_doc_status, d_bad_ids,
	                            reinterpret_cast<cudaStream_t>(cuda_stream), total_bytes, &launches);
	if (gpu_launches) *gpu_launches = launches;
	release_ws(ds, w);
	return rc;
}

/* Host-buffer decode on the encoding's first device.  On success r->bytes (pinned) holds the bytes, r->byte_off / status / bad_ids
 * the per-document arrays. */
static int decode_impl(jtk_encoding *e, const int32_t *ids, const int64_t *tok_off, int64_t ndocs, jtk_result *r) {
	jtk_device_state *ds = e->devs[0];
	CUDA_TRY(cudaSetDevice(ds->device));
	const int64_t nids = tok_off[ndocs];
	if (tok_off[0] != 0) return set_error(JTK_E_ARG, "tok_off[0] must be 0");
	for (int64_t d = 0; d < ndocs; d++)
		if (tok_off[d + 1] < tok_off[d]) return set_error(JTK_E_ARG, "tok_off is not monotone");
	cudaStream_t st;
	CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	jtk_workspace *w = acquire_ws(ds);
	int32_t *d_ids = nullptr, *d_status = nullptr, *d_bad = nullptr;
	int64_t *d_tok_off = nullptr, *d_byte_off = nullptr;
	uint8_t *d_out = nullptr;
	int rc = JTK_OK;
	do {
		cudaError_t ce;
#define DTRY(expr)                                                                       \
	if ((ce = (expr)) != cudaSuccess) {                                                  \
		rc = set_error(JTK_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(ce)); \
		break;                                                                           \
	}
		DTRY(cudaMalloc(&d_ids, sizeof(int32_t) * (size_t) std::max<int64_t>(nids, 4)));
		DTRY(cudaMalloc(&d_tok_off, sizeof(int64_t) * (size_t) (ndocs + 1)));
		DTRY(cudaMalloc(&d_byte_off, sizeof(int64_t) * (size_t) (ndocs + 1)));
		DTRY(cudaMalloc(&d_status, sizeof(int32_t) * (size_t) (ndocs + 1)));
		DTRY(cudaMalloc(&d_bad, sizeof(int32_t) * (size_t) (ndocs + 1)));
		if (nids > 0) DTRY(cudaMemcpyAsync(d_ids, ids, sizeof(int32_t) * (size_t) nids, cudaMemcpyHostToDevice, st));
		DTRY(cudaMemcpyAsync(d_tok_off, tok_off, sizeof(int64_t) * (size_t) (ndocs + 1), cudaMemcpyHostToDevice, st));
		DTRY(cudaMemsetAsync(d_status, 0, sizeof(int32_t) * (size_t) (ndocs + 1), st));
		int64_t total = 0, launches = 0;
		if ((rc = decode_device_impl(ds, w, d_ids, nids, d_tok_off, ndocs, nullptr, 0, nullptr, nullptr, nullptr, st, &total, &launches)) != JTK_OK) break;
		DTRY(cudaMalloc(&d_out, (size_t) std::max<int64_t>(total, 1) + 16));
		int64_t launches2 = 0;
		if ((rc = decode_device_impl(ds, w, d_ids, nids, d_tok_off, ndocs, d_out, total, d_byte_off, d_status, d_bad, st, &total, &launches2)) != JTK_OK) break;
		r->nbytes = total;
		r->launches = launches + launches2;
		if ((rc = pinned_get(e, total, &r->bytes)) != JTK_OK) break;
		if ((rc = pinned_get(e, sizeof(int64_t) * (ndocs + 1), &r->byte_off)) != JTK_OK) break;
		if ((rc = pinned_get(e, sizeof(int32_t) * (ndocs + 1), &r->status)) != JTK_OK) break;
		if ((rc = pinned_get(e, sizeof(int32_t) * (ndocs + 1), &r->bad_ids)) != JTK_OK) break;
		if (total > 0) DTRY(cudaMemcpyAsync(r->bytes.p, d_out, (size_t) total, cudaMemcpyDeviceToHost, st));
		DTRY(cudaMemcpyAsync(r->byte_off.p, d_byte_off, sizeof(int64_t) * (size_t) (ndocs + 1), cudaMemcpyDeviceToHost, st));
		DTRY(cudaMemcpyAsync(r->status.p, d_status, sizeof(int32_t) * (size_t) (ndocs + 1), cudaMemcpyDeviceToHost, st));
		DTRY(cudaMemcpyAsync(r->bad_ids.p, d_bad, sizeof(int32_t) * (size_t) (ndocs + 1), cudaMemcpyDeviceToHost, st));
		DTRY(cudaStreamSynchronize(st));
#undef DTRY
	} while (0);
	cudaFree(d_ids);
	cudaFree(d_tok_off);
	cudaFree(d_byte_off);
	cudaFree(d_status);
	cudaFree(d_bad);
	cudaFree(d_out);
	cudaStreamDestroy(st);
	release_ws(ds, w);
	return rc;
}

extern "C" int jtk_decode_batch(jtk_encoding *e, const int32_t *ids, const int64_t *tok_off, int64_t ndocs, jtk_result **out) {
	if (!e || !tok_off || !out || ndocs < 0) return set_error(JTK_E_ARG, "null argument");
	*out = nullptr;
	if (tok_off[ndocs] > 0 && !ids) return set_error(JTK_E_ARG, "ids is null");
	jtk_result *r = new jtk_result();
	r->enc = e;
	r->ndocs = ndocs;
	r->ntokens = tok_off[ndocs];
	int rc = decode_impl(e, ids, tok_off, ndocs, r);
	if (rc != JTK_OK) {
		std::string keep = g_last_error;
		jtk_result_free(r);
		return set_error(rc, keep);
	}
	*out = r;
	return JTK_OK;
}

/* ------------------------------------------------------------------ encode(text, maxTokens) */
/* new String(bytes, UTF_8) as the JDK decodes it: UTF-16 code units, one U+FFFD per malformed or truncated
 * sequence (the JDK is not part of /root/reference; this is what String.startsWith in the reference's
 * back-off loop, GptBytePairEncoding.java:90-100, compares).  In the JTokkit integration this loop runs in the
 * Java shim with the JVM's own decoder; this C version serves non-JVM callers. */
static void java_utf16(const uint8_t *src, int64_t sl, std::vector<uint16_t> *dst) {
	dst->clear();
	const uint16_t REPL = 0xFFFD;
	auto nc = [](uint8_t b) { return (b & 0xC0) != 0x80; };
	int64_t sp = 0;
	while (sp < sl) {
		uint8_t b1 = src[sp++];
		if (b1 < 0x80) {
			dst->push_back(b1);
		} else if ((b1 & 0xE0) == 0xC0 && (b1 & 0x1E) != 0) {
			if (sp >= sl) {
				dst->push_back(REPL);
				break;
			}
			uint8_t b2 = src[sp++];
			if (nc(b2)) {
				dst->push_back(REPL);
				sp--;
			} else dst->push_back((uint16_t) (((b1 & 0x1F) << 6) | (b2 & 0x3F)));
		} else if ((b1 & 0xF0) == 0xE0) {
			if (sp + 1 < sl) {
				uint8_t b2 = src[sp], b3 = src[sp + 1];
				if ((b1 == 0xE0 && (b2 & 0xE0) == 0x80) || nc(b2) || nc(b3)) {
					dst->push_back(REPL);
					sp += ((b1 == 0xE0 && (b2 & 0xE0) == 0x80) || nc(b2)) ? 0 : 1;
				} else {
					uint16_t c = (uint16_t) (((b1 & 0x0F) << 12) | ((b2 & 0x3F) << 6) | (b3 & 0x3F));
					dst->push_back((c >= 0xD800 && c <= 0xDFFF) ? REPL : c);
					sp += 2;
				}
				continue;
			}
			dst->push_back(REPL);
			if (sp < sl && ((b1 == 0xE0 && (src[sp] & 0xE0) == 0x80) || nc(src[sp]))) continue;
			break;
		} else if ((b1 & 0xF8) == 0xF0) {
			if (sp + 2 < sl) {
				uint8_t b2 = src[sp], b3 = src[sp + 1], b4 = src[sp + 2];
				uint32_t uc = ((uint32_t) (b1 & 0x07) << 18) | ((uint32_t) (b2 & 0x3F) << 12) | ((uint32_t) (b3 & 0x3F) << 6) | (b4 & 0x3F);
				if (nc(b2) || nc(b3) || nc(b4) || uc < 0x10000 || uc > 0x10FFFF) {
					dst->push_back(REPL);
					if (b1 > 0xF4 || (b1 == 0xF0 && (b2 < 0x90 || b2 > 0xBF)) || (b1 == 0xF4 && (b2 & 0xF0) != 0x80) || nc(b2)) sp += 0;
					else if (nc(b3)) sp += 1;
					else sp += 2;
				} else {
					uc -= 0x10000;
					dst->push_back((uint16_t) (0xD800 + (uc >> 10)));
					dst->push_back((uint16_t) (0xDC00 + (uc & 0x3FF)));
					sp += 3;
				}
				continue;
			}
			dst->push_back(REPL);
			if (b1 > 0xF4 || (sp < sl && ((b1 == 0xF0 && (src[sp] < 0x90 || src[sp] > 0xBF)) || (b1 == 0xF4 && (src[sp] & 0xF0) != 0x80) || nc(src[sp])))) continue;
			sp++;
			if (sp < sl && nc(src[sp])) continue;
			break;
		} else {
			dst->push_back(REPL);
		}
	}
}

extern "C" int jtk_encode_max_tokens(jtk_encoding *e, const uint8_t *utf8, int64_t nbytes, int32_t max_tokens, uint32_t flags, int32_t **ids,
                                     int64_t *num_ids, int32_t *truncated, int32_t *doc_status) {
	if (!e || !ids || !num_ids || !truncated || !doc_status || nbytes < 0 || (nbytes > 0 && !utf8)) return set_error(JTK_E_ARG, "null argument");
	*ids = nullptr;
	*num_ids = 0;
	*truncated = 0;
	*doc_status = 0;
	/* full device encode (the reference stops the find loop early, :79; the first min(maxTokens, total) tokens are the same) */
	const int64_t doc_off[2] = {0, nbytes};
	jtk_result *r = nullptr;
	int rc = jtk_encode_batch(e, utf8, doc_off, 1, flags & ~JTK_COUNT_ONLY, &r);
	if (rc != JTK_OK) return rc;
	*doc_status = jtk_result_doc_status(r)[0];
	int64_t k = std::min<int64_t>(std::max<int32_t>(max_tokens, 0), jtk_result_num_tokens(r));
	std::vector<int32_t> tok(jtk_result_ids(r), jtk_result_ids(r) + k);
	jtk_result_free(r);
	if (*doc_status & JTK_DOC_HAS_SPECIAL) return JTK_OK; /* the shim throws UnsupportedOperationException */
	/* decode the clipped prefix once on the device; token byte offsets give every shorter prefix */
	jtk_result dr;
	dr.enc = e;
	const int64_t tok_off[2] = {0, k};
	rc = decode_impl(e, tok.data(), tok_off, 1, &dr);
	/* byte offset of every token of the clipped prefix (lengths from the registration-time tables: the bytes themselves came from the device) */
	std::vector<int64_t> id_off((size_t) k + 1, 0);
	for (int64_t i = 0; i < k && rc == JTK_OK; i++) {
		const jtk_host_tables &h = e->host;
		int64_t len = -1;
		uint32_t sl = jtk_hash_pair(tok[(size_t) i], 0) & h.mask_d;
		for (;;) {
			const uint32_t v = h.dec_keys[2 * sl + 1];
			if (v == 0) break;
			if (h.dec_keys[2 * sl] == (uint32_t) tok[(size_t) i]) {
				len = (int64_t) h.dec_off[v] - h.dec_off[v - 1];
				break;
			}
			sl = (sl + 1) & h.mask_d;
		}
		if (len < 0) rc = set_error(JTK_E_ARG, "internal: encode produced an id the decode table does not know");
		id_off[(size_t) i + 1] = id_off[(size_t) i] + std::max<int64_t>(len, 0);
	}
	std::vector<uint8_t> dbytes;
	if (rc == JTK_OK) dbytes.assign(static_cast<uint8_t *>(dr.bytes.p), static_cast<uint8_t *>(dr.bytes.p) + dr.nbytes);
	pinned_put(e, dr.bytes);
	pinned_put(e, dr.byte_off);
	pinned_put(e, dr.status);
	pinned_put(e, dr.bad_ids);
	if (rc != JTK_OK) return rc;
	std::vector<uint16_t> t16, d16;
	java_utf16(utf8, nbytes, &t16);
	int64_t keep = 0;
	for (int64_t drop = 0; drop <= k; drop++) { /* :90-100 */
		java_utf16(dbytes.data(), id_off[(size_t) (k - drop)], &d16);
		if (d16.size() <= t16.size() && std::equal(d16.begin(), d16.end(), t16.begin())) {
			keep = k - drop;
			*truncated = t16.size() > d16.size();
			break;
		}
	}
	*ids = static_cast<int32_t *>(malloc(sizeof(int32_t) * (size_t) std::max<int64_t>(keep, 1)));
	if (!*ids) return set_error(JTK_E_NOMEM, "malloc failed");
	memcpy(*ids, tok.data(), sizeof(int32_t) * (size_t) keep);
	*num_ids = keep;
	return JTK_OK;
}
