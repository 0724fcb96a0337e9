/*
 * Kernel-side declarations shared between jtk_kernels.cu and the C ABI implementation.
 */
#ifndef JTK_KERNELS_CUH
#define JTK_KERNELS_CUH

#include <cuda_runtime.h>

#include "../../include/jtokkit_b200.h"
#include "jtk_common.h"

/* A piece longer than JTK_LONG_PIECE bytes, deferred by the tile kernel. */
struct jtk_long_piece {
	int64_t start;     /* global byte position */
	int64_t end;       /* filled by jtk_long_bounds_kernel */
	int64_t insert_at; /* index in the tile kernel's id stream where its tokens belong */
	int64_t count;     /* tokens produced by the long-piece kernel */
	int64_t scratch;   /* offset of its scratch area (in parts) */
	int32_t doc;       /* unused by the kernels; for diagnostics */
	int32_t flags;
};

/* Header read back by the host after a batch. */
struct jtk_batch_header {
	unsigned long long total_tokens;
	unsigned int n_long;
	unsigned int overflow;   /* ids capacity exceeded */
	unsigned int long_next;  /* work counter of the long-piece kernel */
	unsigned int violations; /* long-piece rounds that had to fall back to one-merge-at-a-time */
	unsigned int rx_ticket;  /* slice counter of jtk_general_slice_kernel */
	unsigned int rx_ticket2; /* document counter of jtk_general_stitch_kernel */
};

/* Counters of one sub-batch.  Two sub-batches are in flight at a time (the split+lookup kernel of one runs while the merge /
 * gather kernels of the previous one finish), so there are two of these and two sets of per-sub-batch buffers ("lanes"). */
struct jtk_sub_header {
	unsigned int ticket;     /* tile ticket counter */
	unsigned int n_med8, n_med32;   /* medium pieces of the sub-batch (65..256 bytes / 257..JTK_LONG_PIECE bytes) */
	unsigned int cursor8, cursor32; /* work counters of jtk_merge_medium_kernel */
	unsigned int short_cnt[JTK_SHORT_PIECE + 2];  /* unresolved short pieces of the sub-batch by length */
	unsigned int short_base[JTK_SHORT_PIECE + 2]; /* exclusive scan of short_cnt */
	unsigned int short_cur[JTK_SHORT_PIECE + 2];  /* scatter cursors */
	unsigned int short_next[4];                   /* work counters of the three jtk_merge_short_kernel launches */
	unsigned int pad;
};

#define JTK_RECN (JTK_TILE + JTK_FWD_HALO) /* per-tile slots of rec / slowtok: pieces + unresolved pieces <= JTK_RECN */
#define JTK_QCAP (JTK_RECN / 2)            /* an unresolved piece has at least two bytes */
#define JTK_REC_MIN_ID (-(1 << 30))        /* token ids below this are rejected at registration; the space encodes piece records */
#define JTK_GROUP8_PIECE 256                /* up to this length a medium piece is merged by 8 lanes, above by a warp */
#define JTK_MED8_PER_TILE (JTK_RECN / (JTK_SHORT_PIECE + 1) + 1)
#define JTK_MED32_PER_TILE (JTK_RECN / (JTK_GROUP8_PIECE + 1) + 1)
#define JTK_DEFAULT_SUB_TILES ((512 << 20) / JTK_TILE) /* largest sub-batch: 512 MiB of input (measured on the 1 GiB corpus: 128 MiB 14.2 ms, 256 MiB 13.7, 512 MiB 13.5) */
#define JTK_FIRST_SUB_TILES ((16 << 20) / JTK_TILE)    /* a small first sub-batch warms the piece memo */

struct jtk_encode_args {
	jtk_tables T;
	const uint8_t *bytes;
	int64_t total;
	const int64_t *doc_off;
	int64_t ndocs;
	int64_t ntiles;
	int64_t tile_begin, tile_end; /* the sub-batch */
	/* per tile, whole batch */
	const int32_t *tile_first_doc;
	int32_t *npieces;      /* pieces that start in the tile */
	int32_t *nslow;        /* of those, pieces the whole-piece lookup did not resolve */
	int32_t *tile_count;   /* tokens produced by the tile */
	int32_t *tile_slow_used; /* tokens in the dense front part of the tile's slowtok slice */
	int64_t *tile_base;    /* ntiles + 1: exclusive scan of tile_count */
	int64_t *tile_first_b; /* first piece start in the tile (global position) or -1 */
	/* per tile of the sub-batch (index tile - tile_begin) */
	int32_t *rec;          /* JTK_RECN per tile: one record per piece, in order */
	int32_t *slowtok;      /* JTK_RECN per tile: tokens of merged / memoised pieces, densely packed from the front (records hold the offset) */
	uint16_t *slowq;       /* JTK_QCAP per tile: piece indices of the unresolved short pieces (<= JTK_SHORT_PIECE bytes) */
	uint32_t *shortlist;   /* JTK_QCAP per tile (one list for the sub-batch): unresolved short pieces sorted by length */
	uint32_t *med8;        /* JTK_MED8_PER_TILE per tile: (tile index << 14 | piece index) of unresolved pieces of 33..256 bytes */
	uint32_t *med32;       /* JTK_MED32_PER_TILE per tile: same for 257..JTK_LONG_PIECE bytes */
	jtk_batch_header *hdr;
	jtk_sub_header *sub;   /* the lane's counters */
	int32_t *ids;
	int64_t ids_cap;
	int64_t *tok_off;
	int32_t *doc_status;
	uint32_t flags;
	jtk_long_piece *long_list;
	int64_t long_cap;
	uint8_t *piece_flags;  /* debug: one byte per input byte, 1 where a piece starts (nullable) */
	jtk_memo_entry *memo;  /* per-call piece memo (nullable) */
	uint32_t memo_mask, memo_epoch;
	/* JTK_PAT_GENERAL only: one bit per input byte (word g / 32, bit g % 32), written by the jtk_general_* kernels:
	 * piece starts, and which of those pieces are gaps (text the pattern did not match: no tokens) */
	uint32_t *rx_start, *rx_skip;
	int64_t rx_words;
	void *rx_stacks;       /* JTK_RX_THREADS backtrack stacks of JTK_RX_STACK frames */
	uint32_t *rx_spec;     /* three more bit arrays of rx_words words: speculative match starts / ends / resume positions per slice */
	int64_t *rx_rec;       /* 4 * rx_slices + ndocs + 1 int64: per-slice exit / crossing match / join position, per-document exit */
	int64_t rx_slices;
	/* host side only: L2 access-policy window over the hot tables (0 bytes = none) */
	const void *l2_base;
	size_t l2_bytes;
};


cudaError_t jtk_launch_tile_first_doc(const int64_t *doc_off, int64_t ndocs, int64_t ntiles, int32_t *out, cudaStream_t st);
/* Side streams on which the merge kernels of a sub-batch run next to each other: each of them alone leaves most of the GPU
 * idle (few pieces, long dependent chains), together they fill it.  nullptr = everything on the caller's stream. */
struct jtk_side_streams {
	cudaStream_t s[3];
	cudaEvent_t fork, join[3];
};
/* the kernels of one sub-batch in two parts, so that the second part of sub-batch k can run (on its own stream, other lane of
 * buffers) next to the first part of sub-batch k + 1; k0/k1 (nullable) bracket the split+lookup kernel */
cudaError_t jtk_launch_split(const jtk_encode_args &a, int num_sms, int ctas_per_sm, cudaEvent_t k0, cudaEvent_t k1, cudaStream_t st);
/* marks (nullable, development aid): four events recorded after the scatter, after the merge kernels have joined, after the scan and after the gather */
cudaError_t jtk_launch_post(const jtk_encode_args &a, int num_sms, cudaStream_t st, const jtk_side_streams *side, cudaEvent_t *marks = nullptr);
cudaError_t jtk_launch_finalize(const jtk_encode_args &a, cudaStream_t st);
/* JTK_PAT_GENERAL: Matcher.find() over every document before the sub-batches, in three passes (speculative per slice, stitch per
 * document, finish per word: jtk_regex.h); rx_start / rx_skip / rx_spec must be zeroed */
#define JTK_RX_THREADS (148 * 128)
cudaError_t jtk_launch_general_split(const jtk_encode_args &a, cudaStream_t st);
cudaError_t jtk_encode_kernel_setup();

/* long-piece path */
cudaError_t jtk_launch_long_bounds(const jtk_encode_args &a, unsigned int n_long, cudaStream_t st);
/* scratch: JTK_LONG_SCRATCH_ARRAYS int32 arrays of `stride` elements each (stride >= total bytes of the long pieces); the first array holds the tokens afterwards */
#define JTK_LONG_SCRATCH_ARRAYS 8
cudaError_t jtk_launch_long_merge(const jtk_encode_args &a, unsigned int n_long, int32_t *scratch, int64_t stride, int num_sms, cudaStream_t st);
/* list: device copy sorted by start with scratch offsets filled; cum[i] = tokens of long pieces 0..i-1 (n_long + 1 entries) */
cudaError_t jtk_launch_long_insert(const jtk_long_piece *list, const int64_t *cum, unsigned int n_long, const int32_t *scr_tok, const int32_t *ids_in,
                                   int32_t *ids_out, int64_t total_in, cudaStream_t st);
cudaError_t jtk_launch_long_fix_offsets(const jtk_long_piece *list, const int64_t *cum, unsigned int n_long, const int64_t *doc_off, int64_t ndocs,
                                        int64_t *tok_off, cudaStream_t st);

/* Special-token ENCODING (not in the reference: README.md:46; semantics of tiktoken's allowed_special="all").  Every
 * occurrence of a special token cuts its document into segments: text, token, text, ...  The segments are encoded as
 * documents of their own by the ordinary pipeline (the text of a special token yields throw-away tokens), then a gather
 * replaces each special segment by the token's id. */
struct jtk_special_args {
	jtk_tables T;
	const uint8_t *bytes;
	int64_t total;
	const int64_t *doc_off;
	int64_t ndocs;
	int64_t *match_base;    /* ndocs + 1: matches per document, then their exclusive scan */
	int64_t nseg;           /* ndocs + 2 * matches */
	int64_t *seg_off;       /* nseg + 1: the segments as document offsets */
	int32_t *seg_special;   /* nseg: special-token index + 1, 0 for text */
	/* after the ordinary encode of the segments */
	const int32_t *seg_ids;     /* tokens of all segments */
	const int64_t *seg_tok_off; /* nseg + 1 */
	const int32_t *seg_status;  /* nseg */
	int64_t *shift;         /* nseg + 1: per segment (1 - its token count) for special segments, 0 for text; then the exclusive scan */
	int32_t *ids;           /* final tokens */
	int64_t *tok_off;       /* ndocs + 1 */
	int32_t *doc_status;    /* ndocs */
};
cudaError_t jtk_launch_special_count(const jtk_special_args &a, cudaStream_t st);
cudaError_t jtk_launch_special_fill(const jtk_special_args &a, cudaStream_t st);
cudaError_t jtk_launch_special_shift(const jtk_special_args &a, cudaStream_t st);
cudaError_t jtk_launch_special_gather(const jtk_special_args &a, int64_t nseg_tokens, cudaStream_t st);
/* in-place exclusive scan of data[0..n) (int64); *total (device) receives the sum; block_sums: jtk_scan_blocks(n) scratch */
cudaError_t jtk_launch_exclusive_scan(int64_t *data, int64_t n, int64_t *block_sums, int64_t *total, cudaStream_t st);

/* decode path: ids -> bytes (GptBytePairEncoding.decodeBytes / decodeToken, :136-151,302-314) in two passes over tiles of 4 096 tokens */
struct jtk_decode_args {
	jtk_tables T;
	const int32_t *ids; /* 16-byte aligned */
	int64_t nids;
	const int64_t *tok_off;
	int64_t ndocs;
	int64_t *tile_bytes;          /* size query only: jtk_decode_tiles(nids) + 1: bytes per tile, then their exclusive scan */
	int64_t ntiles;               /* max(jtk_decode_tiles(nids), 1) */
	unsigned long long *tile_state; /* ntiles words, zeroed: published byte counts / prefixes of the single-pass kernel */
	int64_t *tile_first_doc;      /* ntiles: scratch */
	unsigned int *ticket;         /* zeroed */
	unsigned int *overflow;       /* zeroed; set when the bytes do not fit out_capacity */
	long long *total_out;         /* device: receives the byte count */
	int64_t out_capacity;
	unsigned long long *bad_pos;  /* ndocs, preset to ~0: position of the first unknown id of a document */
	uint8_t *out;
	int64_t *byte_off; /* ndocs + 1 */
	int32_t *doc_status;
	int32_t *bad_ids;
};
int64_t jtk_scan_blocks(int64_t n);
int64_t jtk_decode_tiles(int64_t nids);
cudaError_t jtk_launch_decode_count(const jtk_decode_args &a, int64_t *block_sums, int64_t *total, cudaStream_t st);
cudaError_t jtk_launch_decode_fused(const jtk_decode_args &a, cudaStream_t st);

#endif
