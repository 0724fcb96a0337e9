/*
 * jtokkit_b200 - C ABI of the B200-native JTokkit encode path.
 *
 * This header is the drop-in boundary: the functions below are exactly what a JTokkit-side FFI
 * (Panama FFM or JNI) binds in place of the pure-Java engine
 *   com.knuddels.jtokkit.GptBytePairEncoding            (lib/src/main/java/com/knuddels/jtokkit/GptBytePairEncoding.java)
 * that EncodingFactory.fromParameters (EncodingFactory.java:117-119) constructs.  Plain C types only,
 * no exceptions cross the boundary, no CUDA or torch types appear in the signatures (streams and device
 * pointers travel as void* / raw pointers).
 *
 * There is NO CPU fallback: every encode/decode entry point runs CUDA kernels compiled for sm_100a and
 * fails with JTK_E_CUDA when no usable device exists.
 */
#ifndef JTOKKIT_B200_H
#define JTOKKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- return codes --------------------------------------------------------------------------- */
#define JTK_OK 0
#define JTK_E_ARG (-1)                 /* bad argument (null pointer, offsets not monotone, ...) */
#define JTK_E_CUDA (-2)                /* CUDA error or no device; see jtk_last_error() */
#define JTK_E_PATTERN_UNSUPPORTED (-3) /* split pattern outside the supported set; registration fails, nothing runs on the CPU */
#define JTK_E_NOMEM (-4)
#define JTK_E_CAPACITY (-5)            /* caller-provided output buffer too small */

/* ---- per-document status (bit flags), mapped back to the reference's exceptions by the host shim - */
#define JTK_DOC_OK 0
#define JTK_DOC_HAS_SPECIAL 1   /* text.contains(specialToken): UnsupportedOperationException, GptBytePairEncoding.java:52-56 */
#define JTK_DOC_UNKNOWN_BYTES 2 /* a final part is not in the vocabulary: IllegalArgumentException, TokenEncoder.java:64-71 */
#define JTK_DOC_UNKNOWN_ID 4    /* decode: IllegalArgumentException("Unknown token for decoding"), GptBytePairEncoding.java:313 */
#define JTK_DOC_PATTERN_STACK 8 /* general split pattern only: the backtracking stack overflowed (java.util.regex: StackOverflowError); the document has no ids */

/* ---- encode flags ---------------------------------------------------------------------------- */
#define JTK_ENCODE_ORDINARY 0u /* encodeOrdinary: GptBytePairEncoding.java:61-64,71-103 */
#define JTK_CHECK_SPECIAL 1u   /* encode: adds the special-token guard of encodeInternal, :47-59 */
#define JTK_COUNT_ONLY 2u      /* countTokens / countTokensOrdinary, :121-129: token offsets only, no ids */
#define JTK_TIME_KERNEL 0x100u /* device-resident call only: bracket the tile kernel with CUDA events (jtk_device_info.tile_kernel_ms) */

/* java.util.regex.Pattern flag bits accepted in jtk_params.pattern_flags (EncodingFactory.java:129) */
#define JTK_RE_CASE_INSENSITIVE 0x02
#define JTK_RE_UNICODE_CASE 0x40
#define JTK_RE_UNICODE_CHARACTER_CLASS 0x100

typedef struct jtk_encoding jtk_encoding; /* immutable after create; any thread may use it concurrently */
typedef struct jtk_result jtk_result;     /* library-owned result of a host-buffer batch call */

/*
 * Registration input = com.knuddels.jtokkit.api.GptBytePairEncodingParams (api/GptBytePairEncodingParams.java:36-46)
 * flattened by the host shim: pattern.pattern() as UTF-8 + pattern.flags(); Map<byte[],Integer> encoder as
 * (vocab_bytes, vocab_off[vocab_size+1], vocab_ranks[vocab_size]); Map<String,Integer> specialTokensEncoder as
 * UTF-8 (special_bytes, special_off[special_size+1], special_ids[special_size]).  Empty maps are legal
 * (BaseEncodingRegistryTest.java:110-125).  Everything is copied; the caller keeps ownership.
 */
typedef struct jtk_params {
	const char *name;
	const char *pattern;
	int32_t pattern_flags;
	const uint8_t *vocab_bytes;
	const int64_t *vocab_off;
	const int32_t *vocab_ranks;
	int64_t vocab_size;
	const uint8_t *special_bytes;
	const int64_t *special_off;
	const int32_t *special_ids;
	int64_t special_size;
} jtk_params;

/* Replaces `new GptBytePairEncoding(params)` (GptBytePairEncoding.java:30-35, reached from
 * EncodingFactory.fromParameters :117-119 and AbstractEncodingRegistry.registerGptBytePairEncoding :63-66).
 * Compiles the split pattern into device tables, flattens the vocabulary into the byte-keyed piece table and
 * the pair table, and replicates them on every listed device (devices == NULL: device 0 only). */
int jtk_encoding_create(const jtk_params *params, const int *devices, int ndev, jtk_encoding **out);

/* Replaces EncodingFactory.r50kBase/p50kBase/p50kEdit/cl100kBase (EncodingFactory.java:60-109) including
 * loadMergeableRanks (:139-164): `name` selects pattern + special tokens, `tiktoken_path` is the vocabulary
 * file in the reference's resource format ("<base64> <rank>\n"). */
int jtk_encoding_create_builtin(const char *name, const char *tiktoken_path, const int *devices, int ndev, jtk_encoding **out);

/* Every jtk_result of the encoding must have been freed before (a result's pinned buffers return to the encoding's pool). */
void jtk_encoding_destroy(jtk_encoding *enc);
const char *jtk_encoding_name(const jtk_encoding *enc); /* Encoding.getName(), api/Encoding.java:189 */
int jtk_encoding_num_devices(const jtk_encoding *enc);

/*
 * The new batch entry point (the shape of the JMH harness' encodeAll(Encoding, List<String>),
 * benchmark/.../AbstractBenchmark.java:37).  Documents are the UTF-8 bytes utf8[doc_off[d] .. doc_off[d+1]),
 * doc_off[0] == 0.  HOST buffers; host<->device copies happen inside the call.  The batch is cut into byte-balanced
 * contiguous document ranges ("chunks", jtk_plan_chunks); chunk c runs on device c % ndev on that device's own streams
 * (like one task per document on the thread pool of AbstractMultiThreadedBenchmark.java:34-45, with a range of documents as
 * the task and a GPU as the worker).  No collective: a chunk's token count is published when its kernels have run, the
 * prefix over earlier chunks is its position in the one pinned result buffer, and its ids are copied there directly.
 * Replaces Encoding.encode / encodeOrdinary / countTokens / countTokensOrdinary (api/Encoding.java:29,80,127,147) for a
 * whole batch.
 */
int jtk_encode_batch(jtk_encoding *enc, const uint8_t *utf8, const int64_t *doc_off, int64_t ndocs, uint32_t flags, jtk_result **out);

/* The chunk plan jtk_encode_batch uses for `ndev` devices (pure host function, no device needed): cuts[0 .. n] with
 * cuts[0] == 0 and cuts[n] == ndocs; chunk c = documents [cuts[c], cuts[c + 1]) runs on device c % ndev.  chunk_bytes <= 0
 * selects the library default (64 MiB, JTK_CHUNK_MB).  Returns n, or the required capacity (> cuts_capacity - 1) without
 * writing when `cuts` is too small, or JTK_E_ARG. */
int64_t jtk_plan_chunks(const int64_t *doc_off, int64_t ndocs, int ndev, int64_t chunk_bytes, int64_t *cuts, int64_t cuts_capacity);

/* Special-token ENCODING: every occurrence of a registered special token becomes its id, the text between occurrences is
 * encoded like encodeOrdinary.  NOT in the reference (README.md:46 lists it as not started; GptBytePairEncoding.java:52-56
 * throws instead) - behind its own entry point for that reason; semantics of the upstream the reference mirrors, tiktoken's
 * encode(text, allowed_special="all"): leftmost-first, non-overlapping, the longest token where several start at the same
 * position.  Same result object as jtk_encode_batch; first device of the encoding only. */
int jtk_encode_batch_special(jtk_encoding *enc, const uint8_t *utf8, const int64_t *doc_off, int64_t ndocs, uint32_t flags, jtk_result **out);

int64_t jtk_result_num_docs(const jtk_result *r);
int64_t jtk_result_num_tokens(const jtk_result *r);
const int32_t *jtk_result_ids(const jtk_result *r);           /* num_tokens ids, NULL for JTK_COUNT_ONLY */
const int64_t *jtk_result_token_offsets(const jtk_result *r); /* ndocs + 1 */
const int32_t *jtk_result_doc_status(const jtk_result *r);    /* ndocs, JTK_DOC_* bits */
double jtk_result_device_ms(const jtk_result *r);             /* kernel time (CUDA events), max over devices */
int64_t jtk_result_gpu_launches(const jtk_result *r);         /* kernels launched for this batch */
void jtk_result_free(jtk_result *r);

/*
 * Device-resident variant: all pointers are device memory on `device` (one of the encoding's devices),
 * work is enqueued on `cuda_stream` (a cudaStream_t passed as void*, NULL = default stream) and the call
 * returns after the stream has been synchronised once to read back the totals.
 *   d_ids        capacity ids_capacity (nbytes is always enough; NULL allowed with JTK_COUNT_ONLY)
 *   d_tok_off    ndocs + 1 token offsets
 *   d_doc_status ndocs status words (must be zeroed by the caller or by a previous call; bits are OR-ed in)
 */
typedef struct jtk_device_info {
	int64_t num_tokens;
	int64_t num_long_pieces; /* pieces longer than the in-tile limit, handled by the long-piece kernels */
	int64_t gpu_launches;
	int32_t reserved;
	float tile_kernel_ms; /* duration of the dominant kernel (jtk_split_lookup_kernel, summed over the sub-batches) when JTK_TIME_KERNEL is set, else 0 */
} jtk_device_info;

int jtk_encode_batch_device(jtk_encoding *enc, int device, const uint8_t *d_utf8, int64_t nbytes, const int64_t *d_doc_off, int64_t ndocs,
                            uint32_t flags, int32_t *d_ids, int64_t ids_capacity, int64_t *d_tok_off, int32_t *d_doc_status, void *cuda_stream,
                            jtk_device_info *info);

/* Debug / test entry: the piece boundaries the split kernel finds (replaces the matcher.find() loop alone,
 * GptBytePairEncoding.java:77-80).  d_piece_flags receives one byte per input byte: 1 where a piece starts. */
int jtk_split_batch_device(jtk_encoding *enc, int device, const uint8_t *d_utf8, int64_t nbytes, const int64_t *d_doc_off, int64_t ndocs,
                           uint8_t *d_piece_flags, void *cuda_stream);

/*
 * Replaces Encoding.decodeBytes (api/Encoding.java:181; GptBytePairEncoding.java:136-151,302-314) for a batch:
 * ids[tok_off[d] .. tok_off[d+1]) -> bytes.  HOST buffers.  The result reuses jtk_result: jtk_result_bytes /
 * jtk_result_byte_offsets; unknown ids set JTK_DOC_UNKNOWN_ID and the first offending id per document is
 * reported through jtk_result_bad_ids.  (String construction, i.e. Encoding.decode's new String(bytes, UTF_8),
 * stays on the JVM side.)
 */
int jtk_decode_batch(jtk_encoding *enc, const int32_t *ids, const int64_t *tok_off, int64_t ndocs, jtk_result **out);
/* Device-resident variant of the same: all pointers are device memory on `device`, d_ids 16-byte aligned; work is enqueued on
 * `cuda_stream` and the call returns when the bytes are written.  d_out == NULL: size query (*total_bytes only).
 * d_out too small: JTK_E_CAPACITY with *total_bytes set.  d_doc_status must be zeroed by the caller. */
int jtk_decode_batch_device(jtk_encoding *enc, int device, const int32_t *d_ids, int64_t nids, const int64_t *d_tok_off, int64_t ndocs, uint8_t *d_out,
                            int64_t out_capacity, int64_t *d_byte_off, int32_t *d_doc_status, int32_t *d_bad_ids, void *cuda_stream, int64_t *total_bytes,
                            int64_t *gpu_launches);
const uint8_t *jtk_result_bytes(const jtk_result *r);
const int64_t *jtk_result_byte_offsets(const jtk_result *r); /* ndocs + 1 */
const int32_t *jtk_result_bad_ids(const jtk_result *r);      /* ndocs */

/*
 * Replaces Encoding.encode(String, int maxTokens) / encodeOrdinary(String, int) (api/Encoding.java:61,107;
 * GptBytePairEncoding.java:42-45,66-69,79,90-100,110-119) for one text: full device encode, clip to
 * maxTokens, then the reference's back-off to a token prefix that decodes to a prefix of the text.
 * *ids is malloc'ed by the library (free with jtk_free); returns JTK_OK and *doc_status.
 * Known divergence: the reference stops its find() loop once maxTokens tokens exist (:79), so text AFTER that point is never
 * looked at; this call encodes the whole text, so *doc_status also reports JTK_DOC_UNKNOWN_BYTES (custom vocabularies
 * without all 256 bytes) or JTK_DOC_PATTERN_STACK (general patterns) that stem from the unread tail, where the reference
 * would return normally.  The predefined encodings cannot produce either status.  The returned ids are always the
 * reference's.
 */
int jtk_encode_max_tokens(jtk_encoding *enc, const uint8_t *utf8, int64_t nbytes, int32_t max_tokens, uint32_t flags, int32_t **ids,
                          int64_t *num_ids, int32_t *truncated, int32_t *doc_status);
void jtk_free(void *p);

/* Pinned host memory for callers that want the fast H2D/D2H path (FFM callers wrap it in a MemorySegment). */
void *jtk_host_alloc(int64_t nbytes);
void jtk_host_free(void *p);

/* Thread-local description of the last error returned on this thread. */
const char *jtk_last_error(void);

/* Library / build identification ("jtokkit_b200 <version> sm_100a unicode <ver>"). */
const char *jtk_version(void);

#ifdef __cplusplus
}
#endif
#endif /* JTOKKIT_B200_H */
