/*
 * ORACLE - TEST INFRASTRUCTURE ONLY.  A CPU restatement of JTokkit's encode hot path
 * (lib/src/main/java/com/knuddels/jtokkit/GptBytePairEncoding.java) used as the checker for the
 * CUDA path and as the timed "port" CPU baseline.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (jtokkit_b200/) never does.
 *
 * Parity pin: checked against the reference's golden vectors tests/golden/{cl100k_base,r50k_base,
 * p50k_base,p50k_edit}_encodings.csv (copied from lib/src/test/resources, asserted by
 * lib/src/test/java/com/knuddels/jtokkit/reference/ (the four ...BaseTest classes, lines 19-111)) and cross-checked against
 * tiktoken 0.12.0 built from the same vocab files and regex strings (tests/test_oracle.py).
 * The reference itself (Java) cannot run in this image (no JVM), so oracle/_ref does not exist.
 */
#ifndef JO_ORACLE_H
#define JO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jo_encoding jo_encoding;

/* status codes (negative return values) */
#define JO_E_SPECIAL (-1)      /* UnsupportedOperationException, GptBytePairEncoding.java:52-56 */
#define JO_E_UNKNOWN_BYTES (-2) /* IllegalArgumentException "Unknown token for encoding", TokenEncoder.java:64-71 */
#define JO_E_UNKNOWN_ID (-3)    /* IllegalArgumentException "Unknown token for decoding", GptBytePairEncoding.java:313 */
#define JO_E_CAPACITY (-4)

/* merge algorithm selection */
#define JO_MERGE_LITERAL 0 /* the reference's O(n^2) loop, GptBytePairEncoding.java:200-275 */
#define JO_MERGE_HEAP 1    /* exact (rank, position) heap with lazy invalidation (same result, O(n log n)) */
#define JO_MERGE_AUTO 2    /* literal for pieces <= 256 bytes, heap above */

/* Mirrors new GptBytePairEncoding(GptBytePairEncodingParams) (GptBytePairEncoding.java:30-35):
 * pattern + java.util.regex flag bits, the encoder map flattened to (keys, key_off[n+1], ranks[n]),
 * the special-token map flattened to (spec, spec_off[m+1], spec_ids[m]). */
jo_encoding *jo_create(const char *pattern, int flags, const uint8_t *keys, const int64_t *key_off, const int32_t *ranks, int64_t nkeys,
                       const uint8_t *spec, const int64_t *spec_off, const int32_t *spec_ids, int64_t nspec, char *err, int errlen);
void jo_destroy(jo_encoding *enc);

/* The matcher.find() loop alone (GptBytePairEncoding.java:77-80): byte offsets of every match. */
int64_t jo_split(const jo_encoding *enc, const uint8_t *text, int64_t n, int64_t *starts, int64_t *ends, int64_t cap);

/* encodeOrdinary / encode (check_special) without maxTokens.  Returns the token count or a JO_E_*. */
int64_t jo_encode(const jo_encoding *enc, const uint8_t *text, int64_t n, int check_special, int merge_algo, int32_t *out, int64_t cap);

/* encode(text, maxTokens) / encodeOrdinary(text, maxTokens) incl. the back-off loop (:90-100). */
int64_t jo_encode_max(const jo_encoding *enc, const uint8_t *text, int64_t n, int check_special, int max_tokens, int32_t *out, int64_t cap,
                      int *truncated);

/* text.contains(specialToken) for any special token (:52-56). */
int jo_contains_special(const jo_encoding *enc, const uint8_t *text, int64_t n);

/* special-token encoding as tiktoken's encode(text, allowed_special="all") (the reference has none, README.md:46) */
int64_t jo_encode_with_special(const jo_encoding *enc, const uint8_t *text, int64_t n, int merge_algo, int32_t *out, int64_t cap);

/* decodeBytes (:136-151).  Returns byte count or JO_E_UNKNOWN_ID (bad id in *bad_id). */
int64_t jo_decode_bytes(const jo_encoding *enc, const int32_t *ids, int64_t n, uint8_t *out, int64_t cap, int32_t *bad_id);

/* One bytePairMerge call on a raw piece (for merge-variant tests). */
int64_t jo_merge_piece(const jo_encoding *enc, const uint8_t *piece, int64_t n, int merge_algo, int32_t *out, int64_t cap);

/* new String(bytes, UTF_8) -> UTF-16 code units with the JDK's U+FFFD replacement (for the back-off loop). */
int64_t jo_java_utf8_to_utf16(const uint8_t *b, int64_t n, uint16_t *out);

/* Per-document thread pool, one task per document as AbstractMultiThreadedBenchmark.java:34-45.
 * Tokens of document d are written to ids[doc_off[d] ...] (a document never has more tokens than bytes);
 * counts[d] receives the token count or a JO_E_* status.  Returns the total token count. */
int64_t jo_encode_batch(const jo_encoding *enc, const uint8_t *utf8, const int64_t *doc_off, int64_t ndocs, int nthreads, int check_special,
                        int merge_algo, int32_t *ids, int64_t *counts);

#ifdef __cplusplus
}
#endif
#endif
