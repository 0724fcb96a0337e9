/*
 * ORACLE - TEST INFRASTRUCTURE ONLY.  Never linked into or called by the product path
 * (jtokkit_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * jo_regex: a small backtracking matcher restating the java.util.regex semantics the reference relies
 * on for its split step (GptBytePairEncoding.java:77-80: pattern.matcher(text) / matcher.find()):
 *   - matching over code points, ordered alternation (first alternative that leads to an overall
 *     match wins), greedy / lazy / possessive quantifiers with backtracking, negative and positive
 *     look-ahead, inline (?i) / (?i:...) groups
 *   - find() resumes at the end of the previous match; after an empty match it advances by one
 *     (java.util.regex.Matcher.find()); unmatched characters are skipped silently
 *   - flags use the java.util.regex.Pattern bit values (CASE_INSENSITIVE 0x02, UNICODE_CASE 0x40,
 *     UNICODE_CHARACTER_CLASS 0x100 which implies UNICODE_CASE), as passed at EncodingFactory.java:129
 * The JDK itself is not under /root/reference; Unicode classes are pinned to the generated 15.0 tables.
 */
#ifndef JO_REGEX_H
#define JO_REGEX_H
#include <stdint.h>

#define JO_RE_CASE_INSENSITIVE 0x02
#define JO_RE_UNICODE_CASE 0x40
#define JO_RE_UNICODE_CHARACTER_CLASS 0x100

typedef struct jo_regex jo_regex;

/* Compile `pattern` (UTF-8, NUL terminated).  Returns NULL and fills err on unsupported syntax. */
jo_regex *jo_regex_compile(const char *pattern, int flags, char *err, int errlen);
void jo_regex_free(jo_regex *re);

/* Try to find the next match in cps[0..n) searching from `from`.  Returns 1 and sets *ms,*me
 * (code point indices) or 0. */
int jo_regex_search(const jo_regex *re, const uint32_t *cps, int64_t n, int64_t from, int64_t *ms, int64_t *me);

/* Unicode class predicates (Unicode 15.0 tables) shared with the rest of the oracle. */
int jo_uc_is_letter(uint32_t cp);
int jo_uc_is_number(uint32_t cp);
int jo_uc_is_space(uint32_t cp);

#endif
