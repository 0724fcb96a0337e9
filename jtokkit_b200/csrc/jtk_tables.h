/*
 * Host-side registration work: GptBytePairEncodingParams -> flat tables that the kernels read.
 * Replaces what the reference does at construction time:
 *   EncodingFactory.java:121-137  fromPredefinedParameters (Pattern.compile + loadMergeableRanks)
 *   EncodingFactory.java:139-164  loadMergeableRanks (.tiktoken parser)
 *   GptBytePairEncoding.java:30-35 / TokenEncoder.java:38-45  building the two hash maps
 * Pure C++ (no CUDA) so that the same tables can be built for the device upload and for host-side tests.
 */
#ifndef JTK_TABLES_H
#define JTK_TABLES_H

#include <string>
#include <vector>

#include "../../include/jtokkit_b200.h"
#include "jtk_common.h"

struct jtk_host_tables {
	std::string name;
	int32_t pattern_kind = 0;
	int32_t max_token_len = 0;
	std::vector<uint8_t> ascii_cls;
	std::vector<uint16_t> cp_stage1;
	std::vector<uint8_t> cp_stage2;
	std::vector<uint32_t> lut_sp;
	std::vector<uint8_t> cls2, bmp_nib;
	std::vector<jtk_slot_a> tab_a;
	uint32_t mask_a = 0;
	std::vector<jtk_slot> tab_b;
	uint32_t mask_b = 0;
	std::vector<uint32_t> long_filter;
	std::vector<uint8_t> tok_bytes;
	std::vector<uint32_t> tok_off;
	std::vector<int32_t> tok_rank; /* by token index (host only) */
	std::vector<int32_t> byte_id;
	std::vector<int32_t> bytepair;
	std::vector<jtk_slot> pair;
	uint32_t mask_p = 0;
	std::vector<uint32_t> bigram_bits, trigram_bits;
	int32_t nspecial = 0;
	int32_t special_has_empty = 0;
	std::vector<uint8_t> special_bytes;
	std::vector<uint32_t> special_off;
	std::vector<int32_t> special_ids;
	uint32_t special_first[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	std::vector<uint32_t> dec_keys;
	uint32_t mask_d = 0;
	std::vector<uint8_t> dec_bytes;
	std::vector<uint32_t> dec_off;
	std::vector<uint32_t> dec_direct; /* quadruples (offset << 8 | length, first twelve bytes) by id, empty when the ids are not small non-negative numbers */
	/* JTK_PAT_GENERAL: the compiled split program (raw jtk_rx_inst / jtk_rx_set arrays, 16-byte aligned by the vector) */
	std::vector<uint8_t> rx_inst, rx_sets;
	std::vector<uint32_t> rx_ranges;
	int32_t rx_ninst = 0;
	uint32_t rx_first[8] = {~0u, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u, ~0u};
	/* ... and its DFA (empty when the pattern has none; rx_dfa_why says what stood in the way) */
	std::vector<uint16_t> rx_dfa_trans, rx_dfa_stage1;
	std::vector<uint8_t> rx_dfa_stage2, rx_dfa_stay;
	int32_t rx_dfa_nsym = 0, rx_dfa_nstates = 0, rx_dfa_start = 0, rx_dfa_start_bol = 0, rx_dfa_acc_lo = 0;
	std::string rx_dfa_why;
	/* statistics (reported by DESIGN.md / tests) */
	int64_t n_tokens = 0, n_pairs = 0;
	int32_t max_probe_a = 0, max_probe_b = 0, max_probe_p = 0;
};

/* Returns JTK_OK or a JTK_E_* code with a message in *err. */
int jtk_build_host_tables(const jtk_params *params, jtk_host_tables *out, std::string *err);

/* A jtk_tables whose pointers alias the host vectors (for host-side tests of the __host__ __device__ code). */
jtk_tables jtk_host_view(const jtk_host_tables &h);

/* Parser for the reference's resource format (EncodingFactory.java:139-164). */
int jtk_load_tiktoken_file(const char *path, std::vector<uint8_t> *bytes, std::vector<int64_t> *off, std::vector<int32_t> *ranks, std::string *err);

/* Pattern strings + special tokens of the predefined encodings (EncodingFactory.java:18-53,63,77,91,105). */
struct jtk_builtin_def {
	const char *name;
	const char *pattern;
	int nspecial;
	const char *special[5];
	int32_t special_ids[5];
};
const jtk_builtin_def *jtk_find_builtin(const char *name);

/* Unicode version of the generated class tables (unicode_ranges.inc). */
const char *jtk_unicode_version();

#endif
