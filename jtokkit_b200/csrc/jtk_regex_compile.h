/* Host-side compiler for general split patterns (see jtk_regex.h for the program format and the supported subset). */
#ifndef JTK_REGEX_COMPILE_H
#define JTK_REGEX_COMPILE_H

#include <string>
#include <vector>

#include "../../include/jtokkit_b200.h"
#include "jtk_regex.h"

struct jtk_rx_compiled {
	std::vector<jtk_rx_inst> inst;
	std::vector<jtk_rx_set> sets;
	std::vector<uint32_t> ranges;
	bool nullable = false; /* the pattern can match the empty string */
	uint32_t first[8];     /* bytes with which a match can begin (every byte when nullable): lets the search skip the others */
};

/* pattern: UTF-8 java.util.regex source; flags: Pattern flag bits.  Returns JTK_OK or JTK_E_PATTERN_UNSUPPORTED with *err. */
int jtk_rx_compile(const char *pattern, int flags, jtk_rx_compiled *out, std::string *err);

/* The program determinised (jtk_dfa.cpp): transition table over code-point classes, leftmost-first semantics. */
struct jtk_rx_dfa_host {
	std::vector<uint16_t> trans; /* nstates x nsym: next state (0 = dead) | 0x8000 when a match ends BEFORE the character read */
	int nstates = 0, nsym = 0;   /* symbols: the code point classes, then "end of text" */
	int start = 0, start_bol = 0; /* start state inside a document / at its first byte ('^') */
	int acc_lo = 0;              /* states >= acc_lo hold nothing but MATCH: the run ends there without another read */
	std::vector<uint16_t> stage1; /* (code point >> 8) -> block, 8192 entries (four-byte sequences reach 0x1FFFFF) */
	std::vector<uint8_t> stage2;  /* block * 256 + (code point & 255) -> class */
	std::vector<uint8_t> stay;    /* nstates x 128 (empty above 256 states): what an ASCII byte does in a state - 1: the state loops on it, 2: loops and a
	                               * match ends before it, 3 (row of `start` only): the attempt dies on it at once, 0: anything else */
};

/* view: a jtk_tables with cp_stage1 / cp_stage2 set (what jtk_rx_in_set reads).  False with *why when the program has no DFA form. */
bool jtk_rx_build_dfa(const jtk_rx_compiled &prog, const jtk_tables &view, jtk_rx_dfa_host *out, std::string *why);

#endif
