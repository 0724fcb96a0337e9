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

#endif
