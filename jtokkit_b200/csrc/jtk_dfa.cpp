/* Split-pattern DFA: the compiled program of jtk_regex.cpp (a Pike-style instruction list) determinised at registration
 * into one transition table over code-point classes, with java.util.regex's leftmost-FIRST semantics:
 *
 *  - a DFA state is the ORDERED list of program positions the backtracking matcher would try, most preferred first
 *    (ordered alternation, greedy before lazy); a position reached twice keeps its first (preferred) occurrence;
 *  - when a thread reaches MATCH every less preferred thread is dropped: whatever the remaining (more preferred)
 *    threads match later replaces it, so the LAST match position seen before the state dies is the answer
 *    (Matcher.find() at a fixed start position);
 *  - a one-character look-ahead (?=[set]) / (?![set]) is a thread that waits for the next character's class (or the end of
 *    the text) without consuming it; a possessive repeat of one class X*+ / X++ / X?+ / X{n,m}+ is the greedy repeat whose
 *    exits carry (?!X), because the longest run is the only one a possessive quantifier ever tries;
 *  - '^' needs one extra start state (search begins at the document start or not).
 *
 * Outside this form ('$', \b, look-ahead over more than one character, more than 64 distinct sets, table too large) the
 * pattern keeps the backtracking program on the device (jtk_regex.h); both run under the same sliced find() passes.
 * Replaces: java.util.regex.Pattern.compile at EncodingFactory.java:129 for custom GptBytePairEncodingParams patterns. */
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "jtk_regex_compile.h"

namespace {

enum { D_CHAR, D_EPS, D_EPS2, D_LOOK1, D_BOL, D_EOI /* \\z: passes at the end of the text only */, D_MATCH };

struct dnode {
	int type, set, neg, a, b;
};

struct builder {
	const jtk_rx_compiled &prog;
	std::vector<dnode> nodes;
	std::vector<int> node_of_pc;
	std::string why;

	explicit builder(const jtk_rx_compiled &p) : prog(p) {}

	int add(int type, int set = -1, int neg = 0, int a = -1, int b = -1) {
		nodes.push_back({type, set, neg, a, b});
		return (int) nodes.size() - 1;
	}

	/* program -> epsilon-NFA: one entry node per instruction, REP expanded into copies */
	bool build_nfa() {
		const auto &code = prog.inst;
		const int n = (int) code.size();
		node_of_pc.assign((size_t) n + 1, -1);
		for (int pc = 0; pc < n; pc++) node_of_pc[(size_t) pc] = add(D_EPS); /* placeholder: rewritten below */
		for (int pc = 0; pc < n; pc++) {
			const jtk_rx_inst &in = code[(size_t) pc];
			const int self = node_of_pc[(size_t) pc];
			const int next = pc + 1 < n ? node_of_pc[(size_t) pc + 1] : -1;
			switch (in.op) {
			case JTK_RX_SET: nodes[(size_t) self] = {D_CHAR, in.a, 0, next, -1}; break;
			case JTK_RX_SPLIT: nodes[(size_t) self] = {D_EPS2, -1, 0, node_of_pc[(size_t) in.a], node_of_pc[(size_t) in.b]}; break;
			case JTK_RX_JMP: nodes[(size_t) self] = {D_EPS, -1, 0, node_of_pc[(size_t) in.a], -1}; break;
			case JTK_RX_BOL: nodes[(size_t) self] = {D_BOL, -1, 0, next, -1}; break;
			case JTK_RX_EOL:
				if (in.a == 0) { /* '$' / \\Z look two characters ahead (\\r\\n before the end) */
					why = "'$'";
					return false;
				}
				nodes[(size_t) self] = {D_EOI, -1, 0, next, -1};
				break;
			case JTK_RX_MATCH: nodes[(size_t) self] = {D_MATCH, -1, 0, -1, -1}; break;
			case JTK_RX_LOOK: {
				/* the sub-program must test exactly one character: SET MATCH, or a repeat that needs exactly one */
				const jtk_rx_inst &s0 = code[(size_t) in.b];
				const bool one = in.b + 1 < n && code[(size_t) in.b + 1].op == JTK_RX_MATCH &&
				                 (s0.op == JTK_RX_SET || (s0.op == JTK_RX_REP && s0.b == 1));
				if (!one) {
					why = "look-ahead over more than one character";
					return false;
				}
				nodes[(size_t) self] = {D_LOOK1, s0.a, in.a, next, -1};
				break;
			}
			case JTK_RX_REP: {
				const int set = in.a, mn = in.b, mx = in.c, mode = in.d;
				if (mn > 256 || mx > 256) {
					why = "counted repeat beyond 256";
					return false;
				}
				/* exit of the repeat before its maximum is reached: possessive repeats only leave when no further character fits */
				auto exit_node = [&]() { return mode == 2 ? add(D_LOOK1, set, 1, next) : next; };
				/* build back to front: tail = what follows the optional part */
				int tail;
				if (mx < 0) { /* L: SPLIT(CHAR -> L, exit) */
					const int l = add(D_EPS2);
					const int body = add(D_CHAR, set, 0, l);
					const int ex = exit_node();
					nodes[(size_t) l].a = mode == 1 ? ex : body;
					nodes[(size_t) l].b = mode == 1 ? body : ex;
					tail = l;
				} else {
					tail = next; /* after the last optional copy the exit is unconditional */
					for (int k = mx - 1; k >= mn; k--) {
						const int body = add(D_CHAR, set, 0, tail);
						const int ex = exit_node();
						tail = mode == 1 ? add(D_EPS2, -1, 0, ex, body) : add(D_EPS2, -1, 0, body, ex);
					}
				}
				for (int k = 0; k < mn; k++) tail = add(D_CHAR, set, 0, tail);
				nodes[(size_t) self] = {D_EPS, -1, 0, tail, -1};
				break;
			}
			default: why = in.op == JTK_RX_WORDB ? "\\b" : in.op == JTK_RX_LOOKB ? "look-behind" : "unknown instruction"; return false;
			}
		}
		return true;
	}
};

struct dfa_state {
	std::vector<int> list;
	bool bol;
	bool operator<(const dfa_state &o) const { return bol != o.bol ? bol < o.bol : list < o.list; }
};

} /* namespace */

bool jtk_rx_build_dfa(const jtk_rx_compiled &prog, const jtk_tables &view, jtk_rx_dfa_host *out, std::string *why) {
	builder B(prog);
	if (!B.build_nfa()) {
		*why = B.why;
		return false;
	}
	const std::vector<dnode> &N = B.nodes;
	/* ---- code point classes: two code points are equivalent when every set the automaton reads agrees on them ---- */
	std::vector<int> used_sets;
	std::vector<int> set_slot(prog.sets.size(), -1);
	for (const dnode &d : N)
		if ((d.type == D_CHAR || d.type == D_LOOK1) && set_slot[(size_t) d.set] < 0) {
			set_slot[(size_t) d.set] = (int) used_sets.size();
			used_sets.push_back(d.set);
		}
	if (used_sets.size() > 64) {
		*why = "more than 64 distinct character sets";
		return false;
	}
	jtk_rx_program P;
	memset(&P, 0, sizeof(P));
	P.inst = prog.inst.data();
	P.ninst = (int32_t) prog.inst.size();
	P.sets = prog.sets.data();
	P.ranges = prog.ranges.data();
	constexpr uint32_t CP_END = 0x110000u; /* one more signature stands for everything a four-byte sequence can encode beyond U+10FFFF */
	std::unordered_map<uint64_t, int> class_of_sig;
	std::vector<uint64_t> sig_of_class;
	std::vector<uint8_t> cls((size_t) CP_END + 1);
	for (uint32_t cp = 0; cp <= CP_END; cp++) {
		uint64_t sig = 0;
		for (size_t k = 0; k < used_sets.size(); k++)
			if (jtk_rx_in_set(P, view, used_sets[k], cp)) sig |= 1ull << k;
		auto it = class_of_sig.find(sig);
		if (it == class_of_sig.end()) {
			if (sig_of_class.size() >= 255) {
				*why = "more than 255 code point classes";
				return false;
			}
			it = class_of_sig.emplace(sig, (int) sig_of_class.size()).first;
			sig_of_class.push_back(sig);
		}
		cls[cp] = (uint8_t) it->second;
	}
	const int nclass = (int) sig_of_class.size(), nsym = nclass + 1, eof = nclass;
	auto in_class = [&](int sym, int set) { return sym != eof && ((sig_of_class[(size_t) sym] >> set_slot[(size_t) set]) & 1ull); };

	/* ---- subset construction over ordered thread lists ---- */
	std::map<dfa_state, int> id_of;
	std::vector<dfa_state> states;
	auto intern = [&](dfa_state &&s) {
		if (s.list.empty()) return 0;
		auto it = id_of.find(s);
		if (it != id_of.end()) return it->second;
		const int id = (int) states.size();
		id_of.emplace(s, id);
		states.push_back(std::move(s));
		return id;
	};
	states.push_back({{}, false}); /* 0: dead */
	std::vector<char> seen_post(N.size()), seen_pre(N.size());
	struct ctx {
		std::vector<int> out;
		bool cut = false, mb = false;
	};
	/* leaves (CHAR / LOOK1 / MATCH) reachable from node v without consuming, in preference order, appended to c.out */
	auto closure = [&](int v0, bool bol, ctx &c, std::vector<char> &seen, auto &&leaf) {
		/* explicit DFS keeping preference order: push b then a so that a is expanded first */
		std::vector<int> st(1, v0);
		while (!st.empty() && !c.cut) {
			const int v = st.back();
			st.pop_back();
			if (v < 0 || seen[(size_t) v]) continue;
			seen[(size_t) v] = 1;
			const dnode &d = N[(size_t) v];
			switch (d.type) {
			case D_EPS: st.push_back(d.a); break;
			case D_EPS2:
				st.push_back(d.b);
				st.push_back(d.a);
				break;
			case D_BOL:
				if (bol) st.push_back(d.a);
				break;
			default: leaf(v);
			}
		}
	};
	auto start_state = [&](bool bol) {
		ctx c;
		std::fill(seen_post.begin(), seen_post.end(), 0);
		closure(B.node_of_pc[0], bol, c, seen_post, [&](int v) {
			c.out.push_back(v);
			if (N[(size_t) v].type == D_MATCH) c.cut = true;
		});
		return intern({std::move(c.out), bol});
	};
	out->start = start_state(false);
	out->start_bol = start_state(true);
	constexpr int MAX_STATES = 4096;
	std::vector<uint16_t> trans;
	for (size_t si = 0; si < states.size(); si++) {
		if (states.size() > (size_t) MAX_STATES) {
			*why = "more than 4096 DFA states";
			return false;
		}
		trans.resize((si + 1) * (size_t) nsym, 0);
		if (si == 0) continue;
		for (int sym = 0; sym < nsym; sym++) {
			const dfa_state cur = states[si]; /* (copy: intern may reallocate) */
			ctx c;
			std::fill(seen_post.begin(), seen_post.end(), 0);
			std::fill(seen_pre.begin(), seen_pre.end(), 0);
			auto post_leaf = [&](int v) {
				c.out.push_back(v);
				if (N[(size_t) v].type == D_MATCH) c.cut = true;
			};
			/* a thread at the position BEFORE the character `sym` */
			std::vector<int> pre(cur.list.rbegin(), cur.list.rend()); /* work stack, most preferred on top */
			while (!pre.empty() && !c.cut) {
				const int v = pre.back();
				pre.pop_back();
				const dnode &d = N[(size_t) v];
				if (d.type == D_CHAR || d.type == D_LOOK1 || d.type == D_EOI || d.type == D_MATCH) {
					if (seen_pre[(size_t) v] == 2) continue;
					seen_pre[(size_t) v] = 2;
				}
				if (d.type == D_MATCH) {
					c.mb = true;
					c.cut = true;
				} else if (d.type == D_CHAR) {
					if (in_class(sym, d.set)) closure(d.a, false, c, seen_post, post_leaf);
				} else if (d.type == D_LOOK1 || d.type == D_EOI) {
					if (d.type == D_EOI ? sym == eof : in_class(sym, d.set) != (d.neg != 0)) {
						/* goes on at the same position: its leaves are tried right here, before the less preferred threads */
						std::vector<int> leaves;
						ctx tmp;
						std::vector<char> seen_eps(N.size(), 0);
						closure(d.a, cur.bol, tmp, seen_eps, [&](int u) { leaves.push_back(u); });
						for (auto it = leaves.rbegin(); it != leaves.rend(); ++it) pre.push_back(*it);
					}
				}
			}
			const bool mb = c.mb;
			const int to = intern({std::move(c.out), false});
			trans[si * (size_t) nsym + (size_t) sym] = (uint16_t) (to | (mb ? 0x8000 : 0));
		}
	}
	if (states.size() > (size_t) MAX_STATES || states.size() * (size_t) nsym > (1u << 19)) {
		*why = "DFA table too large";
		return false;
	}
	/* ---- the state that holds nothing but MATCH moves to the end: reaching it ends the run without another read ---- */
	int nstates = (int) states.size();
	int acc = -1;
	for (int s = 1; s < nstates; s++)
		if (!states[(size_t) s].bol && states[(size_t) s].list.size() == 1 && N[(size_t) states[(size_t) s].list[0]].type == D_MATCH) acc = s;
	std::vector<int> perm((size_t) nstates);
	for (int s = 0; s < nstates; s++) perm[(size_t) s] = s;
	if (acc >= 0 && acc != nstates - 1) std::swap(perm[(size_t) acc], perm[(size_t) nstates - 1]);
	out->trans.assign((size_t) nstates * (size_t) nsym, 0);
	for (int s = 0; s < nstates; s++)
		for (int sym = 0; sym < nsym; sym++) {
			const uint16_t t = trans[(size_t) s * (size_t) nsym + (size_t) sym];
			out->trans[(size_t) perm[(size_t) s] * (size_t) nsym + (size_t) sym] = (uint16_t) (perm[(size_t) (t & 0x7FFF)] | (t & 0x8000));
		}
	out->start = perm[(size_t) out->start];
	out->start_bol = perm[(size_t) out->start_bol];
	out->acc_lo = acc >= 0 ? nstates - 1 : nstates;
	out->nstates = nstates;
	out->nsym = nsym;
	/* ---- runs: what an ASCII byte does in a state, for the sequential passes that step over long runs four bytes at a time ---- */
	out->stay.clear();
	if (nstates <= 256) {
		out->stay.assign((size_t) nstates * 128, 0);
		for (int s = 1; s < nstates; s++)
			for (int b = 0; b < 128; b++) {
				const uint16_t t = out->trans[(size_t) s * (size_t) nsym + cls[(size_t) b]];
				const int ns = t & 0x7FFF, mb = t >> 15;
				uint8_t code = 0;
				if (ns == s && s < out->acc_lo) code = mb ? 2 : 1;
				else if (s == out->start && ns == 0 && !mb) code = 3;
				out->stay[(size_t) s * 128 + (size_t) b] = code;
			}
	}
	/* ---- class lookup: ASCII directly, the rest through a two-level table with shared blocks ---- */
	out->stage1.assign(0x200000 >> 8, 0);
	out->stage2.clear();
	std::map<std::string, int> block_of;
	for (uint32_t hi = 0; hi < (0x200000u >> 8); hi++) {
		std::string blk(256, (char) cls[CP_END]);
		if (hi < (CP_END >> 8)) blk.assign(reinterpret_cast<const char *>(cls.data()) + ((size_t) hi << 8), 256);
		auto it = block_of.find(blk);
		if (it == block_of.end()) {
			it = block_of.emplace(blk, (int) (out->stage2.size() >> 8)).first;
			out->stage2.insert(out->stage2.end(), blk.begin(), blk.end());
		}
		out->stage1[hi] = (uint16_t) it->second;
	}
	return true;
}
