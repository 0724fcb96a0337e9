#!/usr/bin/env python3
"""Throughput of the general split-pattern path - as a DFA over code-point classes (jtk_dfa.cpp) and as the backtracking program
(JTK_RX_DFA=0), both under the sliced find() passes - next to the rule-based path of the predefined pattern, same vocabulary and
text (development probe)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth

p = jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE)
size = (int(sys.argv[1]) if len(sys.argv) > 1 else 64) << 20
data, off = synth.config3_multilingual(torch.device("cpu"), total=size, seed=11)
d, o = data.numpy(), off.numpy()
variants = [("predefined cl100k pattern (class tables + bit-parallel rules)", p.get_pattern().pattern(), 0x100, "1"),
            ("same pattern with CASE_INSENSITIVE, DFA", p.get_pattern().pattern(), 0x102, "1"),
            ("same pattern with CASE_INSENSITIVE, backtracking program", p.get_pattern().pattern(), 0x102, "0"),
            (r"general: \w+|\s+|[^\w\s]+, DFA", r"\w+|\s+|[^\w\s]+", 0, "1"),
            (r"general: \w+|\s+|[^\w\s]+, backtracking program", r"\w+|\s+|[^\w\s]+", 0, "0"),
            (r"general with '$' (no DFA form): \w+$|\w+|\s+|[^\w\s]+", r"\w+$|\w+|\s+|[^\w\s]+", 0, "1")]
from tools.gpu_probe import run
dev = torch.device("cuda", 0)
d_dev, o_dev = data.to(dev), off.to(dev)
ref_ids = {}
for label, pat, flags, dfa in variants:
    os.environ["JTK_RX_DFA"] = dfa
    enc = jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("probe", jt.Pattern.compile(pat, flags), p.encoder, p.special_tokens_encoder))
    res = enc.encode_packed(d, o, ordinary=True)
    same = ref_ids.setdefault((pat, flags), res.ids.copy())
    print("%s: %d tokens, flagged documents %d, ids identical to the first variant of this pattern: %s" %
          (label, res.ids.size, int((res.doc_status != 0).sum()), same.size == res.ids.size and bool((same == res.ids).all())), flush=True)
    res.close()
    run(enc, d_dev, o_dev, "  device-resident", steps=3)

# adversarial for the sliced matcher: one match (or one gap) per 1 MiB document
os.environ["JTK_RX_DFA"] = "1"
enc = jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("probe_long", jt.Pattern.compile(r"[a-z]+|\d{1,3}"), p.encoder, p.special_tokens_encoder))
for name, doc in (("one 1 MiB match per document", b"a" * (1 << 20)), ("one 1 MiB gap per document", b"!" * (1 << 20))):
    blob, o2 = jt.pack_documents([doc] * 8)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        res = enc.encode_packed(blob, o2, ordinary=True)
        ts.append(time.perf_counter() - t0)
        n = res.ids.size
        res.close()
    print("%-40s 8 x 1 MiB: %.1f ms host-to-host, %d tokens" % (name, min(ts[1:]) * 1e3, n), flush=True)
