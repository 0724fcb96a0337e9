#!/usr/bin/env python3
"""Adversarial documents for the general-pattern path (one 1 MiB match / one 1 MiB gap per document), for ncu launch lists:
tools/general_long_one.py [match|gap]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jtokkit_b200 as jt

kind = sys.argv[1] if len(sys.argv) > 1 else "match"
p = jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE)
enc = jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("probe_long", jt.Pattern.compile(r"[a-z]+|\d{1,3}"), p.encoder, p.special_tokens_encoder))
doc = (b"a" if kind == "match" else b"!") * (1 << 20)
blob, off = jt.pack_documents([doc] * 8)
for _ in range(2):
    t0 = time.perf_counter()
    res = enc.encode_packed(blob, off, ordinary=True)
    print("%s: %.1f ms host-to-host, %d tokens" % (kind, (time.perf_counter() - t0) * 1e3, res.ids.size), flush=True)
    res.close()
