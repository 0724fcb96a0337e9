#!/usr/bin/env python3
"""Small workload that touches every kernel once (a quick end-to-end check after kernel changes): predefined and general patterns, long pieces,
special-token encoding, count-only, decode, maxTokens."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth

enc = jt.EncodingFactory.cl100k_base()
data, off = synth.config3_multilingual(torch.device("cpu"), total=3 << 20, seed=5)
res = enc.encode_packed(data.numpy(), off.numpy())
n1 = res.ids.size
docs = ["a" * 5000, "中文" * 3000, " " * 4000 + "x", "hello <|endoftext|> world", "", "1234567890" * 500, "ab" * 2500]
r2 = enc.encode_ordinary_batch(docs)
r3 = enc.encode_with_special_tokens_batch(docs)
cnt = enc.count_tokens_batch(["hello world", "x" * 3000])
dec = enc.decode_bytes_batch([r2.tokens(0), r2.tokens(1)])
mt = enc.encode("hello wonderful world of tokens", 3)
p = jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE)
g = jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("san", jt.Pattern.compile(r"\w+|\s+"), p.encoder, p.special_tokens_encoder))
r4 = g.encode_ordinary_batch(docs + [bytes(data.numpy()[:200000]).decode("utf-8", "ignore")])
r50 = jt.EncodingFactory.r50k_base()
r5 = r50.encode_ordinary_batch(docs)
print("ok", n1, r2.ids.size, r3.ids.size, list(cnt), len(dec[0]), mt.get_tokens(), r4.ids.size, r5.ids.size)
