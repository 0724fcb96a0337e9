#!/usr/bin/env python3
"""BASELINE.json configs 3 (countTokens on 10 M chat-length strings) and 4 (adversarial 1 MiB single-piece documents) at full
size on one GPU, device-resident, with the size-independent checks the tests use (development probe, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth
from tools.gpu_probe import run

dev = torch.device("cuda", 0)
enc = jt.EncodingFactory.cl100k_base()
d, off = synth.config4_chat(dev, total=2560 << 20)
print("config 3 (chat strings): %d strings, %.2f GB" % (off.numel() - 1, d.numel() / 1e9), flush=True)
run(enc, d, off, "chat count-only", steps=3, count_only=True)
run(enc, d, off, "chat encode", steps=3)
docs = synth.config5_adversarial(n=1 << 20)
names = ["a x 2^20", "random [a-z]", "spaces", "'!' x 2^20", "'ab' x 2^19", "random CJK", "newlines", "digits"]
for name, doc in zip(names, docs):
    blob, o = jt.pack_documents([doc] * 8)
    times = []
    for _ in range(4):
        t0 = time.perf_counter()
        res = enc.encode_packed(blob, o, ordinary=True)
        times.append(time.perf_counter() - t0)
    dt = min(times[1:])
    first = res.tokens(0)
    ok = all(res.tokens(i) == first for i in range(1, 8)) and enc.decode_bytes(first) == doc
    print("config 4 %-14s 8 x 1 MiB: %8.1f ms host-to-host, %8d tokens per document, round trip %s" % (name, dt * 1e3, len(first), "ok" if ok else "FAILED"), flush=True)
