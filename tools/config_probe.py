#!/usr/bin/env python3
"""BASELINE.json configs[3] (countTokens on 10 M chat-length strings) and configs[4] (adversarial 1 MiB single-piece documents, eight
per class PER GPU) at full size through ONE handle that spans --gpus devices (jtk_encode_batch: in-library chunk sharding), host
buffers in and out, every result compared with the oracle (counts of all strings; ids of every adversarial document from the exact
heap merge).  Development / evidence probe, not the bench:  python tools/config_probe.py --gpus 8 [--chat-mib 2560]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import jtokkit_b200 as jt
from jtokkit_b200 import synth
from oracle import jo

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--chat-mib", type=int, default=2560)
args = ap.parse_args()
G = args.gpus
enc = jt.Encoding(jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE), devices=list(range(G)))
orc = jo.OracleEncoding.builtin("cl100k_base")
threads = os.cpu_count() or 1

# ---- configs[3]: countTokens only
d, off = synth.config4_chat(torch.device("cuda", 0), total=args.chat_mib << 20)
h = torch.empty(d.numel(), dtype=torch.uint8).pin_memory()
h.copy_(d)
hn, on = h.numpy(), off.cpu().numpy()
del d
print("configs[3] countTokens: %d strings, %.2f GB, %d GPU(s)" % (on.size - 1, hn.size / 1e9, G), flush=True)
times = []
for _ in range(4):
    t0 = time.perf_counter()
    res = enc.encode_packed(hn, on, count_only=True)
    times.append(time.perf_counter() - t0)
    counts = res.counts().copy()
    kernel_ms = res.device_ms
    res.close()
dt = min(times[1:])
t0 = time.perf_counter()
_, exp, _ = orc.encode_batch(hn, on, threads, check_special=True)
t_or = time.perf_counter() - t0
ok = bool(np.array_equal(counts, exp))
print("  host to host %.1f ms = %.1f GB/s input, %.2f G tokens/s (kernel time max over devices %.1f ms); oracle port on %d cores %.1f s; every count identical: %s"
      % (dt * 1e3, hn.size / dt / 1e9, counts.sum() / dt / 1e9, kernel_ms, threads, t_or, ok), flush=True)
assert ok

# ---- configs[4]: adversarial single-piece documents, eight per class per GPU
names = ["a x 2^20", "random [a-z]", "spaces", "'!' x 2^20", "'ab' x 2^19", "random CJK", "newlines", "digits"]
for name, doc in zip(names, synth.config5_adversarial(n=1 << 20)):
    blob, o = jt.pack_documents([doc] * (8 * G))
    pin = torch.empty(blob.size, dtype=torch.uint8).pin_memory()
    pin.numpy()[:] = blob
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        res = enc.encode_packed(pin.numpy(), o, ordinary=True)
        times.append(time.perf_counter() - t0)
    exp = orc.encode_ordinary(doc, jo.MERGE_HEAP)
    ok = all(res.tokens(i) == exp for i in range(8 * G))
    print("configs[4] %-14s %3d x 1 MiB on %d GPU(s): %8.1f ms host to host, %8d tokens per document, ids identical to the oracle's heap merge: %s"
          % (name, 8 * G, G, min(times[1:]) * 1e3, len(exp), ok), flush=True)
    assert ok
