#!/usr/bin/env python3
"""Device-resident decode throughput (jtk_decode_batch_device) on the ids of the multilingual corpus: tools/decode_probe.py [MiB]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth

size = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
dev = torch.device("cuda", 0)
enc = jt.EncodingFactory.cl100k_base()
data, off = synth.config3_multilingual(dev, total=size)
n, nd = data.numel(), off.numel() - 1
d_in = torch.zeros(n + 80, dtype=torch.uint8, device=dev)
d_in[:n] = data
d_ids = torch.empty(n, dtype=torch.int32, device=dev)
d_tok = torch.empty(nd + 1, dtype=torch.int64, device=dev)
d_st = torch.zeros(nd + 1, dtype=torch.int32, device=dev)
ntok, _, _, _ = enc.encode_device(d_in[:n], off, d_ids, d_tok, d_st)
d_out = torch.empty(n + 64, dtype=torch.uint8, device=dev)
d_boff = torch.empty(nd + 1, dtype=torch.int64, device=dev)
d_dst = torch.zeros(nd + 1, dtype=torch.int32, device=dev)
d_bad = torch.empty(nd + 1, dtype=torch.int32, device=dev)
ms = []
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    total, nl = enc.decode_device(d_ids[:ntok], d_tok, d_out[:n], d_boff, d_dst, d_bad)
    e1.record()
    torch.cuda.synchronize()
    if i:
        ms.append(e0.elapsed_time(e1))
ok = total == n and bool(torch.equal(d_out[:n], data)) and bool(torch.equal(d_boff, off))
algo = 4 * ntok + n + 16 * (nd + 1)
t = min(ms)
print("decode %d tokens -> %d bytes: %.3f ms (best of 5, %d launches), %.0f GB/s of algorithmic bytes, %.1f G tokens/s, round trip %s" %
      (ntok, total, t, nl, algo / t / 1e6, ntok / t / 1e6, "ok" if ok else "FAILED"))
