#!/usr/bin/env python3
"""Per-language / per-config device-resident throughput of the tile kernel (development probe, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth

def run(enc, data, doc_off, label, steps=3, count_only=False):
    dev = data.device
    n = data.numel()
    pad = (-n) % 16
    d_in = torch.zeros(n + pad + 64, dtype=torch.uint8, device=dev); d_in[:n] = data
    d_ids = torch.empty(n + 16, dtype=torch.int32, device=dev)
    d_tok = torch.empty(doc_off.numel(), dtype=torch.int64, device=dev)
    d_st = torch.zeros(doc_off.numel(), dtype=torch.int32, device=dev)
    ms, tot = [], []
    for i in range(steps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ntok, nlong, nl, kms = enc.encode_device(d_in[:n], doc_off, d_ids, d_tok, d_st, time_kernel=True, count_only=count_only)
        e1.record(); torch.cuda.synchronize()
        if i: ms.append(kms); tot.append(e0.elapsed_time(e1))
    kms, t = min(ms), min(tot)
    print("%-28s %7.1f MB %7d docs %9d tok %5.2f B/tok long %4d | split+lookup %7.3f ms %6.1f GB/s | whole call %7.3f ms %6.1f GB/s %7.1f Mtok/s" %
          (label, n / 1e6, doc_off.numel() - 1, ntok, n / max(ntok, 1), nlong, kms, n / kms / 1e6, t, n / t / 1e6, ntok / t / 1e3), flush=True)

def main():
    dev = torch.device("cuda", 0)
    enc = jt.EncodingFactory.cl100k_base()
    size = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 128 << 20
    for lang in ["english", "latin", "cyrillic", "cjk", "semitic", "indic"]:
        d, off = synth.generate(size, 77, dev, mix=[(lang, 1.0)])
        run(enc, d, off, lang)
    d, off = synth.config3_multilingual(dev, total=size); run(enc, d, off, "multilingual mix")
    d, off = synth.config4_chat(dev, total=size); run(enc, d, off, "chat 256B docs"); run(enc, d, off, "chat 256B docs count-only", count_only=True)
    r50 = jt.EncodingFactory.r50k_base()
    d, off = synth.config2_english_64mib(dev, total=size); run(r50, d, off, "r50k english 64K docs")

if __name__ == "__main__":
    main()
