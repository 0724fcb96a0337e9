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
    ms = []
    for i in range(steps + 1):
        ntok, nlong, nl, kms = enc.encode_device(d_in[:n], doc_off, d_ids, d_tok, d_st, time_kernel=True, count_only=count_only)
        if i: ms.append(kms)
    kms = min(ms)
    print("%-28s %8.1f MB %7d docs  %9d tok  %6.2f B/tok  long %5d  kernel %8.3f ms  %7.2f GB/s in  %7.1f Mtok/s" %
          (label, n / 1e6, doc_off.numel() - 1, ntok, n / max(ntok, 1), nlong, kms, n / kms / 1e6, ntok / kms / 1e3), flush=True)

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
