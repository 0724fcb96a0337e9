#!/usr/bin/env python3
"""One configuration through the device-resident path (for ncu launch lists): tools/gpu_one.py <lang|mix|chat> <MiB> [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth
from tools.gpu_probe import run

kind, size = sys.argv[1], int(sys.argv[2]) << 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
enc = jt.EncodingFactory.cl100k_base()
if kind == "mix":
    d, off = synth.config3_multilingual(dev, total=size)
elif kind == "chat":
    d, off = synth.config4_chat(dev, total=size)
else:
    d, off = synth.generate(size, 77, dev, mix=[(kind, 1.0)])
run(enc, d, off, kind, steps=steps)
