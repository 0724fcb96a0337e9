#!/usr/bin/env python3
"""The cl100k pattern string registered as a GENERAL pattern (CASE_INSENSITIVE) through the device-resident path, for ncu launch
lists of the jtk_general_* kernels: tools/general_one.py <MiB> [steps]   (JTK_RX_DFA=0: the backtracking program)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth
from tools.gpu_probe import run

size = (int(sys.argv[1]) if len(sys.argv) > 1 else 128) << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p = jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE)
enc = jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("general_one", jt.Pattern.compile(p.get_pattern().pattern(), 0x102), p.encoder, p.special_tokens_encoder))
dev = torch.device("cuda", 0)
d, off = synth.config3_multilingual(dev, total=size, seed=11)
run(enc, d, off, "cl100k string as a general pattern", steps=steps)
