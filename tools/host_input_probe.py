import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth
enc = jt.EncodingFactory.cl100k_base()
for size in (64, 256):
    data, off = synth.config3_multilingual(torch.device("cpu"), total=size << 20, seed=11)
    d, o = data.numpy(), off.numpy()
    dp = data.pin_memory().numpy()
    for label, arr in (("pageable", d), ("pinned", dp)):
        for copy in (True, False):
            ts = []
            for _ in range(4):
                t0 = time.perf_counter()
                res = enc.encode_packed(arr, o, copy=copy)
                ts.append(time.perf_counter() - t0)
                dev_ms = res.device_ms
                if not copy: res.close()
            print("%4d MiB %-8s copy=%-5s best %.1f ms (%.2f GB/s), device %.1f ms, all %s" % (size, label, copy, min(ts[1:]) * 1e3, arr.size / min(ts[1:]) / 1e9, dev_ms, ["%.0f" % (t * 1e3) for t in ts]), flush=True)
