#!/usr/bin/env python3
"""PCIe copy rates of the box (pinned host memory): H2D alone, D2H alone, both at once.  Context for bench.py's e2e number."""
import sys, torch, time
if len(sys.argv) > 1:
    torch.cuda.set_device(int(sys.argv[1]))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
t = timed(h2d); print("H2D alone   %.1f GB/s" % (n / t / 1e9))
t = timed(d2h); print("D2H alone   %.1f GB/s" % (n / t / 1e9))
t = timed(both); print("both        %.1f GB/s each direction (%.1f ms per GiB pair)" % (n / t / 1e9, t * 1e3))
for chunk in (16, 64):
    c = chunk << 20
    def chunked():
        for o in range(0, n, c):
            with torch.cuda.stream(s2): h_out[o:o + c].copy_(d_b[o:o + c], non_blocking=True)
    t = timed(chunked); print("D2H in %d MiB chunks %.1f GB/s" % (chunk, n / t / 1e9))
