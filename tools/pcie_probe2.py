#!/usr/bin/env python3
"""Chunked concurrent H2D / D2H like the host pipeline: per-direction rates for a few chunk-size pairs."""
import torch, time
n_in, n_out = 1 << 30, int(1.5 * (1 << 30))
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(cin, cout, label):
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(s1):
        e[0].record()
        for o in range(0, n_in, cin): d_in[o:o + cin].copy_(h_in[o:o + cin], non_blocking=True)
        e[1].record()
    with torch.cuda.stream(s2):
        e[2].record()
        for o in range(0, n_out, cout): h_out[o:o + cout].copy_(d_out[o:o + cout], non_blocking=True)
        e[3].record()
    torch.cuda.synchronize()
    t_in, t_out = e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])
    print("%-34s H2D 1.00 GiB in %6.2f ms (%5.1f GB/s)   D2H 1.50 GiB in %6.2f ms (%5.1f GB/s)   both done after %.2f ms" %
          (label, t_in, n_in / t_in / 1e6, t_out, n_out / t_out / 1e6, max(e[0].elapsed_time(e[1]), e[0].elapsed_time(e[3]))))
for _ in range(2):
    run(n_in, n_out, "single copies")
run(64 << 20, 96 << 20, "64 MiB in / 96 MiB out chunks")
run(16 << 20, 24 << 20, "16 MiB in / 24 MiB out chunks")
run(4 << 20, 6 << 20, "4 MiB in / 6 MiB out chunks")
