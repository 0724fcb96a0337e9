#!/usr/bin/env python3
"""Aggregate host<->device bandwidth of the box with G GPUs busy at once (one process, one thread per GPU, pinned memory):
H2D alone, D2H alone, both directions.  The end-to-end numbers of bench.py are bound by these, not by the kernels.
Usage: python tools/pcie_multi_probe.py [G]"""
import sys, threading, time
import torch

G = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
N = 512 << 20
bufs = []
for g in range(G):
    torch.cuda.set_device(g)
    bufs.append((torch.empty(N, dtype=torch.uint8).pin_memory(), torch.empty(N, dtype=torch.uint8).pin_memory(),
                 torch.empty(N, dtype=torch.uint8, device="cuda:%d" % g), torch.empty(N, dtype=torch.uint8, device="cuda:%d" % g),
                 torch.cuda.Stream(device=g), torch.cuda.Stream(device=g)))


def work(g, mode, reps, barrier):
    hi, ho, di, do, s1, s2 = bufs[g]
    torch.cuda.set_device(g)
    barrier.wait()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                di.copy_(hi, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize(g)


for mode in ("h2d", "d2h", "both"):
    for reps in (2, 6):  # first round warms up
        barrier = threading.Barrier(G + 1)
        th = [threading.Thread(target=work, args=(g, mode, reps, barrier)) for g in range(G)]
        for t in th:
            t.start()
        barrier.wait()
        t0 = time.perf_counter()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
    moved = G * reps * N * (2 if mode == "both" else 1)
    print("%d GPU(s) %-5s aggregate %6.1f GB/s (%5.1f GB/s per GPU%s)" % (G, mode, moved / dt / 1e9, moved / dt / 1e9 / G, " summed over both directions" if mode == "both" else ""), flush=True)
