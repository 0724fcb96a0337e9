"""The N > 1 plumbing on CPU: world_size-2 gloo run of the benchmark's reduction (max of times, sum of work) and the
byte-balanced document sharding that jtk_encode_batch applies across an encoding's devices."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import bench
    from jtokkit_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank owns a byte-balanced contiguous range of documents of one shared batch
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 5000, size=1000)
    doc_off = np.zeros(1001, dtype=np.int64)
    doc_off[1:] = np.cumsum(lens)
    cuts = sharding.byte_balanced_cuts(doc_off, world)
    d0, d1 = cuts[rank], cuts[rank + 1]
    my_bytes = float(doc_off[d1] - doc_off[d0])
    # pretend the rank took (rank + 1) ms and produced one token per four bytes
    stats = torch.tensor([float(rank + 1), float(rank + 2), my_bytes / 4, my_bytes, 5.0, float(d1 - d0), 0.5], dtype=torch.float64)
    mx, sm = bench.reduce_over_ranks(stats, world)
    out[rank] = (mx.tolist(), sm.tolist(), int(d0), int(d1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduction_and_sharding():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (mx0, sm0, a0, b0), (mx1, sm1, a1, b1) = out[0], out[1]
    assert mx0 == mx1 and sm0 == sm1                      # every rank sees the same reduced statistics
    assert mx0[0] == 2.0 and mx0[1] == 3.0                # time = max over ranks
    assert a0 == 0 and b0 == a1 and b1 == 1000            # contiguous, disjoint, complete document ranges
    total = sm0[3]
    assert abs(total - float(np.sum(np.random.default_rng(0).integers(0, 5000, size=1000)))) < 1e-6  # work = sum over ranks
    assert sm0[5] == 1000.0


def test_byte_balanced_cuts_edge_cases():
    from jtokkit_b200 import sharding
    off = np.array([0, 10, 10, 10, 1000, 1001], dtype=np.int64)
    for g in (1, 2, 3, 8):
        cuts = sharding.byte_balanced_cuts(off, g)
        assert cuts[0] == 0 and cuts[-1] == 5 and all(cuts[i] <= cuts[i + 1] for i in range(g))
    assert sharding.byte_balanced_cuts(np.array([0], dtype=np.int64), 4).tolist() == [0, 0, 0, 0, 0]  # empty batch
    even = np.arange(0, 801, 100, dtype=np.int64)
    assert sharding.byte_balanced_cuts(even, 4).tolist() == [0, 2, 4, 6, 8]
