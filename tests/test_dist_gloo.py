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


def test_chunk_plan_of_the_in_library_multi_gpu_batch():
    """jtk_plan_chunks (pure host function of the C ABI): the plan jtk_encode_batch runs for G devices - contiguous, complete,
    whole documents, chunk c on device c % G, devices byte balanced, a single device reduces to one ramped chunk list."""
    from jtokkit_b200 import sharding
    rng = np.random.default_rng(3)
    lens = np.exp(rng.uniform(np.log(1024), np.log(65536), size=70000)).astype(np.int64)  # ~1 GiB of 1-64 KiB documents
    off = np.zeros(lens.size + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    total = int(off[-1])
    for G in (1, 2, 4, 8):
        cuts = sharding.plan_chunks(off, G)
        assert cuts[0] == 0 and cuts[-1] == lens.size and np.all(np.diff(cuts) > 0)
        per_dev = np.zeros(G, dtype=np.int64)
        sizes = off[cuts[1:]] - off[cuts[:-1]]
        for c, sz in enumerate(sizes):
            per_dev[sharding.device_of_chunk(c, G)] += sz
        assert per_dev.sum() == total
        assert per_dev.max() - per_dev.min() <= 0.03 * total / G + (8 << 20), (G, per_dev)
        assert sizes.max() <= (64 << 20) + 65536 and sizes[0] <= (8 << 20) + 65536  # ramp from 8 MiB up to the 64 MiB default
    # small chunk size, ragged documents, empty documents, an oversized document, an empty batch
    off2 = np.array([0, 0, 10, 10, 5_000_000, 5_000_001, 5_000_001], dtype=np.int64)
    cuts = sharding.plan_chunks(off2, 3, chunk_bytes=1 << 20)
    assert cuts[0] == 0 and cuts[-1] == 6 and np.all(np.diff(cuts) > 0)
    assert sharding.plan_chunks(np.array([0], dtype=np.int64), 4).tolist() == [0]
    import pytest
    with pytest.raises(ValueError):
        sharding.plan_chunks(np.array([0, 5, 3], dtype=np.int64), 2)
