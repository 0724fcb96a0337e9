"""Document sharding of one batch over the GPUs of a box.

jtk_encode_batch cuts a batch into byte-balanced contiguous document ranges ("chunks") and runs chunk c on device
c % ndev (jtk_capi.cu: plan_chunks / run_device); `plan_chunks` below returns exactly that plan through the C ABI
(jtk_plan_chunks, a pure host function).  `byte_balanced_cuts` is the one-range-per-shard rule for callers that drive one
process per GPU themselves (the rule of SURVEY.md section 8e: shard g of G gets the documents
[lower_bound(doc_off, g * total / G), lower_bound(doc_off, (g + 1) * total / G))).  Neither involves a collective:
documents encode independently (GptBytePairEncoding.java:71-103 keeps no state across calls)."""
import ctypes as C

import numpy as np


def plan_chunks(doc_off, ndev, chunk_bytes=0):
    """Chunk cuts (numpy int64, n + 1 entries) of jtk_encode_batch for `ndev` devices; chunk c runs on device c % ndev."""
    from . import _capi
    doc_off = np.ascontiguousarray(doc_off, dtype=np.int64)
    cap = doc_off.size + 1
    cuts = np.zeros(cap, dtype=np.int64)
    n = _capi.lib().jtk_plan_chunks(doc_off.ctypes.data_as(C.c_void_p), doc_off.size - 1, int(ndev), int(chunk_bytes), cuts.ctypes.data_as(C.c_void_p), cap)
    if n < 0:
        raise ValueError(_capi.last_error())
    return cuts[:n + 1].copy()


def device_of_chunk(c, ndev):
    return c % ndev


def byte_balanced_cuts(doc_off, nshards):
    doc_off = np.asarray(doc_off, dtype=np.int64)
    ndocs = doc_off.size - 1
    total = int(doc_off[-1])
    cuts = np.zeros(nshards + 1, dtype=np.int64)
    for g in range(1, nshards):
        target = total // nshards * g
        cuts[g] = min(ndocs, max(int(cuts[g - 1]), int(np.searchsorted(doc_off, target, side="left"))))
    cuts[nshards] = ndocs
    return cuts


def concat_shards(results):
    """results: list of (ids, token_offsets) per shard in document order -> (ids, token_offsets) of the whole batch."""
    ids = np.concatenate([r[0] for r in results]) if results else np.zeros(0, dtype=np.int32)
    offs = [np.zeros(1, dtype=np.int64)]
    base = 0
    for _, t in results:
        offs.append(np.asarray(t[1:], dtype=np.int64) + base)
        base += int(t[-1])
    return ids, np.concatenate(offs)
