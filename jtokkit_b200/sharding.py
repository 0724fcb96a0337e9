"""Byte-balanced document sharding: with doc_off the prefix sums of the document lengths, shard g of G
gets the contiguous documents [lower_bound(doc_off, g * total / G), lower_bound(doc_off, (g + 1) * total / G)).

This is the Python statement of the rule jtk_encode_batch applies in C when an encoding spans several devices
(jtk_capi.cu, jtk_encode_batch); callers that drive one process per GPU use it to pick their own range.  No
collective is involved: documents encode independently and the host only concatenates per-shard arrays."""
import numpy as np


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
