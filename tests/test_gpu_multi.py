"""In-process multi-GPU path of jtk_encode_batch (byte-balanced document ranges, one worker per device, no collective);
skipped on boxes with a single GPU (the one-process-per-GPU path is what bench.py measures)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_devices_match_one_device(oracles):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import jtokkit_b200 as jt
    from jtokkit_b200 import synth
    data, off = synth.config3_multilingual(torch.device("cpu"), total=48 << 20, seed=7)
    d, o = data.numpy(), off.numpy()
    one = jt.Encoding(jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE), devices=[0]).encode_packed(d, o)
    two = jt.Encoding(jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE), devices=[0, 1]).encode_packed(d, o)
    assert np.array_equal(one.ids, two.ids) and np.array_equal(one.token_offsets, two.token_offsets)
    assert np.array_equal(one.doc_status, two.doc_status)
    counts = jt.Encoding(jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE), devices=[1, 0]).encode_packed(d, o, count_only=True)
    assert np.array_equal(counts.token_offsets, one.token_offsets)
    # spot check against the oracle
    exp = oracles["cl100k_base"]
    for k in range(0, o.size - 1, 97):
        assert two.tokens(k) == exp.encode(bytes(d[o[k]:o[k + 1]]))
