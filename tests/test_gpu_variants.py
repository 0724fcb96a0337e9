"""GPU tests of the call contract around the kernels: concurrent callers on one handle (the reference's Encoding is thread-safe
and shared, README.md:94-95) and the development switches, which must never change a result."""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_concurrent_callers_on_one_handle(gpu_encodings, oracles):
    from jtokkit_b200 import synth
    import torch
    enc, o = gpu_encodings["cl100k_base"], oracles["cl100k_base"]
    data, off = synth.config3_multilingual(torch.device("cpu"), total=6 << 20)
    utf8, off = data.numpy(), off.numpy()
    nd = off.size - 1
    exp_ids, exp_counts, _ = o.encode_batch(utf8, off, 8, check_special=True)
    expected = [exp_ids[off[d]:off[d] + exp_counts[d]].tolist() for d in range(nd)]
    errors = []

    def worker(k):
        try:
            lo, hi = k * nd // 6, (k + 1) * nd // 6
            for _ in range(3):
                sub_off = off[lo:hi + 1] - off[lo]
                res = enc.encode_packed(utf8[off[lo]:off[hi]], sub_off)
                got = res.to_lists()
                if got != expected[lo:hi] or res.doc_status.any():
                    errors.append("slice %d differs" % k)
                text = bytes(utf8[off[lo]:off[lo + 1]]).decode("utf-8")
                if enc.count_tokens(text) != len(expected[lo]) or enc.decode(expected[lo]) != text:
                    errors.append("single-call %d differs" % k)
        except Exception as e:  # noqa: BLE001 - reported below
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


SCRIPT = r"""
import hashlib, json, sys
sys.path.insert(0, %r)
import torch
import jtokkit_b200 as jt
from jtokkit_b200 import synth
enc = jt.EncodingFactory.cl100k_base()
data, off = synth.config3_multilingual(torch.device("cpu"), total=24 << 20)
res = enc.encode_packed(data.numpy(), off.numpy())
docs = synth.config5_adversarial(n=1 << 16)
blob, o = jt.pack_documents(docs)
res2 = enc.encode_packed(blob, o, ordinary=True)
print(json.dumps({"n": int(res.ids.size), "ids": hashlib.sha256(res.ids.tobytes()).hexdigest(), "off": hashlib.sha256(res.token_offsets.tobytes()).hexdigest(),
                  "n2": int(res2.ids.size), "ids2": hashlib.sha256(res2.ids.tobytes()).hexdigest()}))
"""


def run_variant(env):
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_development_switches_do_not_change_results():
    """JTK_MEMO / JTK_SIDE_STREAMS / JTK_PIPELINE / JTK_SUB_TILES / JTK_CHUNK_MB / JTK_MEMO_LOG2 only move work around."""
    base = run_variant({})
    assert base["n"] > 0 and base["n2"] > 0
    for env in [{"JTK_MEMO": "0"}, {"JTK_SIDE_STREAMS": "0"}, {"JTK_PIPELINE": "1"}, {"JTK_SUB_TILES": "256", "JTK_CHUNK_MB": "3"}, {"JTK_MEMO_LOG2": "16"},
                {"JTK_PIPELINE": "1", "JTK_SUB_TILES": "128", "JTK_SPLIT_CTAS": "4"}]:
        assert run_variant(env) == base, env


def test_dfa_switch_does_not_change_results():
    """JTK_RX_DFA=0 keeps the backtracking program for a general pattern that has a DFA: same ids, offsets and statuses on the
    multilingual corpus (the predefined cl100k string registered with CASE_INSENSITIVE, and a Unicode-category pattern)."""
    import torch
    import jtokkit_b200 as jt
    from jtokkit_b200 import synth
    p = jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE)
    data, off = synth.config3_multilingual(torch.device("cpu"), total=12 << 20, seed=21)
    d, o = data.numpy(), off.numpy()
    prev = os.environ.get("JTK_RX_DFA")
    try:
        for pat, flags in [(p.get_pattern().pattern(), 0x102), (r"\p{Lu}?\p{Ll}+|\p{N}{1,3}|\s+(?!\S)|\s+|[^\s\p{L}\p{N}]+", 0x100)]:
            out = []
            for sw in ("1", "0"):
                os.environ["JTK_RX_DFA"] = sw
                enc = jt.EncodingFactory.from_parameters(jt.GptBytePairEncodingParams("sw" + sw, jt.Pattern.compile(pat, flags), p.encoder, p.special_tokens_encoder))
                res = enc.encode_packed(d, o, ordinary=True)
                out.append((res.ids.copy(), res.token_offsets.copy(), res.doc_status.copy()))
                res.close()
                enc.close()
            assert all(np.array_equal(a, b) for a, b in zip(out[0], out[1])), pat
            assert out[0][0].size > 0 and not out[0][2].any()
    finally:
        if prev is None:
            os.environ.pop("JTK_RX_DFA", None)
        else:
            os.environ["JTK_RX_DFA"] = prev
