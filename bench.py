#!/usr/bin/env python3
"""Benchmark of the JTokkit encode hot path on B200 (BASELINE.json metric: cl100k_base encode tokens/s and input GB/s).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the reference algorithm's CPU restatement on the host cores

A "step" is one pass of the hot path over one batch of synthetic input: cl100k_base encode of a 1 GiB synthetic
multilingual corpus (documents 1-64 KiB), configs[2] of BASELINE.json, one corpus per GPU (weak scaling: documents are
independent, there is no data-path collective).
  value    : whole-job tokens/s, corpus resident in HBM, CUDA-event timed on the launching stream, max over ranks
  e2e      : same metric through the host-buffer C-ABI call (jtk_encode_batch) from pinned host memory, H2D + D2H inside
  roofline : algorithmic bytes of the tile kernel / its CUDA-event duration vs the measured HBM copy peak
  strong   : ONE 1 GiB corpus (rank 0's) encoded through jtk_encode_batch on a handle that spans all N GPUs of the box
             (in-library sharding: byte-balanced document chunks, chunk c on device c % N, no collective), host buffers in,
             host result out, checked id by id against the single-device result; the other ranks wait on a host-side barrier
  cpu_baseline : the oracle's C restatement of the reference algorithm ("port"; this image has no JVM, so JTokkit itself
             cannot run) on all host cores over the same corpus (all of it unless the host is too slow for ~30 s), which
             doubles as the parity check: a mismatch makes the bench exit non-zero
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEFAULT_BYTES = 1 << 30
REFERENCE_SAMPLE_BYTES = DEFAULT_BYTES  # the reference arm times the same 1 GiB corpus as the CUDA arm (~3 s per step on 16 cores)
METRIC = "cl100k_base encode tokens/s (1 GiB synthetic multilingual corpus per GPU)"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm = sorted(float(r[1]) for r in rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(rows)}


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process (and thus its pinned host buffers, first touch) to the CPUs NVML reports as local to the GPU:
    with several ranks per box the host<->device copies of the e2e leg otherwise cross the socket interconnect.
    Returns the previous affinity so that the CPU baseline can use every core again."""
    try:
        import pynvml
        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= before
        if cpus:
            os.sched_setaffinity(0, cpus)
        return before
    except Exception:
        return None


def reduce_over_ranks(stats, world):
    """(max over ranks, sum over ranks) of a 1-D float64 tensor; times use the max, work uses the sum."""
    if world <= 1:
        return stats, stats
    import torch.distributed as dist
    mx = stats.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = stats.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx, sm


def cpu_port_throughput(utf8_np, doc_off_np, sample_bytes, threads):
    """Times the oracle port (reference algorithm restatement) with `threads` host threads on the first documents of the
    corpus up to sample_bytes.  Returns (tokens/s, bytes, tokens, seconds, ids, counts)."""
    import numpy as np
    from oracle import jo
    jo.build()
    enc = jo.OracleEncoding.builtin("cl100k_base")
    nd = int(np.searchsorted(doc_off_np, sample_bytes, side="right")) - 1
    nd = max(nd, 1)
    nb = int(doc_off_np[nd])
    t0 = time.perf_counter()
    ids, counts, total = enc.encode_batch(utf8_np[:nb], doc_off_np[:nd + 1], threads, check_special=True)
    dt = time.perf_counter() - t0
    return total / dt, nb, int(total), dt, ids, counts, nd


def cpu_tiktoken_throughput(utf8_np, doc_off_np, sample_bytes, threads):
    """SURVEY.md §8d CPU baseline (2): tiktoken's encode_ordinary_batch exactly as the reference's own benchmark/bench.py:28 calls
    it, built from the reference's regex string and vocabulary file.  Context only (it is the upstream JTokkit mirrors, not
    JTokkit); None when tiktoken is not importable."""
    try:
        import tiktoken
        from tiktoken.load import load_tiktoken_bpe
    except ImportError:
        return None
    import numpy as np
    import jtokkit_b200 as jt
    params = jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE)
    path = os.path.join(os.path.dirname(os.path.abspath(jt.__file__)), "data", "cl100k_base.tiktoken")
    enc = tiktoken.Encoding("cl100k_jtokkit", pat_str=params.get_pattern().pattern(), mergeable_ranks=load_tiktoken_bpe(path), special_tokens={})
    nd = max(1, int(np.searchsorted(doc_off_np, sample_bytes, side="right")) - 1)
    raw = utf8_np[:int(doc_off_np[nd])].tobytes()
    docs = [raw[int(doc_off_np[d]):int(doc_off_np[d + 1])].decode("utf-8") for d in range(nd)]
    t0 = time.perf_counter()
    out = enc.encode_ordinary_batch(docs, num_threads=threads)
    dt = time.perf_counter() - t0
    ntok = sum(len(x) for x in out)
    return {"value": ntok / dt, "unit": "tokens/s", "cores": threads, "kind": "tiktoken %s encode_ordinary_batch" % tiktoken.__version__,
            "sample": "first %d bytes / %d documents (%d tokens) of the same corpus, %.1f s, %.3f GB/s input" % (len(raw), nd, ntok, dt, len(raw) / dt / 1e9)}


WORKLOAD = "cl100k_base encode (special-token guard on), 1 GiB synthetic multilingual corpus per GPU, documents log-uniform 1-64 KiB"


def workload_config(nbytes, ndocs, ntok):
    """The same dictionary in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "bytes_per_gpu": int(nbytes), "docs_per_gpu": int(ndocs), "tokens_per_gpu": int(ntok),
            "l2": "input (1 GiB) and output (~1.5 GiB) are far larger than the 126 MB L2; no flush needed"}


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores (oracle port: no JVM in this image)."""
    import numpy as np
    import torch
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return
    from jtokkit_b200 import synth
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    data, doc_off = synth.config3_multilingual(dev, total=args.bytes)
    utf8, off = data.cpu().numpy(), doc_off.cpu().numpy()
    threads = os.cpu_count() or 1
    times, tokens, nbytes = [], 0, 0
    for step in range(args.warmup + args.steps):
        tps, nbytes, tokens, dt, _, _, _ = cpu_port_throughput(utf8, off, int(utf8.size), threads)
        if step >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    value = tokens / dt
    sample = "the whole corpus of the CUDA arm's rank 0 per step: %d bytes, %d documents, %d tokens" % (nbytes, off.size - 1, tokens)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(nbytes, off.size - 1, tokens),
        "input_gb_per_s": nbytes / dt / 1e9,
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of the reference algorithm (regex find loop + HashMap-keyed O(n^2) bytePairMerge, one task per document "
                "on a thread pool like AbstractMultiThreadedBenchmark.java:34-45); JTokkit itself cannot run here (no JVM)",
    }))


def run_ours(args):
    import numpy as np
    import torch
    import jtokkit_b200 as jt
    from jtokkit_b200 import sharding, synth
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the encode path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = bind_to_gpu_numa_node(local)
    dev = torch.device("cuda", local)
    jt.EncodingFactory.devices = [local]
    enc = jt.EncodingFactory.cl100k_base()
    # ranks that only wait (strong-scaling leg) must wait on the HOST: an NCCL barrier would spin on their GPU
    host_group = torch.distributed.new_group(backend="gloo") if world > 1 else None

    # ---- workload: one 1 GiB corpus per GPU (different seed per rank)
    data, doc_off = synth.config3_multilingual(dev, total=args.bytes, seed=3003 + rank)
    pad = (-data.numel()) % 16
    d_in = torch.zeros(data.numel() + pad + 64, dtype=torch.uint8, device=dev)
    d_in[:data.numel()] = data
    nbytes, ndocs = data.numel(), doc_off.numel() - 1
    del data
    d_view = d_in[:nbytes]
    d_ids = torch.empty(nbytes + 16, dtype=torch.int32, device=dev)
    d_tok_off = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(ndocs + 1, dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value, roofline)
    for _ in range(args.warmup):
        enc.encode_device(d_view, doc_off, d_ids, d_tok_off, d_status, time_kernel=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches, ntok, nlong = [], 0, 0, 0
    e0.record()
    for _ in range(args.steps):
        ntok, nlong, nl, kms = enc.encode_device(d_view, doc_off, d_ids, d_tok_off, d_status, time_kernel=True)
        kernel_ms.append(kms)
        launches += nl
    e1.record()
    barrier()
    t_wall1 = time.time()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    assert int(d_status.max().item()) == 0, "synthetic corpus must not trip the special-token guard"

    # ---- decode: the step's 404 M ids back to bytes, device resident; must reproduce the corpus
    d_dec = torch.empty(nbytes + 64, dtype=torch.uint8, device=dev)
    d_boff = torch.empty(ndocs + 1, dtype=torch.int64, device=dev)
    d_dst = torch.zeros(ndocs + 1, dtype=torch.int32, device=dev)
    d_dbad = torch.empty(ndocs + 1, dtype=torch.int32, device=dev)
    ids_view = d_ids[:ntok]
    for _ in range(2):
        enc.decode_device(ids_view, d_tok_off, d_dec[:nbytes], d_boff, d_dst, d_dbad)
    torch.cuda.synchronize()
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    q0.record()
    dec_launches = 0
    for _ in range(args.steps):
        dec_total, nl = enc.decode_device(ids_view, d_tok_off, d_dec[:nbytes], d_boff, d_dst, d_dbad)
        dec_launches += nl
    q1.record()
    torch.cuda.synchronize()
    dec_ms = q0.elapsed_time(q1) / args.steps
    dec_ok = dec_total == nbytes and bool(torch.equal(d_dec[:nbytes], d_view)) and bool(torch.equal(d_boff, doc_off.to(dev))) and int(d_dst.max().item()) == 0
    dec_algo = 4 * ntok + nbytes + 16 * (ndocs + 1)
    del d_dec

    # ---- end to end through the host-buffer C-ABI call, pinned host input
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_in.copy_(d_view)
    h_off = doc_off.cpu()
    h_in_np, h_off_np = h_in.numpy(), h_off.numpy()
    r = None
    for _ in range(max(2, min(args.warmup, 3))):
        r = enc.encode_packed(h_in_np, h_off_np, copy=False)
        r.close()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        # the step's result (ids, token offsets, status) is in host memory when the call returns; the previous
        # step's buffers go back to the library's pinned pool
        if r is not None:
            r.close()
        r = enc.encode_packed(h_in_np, h_off_np, copy=False)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_tokens = int(r.token_offsets[-1])
    assert e2e_tokens == ntok, (e2e_tokens, ntok)
    assert np.array_equal(r.ids[:4096], d_ids[:4096].cpu().numpy())

    # ---- reduce over ranks: time = max, tokens / bytes = sum
    stats = torch.tensor([dev_ms, e2e_s, float(ntok), float(nbytes), float(launches), float(ndocs), float(sum(kernel_ms) / len(kernel_ms))],
                         dtype=torch.float64, device=dev)
    mx, sm = reduce_over_ranks(stats, world)

    # ---- strong scaling: ONE corpus (rank 0's) through ONE handle that spans all N GPUs (jtk_encode_batch shards it in the library)
    strong = None
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier(group=host_group)  # every rank's GPU is idle from here on
    if rank == 0:
        if all_cpus:
            os.sched_setaffinity(0, all_cpus)
        enc_all = enc if world == 1 else jt.Encoding(jt.EncodingFactory.predefined_params(jt.EncodingType.CL100K_BASE), devices=list(range(world)))
        # the same call on ONE device while every other GPU of the box is idle: the denominator of the strong-scaling speed-up
        rs = enc.encode_packed(h_in_np, h_off_np, copy=False)  # (warm-up: `r` still holds a result buffer, this call allocates a second one)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rs.close()
            rs = enc.encode_packed(h_in_np, h_off_np, copy=False)
        alone_s = (time.perf_counter() - t0) / args.steps
        rs.close()
        for _ in range(2):
            rs = enc_all.encode_packed(h_in_np, h_off_np, copy=False)
            rs.close()
        rs = None
        kernel_ms_strong = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            if rs is not None:
                rs.close()
            rs = enc_all.encode_packed(h_in_np, h_off_np, copy=False)
            kernel_ms_strong += rs.device_ms
        strong_s = (time.perf_counter() - t0) / args.steps
        same = bool(np.array_equal(rs.ids, r.ids)) and bool(np.array_equal(rs.token_offsets, r.token_offsets)) and not rs.doc_status.any()
        plan = sharding.plan_chunks(h_off_np, world)
        strong = {"n_gpus": world, "api": "jtk_encode_batch on one handle over %d devices (in-library chunk sharding, no collective)" % world,
                  "corpus": "rank 0's 1 GiB corpus (%d bytes, %d documents)" % (nbytes, ndocs), "chunks": int(plan.size - 1),
                  "ms_per_step": strong_s * 1e3, "value": ntok / strong_s, "unit": "tokens/s", "input_gb_per_s": nbytes / strong_s / 1e9,
                  "kernel_ms_per_step_max_over_devices": kernel_ms_strong / args.steps,
                  "one_device_alone_ms_per_step": alone_s * 1e3, "speedup_vs_one_device": alone_s / strong_s,
                  "h2d_bytes_per_step": int(nbytes + 8 * (ndocs + 1)), "d2h_bytes_per_step": int(4 * ntok + 12 * (ndocs + 1)),
                  "parity": "ids, token offsets and statuses identical to the single-device result" if same else "MISMATCH vs single device"}
        rs.close()
        if world > 1:
            enc_all.close()
    if world > 1:
        torch.distributed.barrier(group=host_group)
    if rank != 0:
        return
    if strong and strong["parity"].startswith("MISMATCH"):
        print(json.dumps({"error": "multi-device result differs from the single-device result", "strong": strong}), flush=True)
        raise SystemExit(3)
    dev_ms_max, e2e_s_max = float(mx[0]), float(mx[1])
    tokens_all, bytes_all, launches_all, ndocs_all = float(sm[2]), float(sm[3]), int(sm[4]), float(sm[5])
    value = tokens_all * args.steps / (dev_ms_max * 1e-3)
    e2e_value = tokens_all * args.steps / e2e_s_max

    # ---- roofline.  The headline fraction is the conservative one: the algorithmic bytes of the WHOLE step (bytes in + ids out + offsets,
    # SURVEY.md section 8d) over the whole step's device time.  The dominant kernel is reported next to it with the part of those bytes it
    # moves itself (it reads the input; the ids are written by the gather kernel) over its own CUDA-event time.  DRAM traffic: ncu
    # dram__bytes_read + write summed over every kernel of one step / over the dominant kernel's launches (profiles/r2_step_traffic.json).
    peak, peak_src = measured_peaks()
    algo_bytes = nbytes + 4 * ntok + 16 * (ndocs + 1)
    kms = sum(kernel_ms) / len(kernel_ms)
    step_ms = dev_ms / args.steps
    achieved = algo_bytes / (step_ms * 1e-3) / 1e9
    traffic, kernel_traffic = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_step_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_step")
        kernel_traffic = tj.get("per_kernel", {}).get("jtk_split_lookup_kernel<0>", {}).get("dram_bytes")
    kernel_algo = nbytes + 8 * (ndocs + 1)  # what the split+lookup kernel reads of the algorithmic bytes

    # ---- CPU baseline on a bounded sample of the same corpus (rank 0, N=1 only) + parity of that sample
    cpu = None
    parity = None
    tk = None
    if world == 1 and not args.no_cpu_baseline:
        if all_cpus:
            os.sched_setaffinity(0, all_cpus)
        threads = os.cpu_count() or 1
        probe = cpu_port_throughput(h_in_np, h_off_np, 8 << 20, threads)
        # the whole corpus when the host gets through it in ~30 s (16 cores: ~3 s), else a prefix of it
        sample_bytes = int(min(nbytes, max(16 << 20, probe[1] / probe[3] * 30.0)))
        tps, sb, st, dt, o_ids, o_counts, nd = cpu_port_throughput(h_in_np, h_off_np, sample_bytes, threads)
        whole = nd == ndocs
        cpu = {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port",
               "sample": "%s: %d bytes / %d documents (%d tokens), %.1f s, %.3f GB/s input" %
                         ("the whole corpus" if whole else "a prefix of the same corpus", sb, nd, st, dt, sb / dt / 1e9)}
        ok = bool(np.array_equal(o_counts[:nd], np.diff(r.token_offsets[:nd + 1])))
        for d in range(nd):  # every document of the sample, id by id
            if not ok:
                break
            c = int(o_counts[d])
            ok = bool(np.array_equal(o_ids[h_off_np[d]:h_off_np[d] + c], r.ids[r.token_offsets[d]:r.token_offsets[d] + c]))
        parity = "bit-exact vs oracle: counts and ids of %s (%d of %d documents, %d tokens)" % ("every document" if whole else "a prefix", nd, ndocs, st) \
            if ok else "MISMATCH vs oracle"
        del o_ids
        tk = cpu_tiktoken_throughput(h_in_np, h_off_np, 64 << 20, threads)

    out = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(nbytes, ndocs, ntok), "long_pieces": nlong,
        "input_gb_per_s": bytes_all * args.steps / (dev_ms_max * 1e-3) / 1e9,
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": int(nbytes + 8 * (ndocs + 1)),
                "d2h_bytes_per_step": int(4 * ntok + 12 * (ndocs + 1)), "input_gb_per_s": bytes_all * args.steps / e2e_s_max / 1e9,
                "ms_per_step": e2e_s_max / args.steps * 1e3, "api": "jtk_encode_batch (host buffers, pinned input)",
                # the link, not the kernels, bounds this number: PCIe rates measured on this pool (profiles/r1_pcie.txt: one GPU, both directions
                # busy, 43-49 GB/s each; profiles/r2_pcie_8gpu.txt: eight GPUs, 92 GB/s device-to-host in total)
                "link_floor_ms": max(nbytes / 46e9, 4 * ntok / 46e9) * 1e3 if world == 1 else max(bytes_all / 184e9, 4 * tokens_all / 92e9) * 1e3,
                "link_floor_note": "max over directions of bytes / measured link rate while both directions are busy (1 GPU: 46 GB/s each; 8 GPUs: 184 GB/s in, 92 GB/s out in total)"},
        "gpu_launches": launches_all,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "scope": "whole step: all %d kernel launches of one pass over the corpus" % int(launches // args.steps),
                     "algorithmic_bytes_per_step": int(algo_bytes), "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                     "traffic_over_algorithmic": (traffic / algo_bytes) if traffic else None,
                     "dominant_kernel": {"name": "jtk_split_lookup_kernel", "ms_per_step": kms, "share_of_step": kms / step_ms,
                                         "algorithmic_bytes": int(kernel_algo), "achieved": kernel_algo / (kms * 1e-3) / 1e9,
                                         "frac": kernel_algo / (kms * 1e-3) / 1e9 / peak, "traffic": kernel_traffic,
                                         "bound_in_practice": "instruction issue (integer work per byte), not HBM: see profiles/r2_split_lookup_ncu_full.txt"}},
        "strong": strong,
        "decode": {"api": "jtk_decode_batch_device (ids of the step, resident in HBM -> bytes)", "ms_per_step": dec_ms, "tokens_per_s": ntok / (dec_ms * 1e-3),
                   "algorithmic_bytes": int(dec_algo), "achieved_gb_per_s": dec_algo / (dec_ms * 1e-3) / 1e9, "frac_of_hbm_peak": dec_algo / (dec_ms * 1e-3) / 1e9 / peak,
                   "gpu_launches_per_step": int(dec_launches // args.steps),
                   "parity": "decoded bytes and document offsets identical to the corpus" if dec_ok else "MISMATCH: decode(encode(x)) != x"},
        "cpu_baseline": cpu,
        "cpu_tiktoken": tk,
        "parity": parity,
        "clocks": clocks,
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(out), flush=True)
    if (parity is not None and parity.startswith("MISMATCH")) or not dec_ok:
        raise SystemExit(3)  # a fast result that differs from the reference algorithm is not a result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=DEFAULT_BYTES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
