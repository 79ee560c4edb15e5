#!/usr/bin/env python
"""bench.py — headline benchmark of the two hot paths (contract in the task statement; metric from BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Primary line: Hamming mAP@k queries/s on the MS-COCO shape (BASELINE.json configs[2]: 5k queries x 117k database rows,
80 labels, 128-bit codes, k = 5000 — the configuration quoted "1/2/4/8 B200" like the metric).  One step = one full
evaluation of all queries: bit-pack the float32 +-1 codes / multi-hot labels as the reference hands them over, then the
counting-sort evaluator (stage A histogram, stage S scan, stage B AP, finalize).  With N > 1 GPUs the database rows are
split into N contiguous shards (queries replicated) and the stages exchange shard totals / per-query partials with two
NCCL all-gathers: total work is fixed, so "scaling" is "strong".

`value`  : device-resident float32 inputs, CUDA-event time per step, max over ranks, L2 flushed between steps.
`e2e`    : the same evaluation through the host-buffer C-ABI call (`b200_maphashing_host`; pinned host float32 inputs,
           H2D + pack + stages + D2H inside the timed region).
`extras` : SWT images/s (BASELINE configs[0] and [3]) with its HBM roofline, and the other Hamming shapes.
`--impl reference` times the reference's own CPU algorithm (literal torch restatement of calculate_maphashing — the
reference's third-party stack is not installable offline) on a bounded query sample.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "hamming_map_at_k_queries_per_sec"
WORKLOADS = {
    # name: (description from BASELINE.json, Q, N, bits, labels, k, label density)
    "c1": ("MIRFlickr-25K shape: 2k queries x 18k db, 24 labels, 64-bit codes, mAP@5000", 2000, 18000, 64, 24, 5000, 0.10),
    "c2": ("PASCAL VOC shape: 5k queries x 11.5k db, 20 labels, 64-bit codes, mAP@all", 5000, 11500, 64, 20, None, 0.10),
    "c2_32": ("PASCAL VOC shape, 32-bit codes", 5000, 11500, 32, 20, None, 0.10),
    "c2_128": ("PASCAL VOC shape, 128-bit codes", 5000, 11500, 128, 20, None, 0.10),
    "c3": ("MS-COCO shape: 5k queries x 117k db, 80 labels, 128-bit codes, mAP@5000", 5000, 117000, 128, 80, 5000, 0.036),
    "c3_all": ("MS-COCO shape, mAP@all (k = 117000)", 5000, 117000, 128, 80, None, 0.036),
    "c5": ("Scale-out: 1M-image db, 64-bit codes, 10k queries, mAP@5000", 10000, 1000000, 64, 80, 5000, 0.036),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(key):
    """DRAM bytes per launch of a kernel from the committed ncu captures (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f).get(key)
        return (int(e["bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def make_problem(name, seed=0):
    """Synthetic evaluator inputs: multi-hot labels Bernoulli(p) with >= 1 tag, codes =
    sign(labels . W + noise) so that rankings are informative.  float32, CPU, torch.Generator().manual_seed(seed)."""
    _, nq, n, bits, nlab, k, p = WORKLOADS[name]
    g = torch.Generator().manual_seed(seed)

    def labels(rows):
        lab = (torch.rand(rows, nlab, generator=g) < p).float()
        empty = lab.sum(1) == 0
        lab[empty, torch.randint(0, nlab, (int(empty.sum()),), generator=g)] = 1.0
        return lab

    ql, rl = labels(nq), labels(n)
    w = torch.randn(nlab, bits, generator=g)

    def codes(lab):
        z = lab @ w + 0.8 * torch.randn(lab.shape[0], bits, generator=g)
        return torch.where(z > 0, 1.0, -1.0).float()

    return codes(ql), ql, codes(rl), rl, (n if k is None else k)


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            pass
    return local_rank


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
def literal_step(q, ql, r, rl, k):
    from oracle.eval_ref import maphashing_literal

    return maphashing_literal(q, ql, r, rl, k)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    desc, nq, n, bits, nlab, _, _ = WORKLOADS[name]
    q, ql, r, rl, k = make_problem(name)
    cores = os.cpu_count() or 1
    # bounded sample: as many queries as take about 2 s per step with the better thread setting
    probe = min(8, nq)
    best = None
    for threads in sorted({1, cores}):
        torch.set_num_threads(threads)
        literal_step(q[:2], ql[:2], r, rl, k)
        t0 = time.perf_counter()
        literal_step(q[:probe], ql[:probe], r, rl, k)
        per_q = (time.perf_counter() - t0) / probe
        if best is None or per_q < best[1]:
            best = (threads, per_q)
    threads, per_q = best
    torch.set_num_threads(threads)
    budget = 90.0 / max(1, args.steps + args.warmup)                     # whole run well inside a few minutes
    sample = int(max(4, min(nq, min(2.0, budget) / per_q)))
    for _ in range(args.warmup):
        literal_step(q[:sample], ql[:sample], r, rl, k)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        literal_step(q[:sample], ql[:sample], r, rl, k)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: {desc}", "queries": nq, "database": n, "code_bits": bits, "labels": nlab, "top_k": k},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of {nq} queries per step against the full database; literal torch restatement of "
                                   "calculate_maphashing (accuracy_calculator.py:203-231); host has "
                                   f"{cores} logical cores, {threads} torch thread(s) was the faster setting"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- own arm (GPU)
def l2_flusher(device):
    buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)    # > 126 MB L2
    return lambda: buf.zero_()


def bench_hamming(name, args, world, rank, device, dist, with_e2e=True, with_cpu=True, check=True):
    from image_retrieval_wavelet_b200 import _cabi
    from image_retrieval_wavelet_b200.engine import hamming as H
    from image_retrieval_wavelet_b200.engine.dist import ShardedHammingEvaluator, shard_bounds

    desc, nq, n, bits, nlab, _, _ = WORKLOADS[name]
    q, ql, r, rl, k = make_problem(name)
    b0, b1 = shard_bounds(n, world)[rank]
    dq, dql = q.to(device), ql.to(device)
    dr, drl = r[b0:b1].contiguous().to(device), rl[b0:b1].contiguous().to(device)
    ev = ShardedHammingEvaluator(mode="hist")
    flush = l2_flusher(device)

    def step():
        qc, qlp = H.pack_codes(dq, on_nonbinary="sign"), H.pack_labels_unchecked(dql)
        dc, dlp = H.pack_codes(dr, on_nonbinary="sign"), H.pack_labels_unchecked(drl)
        return ev.evaluate(qc, qlp, [(dc, dlp, b0)], n, k)

    m, ap, tsum = step()
    torch.cuda.synchronize()
    checked = None
    if check and rank == 0:
        from oracle import c_oracle

        sub = np.random.default_rng(0).choice(nq, 16, replace=False)
        m0, ap0, ts0 = c_oracle.maphashing(q[sub].numpy(), ql[sub].numpy(), r.numpy(), rl.numpy(), k)
        ok = np.array_equal(tsum.cpu().numpy()[sub].astype(np.int64), ts0) and np.abs(ap.cpu().numpy()[sub] - ap0).max() <= 1e-6
        if not ok:
            raise SystemExit(f"bench: {name}: CUDA result differs from the oracle on the check sample")
        checked = "16-query sample bit-exact (hits) / 1e-6 (AP) against the C oracle"
    for _ in range(max(args.warmup, 3)):
        step()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(gpu_index(int(os.environ.get("LOCAL_RANK", "0"))))
    sampler.start()
    launches0 = _cabi.launch_count()
    ev.timeline = []
    starts, ends = [], []
    for _ in range(args.steps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ev.timeline.append(("start", s))
        out = step()
        e.record()
        ev.timeline.append(("end", e))
        starts.append(s), ends.append(e)
        torch.cuda.synchronize()          # no CPU run-ahead: every step starts from an idle device, like the SWT loop
    if dist is not None:
        dist.barrier()
    launches = _cabi.launch_count() - launches0
    timeline, ev.timeline = ev.timeline, None
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([total_ms], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    # keep the GPU busy a little longer (about 0.25 s) so that the clock sampler sees the kernels under load.  The step
    # count comes from the all-reduced time: every rank runs the SAME number of steps (they contain collectives).
    for _ in range(int(min(400, max(3, 250.0 / max(ms_per_step, 1e-3))))):
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    value = nq / (ms_per_step * 1e-3)
    # per-stage device times from the marks recorded inside the timed steps
    stage_ms = {}
    for (n0, e0), (n1, e1) in zip(timeline[:-1], timeline[1:]):
        if n1 != "start":                                   # "begin" = pack + plan (from the step's start to stage A's launch)
            stage_ms.setdefault("pack" if n1 == "begin" else ("tail" if n1 == "end" else n1), []).append(e0.elapsed_time(e1))
    stage_avg = {kname: float(np.mean(v)) for kname, v in stage_ms.items()}
    cw, lw = _cabi.code_words(bits), _cabi.label_words(nlab)
    rows = b1 - b0
    algo_bytes = rows * (cw + lw) * 8 + nq * (cw + lw) * 8 + nq * 12           # SURVEY.md §8d compulsory traffic, per launch
    dom = max(("hist", "ap"), key=lambda s_: stage_avg.get(s_, 0.0))
    dom_ms = stage_avg.get(dom, float("nan"))
    peak, peak_src = measured_peaks()
    achieved = algo_bytes / (dom_ms * 1e-3) / 1e9
    sm_mhz = clocks.get("sm_mhz") or 1900.0
    pair_rate = nq * rows / (dom_ms * 1e-3)
    dom_kernel = "hamming_hist_kernel" if dom == "hist" else ("hamming_rank_kernel" if 4 * k <= n else "hamming_walk_kernel")
    traffic, traffic_src = measured_traffic(f"{dom_kernel}:{name}") if world == 1 else (None, None)
    roofline = {
        "bound": "hbm", "kernel": f"{dom_kernel} (stage {'A' if dom == 'hist' else 'B'})", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": dom_ms,
        "note": "the packed database is L2-resident by design, so this kernel is bound by the integer pipes (ALU + POPC), not HBM; "
                "see issue_bound (POPC pipe measured at 25.5 lane-ops/clk/SM by tools/ubench_int.cu)",
        "issue_bound": {"pairs_per_s": pair_rate, "pairs_per_clk_per_sm": pair_rate / (sm_mhz * 1e6) / 148.0,
                        "popc32_per_pair": 2 * cw, "popc_pipe_peak_pairs_per_clk_per_sm": 25.5 / (2 * cw)},
    }
    result = {
        "value": value, "ms_per_step": ms_per_step, "stage_ms": stage_avg, "roofline": roofline, "clocks": clocks,
        "gpu_launches": int(launches), "checked": checked, "map": float(out[0].item()), "top_k": k,
        "config": {"workload": f"{name}: {desc}", "queries": nq, "database": n, "code_bits": bits, "labels": nlab, "top_k": k,
                   "sharding": f"database rows in {world} contiguous shard(s), queries replicated, hist exchange (2 all-gathers)"
                   if world > 1 else "single shard", "l2": "flushed between steps (256 MiB memset, untimed)",
                   "timed": "pack (4 launches) + stage A + totals + stage S + stage B + reduce + finalize, CUDA events, max over ranks"},
    }
    # ---- end to end: pinned host float32 -> H2D -> pack -> stages -> D2H
    if with_e2e:
        hq, hql = q.pin_memory(), ql.pin_memory()
        hr, hrl = r[b0:b1].contiguous().pin_memory(), rl[b0:b1].contiguous().pin_memory()
        h2d = (hq.numel() + hql.numel() + hr.numel() + hrl.numel()) * 4
        if world == 1:
            lib = _cabi.load()
            m_out, bad = ctypes.c_double(), ctypes.c_int()

            def e2e_step():
                rc = lib.b200_maphashing_host(hq.data_ptr(), hql.data_ptr(), hr.data_ptr(), hrl.data_ptr(), nq, n, bits, nlab, 0, k,
                                              None, None, ctypes.addressof(m_out), ctypes.addressof(bad))
                _cabi.check(rc, "b200_maphashing_host")
                return m_out.value
            d2h = 8 + 8
            api = "b200_maphashing_host (C-ABI, host buffers)"
        else:
            def e2e_step():
                a, b_, c_, d_ = (t_.to(device, non_blocking=True) for t_ in (hq, hql, hr, hrl))
                qc, qlp = H.pack_codes(a, on_nonbinary="sign"), H.pack_labels_unchecked(b_)
                dc, dlp = H.pack_codes(c_, on_nonbinary="sign"), H.pack_labels_unchecked(d_)
                return ev.evaluate(qc, qlp, [(dc, dlp, b0)], n, k)[0].item()
            d2h = 8
            api = "ShardedHammingEvaluator.evaluate on pinned host tensors (H2D + pack + stages + .item())"
        for _ in range(3):
            e2e_step()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_map = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        result["e2e"] = {"value": nq / (float(t.item()) / args.steps), "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
                         "d2h_bytes_per_step": int(d2h), "ms_per_step": float(t.item()) / args.steps * 1e3, "api": api,
                         "map": float(e2e_map)}
    # ---- CPU baseline (rank 0, single GPU runs only): C/OpenMP port of the reference loop on a bounded query sample
    if with_cpu and rank == 0 and world == 1:
        from oracle import c_oracle

        threads = c_oracle.num_threads()
        probe = min(nq, 2 * threads)
        t0 = time.perf_counter()
        c_oracle.maphashing(q[:probe].numpy(), ql[:probe].numpy(), r.numpy(), rl.numpy(), k)
        per_q = (time.perf_counter() - t0) / probe
        sample = int(max(threads, min(nq, 12.0 / per_q)))
        t0 = time.perf_counter()
        c_oracle.maphashing(q[:sample].numpy(), ql[:sample].numpy(), r.numpy(), rl.numpy(), k)
        dt = time.perf_counter() - t0
        result["cpu_baseline"] = {"value": sample / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                                  "sample": f"{sample} of {nq} queries against the full database ({dt:.1f} s); oracle/c/oracle.c "
                                            "(float dot products, comparison sort, AP) with OpenMP over queries"}
    return result


def bench_swt(shape, wavelet, level, dtype, args, device, with_cpu=False):
    from image_retrieval_wavelet_b200 import _cabi
    from image_retrieval_wavelet_b200.transforms import swt2

    b, c, h, w = shape
    g = torch.Generator().manual_seed(0)
    x8 = torch.randint(0, 256, (min(b, 8), c, h, w), dtype=torch.uint8, generator=g)
    x8 = x8.repeat((b + x8.shape[0] - 1) // x8.shape[0], 1, 1, 1)[:b].contiguous()
    x = (x8.to(device) if dtype == "u8" else (x8.float() / 255.0).to(device)).contiguous()
    out = torch.empty((b, c, 4, h, w), dtype=torch.float32, device=device)
    flush = l2_flusher(device)
    for _ in range(max(args.warmup, 3)):
        swt2(x, wavelet, level, out=out)
    torch.cuda.synchronize()
    times = []
    launches0 = _cabi.launch_count()
    for _ in range(args.steps):
        flush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        swt2(x, wavelet, level, out=out)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    ms = float(np.mean(times))
    px = b * c * h * w
    algo = px * ((1 if dtype == "u8" else 4) + 16)                    # 17 / 20 bytes per image-channel pixel
    peak, peak_src = measured_peaks()
    res = {
        "workload": f"SWT {wavelet} level {level} on {b}x{c}x{h}x{w} {dtype}", "images_per_s": b / (ms * 1e-3), "ms": ms,
        "ms_min": float(np.min(times)), "gpu_launches_per_step": (_cabi.launch_count() - launches0) // args.steps,
        "roofline": {"bound": "hbm", "kernel": "swt2_tile_kernel", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": algo / (ms * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": algo,
                     "peak_source": peak_src},
    }
    res["roofline"]["traffic"], res["roofline"]["traffic_source"] = measured_traffic("swt2_tile_kernel:" + res["workload"])
    if with_cpu:
        from oracle import c_oracle, filters

        lo, hi = filters.filter_bank(wavelet)
        sample = x8[:min(b, 64)].numpy()
        c_oracle.swt2(sample[:2], lo, hi, level)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 5.0:
            c_oracle.swt2(sample, lo, hi, level)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        res["cpu_baseline"] = {"value": sample.shape[0] / dt, "unit": "images/s", "cores": c_oracle.num_threads(), "kind": "port",
                               "sample": f"{sample.shape[0]} images x {reps} repetitions; oracle/c/oracle.c periodised a-trous loops "
                                         "(CPU restatement of pywt.swt2), OpenMP over image channels"}
    return res


def bench_knn(args, device, nq=5000, n=117000, d=768, k=2048):
    """BASELINE configs[2] cosine rerank: get_knn (inner product) on L2-normalised float32 embeddings, tensor-core scorer."""
    from image_retrieval_wavelet_b200 import _cabi
    from image_retrieval_wavelet_b200.engine.get_knn import knn_topk

    g = torch.Generator().manual_seed(0)
    refs = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1).to(device)
    qs = torch.nn.functional.normalize(torch.randn(nq, d, generator=g), dim=1).to(device)
    flush = l2_flusher(device)
    res = {"workload": f"cosine k-NN: {nq} queries x {n} refs, D={d}, k={k} (get_knn, inner product)"}
    for label, env in (("tensor_core", "1"), ("simt_fp32", "0")):
        os.environ["B200_KNN_TC"] = env
        for _ in range(2):
            knn_topk(refs, qs, k, "cosine")
        torch.cuda.synchronize()
        times, l0 = [], _cabi.launch_count()
        steps = min(args.steps, 5)
        for _ in range(steps):
            flush()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            knn_topk(refs, qs, k, "cosine")
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        ms = float(np.mean(times))
        res[label] = {"ms": ms, "queries_per_s": nq / (ms * 1e-3), "gpu_launches_per_step": (_cabi.launch_count() - l0) // steps}
    os.environ.pop("B200_KNN_TC", None)
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["bf16_tflops_sustained"]), "measured sustained bf16 (MEASURED_PEAKS.json)"
    except Exception:
        peak, src = 1400.0, "fallback (B200_PROFILING.md)"
    flops = 2.0 * nq * n * d
    ms = res["tensor_core"]["ms"]
    res["roofline"] = {"bound": "tensor", "kernel": "knn_scores_tc_kernel (+ split, select)", "achieved": flops / (ms * 1e-3) / 1e12,
                       "peak": peak, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / peak, "traffic": None, "peak_source": src,
                       "note": "achieved counts the algorithmic 2*Q*N*D flops over the whole call (split + GEMM + select); the "
                               "GEMM executes 3x that (bf16 hi/lo split for float32-grade scores)"}
    return res


def bench_fix_size(shape, level, wavelet, args, device):
    """fix_size on the device (Pillow-exact bicubic, 518 -> 520) and fix_size + SWT as SWTTransform.forward runs it."""
    from image_retrieval_wavelet_b200.transforms import SWTTransform, resize_u8

    b, c, h, w = shape
    g = torch.Generator().manual_seed(0)
    x8 = torch.randint(0, 256, (min(b, 8), c, h, w), dtype=torch.uint8, generator=g)
    x = x8.repeat((b + x8.shape[0] - 1) // x8.shape[0], 1, 1, 1)[:b].contiguous().to(device)
    f = 1 << level
    ho, wo = -(-h // f) * f, -(-w // f) * f
    t = SWTTransform(level=level, wavelet=wavelet)
    flush = l2_flusher(device)
    res = {"workload": f"fix_size {h}x{w} -> {ho}x{wo} (PIL bicubic) + SWT {wavelet} level {level} on {b}x{c} uint8 planes"}
    for key, fn in (("resize_ms", lambda: resize_u8(x, (ho, wo))), ("resize_plus_swt_ms", lambda: t.forward(x))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.steps):
            flush()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        res[key] = float(np.mean(times))
    res["images_per_s"] = b / (res["resize_plus_swt_ms"] * 1e-3)
    res["note"] = "includes torch's allocation of the outputs; resize = horizontal + vertical pass through a uint8 intermediate"
    return res


def bench_dsch(args, device):
    """DSCH's metrics on the C3 shape (5k x 117k, 128 bit, 80 labels), wall time per call incl. packing."""
    from image_retrieval_wavelet_b200.engine import DSCH

    q, ql, r, rl, _ = make_problem("c3")
    q, ql, r, rl = (t.to(device) for t in (q, ql, r, rl))
    res = {"workload": "DSCH metrics, MS-COCO shape (5000 x 117000, 128 bit, 80 labels), device-resident float inputs"}
    for key, fn in (("pr_curve_ms", lambda: DSCH.pr_curve(q, r, ql, rl)), ("p_topK_ms", lambda: DSCH.p_topK(q, r, ql, rl)),
                    ("radius2_precision_ms", lambda: DSCH.get_precision_recall_by_Hamming_Radius(r, rl, q, ql)),
                    ("mean_average_precision_at_5000_ms", lambda: DSCH.mean_average_precision(q, r, ql, rl, 5000))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        res[key] = (time.perf_counter() - t0) / 3 * 1e3
    try:                                    # calculate_pr_rc_hashing: full ranking of every query, 500 queries of the 5000
        from image_retrieval_wavelet_b200.engine import CustomCalculator

        calc = CustomCalculator(k=None, distance_metric="hamming", with_faiss=False)
        calc.pr_rc_hashing_curves(q[:64], ql[:64], r, rl)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, _, used = calc.pr_rc_hashing_curves(q[:500], ql[:500], r, rl)
        torch.cuda.synchronize()
        res["pr_rc_hashing_500_queries_ms"] = (time.perf_counter() - t0) * 1e3
        res["pr_rc_hashing_queries_used"] = used
    except Exception as exc:
        res["pr_rc_hashing_error"] = repr(exc)
    return res


def bench_dwt(args, device):
    """DWTTransform on the CIFAR shape of config/transform/cifar_dwt.yaml and on a 224 x 224 batch."""
    from image_retrieval_wavelet_b200.transforms import dwt2

    res = {}
    g = torch.Generator().manual_seed(0)
    for key, shape, wv, lv in (("cifar_haar_L2_1024x3x32x32_ms", (1024, 3, 32, 32), "haar", 2),
                               ("db4_L3_64x3x224x224_ms", (64, 3, 224, 224), "db4", 3)):
        x = torch.randint(0, 256, shape, dtype=torch.uint8, generator=g).to(device)
        for _ in range(3):
            dwt2(x, wv, lv)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            dwt2(x, wv, lv)
        e.record()
        torch.cuda.synchronize()
        res[key] = s.elapsed_time(e) / args.steps
    return res


def run_own(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 hot paths have no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    from image_retrieval_wavelet_b200.engine import hamming as H

    if not hasattr(H, "pack_labels_unchecked"):
        raise SystemExit("bench.py: package too old")
    main = bench_hamming(args.workload, args, world, rank, device, dist)
    extras = {}
    if not args.no_extras and world == 1:
        small = argparse.Namespace(**{**vars(args), "steps": min(args.steps, 10)})
        for name in ("c1", "c2_32", "c2", "c2_128", "c3_all", "c5"):
            if name == args.workload:
                continue
            try:
                r_ = bench_hamming(name, small, 1, 0, device, None, with_e2e=name in ("c1", "c2"), with_cpu=False, check=True)
                extras[name] = {kk: r_[kk] for kk in ("value", "ms_per_step", "stage_ms", "map", "checked") if kk in r_}
                extras[name]["unit"] = "queries/s"
                extras[name]["workload"] = r_["config"]["workload"]
                if "e2e" in r_:
                    extras[name]["e2e"] = r_["e2e"]
                extras[name]["issue_bound"] = r_["roofline"]["issue_bound"]
            except Exception as exc:                                   # an extra must never take the headline down
                extras[name] = {"error": repr(exc)}
        # C1 (both input types) and the whole C4 grid: haar/db2/db4/sym4 x levels 1-3 (518 at level 1, fix_size's 520 above)
        swt_cases = [((64, 3, 224, 224), "haar", 1, "u8", True), ((64, 3, 224, 224), "haar", 1, "f32", False)]
        for lv in (1, 2, 3):
            for wv in ("haar", "db2", "db4", "sym4"):
                swt_cases.append(((256, 3, 518, 518) if lv == 1 else (256, 3, 520, 520), wv, lv, "u8", False))
        try:
            extras["knn"] = bench_knn(small, device)
        except Exception as exc:
            extras["knn"] = {"error": repr(exc)}
        for key, fn in (("dsch", lambda: bench_dsch(small, device)), ("dwt", lambda: bench_dwt(small, device)),
                        ("fix_size", lambda: bench_fix_size((256, 3, 518, 518), 2, "haar", small, device))):
            try:
                extras[key] = fn()
            except Exception as exc:
                extras[key] = {"error": repr(exc)}
        extras["swt"] = []
        for shape, wv, lv, dt, cpu in swt_cases:
            try:
                extras["swt"].append(bench_swt(shape, wv, lv, dt, small, device, with_cpu=cpu))
            except Exception as exc:
                extras["swt"].append({"workload": f"SWT {wv} L{lv} {shape} {dt}", "error": repr(exc)})
    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": main["config"], "clocks": main["clocks"], "e2e": main.get("e2e"),
            "gpu_launches": main["gpu_launches"], "roofline": main["roofline"], "cpu_baseline": main.get("cpu_baseline"),
            "stage_ms": main["stage_ms"], "map": main["map"], "checked": main["checked"], "extras": extras,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
