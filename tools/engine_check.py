#!/usr/bin/env python
"""HammingMapEngine on 1 GPU or under torchrun (one rank per GPU, peer-memory exchange):
    python tools/engine_check.py [--time c3]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/engine_check.py [--time c3]
Every rank compares the engine's result (graph and eager) with the unsharded evaluation of the same problem on its own
GPU: hit counts bit-exact, AP and mAP identical; --time <workload>: CUDA-event time per step of a bench.py workload."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from image_retrieval_wavelet_b200.engine import hamming as H
from image_retrieval_wavelet_b200.engine.dist import shard_bounds
from image_retrieval_wavelet_b200.engine.map_engine import HammingMapEngine

ap_ = argparse.ArgumentParser()
ap_.add_argument("--time", default=None)
ap_.add_argument("--steps", type=int, default=20)
args = ap_.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
CASES = ((64, 5001, 64, 24, 300), (33, 20000, 128, 80, None), (16, 70001, 64, 24, 5000), (40, 3000, 32, 20, 3000), (300, 40000, 64, -1, 700),
         (5, 4000, 96, 12, 50))
for nq, n, bits, nlab, k in CASES:
    rng = np.random.default_rng(nq + n)
    q = torch.from_numpy(rng.integers(0, 2, (nq, bits)).astype(np.float32) * 2 - 1).cuda()
    r = torch.from_numpy(rng.integers(0, 2, (n, bits)).astype(np.float32) * 2 - 1).cuda()
    r[:nq] = q
    if nlab > 0:
        ql = torch.from_numpy((rng.random((nq, nlab)) < 0.1).astype(np.float32)).cuda()
        rl = torch.from_numpy((rng.random((n, nlab)) < 0.1).astype(np.float32)).cuda()
    else:
        ql, rl = torch.from_numpy(rng.integers(0, 6, nq)).cuda(), torch.from_numpy(rng.integers(0, 6, n)).cuda()
    m0, ap0, ts0 = H.hamming_map(H.pack_codes(q), H.pack_labels(ql), H.pack_codes(r), H.pack_labels(rl), k)
    b0, b1 = shard_bounds(n, world)[rank]
    rs, rls = r[b0:b1].contiguous(), rl[b0:b1].contiguous()
    for graph in (True, False):
        eng = HammingMapEngine(use_graph=graph)
        for rep in range(3 if graph else 1):
            m, ap, ts = eng.evaluate(q, ql, rs, rls, k, n_total=n)
        same = bool(torch.equal(ts.cpu(), ts0.cpu())) and float((ap - ap0).abs().max()) <= 1e-12 and abs(m - float(m0)) <= 1e-12
        ok &= same
        if rank == 0:
            print(f"world {world} graph={int(graph)} Q={nq} N={n} B={bits} k={k}: mAP {m:.9f} vs {float(m0):.9f} {'OK' if same else 'MISMATCH'} {eng.last_info}",
                  flush=True)
        eng.close()
if args.time:
    import bench

    q, ql, r, rl, k = bench.make_problem(args.time)
    n = r.shape[0]
    b0, b1 = shard_bounds(n, world)[rank]
    dq, dql, dr, drl = q.cuda(), ql.cuda(), r[b0:b1].contiguous().cuda(), rl[b0:b1].contiguous().cuda()
    eng = HammingMapEngine()
    flush = bench.l2_flusher(torch.device("cuda", local))
    for _ in range(5):
        m, _, _ = eng.evaluate(dq, dql, dr, drl, k, n_total=n)
    times = []
    for _ in range(args.steps):
        flush()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        m, _, _ = eng.evaluate(dq, dql, dr, drl, k, n_total=n)
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    t = torch.tensor([sum(times) / len(times), min(times)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"TIME {args.time} world {world}: mean {t[0].item():.4f} ms min {t[1].item():.4f} ms per step, mAP {m:.9f} info {eng.last_info}", flush=True)
    eng.close()
if world > 1:
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(int(flag.item()))
if rank == 0:
    print("ALL OK" if ok else "FAILED", flush=True)
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
