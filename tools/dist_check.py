#!/usr/bin/env python
"""Real multi-process check of the sharded evaluator over NCCL (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
Every rank evaluates its contiguous shard in both exchange modes ("hist": two all-gathers of histograms / partials;
"lists": all-gather of per-shard top-k lists, merge, ranked AP) and compares with the unsharded evaluation of the same
problem on its own GPU: integer artefacts bit-exact, AP identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from image_retrieval_wavelet_b200.engine import hamming as H
from image_retrieval_wavelet_b200.engine.dist import ShardedHammingEvaluator, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for nq, n, bits, nlab, k in ((64, 5001, 64, 24, 300), (33, 20000, 128, 80, None), (16, 70001, 64, 24, 5000), (40, 3000, 32, 20, 3000)):
    rng = np.random.default_rng(nq + n)
    q = torch.from_numpy(rng.integers(0, 2, (nq, bits)).astype(np.float32) * 2 - 1).cuda()
    r = torch.from_numpy(rng.integers(0, 2, (n, bits)).astype(np.float32) * 2 - 1).cuda()
    r[:nq] = q
    ql = torch.from_numpy((rng.random((nq, nlab)) < 0.1).astype(np.float32)).cuda()
    rl = torch.from_numpy((rng.random((n, nlab)) < 0.1).astype(np.float32)).cuda()
    qc, qlp = H.pack_codes(q), H.pack_labels(ql)
    single = ShardedHammingEvaluator(group=dist.new_group([rank]), mode="hist")      # world-1 group: unsharded reference
    single.world, single.rank, single._dist = 1, 0, None
    m0, ap0, ts0 = single.evaluate(qc, qlp, [(H.pack_codes(r), H.pack_labels(rl), 0)], n, k)
    b0, b1 = shard_bounds(n, world)[rank]
    shard = [(H.pack_codes(r[b0:b1].contiguous()), H.pack_labels(rl[b0:b1].contiguous()), b0)]
    for mode in ("hist", "lists"):
        ev = ShardedHammingEvaluator(mode=mode)
        m, ap, ts = ev.evaluate(qc, qlp, shard, n, k)
        same = bool(torch.equal(ts.cpu(), ts0.cpu())) and float((ap - ap0).abs().max()) <= 1e-12 and abs(float(m) - float(m0)) <= 1e-12
        ok &= same
        if rank == 0:
            print(f"world {world} {mode:5s} Q={nq} N={n} B={bits} k={k}: mAP {float(m):.9f} vs {float(m0):.9f}  {'OK' if same else 'MISMATCH'}"
                  f"  ({ev.collectives} collectives)", flush=True)
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL OK" if int(flag.item()) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
