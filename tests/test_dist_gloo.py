"""The sharded evaluator's collective choreography at world size 2 over gloo on CPU.

Each rank owns one contiguous shard of the database (plus, in the second test, two shards per rank), runs the real stage
programs through the CPU simulator backend and exchanges shard totals / per-query partials with torch.distributed —
the same code path ``bench.py --gpus N`` takes over NCCL with the CUDA stages."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import eval_ref
from simlib import multi_hot, pack_bits, pack_labels_np, pm1, words


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem(seed, nq, n, bits, nlab):
    rng = np.random.default_rng(seed)
    q, r = pm1(rng, nq, bits), pm1(rng, n, bits)
    r[:nq] = q
    r[:nq, :2] *= -1
    ql, rl = multi_hot(rng, nq, nlab, 0.12), multi_hot(rng, n, nlab, 0.12)
    return q, ql, r, rl


def _packed(codes, labels, bits):
    from image_retrieval_wavelet_b200.engine.hamming import PackedCodes, PackedLabels

    cw = pack_bits(codes, words(bits))
    lw, nlw, mode = pack_labels_np(labels)
    return (PackedCodes(torch.from_numpy(cw.view(np.int64)), codes.shape[0], bits),
            PackedLabels(torch.from_numpy(lw.view(np.int64)), labels.shape[0], nlw, mode))


def _worker(rank, world, port, per_rank, seed, nq, n, bits, nlab, topk, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from image_retrieval_wavelet_b200.engine.dist import ShardedHammingEvaluator, shard_bounds
    from sim_stages import SimStages

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        q, ql, r, rl = _problem(seed, nq, n, bits, nlab)
        qc, qlp = _packed(q, ql, bits)
        bounds = shard_bounds(n, world * per_rank)
        shards = []
        for i in range(per_rank):
            b, e = bounds[rank * per_rank + i]
            dc, dl = _packed(r[b:e], rl[b:e], bits)
            shards.append((dc, dl, b))
        ev = ShardedHammingEvaluator(stages=SimStages(num_sms=2 + rank))      # ranks may even plan differently
        m, ap, tsum = ev.evaluate(qc, qlp, shards, n, topk)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), m=m.numpy(), ap=ap.numpy(), tsum=tsum.numpy(),
                 collectives=ev.collectives)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("per_rank,topk,n", [(1, 300, 3001), (2, None, 2000), (1, 70000, 70001)])
def test_sharded_hist_exchange_world2(tmp_path, per_rank, topk, n):
    nq, bits, nlab, seed = 21, 64, 24, 5
    port = _free_port()
    mp.spawn(_worker, args=(2, port, per_rank, seed, nq, n, bits, nlab, topk, str(tmp_path)), nprocs=2, join=True)
    q, ql, r, rl = _problem(seed, nq, n, bits, nlab)
    m0, ap0, ts0, _, _ = eval_ref.maphashing_exact(q, ql, r, rl, topk, return_details=True)
    results = [np.load(tmp_path / f"rank{i}.npz") for i in range(2)]
    for res in results:
        assert np.array_equal(res["tsum"].astype(np.int64), ts0)
        assert np.abs(res["ap"] - ap0).max() <= 1e-6 and abs(float(res["m"]) - m0) <= 1e-6
        assert int(res["collectives"]) == 2                 # shard totals + per-query partials, nothing else
    assert np.array_equal(results[0]["ap"], results[1]["ap"])   # every rank ends with the identical answer


def test_shard_bounds_cover_the_database():
    from image_retrieval_wavelet_b200.engine.dist import shard_bounds

    for n, s in [(0, 4), (1, 8), (10, 3), (117000, 8), (1000000, 8), (7, 2)]:
        b = shard_bounds(n, s)
        assert len(b) == s and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(s - 1))
        assert all(lo % 2 == 0 for lo, _ in b if lo < n)      # even shard starts: 16-byte tile alignment
