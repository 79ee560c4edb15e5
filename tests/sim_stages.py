"""Test support: the stage backend of ``ShardedHammingEvaluator`` on the CPU simulator (``libb200ret_sim.so``).

Same stage programs and planner as the CUDA library, host memory, no stream — so the collective choreography of
``image_retrieval_wavelet_b200.engine.dist`` can be exercised under gloo without a GPU.
"""
import ctypes

import torch

from image_retrieval_wavelet_b200 import _cabi
from simlib import load_sim


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class SimStages:
    device = torch.device("cpu")

    def __init__(self, num_sms=4):
        self.lib = load_sim()
        self.num_sms = num_sms

    def plan_init(self, q, n, n_total, bits, lw, mode, k):
        plan = _cabi.MapPlan()
        rc = self.lib.sim_map_plan_init(ctypes.byref(plan), q, ctypes.c_longlong(n), ctypes.c_longlong(n_total), bits, lw, mode,
                                        ctypes.c_longlong(k), self.num_sms)
        assert rc == 0, rc
        return plan

    def hist(self, plan, qc, ql, dc, dl, ws):
        assert self.lib.sim_hamming_hist(ctypes.byref(plan), _p(qc), _p(ql), _p(dc), _p(dl), _p(ws)) == 0

    def scan(self, plan, ws, ext, n_shards, shard):
        assert self.lib.sim_hamming_scan(ctypes.byref(plan), _p(ws), _p(ext), n_shards, shard) == 0

    def ap(self, plan, qc, ql, dc, dl, ws, rank_idx, rank_dist, index_base):
        assert self.lib.sim_hamming_ap(ctypes.byref(plan), _p(qc), _p(ql), _p(dc), _p(dl), _p(ws), _p(rank_idx), _p(rank_dist),
                                       ctypes.c_longlong(index_base)) == 0

    def ap_reduce(self, plan, ws, sum_q, hits_q):
        assert self.lib.sim_ap_reduce(ctypes.byref(plan), _p(ws), _p(sum_q), _p(hits_q)) == 0

    def ap_finalize(self, sums, hits, n_parts, stride, q, ap, tsum, m):
        assert self.lib.sim_ap_finalize(_p(sums), _p(hits), n_parts, ctypes.c_longlong(stride), q, _p(ap), _p(tsum), _p(m)) == 0
