"""Hamming mAP@k over a database sharded across GPUs (one process per GPU, ``torch.distributed``).

The reference's only multi-GPU retrieval is faiss' ``index_cpu_to_all_gpus(shards=True)``
(``/root/reference/main/engine/get_knn.py:41-44``): the database is split over the GPUs, every GPU sees all queries and
the per-shard results are merged on the host.  Here the database rows (codes + labels) are split into contiguous
global-index ranges — so the (distance, index) tie-break is (distance, shard, local index) — the packed queries are
replicated, and the merge happens on the devices over NCCL.  Two exact exchange forms:

``mode="hist"``  (default) — the counting-sort evaluator only needs, per query and distance, how many rows (and how
    many relevant rows) every *earlier* shard holds.  One all-gather of the shard totals ``[bins, Qpad]`` (uint32 pairs)
    between stage A and stage S, one all-gather of the per-query partial sums after stage B.  No ranked list crosses
    NVLink.
``mode="lists"`` — the mandated "all-gather per-shard top-k lists, then merge" form: every shard materialises its local
    top-k (distance uint16, global index uint32), the lists are all-gathered, merged by ``b200_merge_topk`` and scored
    by the ranked-AP kernel against the (all-gathered, tiny) packed label table.

Both give bit-identical integer artefacts; ``tests/test_dist_gloo.py`` runs the ``hist`` choreography at world size 2
over gloo on CPU with the stage programs executed by the test-only simulator, and the ``-m gpu`` tests emulate several
shards on one device (several shards per process are supported for exactly that purpose).
"""
import ctypes

import torch

from .. import _cabi
from . import hamming as H


def shard_bounds(n_total, n_shards):
    """Contiguous, even-sized global index ranges ``[(begin, end), ...]`` (the last ones may be short or empty)."""
    per = (n_total + n_shards - 1) // n_shards
    per = (per + 1) // 2 * 2
    return [(min(n_total, per * r), min(n_total, per * (r + 1))) for r in range(n_shards)]


class DeviceStages:
    """The C-ABI stage functions on the current CUDA device and stream."""

    def __init__(self):
        self.lib = _cabi.load()
        _cabi.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _s(self):
        return (_cabi.stream_ptr(),)

    def plan_init(self, q, n, n_total, bits, lw, mode, k):
        plan = _cabi.MapPlan()
        _cabi.check(self.lib.b200_map_plan_init(ctypes.byref(plan), q, n, n_total, bits, lw, mode, k), "b200_map_plan_init")
        return plan

    def hist(self, plan, qc, ql, dc, dl, ws):
        _cabi.check(self.lib.b200_hamming_hist(ctypes.byref(plan), _cabi.ptr(qc), _cabi.ptr(ql), _cabi.ptr(dc), _cabi.ptr(dl),
                                               _cabi.ptr(ws), *self._s()), "b200_hamming_hist")

    def scan(self, plan, ws, ext, n_shards, shard):
        _cabi.check(self.lib.b200_hamming_scan(ctypes.byref(plan), _cabi.ptr(ws), _cabi.ptr(ext), n_shards, shard, *self._s()),
                    "b200_hamming_scan")

    def ap(self, plan, qc, ql, dc, dl, ws, rank_idx, rank_dist, index_base):
        _cabi.check(self.lib.b200_hamming_ap(ctypes.byref(plan), _cabi.ptr(qc), _cabi.ptr(ql), _cabi.ptr(dc), _cabi.ptr(dl),
                                             _cabi.ptr(ws), _cabi.ptr(rank_idx), _cabi.ptr(rank_dist), index_base, *self._s()),
                    "b200_hamming_ap")

    def ap_reduce(self, plan, ws, sum_q, hits_q):
        _cabi.check(self.lib.b200_ap_reduce(ctypes.byref(plan), _cabi.ptr(ws), _cabi.ptr(sum_q), _cabi.ptr(hits_q), *self._s()),
                    "b200_ap_reduce")

    def ap_finalize(self, sums, hits, n_parts, stride, q, ap, tsum, m):
        _cabi.check(self.lib.b200_ap_finalize(_cabi.ptr(sums), _cabi.ptr(hits), n_parts, stride, q, _cabi.ptr(ap), _cabi.ptr(tsum),
                                              _cabi.ptr(m), *self._s()), "b200_ap_finalize")

    def topk(self, plan, qc, dc, ws, idx, dist):
        _cabi.check(self.lib.b200_hamming_topk(ctypes.byref(plan), _cabi.ptr(qc), _cabi.ptr(dc), _cabi.ptr(ws), _cabi.ptr(idx),
                                               _cabi.ptr(dist), *self._s()), "b200_hamming_topk")

    def merge_topk(self, in_idx, in_dist, n_shards, q, k, bits, out_idx, out_dist):
        _cabi.check(self.lib.b200_merge_topk(_cabi.ptr(in_idx), _cabi.ptr(in_dist), n_shards, q, k, bits, _cabi.ptr(out_idx),
                                             _cabi.ptr(out_dist), *self._s()), "b200_merge_topk")

    def ranked_ap_u32(self, idx, q, k, ql, dl, lw, mode, ap, hits, m):
        _cabi.check(self.lib.b200_ranked_ap(_cabi.ptr(idx), 0, q, k, _cabi.ptr(ql), _cabi.ptr(dl), lw, mode, None, _cabi.ptr(ap),
                                            _cabi.ptr(hits), _cabi.ptr(m), *self._s()), "b200_ranked_ap")


class ShardedHammingEvaluator:
    """``calculate_maphashing`` for a database split into contiguous shards, one (or, for emulation, several) per process.

    ``group``: a ``torch.distributed`` process group (``None`` = default group when initialised, else single process).
    ``stages``: stage backend; defaults to :class:`DeviceStages`.
    """

    def __init__(self, group=None, stages=None, mode="hist"):
        if mode not in ("hist", "lists"):
            raise ValueError("mode must be 'hist' or 'lists'")
        self.mode = mode
        self.stages = stages if stages is not None else DeviceStages()
        self.group = group
        import torch.distributed as dist

        self._dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self._dist.get_world_size(group) if self._dist else 1
        self.rank = self._dist.get_rank(group) if self._dist else 0
        self.collectives = 0
        self.timeline = None          # set to [] to record (stage name, CUDA event) marks on the current stream

    def _mark(self, name):
        if self.timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timeline.append((name, ev))

    # ------------------------------------------------------------------ collectives
    def _all_gather(self, local):
        """``local``: list (one per local shard) of equal-shape tensors -> ``[total_shards, ...]`` in shard order."""
        mine = torch.stack(local, dim=0).contiguous()
        if self.world == 1:
            return mine
        flat = mine.view(torch.uint8).reshape(-1) if mine.dtype != torch.uint8 else mine.reshape(-1)
        full = torch.empty(self.world * flat.numel(), dtype=torch.uint8, device=flat.device)
        try:                                           # one ncclAllGather straight into the result
            self._dist.all_gather_into_tensor(full, flat, group=self.group)
        except (RuntimeError, NotImplementedError, AttributeError):      # backends without the tensor form
            out = [torch.empty_like(flat) for _ in range(self.world)]
            self._dist.all_gather(out, flat, group=self.group)
            full = torch.cat(out, dim=0)
        self.collectives += 1
        return full.view(mine.dtype).reshape((self.world * mine.shape[0],) + tuple(mine.shape[1:]))

    # ------------------------------------------------------------------ evaluation
    def evaluate(self, qc, ql, shards, n_total, topk=None):
        """``qc``/``ql``: packed queries (replicated).  ``shards``: list of ``(PackedCodes, PackedLabels, index_base)``
        owned by this process, in ascending index order.  Returns ``(map, ap[Q], tsum[Q])`` (identical on all ranks)."""
        if not shards:
            raise ValueError("every process needs at least one (possibly empty) shard")
        for dc, dl, _ in shards:
            H._check_pair(qc, ql, dc, dl)
        k = int(n_total if topk is None else min(int(topk), n_total))
        if k < 1:
            raise ValueError("topk must be >= 1 and the database non-empty")
        if self.mode == "hist":
            return self._evaluate_hist(qc, ql, shards, n_total, k)
        return self._evaluate_lists(qc, ql, shards, n_total, k)

    def _new(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=qc_device(self.stages))

    def _evaluate_hist(self, qc, ql, shards, n_total, k):
        st = self.stages
        q = qc.rows
        per = len(shards)
        plans, wss = [], []
        for slot, (dc, dl, _) in enumerate(shards):
            plan = st.plan_init(q, dc.rows, n_total, qc.bits, ql.lw, ql.mode, k)
            ws = self._workspace(plan, slot)
            self._mark("begin")
            st.hist(plan, qc.words, ql.words, dc.words, dl.words, ws)
            self._mark("hist")
            plans.append(plan), wss.append(ws)
        tot_items = plans[0].bins * plans[0].Qpad * 2
        tots = [ws[p.off_tot:p.off_tot + 4 * tot_items].view(torch.int32) for p, ws in zip(plans, wss)]
        ext = self._all_gather(tots)                                   # [R, bins * Qpad * 2] int32 (uint32 pairs)
        self._mark("gather_totals")
        n_sh = int(ext.shape[0])
        parts = []
        for i, ((dc, dl, base), plan, ws) in enumerate(zip(shards, plans, wss)):
            st.scan(plan, ws, ext if n_sh > 1 else None, n_sh, self.rank * per + i)
            self._mark("scan")
            st.ap(plan, qc.words, ql.words, dc.words, dl.words, ws, None, None, int(base))
            self._mark("ap")
            part = self._new((2, q), torch.int64)                      # row 0: fixed-point sums; row 1: hits (uint32) in the first half
            hits = part[1].view(torch.int32)[:q]
            st.ap_reduce(plan, ws, part[0], hits)
            parts.append(part)
        allp = self._all_gather(parts)                                 # [R, 2, Q] int64
        self._mark("gather_partials")
        sums = allp[:, 0, :].contiguous()
        hits = torch.stack([allp[r, 1].view(torch.int32)[:q] for r in range(n_sh)], dim=0).contiguous()
        ap = self._new((q,), torch.float64)
        tsum = self._new((q,), torch.int32)
        m = self._new((), torch.float64)
        st.ap_finalize(sums, hits, n_sh, q, q, ap, tsum, m)
        self._mark("finalize")
        return m, ap, tsum

    def _workspace(self, plan, slot):
        """Scratch is cached per (local shard, size): repeated evaluations (multi-k, benchmarking) do not re-allocate."""
        key = (slot, int(plan.workspace_bytes))
        cache = self.__dict__.setdefault("_ws_cache", {})
        buf = cache.get(key)
        if buf is None or len(cache) > 16:
            if len(cache) > 16:
                cache.clear()
            buf = cache[key] = self._new((key[1],), torch.uint8)
        return buf

    def _evaluate_lists(self, qc, ql, shards, n_total, k):
        st = self.stages
        q = qc.rows
        idxs, dists, labs = [], [], []
        bounds = shard_bounds(n_total, self.world * len(shards))       # the partition every process must follow
        per_rows = max(bounds[0][1] - bounds[0][0], 2)
        for dc, dl, base in shards:
            idx = torch.full((q, k), -1, dtype=torch.int32, device=qc.words.device)       # 0xFFFFFFFF padding
            dist = torch.full((q, k), -1, dtype=torch.int16, device=qc.words.device)      # 0xFFFF padding
            kl = min(k, dc.rows)
            if kl > 0:
                plan = st.plan_init(q, dc.rows, dc.rows, qc.bits, 1, _cabi.LABELS_EQUAL, kl)
                ws = self._new((plan.workspace_bytes,), torch.uint8)
                li = torch.empty((q, kl), dtype=torch.int32, device=qc.words.device)
                ld = torch.empty((q, kl), dtype=torch.int16, device=qc.words.device)
                st.topk(plan, qc.words, dc.words, ws, li, ld)
                idx[:, :kl] = li + int(base)
                dist[:, :kl] = ld
            idxs.append(idx), dists.append(dist)
            lab = torch.zeros((per_rows, ql.lw), dtype=torch.int64, device=qc.words.device)
            if dl.words.shape[0] > per_rows:
                raise ValueError("lists mode needs the shard_bounds() partition of the database")
            lab[:dl.words.shape[0]] = dl.words
            labs.append(lab)
        # shards are even-sized (shard_bounds), so concatenating the gathered label blocks restores global row order
        all_idx = self._all_gather(idxs)
        all_dist = self._all_gather(dists)
        all_lab = self._all_gather(labs)
        n_sh = int(all_idx.shape[0])
        table = torch.zeros(((n_total + 1) // 2 * 2 + per_rows, ql.lw), dtype=torch.int64, device=qc.words.device)
        for r in range(n_sh):
            b = bounds[r][0]
            table[b:b + per_rows] = all_lab[r]
        out_idx = torch.empty((q, k), dtype=torch.int32, device=qc.words.device)
        out_dist = torch.empty((q, k), dtype=torch.int16, device=qc.words.device)
        st.merge_topk(all_idx, all_dist, n_sh, q, k, qc.bits, out_idx, out_dist)
        ap = self._new((q,), torch.float64)
        tsum = self._new((q,), torch.int32)
        m = self._new((), torch.float64)
        st.ranked_ap_u32(out_idx, q, k, ql.words, table, ql.lw, ql.mode, ap, tsum, m)
        self.last_ranked = (out_idx, out_dist)
        return m, ap, tsum


def qc_device(stages):
    return getattr(stages, "device", torch.device("cpu"))
