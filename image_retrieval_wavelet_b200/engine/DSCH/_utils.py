"""DSCH's Hamming metrics on the B200 evaluator.

Drop-in for the evaluation functions of ``/root/reference/main/engine/DSCH/_utils.py`` — same names, argument order and
return types — that ``MyEval`` (:1015-1063), ``validate`` (:132-140) and ``DSCH/train.py:66`` call:

* ``mean_average_precision``  :409-450   the counting-sort evaluator (``b200_hamming_map``)
* ``calc_hamming_dist``       :453-466   ``b200_hamming_dist``
* ``pr_curve``                :469-494   stage-A histograms -> counts within every radius (``b200_hamming_radius_counts``)
* ``p_topK``                  :497-514   ranked list + running hit count (``b200_hamming_topk``, ``b200_ranked_cumhits``)
* ``get_precision_recall_by_Hamming_Radius``  :577-594   the same radius counts at ``radius``

The reference materialises ``[Q, N]`` float distance / relevance matrices row by row in Python; here a query never
leaves the kernels: one pass over the bit-packed database per metric.  Ties in the Hamming ranking are ordered by
database index (``torch.argsort`` / ``torch.sort`` leave the order of equal distances implementation-defined).  There
is no CPU fallback: a missing ``libb200ret.so`` or CUDA device raises.
"""
import numpy as np
import torch

from ... import _cabi
from .. import hamming as H


def _to_torch(x):
    return torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else torch.as_tensor(x)


def _pack_pair(qB, rB, qL, rL, binarise_labels=False):
    _cabi.require_cuda()
    qB, rB, qL, rL = (_to_torch(t) for t in (qB, rB, qL, rL))
    if qL.dim() != rL.dim() or qL.dim() not in (1, 2):
        raise NotImplementedError(f"not support: {tuple(qL.shape)}")
    if binarise_labels:
        qL, rL = (qL > 0).float(), (rL > 0).float()
    qc = H.pack_codes(qB)
    dev = qc.words.device
    rc = H.pack_codes(rB, device=dev)
    if qL.dim() == 2 and qL.shape[1] == 1:        # [N, 1] multi-hot with a single tag: overlap == both equal to 1
        qL, rL = torch.cat([qL, torch.zeros_like(qL)], 1), torch.cat([rL, torch.zeros_like(rL)], 1)
    ql, rl = H.pack_labels(qL, device=dev), H.pack_labels(rL, device=dev)
    return qc, rc, ql, rl


def mean_average_precision(qB, rB, qL, rL, topk=None):
    """_utils.py:409-450.  Returns a 0-dim tensor like the reference (``mean_AP / num_query``), or 0.0 without a hit."""
    qc, rc, ql, rl = _pack_pair(qB, rB, qL, rL)
    m, _, _ = H.hamming_map(qc, ql, rc, rl, None if topk is None else int(topk))
    return m.to(torch.float32).cpu()


def calc_hamming_dist(B1, B2):
    """_utils.py:453-466: ``0.5 * (k - B1 @ B2.T)`` as float32 ``[n1, n2]`` (a 1-D ``B1`` is one code)."""
    _cabi.require_cuda()
    B1, B2 = _to_torch(B1), _to_torch(B2)
    if B1.dim() < 2:
        B1 = B1.unsqueeze(0)
    out = H.hamming_dist(H.pack_codes(B1), H.pack_codes(B2))
    return out if B1.is_cuda else out.cpu()


def pr_curve(qB, rB, query_label, retrieval_label):
    """_utils.py:469-494: precision / recall within every Hamming radius 0..num_bit, averaged over the queries whose
    precision at that radius is positive.  Returns ``(P, R)``, float32 CPU tensors ``[num_bit + 1]``."""
    qc, rc, ql, rl = _pack_pair(qB, rB, query_label, retrieval_label)
    cum = H.radius_counts(qc, ql, rc, rl)                          # [Q, bins, 2] exact counts
    total = cum[:, :, 0].to(torch.float32)
    count = cum[:, :, 1].to(torch.float32)
    tsum = count[:, -1:]                                           # relevant rows of the whole database
    total = total + (total == 0).float() * 0.1
    has = tsum > 0                                                 # queries without a relevant row keep all-zero rows
    P = torch.where(has, count / total, torch.zeros_like(count))
    R = torch.where(has, count / torch.where(has, tsum, torch.ones_like(tsum)), torch.zeros_like(count))
    mask = (P > 0).float().sum(dim=0)
    mask = mask + (mask == 0).float() * 0.1
    return (P.sum(dim=0) / mask).cpu(), (R.sum(dim=0) / mask).cpu()


def p_topK(qB, rB, qL, rL, K=None):
    """_utils.py:497-514: precision of the K nearest database rows, K in ``[1, 100, ..., 1000]`` by default; queries
    without any relevant row add nothing but count in the mean.  Returns a float32 CPU tensor ``[len(K)]``."""
    if K is None:
        K = [1, 100, 200, 300, 400, 500, 600, 700, 800, 900, 1000]
    qc, rc, ql, rl = _pack_pair(qB, rB, qL, rL)
    n = rc.rows
    totals = [min(int(k), n) for k in K]
    if not totals or max(totals) < 1:
        return torch.zeros(len(K))
    dev = qc.words.device
    idx = H.hamming_topk(qc, rc, max(totals), raw=True)
    cum = H.ranked_cumhits(idx, ql, rl)                            # [Q, kmax] running hits along the ranking
    has = H.radius_counts(qc, ql, rc, rl)[:, -1, 1] > 0            # tsum > 0 over the WHOLE database (:504-506)
    cols = torch.tensor([max(t, 1) - 1 for t in totals], device=dev)
    hits = cum[:, cols].to(torch.float32) * has[:, None].float()
    tot = torch.tensor(totals, dtype=torch.float32, device=dev)
    p = torch.where(tot > 0, hits / tot.clamp(min=1), torch.zeros_like(hits)).sum(dim=0)
    return (p / qc.rows).cpu()


def get_precision_recall_by_Hamming_Radius(database_output, database_labels, query_output, query_labels, radius=2):
    """_utils.py:577-594 (numpy in, float out): mean over the queries of (relevant rows within ``radius``) / (rows within
    ``radius``), 0 for a query with an empty ball.  A database row matches when it shares an active label with the query
    (the reference turns the query's zeros into -1 in place and tests ``database_labels == label``; this mirror does not
    modify its arguments and accepts query labels that were already rewritten to -1/+1)."""
    qc, rc, ql, rl = _pack_pair(query_output, database_output, query_labels, database_labels, binarise_labels=True)
    r = int(np.floor(radius))
    if r < 0:
        return 0.0
    cum = H.radius_counts(qc, ql, rc, rl)[:, min(r, qc.bits), :].to(torch.float64)
    all_num, match_num = cum[:, 0], cum[:, 1]
    prec = torch.where(all_num > 0, match_num / all_num.clamp(min=1), torch.zeros_like(all_num))
    return float(prec.mean().item())
