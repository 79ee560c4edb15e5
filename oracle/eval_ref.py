"""CPU oracle for HP-EVAL (test infrastructure; see ``oracle/__init__.py``).

Restates the retrieval evaluator of the reference:

* ``label_comparison_fn``     /root/reference/main/engine/accuracy_calculator.py:31-37
* ``calc_hamming_dist``       accuracy_calculator.py:183-186
* ``per_bit_balance`` & co    accuracy_calculator.py:188-200
* ``calculate_maphashing``    accuracy_calculator.py:203-231
* ``calculate_map``           accuracy_calculator.py:156-167 (torchmetrics RetrievalMAP)
* ``get_knn_torch``           /root/reference/main/engine/get_knn.py:9-24,60-71
* PML helpers it calls (``get_label_match_counts``, ``determine_k``) restated from
  pytorch-metric-learning's published source (not vendored in the reference).

Two flavours of the mAP loop:

``maphashing_exact``    integer distances, ranking by (distance, index) — i.e. a
                        *stable* sort — integer hit positions, AP in float64.
                        This is the parity target of the CUDA path.
``maphashing_literal``  the reference's op sequence in torch (fp32 matmul,
                        ``torch.argsort`` with its default ``stable=False``,
                        fp32 mean, Python-float outer sum).  This is what the
                        reference prints and what the CPU baseline times.

Pinned against outputs of the real reference code in ``tests/golden/`` (see
``tests/golden/make_golden_eval.py``).
"""
import numpy as np


# --------------------------------------------------------------------------- labels
def label_rel_ref(query_labels, reference_labels):
    """accuracy_calculator.py:31-37.  2-D x 2-D: share >= 1 active tag; else equality."""
    ql = np.asarray(query_labels)
    rl = np.asarray(reference_labels)
    if ql.ndim > 1 and rl.ndim > 1:
        if ql.ndim == 2 and rl.ndim == 2:
            return (ql.astype(np.float64) @ rl.astype(np.float64).T) > 0
        return (ql.astype(np.float64) * rl.astype(np.float64)).sum(-1) > 0
    return ql[:, None] == rl


def label_match_counts_ref(query_labels, reference_labels):
    """PML ``get_label_match_counts`` with the reference's comparison function:
    (unique query label rows, number of references each one matches)."""
    ql = np.asarray(query_labels)
    uniq = np.unique(ql, axis=0)
    counts = np.array([int(label_rel_ref(uniq[i:i + 1], reference_labels).sum()) for i in range(len(uniq))],
                      dtype=np.int64)
    return uniq, counts


def resolve_topk_ref(topk, reference_labels, ref_includes_query=False):
    """accuracy_calculator.py:204-212."""
    while isinstance(topk, (tuple, list)):
        topk = topk[0] if len(topk) else None
    if isinstance(topk, str) and topk == "max_bin_count":
        _, counts = label_match_counts_ref(reference_labels, reference_labels)
        topk = int(counts.max()) - int(ref_includes_query)
    if topk is not None:
        topk = int(topk)
    return topk


# --------------------------------------------------------------------------- hamming
def hamming_ref(q_codes, r_codes):
    """accuracy_calculator.py:183-186 for +-1 codes, as exact integers [Q, N]."""
    q = np.asarray(q_codes)
    r = np.asarray(r_codes)
    if not (np.all(np.abs(q) == 1) and np.all(np.abs(r) == 1)):
        raise ValueError("hamming_ref needs codes in {-1, +1}")
    b = q.shape[1]
    dot = q.astype(np.int64) @ r.astype(np.int64).T
    return (b - dot) // 2


def pack_codes_ref(codes):
    """+-1 codes [N, B] -> uint64 words [N, ceil(B/64)], bit b of word b//64 set iff code > 0."""
    c = np.asarray(codes)
    n, b = c.shape
    words = (b + 63) // 64
    out = np.zeros((n, words), dtype=np.uint64)
    for j in range(b):
        out[:, j // 64] |= (c[:, j] > 0).astype(np.uint64) << np.uint64(j % 64)
    return out


def popcount_hamming_ref(q_packed, r_packed):
    x = q_packed[:, None, :] ^ r_packed[None, :, :]
    bits = np.unpackbits(x.view(np.uint8), axis=-1)
    return bits.sum(-1).astype(np.int64)


def bit_balance_ref(reference):
    """accuracy_calculator.py:188-200 -> (per-bit balance, mean, min) in float64."""
    frac = (np.asarray(reference) > 0).astype(np.float64).mean(axis=0)
    per_bit = 1.0 - 2.0 * np.abs(frac - 0.5)
    return per_bit, float(per_bit.mean()), float(per_bit.min())


# --------------------------------------------------------------------------- mAP
def ap_from_ranked_relevance(rel_topk):
    """AP of one ranked 0/1 list: mean over hits of (hit ordinal / rank); (ap, tsum)."""
    rel_topk = np.asarray(rel_topk).astype(bool)
    pos = np.flatnonzero(rel_topk)
    tsum = int(pos.shape[0])
    if tsum == 0:
        return 0.0, 0
    ordinal = np.arange(1, tsum + 1, dtype=np.float64)
    return float(np.mean(ordinal / (pos.astype(np.float64) + 1.0))), tsum


def maphashing_exact(query, query_labels, reference, reference_labels, topk=None, ref_includes_query=False,
                     return_details=False):
    """Parity target.  Ranking = ascending (integer Hamming distance, reference index)."""
    topk = resolve_topk_ref(topk, reference_labels, ref_includes_query)
    q = np.asarray(query)
    r = np.asarray(reference)
    nq, n = q.shape[0], r.shape[0]
    k = n if topk is None else max(0, min(int(topk), n))
    aps = np.zeros(nq, dtype=np.float64)
    tsums = np.zeros(nq, dtype=np.int64)
    ranked = np.zeros((nq, k), dtype=np.int64) if return_details else None
    dists = np.zeros((nq, k), dtype=np.int64) if return_details else None
    ql = np.asarray(query_labels)
    rl = np.asarray(reference_labels)
    for i in range(nq):
        rel = label_rel_ref(ql[i:i + 1], rl).reshape(-1)
        d = hamming_ref(q[i:i + 1], r).reshape(-1)
        order = np.argsort(d, kind="stable")[:k]
        aps[i], tsums[i] = ap_from_ranked_relevance(rel[order])
        if return_details:
            ranked[i] = order
            dists[i] = d[order]
    m = float(aps.sum() / nq) if nq else float("nan")
    if return_details:
        return m, aps, tsums, ranked, dists
    return m


def maphashing_literal(query, query_labels, reference, reference_labels, topk=None, ref_includes_query=False):
    """The reference's per-query op sequence in torch (accuracy_calculator.py:203-231),
    including the implementation-defined tie order of ``torch.argsort(stable=False)``."""
    import torch

    topk = resolve_topk_ref(topk, np.asarray(reference_labels), ref_includes_query)
    query = torch.as_tensor(query)
    reference = torch.as_tensor(reference)
    ql_all = torch.as_tensor(query_labels)
    rl_all = torch.as_tensor(reference_labels)
    nbits = query.shape[1]
    two_d = ql_all.ndim > 1 and rl_all.ndim > 1
    ref_t = reference.t()
    rl_t = rl_all.t().float() if two_d else None
    total = 0.0
    for i in range(query.shape[0]):
        if two_d:
            gnd = (torch.matmul(ql_all[i:i + 1].float(), rl_t) > 0).float().squeeze()
        else:
            gnd = (ql_all[i:i + 1].unsqueeze(1) == rl_all).float().squeeze()
        hamm = (0.5 * (nbits - torch.matmul(query[i:i + 1], ref_t))).squeeze()
        order = torch.argsort(hamm)
        tgnd = gnd[order][0:topk]
        tsum = torch.sum(tgnd).int().item()
        if tsum > 0:
            tindex = torch.where(tgnd == 1)[0].float() + 1.0
            count = torch.arange(1, tsum + 1).float()
            total += torch.mean(count / tindex).item()
    return total / query.shape[0]


# --------------------------------------------------------------------------- knn + map
def knn_ref(references, queries, num_k, same_source=False, distance_metric="cosine"):
    """get_knn.py:9-24 + get_knn_torch :60-71 in float64 with index tie-break.
    Returns (indices [Q,k] int64, distances [Q,k] float64)."""
    r = np.asarray(references, dtype=np.float64)
    q = np.asarray(queries, dtype=np.float64)
    num_k = int(num_k) + int(bool(same_source))
    if distance_metric in ("hamming", "cosine"):
        score = q @ r.T
        order = np.argsort(-score, axis=1, kind="stable")[:, :num_k]
    else:
        score = np.sqrt(np.maximum(((q[:, None, :] - r[None, :, :]) ** 2).sum(-1), 0.0))
        order = np.argsort(score, axis=1, kind="stable")[:, :num_k]
    dist = np.take_along_axis(score, order, axis=1)
    if same_source:
        return order[:, 1:], dist[:, 1:]
    return order, dist


def retrieval_map_ref(query_labels, knn_labels, not_lone_query_mask=None):
    """accuracy_calculator.py:156-167: torchmetrics ``RetrievalMAP`` over the knn
    list in its given (similarity-descending) order; queries without a hit score 0
    (``empty_target_action='neg'``); mean over the not-lone queries."""
    ql = np.asarray(query_labels)
    kl = np.asarray(knn_labels)
    nq = ql.shape[0]
    mask = np.ones(nq, dtype=bool) if not_lone_query_mask is None else np.asarray(not_lone_query_mask, dtype=bool)
    aps = []
    for i in range(nq):
        if not mask[i]:
            continue
        if ql.ndim > 1:
            rel = (kl[i].astype(np.float64) * ql[i][None, :].astype(np.float64)).sum(-1) > 0
        else:
            rel = kl[i] == ql[i]
        aps.append(ap_from_ranked_relevance(rel)[0])
    return float(np.mean(aps)) if aps else 0.0


def pr_rc_hashing_ref(query, query_labels, reference, reference_labels, not_lone_query_mask=None):
    """accuracy_calculator.py:235-273 with the tie order fixed to (distance, index): mean precision / recall at every
    rank over the queries that are not lone and have a relevant row.  Returns (precision [N], recall [N], n_queries)."""
    q = np.asarray(query, dtype=np.float64)
    r = np.asarray(reference, dtype=np.float64)
    n = r.shape[0]
    rel_all = label_rel_ref(query_labels, reference_labels)
    keep = np.ones(q.shape[0], bool) if not_lone_query_mask is None else np.asarray(not_lone_query_mask, bool)
    prec, rec, used = np.zeros(n), np.zeros(n), 0
    for i in range(q.shape[0]):
        hamm = 0.5 * (q.shape[1] - q[i] @ r.T)
        order = np.argsort(hamm, kind="stable")
        gnd = rel_all[i][order].astype(np.float64)
        total = gnd.sum()
        if total > 0 and keep[i]:
            cum = np.cumsum(gnd)
            prec += cum / np.arange(1, n + 1)
            rec += cum / total
            used += 1
    if used:
        prec /= used
        rec /= used
    return prec, rec, used


# --------------------------------------------------------------------------- DSCH metrics
def dsch_pr_curve_ref(qB, rB, query_label, retrieval_label):
    """/root/reference/main/engine/DSCH/_utils.py:469-494 on exact integer distances: per query and radius r the
    counts within the ball (total, relevant), p = count / total (0.1 for an empty ball), r = count / tsum, queries
    without a relevant row contribute zero rows; columns are averaged over the queries with P > 0 (0.1 if none).
    float32 arithmetic like the reference.  Returns (P, R) float32 [B+1]."""
    d = hamming_ref(qB, rB)
    rel = label_rel_ref(query_label, retrieval_label)
    nq, nbit = d.shape[0], np.asarray(qB).shape[1]
    P = np.zeros((nq, nbit + 1), dtype=np.float32)
    R = np.zeros((nq, nbit + 1), dtype=np.float32)
    radii = np.arange(nbit + 1)
    for i in range(nq):
        tsum = np.float32(rel[i].sum())
        if tsum == 0:
            continue
        within = d[i][None, :] <= radii[:, None]
        total = within.sum(-1).astype(np.float32)
        total = total + (total == 0).astype(np.float32) * np.float32(0.1)
        count = (within & rel[i][None, :]).sum(-1).astype(np.float32)
        P[i] = count / total
        R[i] = count / tsum
    mask = (P > 0).astype(np.float32).sum(0)
    mask = mask + (mask == 0).astype(np.float32) * np.float32(0.1)
    return P.sum(0, dtype=np.float32) / mask, R.sum(0, dtype=np.float32) / mask


def dsch_p_topk_ref(qB, rB, qL, rL, K=None):
    """_utils.py:497-514 with the tie order fixed to (distance, index): mean over ALL queries of the precision of the
    min(K, N) nearest rows; queries without a relevant row in the whole database add 0.  float64 [len(K)]."""
    if K is None:
        K = [1, 100, 200, 300, 400, 500, 600, 700, 800, 900, 1000]
    d = hamming_ref(qB, rB)
    rel = label_rel_ref(qL, rL)
    nq, n = d.shape
    p = np.zeros(len(K), dtype=np.float64)
    for i in range(nq):
        if rel[i].sum() == 0:
            continue
        order = np.argsort(d[i], kind="stable")
        gnd = rel[i][order]
        for j, k in enumerate(K):
            total = min(int(k), n)
            p[j] += gnd[:total].sum() / total
    return p / nq


def dsch_radius_precision_ref(database_output, database_labels, query_output, query_labels, radius=2):
    """_utils.py:577-594: mean over queries of (#rows within `radius` sharing an active label) / (#rows within
    `radius`), 0 for an empty ball.  Does not modify its arguments (the reference rewrites the query labels' zeros
    to -1 in place)."""
    d = hamming_ref(query_output, database_output)
    rel = label_rel_ref((np.asarray(query_labels) > 0).astype(np.float64), (np.asarray(database_labels) > 0).astype(np.float64))
    prec = []
    for i in range(d.shape[0]):
        ball = d[i] <= radius
        all_num = int(ball.sum())
        prec.append(float((ball & rel[i]).sum()) / all_num if all_num else 0.0)
    return float(np.mean(np.array(prec)))
