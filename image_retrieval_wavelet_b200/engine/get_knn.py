"""Drop-in mirror of ``/root/reference/main/engine/get_knn.py`` on the B200 k-NN kernels.

``get_knn(references, queries, num_k, embeddings_come_from_same_source, with_faiss=True, distance_metric="l2")``
keeps the reference's signature and return order ``(indices, distances)``; ``with_faiss`` is accepted and ignored
(there is one implementation).  "hamming"/"cosine" rank by inner product, largest first; anything else by L2 distance,
smallest first — exactly the branches of ``get_knn_torch`` (:60-71) / ``get_knn_faiss`` (:27-57).  Ties go to the smaller
index (torch.topk / faiss leave the tie order unspecified).
"""
import logging

import torch

from .. import _cabi

LOGGER = logging.getLogger("RETRIEVAL")
MAX_K = 4096


def knn_topk(references, queries, num_k, distance_metric="l2"):
    """``(distances float32 [Q, k], indices int64 [Q, k])`` on the device, like ``get_knn_torch``."""
    _cabi.require_cuda()
    refs = torch.as_tensor(references)
    qs = torch.as_tensor(queries)
    dev = refs.device if refs.is_cuda else (qs.device if qs.is_cuda else torch.device("cuda"))
    refs = refs.to(device=dev, dtype=torch.float32).contiguous()
    qs = qs.to(device=dev, dtype=torch.float32).contiguous()
    if refs.dim() != 2 or qs.dim() != 2 or refs.shape[1] != qs.shape[1]:
        raise ValueError("references [N, D] and queries [Q, D] must share D")
    n, d = int(refs.shape[0]), int(refs.shape[1])
    q = int(qs.shape[0])
    k = int(num_k)
    if k < 1 or k > n:
        # torch.topk: "selected index k out of range"
        raise RuntimeError(f"selected index k out of range (k={k}, references={n})")
    if k > MAX_K:
        raise NotImplementedError(f"k-NN lists longer than {MAX_K} are not supported (faiss-gpu 1.6.5 stops at 2048)")
    if d % 4:                                   # zero columns do not change inner products or distances
        pad = 4 - d % 4
        refs = torch.nn.functional.pad(refs, (0, pad))
        qs = torch.nn.functional.pad(qs, (0, pad))
        d += pad
    idx = torch.empty((q, k), dtype=torch.int64, device=dev)
    score = torch.empty((q, k), dtype=torch.float32, device=dev)
    if q == 0:
        return score, idx
    lib = _cabi.load()
    ws_bytes = lib.b200_knn_workspace_bytes(q, n, d, k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    l2 = 0 if distance_metric in ("hamming", "cosine") else 1
    with torch.cuda.device(dev):
        rc = lib.b200_knn_topk(_cabi.ptr(refs), _cabi.ptr(qs), q, n, d, k, l2, _cabi.ptr(idx), _cabi.ptr(score), _cabi.ptr(ws),
                               ws_bytes, _cabi.stream_ptr())
    _cabi.check(rc, "b200_knn_topk")
    return score, idx


def get_knn_torch(references, queries, num_k, distance_metric="l2"):
    return knn_topk(references, queries, num_k, distance_metric)


def get_knn_faiss(references, queries, num_k, distance_metric="l2"):
    return knn_topk(references, queries, num_k, distance_metric)


def get_knn(references, queries, num_k, embeddings_come_from_same_source, with_faiss=True, distance_metric="l2"):
    num_k += embeddings_come_from_same_source
    LOGGER.info("running k-nn with k=%d" % num_k)
    LOGGER.info("embedding dimensionality is %d" % references.shape[-1])
    LOGGER.info(f"distance metric: {distance_metric}")
    distances, indices = knn_topk(references, queries, num_k, distance_metric)
    if embeddings_come_from_same_source:
        return indices[:, 1:], distances[:, 1:]
    return indices, distances
