"""Evaluator glue on the device: mirrors of the callers of the hot path in ``/root/reference/main/engine/evaluate.py``.

The reference's tester moves every batch of float32 codes to the CPU (``q = q.cpu()``, evaluate.py:42), assembles ``[N, B]``
float32 matrices there and hands them to ``CustomCalculator`` (``get_tester(... device=cpu)``, :76-80); ``evaluate_multi_k``
(:172-245) reuses those embeddings for several ``k``.  Here the codes never leave the GPU as floats:

* :func:`compute_all_embeddings` (evaluate.py:26-64) runs the same loop — ``data_and_label_getter``, forward pass,
  fill-by-slices — but writes into device buffers, and with ``pack=True`` bit-packs every batch right behind the forward
  pass (``b200_pack_codes`` / ``b200_pack_labels`` straight into the final packed arrays: 16 + 16 bytes per COCO row instead
  of 832), returning an :class:`EmbeddingSet`;
* :func:`evaluate_multi_k` evaluates ``maphashing`` (+ bit balance) for every ``k`` of ``k_list`` on ONE packing of the
  codes, ``{k: {metric: value}}`` like the reference's ``results_by_k[k][split]``;
* :func:`evaluate` is the single-``k`` form.

The PML tester itself (split bookkeeping, logging, hooks) is pytorch-metric-learning code and stays the reference's; what
a maintainer changes there is in INTEGRATION.md §2.
"""
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _cabi
from . import hamming as H


@dataclass
class EmbeddingSet:
    """Codes and labels of one split on the device: packed (always) and, optionally, the float rows they came from."""
    codes: H.PackedCodes
    labels: H.PackedLabels
    float_codes: Optional[torch.Tensor] = None
    float_labels: Optional[torch.Tensor] = None

    def __len__(self):
        return self.codes.rows


def _default_getter(batch):
    """get_data, evaluate.py:92-93."""
    return batch["image"].cuda(), batch["label"]


def compute_all_embeddings(dataloader, trunk_model, embedder_model=None, data_and_label_getter=_default_getter, pack=True,
                           keep_float=False, on_nonbinary="sign", device=None):
    """evaluate.py:26-64 on the device.  ``pack=False``: ``(all_q, labels)`` float device tensors, the reference's return
    value minus the ``.cpu()``; ``pack=True``: an :class:`EmbeddingSet` (``keep_float`` also keeps the float rows, e.g. for
    the cosine ``map`` metric).  ``on_nonbinary='sign'`` binarises with ``x > 0`` (an eval-mode hashing head emits exact
    +-1, multi_dino_attention.py:750, so this only matters for raw logits); ``'raise'`` rejects anything else."""
    n = len(dataloader.dataset)
    if n == 0:
        raise ValueError("compute_all_embeddings got an empty dataset (check whatever built this split, e.g. build_fast_eval_subset)")
    _cabi.require_cuda()
    lib = _cabi.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    embedder_model = embedder_model if embedder_model is not None else torch.nn.Identity()
    s = 0
    all_q = labels = None
    codes_w = labels_w = bad = None
    bits = ncol = mode = None
    with torch.no_grad():
        for i, data in enumerate(dataloader):
            img, label = data_and_label_getter(data)
            q = embedder_model(trunk_model(img)).to(dev).float().contiguous()
            label = torch.as_tensor(label).to(dev)
            if label.dim() == 1:
                label = label.unsqueeze(1)
            label = label.contiguous()
            e = s + q.size(0)
            if i == 0:
                bits, ncol = int(q.size(1)), int(label.size(1))
                if not pack or keep_float:
                    all_q = torch.zeros(n, bits, device=dev, dtype=q.dtype)
                    labels = torch.zeros(n, ncol, device=dev, dtype=label.dtype)
                if pack:
                    if bits > _cabi.MAX_CODE_BITS:
                        raise NotImplementedError(f"codes wider than {_cabi.MAX_CODE_BITS} bits cannot be packed; use pack=False")
                    mode = _cabi.LABELS_OVERLAP if ncol > 1 else _cabi.LABELS_EQUAL
                    lw = _cabi.label_words(ncol) if ncol > 1 else 1
                    npad = (n + 1) // 2 * 2 + 2
                    codes_w = torch.zeros((npad, _cabi.code_words(bits)), dtype=torch.int64, device=dev)
                    labels_w = torch.zeros((npad, lw), dtype=torch.int64, device=dev)
                    bad = torch.zeros(2, dtype=torch.int32, device=dev)
            if all_q is not None:
                all_q[s:e] = q
                labels[s:e] = label
            if pack:
                if s % 2:
                    raise ValueError("batches must have an even number of rows except the last one (packed rows are written in "
                                     "16-byte units); use an even batch size")
                rows = e - s
                with torch.cuda.device(dev):
                    st = _cabi.stream_ptr()
                    _cabi.check(lib.b200_pack_codes(_cabi.ptr(q), rows, bits, codes_w[s:].data_ptr(), bad.data_ptr(), st), "b200_pack_codes")
                    if mode == _cabi.LABELS_OVERLAP:
                        lab = label.float().contiguous()
                        _cabi.check(lib.b200_pack_labels(_cabi.ptr(lab), rows, ncol, labels_w[s:].data_ptr(), bad.data_ptr() + 4, st),
                                    "b200_pack_labels")
                    else:
                        flat = label.reshape(-1)
                        kind = 1 if not flat.dtype.is_floating_point else (2 if flat.dtype == torch.float64 else 0)
                        flat = (flat.to(torch.int64) if kind == 1 else flat).contiguous()
                        _cabi.check(lib.b200_pack_labels_scalar(_cabi.ptr(flat), kind, rows, labels_w[s:].data_ptr(), bad.data_ptr() + 4, st),
                                    "b200_pack_labels_scalar")
            s = e
    if s != n:
        raise ValueError(f"the dataloader yielded {s} rows for a dataset of {n}")
    if not pack:
        return all_q, labels
    nbad = bad.cpu()
    if int(nbad[0]) and on_nonbinary == "raise":
        raise ValueError(f"{int(nbad[0])} code entries are not +-1: Hamming ranking is undefined for them")
    if int(nbad[1]):
        raise ValueError(f"{int(nbad[1])} label entries are neither 0 nor 1 (or NaN)")
    lw = int(labels_w.shape[1])
    return EmbeddingSet(H.PackedCodes(codes_w, n, bits), H.PackedLabels(labels_w, n, lw, mode), all_q, labels)


def _as_set(codes, labels, on_nonbinary):
    if isinstance(codes, EmbeddingSet):
        return codes
    t = torch.as_tensor(codes)
    lab = torch.as_tensor(labels)
    if lab.dim() == 2 and lab.shape[1] == 1:
        lab = lab.reshape(-1)
    return EmbeddingSet(H.pack_codes(t, on_nonbinary=on_nonbinary), H.pack_labels(lab))


def evaluate_multi_k(query, query_labels=None, reference=None, reference_labels=None, k_list=(5000,),
                     embeddings_come_from_same_source=False, on_nonbinary="raise", with_bit_balance=True):
    """evaluate.py:172-245 at the embedding level: ``maphashing`` for every ``k`` of ``k_list`` (``None`` = all rows) on ONE
    packing of the codes.  ``query`` / ``reference``: :class:`EmbeddingSet` (from :func:`compute_all_embeddings`) or float
    codes with their labels; ``reference=None`` evaluates the query set against itself.
    Returns ``{k: {"maphashing": float, "bit_balance": float, "worst_bit_balance": float}}``."""
    qs = _as_set(query, query_labels, on_nonbinary)
    rs = qs if reference is None else _as_set(reference, reference_labels, on_nonbinary)
    H._check_pair(qs.codes, qs.labels, rs.codes, rs.labels)
    balance = {}
    if with_bit_balance:
        ones = H.bit_counts(rs.codes).float() / float(max(rs.codes.rows, 1))
        per_bit = 1.0 - 2.0 * (ones - 0.5).abs()                # accuracy_calculator.py:188-194
        balance = {"bit_balance": per_bit.mean().item(), "worst_bit_balance": per_bit.min().item()}
    results = {}
    ws = None
    for k in k_list:
        kk = rs.codes.rows if k is None else int(k)
        # ref_includes_query only changes "max_bin_count" (accuracy_calculator.py:206-210); an integer k is used as is
        m, _, _, ws = H.hamming_map(qs.codes, qs.labels, rs.codes, rs.labels, kk, workspace=ws, return_workspace=True)
        results[k] = {"maphashing": m.item(), **balance}
    return results


def evaluate(query, query_labels=None, reference=None, reference_labels=None, k=5000, **kwargs):
    """Single-k form (evaluate.py:143-169 at the embedding level): ``{metric: value}``."""
    return evaluate_multi_k(query, query_labels, reference, reference_labels, k_list=(k,), **kwargs)[k]
