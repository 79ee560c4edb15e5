#!/usr/bin/env python
"""Generate ``tests/golden/eval_golden.npz`` from the REAL reference code.

Run in the build container only (needs ``/root/reference``; the GPU box never runs
this).  The reference's evaluator imports third-party packages that are absent here
(pytorch_metric_learning, torchmetrics, faiss) but the functions on the hot path —
``CustomCalculator.calculate_maphashing / calc_hamming_dist / label_comparison_fn /
per_bit_balance`` (main/engine/accuracy_calculator.py), ``get_knn_torch``
(main/engine/get_knn.py) and DSCH's independent ``mean_average_precision``
(main/engine/DSCH/_utils.py:409-450) — only use torch.  So the absent modules are
replaced by empty stubs, the reference source files are loaded unmodified from where
they lie, and their outputs on seeded inputs are recorded.

    python tests/golden/make_golden_eval.py
"""
import importlib.util
import logging
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eval_golden.npz")


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_loader import load_reference  # noqa: E402


def load_dsch_map():
    """Extract DSCH's mean_average_precision by exec'ing only that function's source lines."""
    path = os.path.join(REF, "main/engine/DSCH/_utils.py")
    src = open(path).read().splitlines()
    start = next(i for i, line in enumerate(src) if line.startswith("def mean_average_precision"))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def "))
    ns = {"torch": torch}
    exec("\n".join(src[start:end]), ns)
    return ns["mean_average_precision"]


def multi_hot(gen, n, nlab, p):
    lab = (torch.rand(n, nlab, generator=gen) < p).float()
    empty = lab.sum(1) == 0
    lab[empty, torch.randint(0, nlab, (int(empty.sum()),), generator=gen)] = 1.0
    return lab


def pm1(gen, n, b):
    return torch.randint(0, 2, (n, b), generator=gen).float() * 2 - 1


def main():
    acc, knn = load_reference()
    dsch_map = load_dsch_map()
    out = {}
    cases = []

    def calc(k):
        return acc.CustomCalculator(k=k, device=torch.device("cpu"), distance_metric="hamming", with_faiss=False)

    def record(name, q, ql, r, rl, topk, includes=False):
        c = calc(topk)
        val = c.calculate_maphashing(q, ql, r, rl, topk, ref_includes_query=includes)
        out[f"{name}/q"] = q.numpy().astype(np.int8)
        out[f"{name}/r"] = r.numpy().astype(np.int8)
        out[f"{name}/ql"] = ql.numpy().astype(np.float32 if ql.dtype.is_floating_point else np.int64)
        out[f"{name}/rl"] = rl.numpy().astype(np.float32 if rl.dtype.is_floating_point else np.int64)
        flat = topk
        while isinstance(flat, (tuple, list)):
            flat = flat[0] if len(flat) else None
        tk = -1 if flat is None else (-2 if flat == "max_bin_count" else int(flat))
        out[f"{name}/topk"] = np.array([tk, int(includes)], dtype=np.int64)
        out[f"{name}/map_reference"] = np.array(val, dtype=np.float64)
        # same reference code, but with torch.argsort's tie order forced to the stable
        # (index) order: isolates the one implementation-defined step
        real_argsort = torch.argsort
        torch.argsort = lambda x, *a, **kw: real_argsort(x, *a, **{**kw, "stable": True})
        try:
            val_stable = c.calculate_maphashing(q, ql, r, rl, topk, ref_includes_query=includes)
        finally:
            torch.argsort = real_argsort
        out[f"{name}/map_reference_stable"] = np.array(val_stable, dtype=np.float64)
        if not isinstance(topk, (str, list)):
            out[f"{name}/map_dsch"] = np.array(float(dsch_map(q, r, ql, rl, topk)), dtype=np.float64)
        cases.append(name)
        print(f"{name:34s} topk={topk!s:>14}  reference maphashing = {val:.10f}  (stable ties: {val_stable:.10f})")

    # MAP-6: measure_random_baseline.py recipe (seeded random +-1 codes, multi-hot labels)
    for bits in (32, 64, 96, 128):
        g = torch.Generator().manual_seed(bits)
        q, r = pm1(g, 24, bits), pm1(g, 400, bits)
        ql, rl = multi_hot(g, 24, 24, 0.10), multi_hot(g, 400, 24, 0.10)
        for topk in (50, None, 1000):
            record(f"random_b{bits}_k{topk}", q, ql, r, rl, topk)
    # MAP-1: one constant code for every sample (all ties)
    g = torch.Generator().manual_seed(1)
    ql, rl = multi_hot(g, 12, 20, 0.15), multi_hot(g, 150, 20, 0.15)
    record("constant_codes", torch.ones(12, 64), ql, torch.ones(150, 64), rl, 40)
    # MAP-2: tie-free by construction: d(query m, item j) = j + m
    bits, nq = 64, 8
    base = pm1(torch.Generator().manual_seed(2), 1, bits)
    r = base.repeat(bits - nq + 2, 1)
    for j in range(r.shape[0]):
        r[j, :j] *= -1
    q = base.repeat(nq, 1)
    for m in range(nq):
        if m:
            q[m, -m:] *= -1
    g = torch.Generator().manual_seed(3)
    ql, rl = multi_hot(g, nq, 10, 0.3), multi_hot(g, r.shape[0], 10, 0.3)
    for topk in (10, None):
        record(f"tiefree_k{topk}", q, ql, r, rl, topk)
    # shuffled copy of the tie-free DB (index order != distance order)
    perm = torch.randperm(r.shape[0], generator=g)
    record("tiefree_shuffled", q, ql, r[perm], rl[perm], 25)
    # MAP-4: queries with all-zero labels count in the denominator
    g = torch.Generator().manual_seed(4)
    q, r = pm1(g, 10, 32), pm1(g, 120, 32)
    ql, rl = multi_hot(g, 10, 12, 0.2), multi_hot(g, 120, 12, 0.2)
    ql[::3] = 0
    record("zero_label_queries", q, ql, r, rl, 30)
    # MAP-5: 1-D integer labels take the equality branch
    g = torch.Generator().manual_seed(5)
    q, r = pm1(g, 16, 48), pm1(g, 300, 48)
    ql, rl = torch.randint(0, 7, (16,), generator=g), torch.randint(0, 7, (300,), generator=g)
    record("int_labels_k20", q, ql, r, rl, 20)
    record("int_labels_all", q, ql, r, rl, None)
    # "max_bin_count" (batch_map.py path: query == reference, ref_includes_query=True)
    g = torch.Generator().manual_seed(6)
    e = pm1(g, 64, 64)
    lab = multi_hot(g, 64, 8, 0.25)
    record("max_bin_count_self", e, lab, e, lab, "max_bin_count", includes=True)
    lab1 = torch.randint(0, 5, (64,), generator=g)
    record("max_bin_count_self_int", e, lab1, e, lab1, "max_bin_count", includes=True)
    # nested topk list (accuracy_calculator.py:204-205)
    record("nested_topk", q, ql, r, rl, [[15]])

    # primitives
    g = torch.Generator().manual_seed(7)
    c = calc(None)
    for bits in (32, 64, 96, 128):
        a, b = pm1(g, 5, bits), pm1(g, 33, bits)
        out[f"hamming_b{bits}/q"] = a.numpy().astype(np.int8)
        out[f"hamming_b{bits}/r"] = b.numpy().astype(np.int8)
        out[f"hamming_b{bits}/dist"] = c.calc_hamming_dist(a, b).numpy()
    la, lb = multi_hot(g, 6, 80, 0.04), multi_hot(g, 50, 80, 0.04)
    out["labels2d/q"], out["labels2d/r"] = la.numpy(), lb.numpy()
    out["labels2d/rel"] = c.label_comparison_fn(la, lb).numpy()
    ia, ib = torch.randint(0, 4, (6,), generator=g), torch.randint(0, 4, (50,), generator=g)
    out["labels1d/q"], out["labels1d/r"] = ia.numpy(), ib.numpy()
    out["labels1d/rel"] = c.label_comparison_fn(ia, ib).numpy()
    codes = pm1(g, 257, 64)
    codes[:, 3] = 1.0
    codes[:200, 5] = -1.0
    out["balance/codes"] = codes.numpy().astype(np.int8)
    out["balance/per_bit"] = c.per_bit_balance(codes).numpy()
    out["balance/mean"] = np.array(c.calculate_bit_balance(codes))
    out["balance/worst"] = np.array(c.calculate_worst_bit_balance(codes))
    # knn (torch path of the reference)
    refs = torch.nn.functional.normalize(torch.randn(200, 48, generator=g), dim=1)
    qs = torch.nn.functional.normalize(torch.randn(9, 48, generator=g), dim=1)
    out["knn/refs"], out["knn/queries"] = refs.numpy(), qs.numpy()
    for metric in ("cosine", "l2"):
        for same in (False, True):
            src_q = refs[:9] if same else qs
            idx, dist = knn.get_knn(refs, src_q, 10, same, with_faiss=False, distance_metric=metric)
            out[f"knn/{metric}_same{int(same)}/idx"] = idx.numpy()
            out[f"knn/{metric}_same{int(same)}/dist"] = dist.numpy()

    out["cases"] = np.array(cases)
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({os.path.getsize(OUT)} bytes, {len(cases)} mAP cases)")


if __name__ == "__main__":
    main()
