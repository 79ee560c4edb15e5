#!/usr/bin/env python
"""Generate ``tests/golden/dsch_golden.npz`` from the REAL reference code (SURVEY.md §8 f2 metrics).

Build container only (needs ``/root/reference``).  ``pr_curve``, ``p_topK``, ``calc_hamming_dist`` and
``get_precision_recall_by_Hamming_Radius`` of ``main/engine/DSCH/_utils.py`` only use torch / numpy, so their source
lines are exec'ed unmodified out of the file (the module itself imports packages that are absent here);
``CustomCalculator.calculate_pr_rc_hashing`` (main/engine/accuracy_calculator.py:235-273) is run through the same
stubbed import as ``make_golden_eval.py`` and its ``pr_rc.csv`` read back.

    python tests/golden/make_golden_dsch.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden_eval import REF, load_reference, multi_hot, pm1   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dsch_golden.npz")


def load_dsch(names):
    src = open(os.path.join(REF, "main/engine/DSCH/_utils.py")).read().splitlines()
    ns = {"torch": torch, "np": np}
    for name in names:
        start = next(i for i, line in enumerate(src) if line.startswith(f"def {name}("))
        end = next(i for i in range(start + 1, len(src)) if src[i].startswith(("def ", "class ")))
        exec("\n".join(src[start:end]), ns)
    return ns


def stable_sorts():
    """Context: torch.argsort / torch.sort with the tie order forced to the index order."""
    class _Ctx:
        def __enter__(self):
            self.a, self.s = torch.argsort, torch.sort
            torch.argsort = lambda x, *a, **kw: self.a(x, *a, **{**kw, "stable": True})
            torch.sort = lambda x, *a, **kw: self.s(x, *a, **{**kw, "stable": True})

        def __exit__(self, *exc):
            torch.argsort, torch.sort = self.a, self.s
    return _Ctx()


def main():
    ns = load_dsch(["calc_hamming_dist", "pr_curve", "p_topK", "get_precision_recall_by_Hamming_Radius"])
    acc, _ = load_reference()
    calc = acc.CustomCalculator(k=None, device=torch.device("cpu"), distance_metric="hamming", with_faiss=False)
    out, cases = {}, []

    def record(name, q, ql, r, rl, K, radius, zero_queries=()):
        ql = ql.clone()
        for z in zero_queries:
            ql[z] = 0
        out[f"{name}/q"], out[f"{name}/r"] = q.numpy().astype(np.int8), r.numpy().astype(np.int8)
        out[f"{name}/ql"], out[f"{name}/rl"] = ql.numpy(), rl.numpy()
        P, R = ns["pr_curve"](q, r, ql, rl)
        out[f"{name}/pr_P"], out[f"{name}/pr_R"] = P.numpy(), R.numpy()
        out[f"{name}/K"] = np.array(K, dtype=np.int64)
        out[f"{name}/ptopk_reference"] = ns["p_topK"](q, r, ql, rl, K=list(K)).numpy()
        with stable_sorts():
            out[f"{name}/ptopk_reference_stable"] = ns["p_topK"](q, r, ql, rl, K=list(K)).numpy()
        out[f"{name}/radius"] = np.array(radius, dtype=np.int64)
        out[f"{name}/radius_prec"] = np.array(
            [ns["get_precision_recall_by_Hamming_Radius"](r.numpy().copy(), rl.numpy().copy(), q.numpy().copy(),
                                                          ql.numpy().copy(), radius=rad) for rad in radius], dtype=np.float64)
        # calculate_pr_rc_hashing writes pr_rc.csv into the working directory
        mask = torch.ones(q.shape[0], dtype=torch.bool)
        mask[1::4] = False
        out[f"{name}/not_lone"] = mask.numpy()
        import pandas as pd
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                for tag, ctx in (("reference", None), ("reference_stable", stable_sorts())):
                    if os.path.exists("pr_rc.csv"):
                        os.remove("pr_rc.csv")
                    if ctx is None:
                        calc.calculate_pr_rc_hashing(q, ql, r, rl, mask)
                    else:
                        with ctx:
                            calc.calculate_pr_rc_hashing(q, ql, r, rl, mask)
                    df = pd.read_csv("pr_rc.csv")
                    out[f"{name}/prrc_pr_{tag}"] = df["pr"].to_numpy(dtype=np.float64)
                    out[f"{name}/prrc_rc_{tag}"] = df["rc"].to_numpy(dtype=np.float64)
            finally:
                os.chdir(cwd)
        cases.append(name)
        print(f"{name:28s} P[:3]={P[:3].numpy()}  p@K={out[f'{name}/ptopk_reference'][:3]}  P@H<={radius[0]}={out[f'{name}/radius_prec'][0]:.6f}")

    for bits in (32, 64, 128):
        g = torch.Generator().manual_seed(100 + bits)
        q, r = pm1(g, 20, bits), pm1(g, 500, bits)
        # make some database rows close to the queries so that small radii are not empty
        for i in range(20):
            near = q[i].repeat(6, 1)
            flips = torch.randint(0, bits, (6, 3), generator=g)
            for j in range(6):
                near[j, flips[j, :j % 4]] *= -1
            r[i * 6:(i + 1) * 6] = near
        ql, rl = multi_hot(g, 20, 24, 0.10), multi_hot(g, 500, 24, 0.10)
        record(f"dsch_b{bits}", q, ql, r, rl, K=[1, 5, 50, 100, 499, 500, 1000], radius=[0, 2, 5], zero_queries=(3,))
    # small database: every K beyond N clamps
    g = torch.Generator().manual_seed(7)
    q, r = pm1(g, 6, 16), pm1(g, 40, 16)
    ql, rl = multi_hot(g, 6, 5, 0.3), multi_hot(g, 40, 5, 0.3)
    record("dsch_small", q, ql, r, rl, K=[1, 10, 40, 100], radius=[2, 16])

    out["cases"] = np.array(cases)
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({os.path.getsize(OUT)} bytes, {len(cases)} cases)")


if __name__ == "__main__":
    main()
