"""SURVEY.md §8 f2 on a B200: DSCH's Hamming metrics and calculate_pr_rc_hashing through the C-ABI kernels
(b200_hamming_radius_counts, b200_ranked_cumhits, b200_curve_accumulate) against the CPU oracle and against the goldens
recorded from the real reference code (tests/golden/dsch_golden.npz).

Bar: counts within a radius and running hit counts bit-exact; float32 curves within 1e-6 (sum order only)."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_ref
from simlib import multi_hot, pm1

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dsch_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "dsch_golden.npz"))


def _case(g, name):
    return (g[f"{name}/q"].astype(np.float32), g[f"{name}/ql"], g[f"{name}/r"].astype(np.float32), g[f"{name}/rl"])


def _problem(seed, nq, n, bits, nlab, near=True):
    rng = np.random.default_rng(seed)
    q, r = pm1(rng, nq, bits), pm1(rng, n, bits)
    if near:                                                     # populate the small radii
        for i in range(min(nq, n // 4)):
            r[4 * i:4 * i + 4] = q[i]
            for j in range(4):
                r[4 * i + j, rng.integers(0, bits, j)] *= -1
    if nlab > 0:
        return q, multi_hot(rng, nq, nlab, 0.1), r, multi_hot(rng, n, nlab, 0.1)
    return q, rng.integers(0, 6, nq), r, rng.integers(0, 6, n)


def _t(*arrays):
    return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in arrays)


@pytest.mark.parametrize("nq,n,bits,nlab", [(37, 501, 64, 24), (130, 3000, 32, 80), (64, 70000, 128, 130), (9, 1000, 48, -1),
                                            (5, 1, 16, 4)])
def test_radius_counts_bit_exact(nq, n, bits, nlab):
    from image_retrieval_wavelet_b200.engine import hamming as H

    q, ql, r, rl = _problem(nq * 7 + n, nq, n, bits, nlab)
    tq, tql, tr, trl = _t(q, ql, r, rl)
    qc, rc = H.pack_codes(tq), H.pack_codes(tr)
    cum = H.radius_counts(qc, H.pack_labels(tql), rc, H.pack_labels(trl)).cpu().numpy()
    d = eval_ref.hamming_ref(q, r)
    rel = eval_ref.label_rel_ref(ql, rl)
    want_all = np.stack([np.cumsum(np.bincount(d[i], minlength=bits + 1)) for i in range(nq)])
    want_rel = np.stack([np.cumsum(np.bincount(d[i][rel[i]], minlength=bits + 1)) for i in range(nq)])
    assert cum.shape == (nq, bits + 1, 2)
    assert np.array_equal(cum[:, :, 0], want_all) and np.array_equal(cum[:, :, 1], want_rel)


@pytest.mark.parametrize("nq,n,bits,nlab,k", [(37, 501, 64, 24, 501), (70, 3000, 32, 80, 1000), (33, 2000, 128, 200, 33), (9, 1000, 48, -1, 1)])
def test_ranked_cumhits_bit_exact(nq, n, bits, nlab, k):
    from image_retrieval_wavelet_b200.engine import hamming as H

    q, ql, r, rl = _problem(nq + n + k, nq, n, bits, nlab)
    tq, tql, tr, trl = _t(q, ql, r, rl)
    qc, rc = H.pack_codes(tq), H.pack_codes(tr)
    pql, prl = H.pack_labels(tql), H.pack_labels(trl)
    idx = H.hamming_topk(qc, rc, k, raw=True)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (nq, k)
    cum = H.ranked_cumhits(idx, pql, prl).cpu().numpy()
    d = eval_ref.hamming_ref(q, r)
    rel = eval_ref.label_rel_ref(ql, rl)
    order = np.argsort(d, axis=1, kind="stable")[:, :k]
    assert np.array_equal(idx.cpu().numpy().astype(np.int64) & 0xFFFFFFFF, order)
    assert np.array_equal(cum, np.cumsum(np.take_along_axis(rel, order, 1), axis=1))
    # padding entries (0xFFFFFFFF) never count
    padded = idx.clone()
    padded[:, k // 2:] = -1
    cum2 = H.ranked_cumhits(padded, pql, prl).cpu().numpy()
    assert np.array_equal(cum2[:, :k // 2], cum[:, :k // 2])
    if k // 2:
        assert np.all(cum2[:, k // 2:] == cum[:, k // 2 - 1:k // 2])


def test_dsch_metrics_match_reference_goldens(dsch_golden):
    from image_retrieval_wavelet_b200.engine import DSCH

    g = dsch_golden
    for name in g["cases"]:
        q, ql, r, rl = _case(g, name)
        tq, tql, tr, trl = _t(q, ql, r, rl)
        P, R = DSCH.pr_curve(tq, tr, tql, trl)
        assert P.dtype == torch.float32 and tuple(P.shape) == (q.shape[1] + 1,) and not P.is_cuda
        assert np.abs(P.numpy() - g[f"{name}/pr_P"]).max() <= 1e-6, name
        assert np.abs(R.numpy() - g[f"{name}/pr_R"]).max() <= 1e-6, name
        K = [int(k) for k in g[f"{name}/K"]]
        p = DSCH.p_topK(tq, tr, tql, trl, K=K)
        assert p.dtype == torch.float32 and tuple(p.shape) == (len(K),)
        assert np.abs(p.numpy() - g[f"{name}/ptopk_reference_stable"]).max() <= 1e-6, name
        for rad, want in zip(g[f"{name}/radius"], g[f"{name}/radius_prec"]):
            ql_np = ql.copy()
            got = DSCH.get_precision_recall_by_Hamming_Radius(r, rl, q, ql_np, radius=int(rad))
            assert isinstance(got, float) and abs(got - float(want)) <= 1e-12, (name, rad)
            assert np.array_equal(ql_np, ql)                     # arguments are left alone
        rewritten = np.where(ql == 0, -1.0, ql).astype(ql.dtype)  # labels a previous reference call already rewrote
        assert abs(DSCH.get_precision_recall_by_Hamming_Radius(r, rl, q, rewritten, radius=2)
                   - float(g[f"{name}/radius_prec"][list(g[f"{name}/radius"]).index(2)])) <= 1e-12


def test_dsch_metrics_match_oracle_on_larger_inputs():
    from image_retrieval_wavelet_b200.engine import DSCH

    q, ql, r, rl = _problem(11, 150, 20000, 64, 38)
    ql[::7] = 0                                                   # queries without any tag: no relevant row anywhere
    tq, tql, tr, trl = _t(q, ql, r, rl)
    P, R = DSCH.pr_curve(tq.cuda(), tr.cuda(), tql.cuda(), trl.cuda())
    P0, R0 = eval_ref.dsch_pr_curve_ref(q, r, ql, rl)
    assert np.abs(P.numpy() - P0).max() <= 1e-6 and np.abs(R.numpy() - R0).max() <= 1e-6
    p = DSCH.p_topK(tq, tr, tql, trl)                             # default K = 1, 100, ..., 1000
    assert np.abs(p.numpy() - eval_ref.dsch_p_topk_ref(q, r, ql, rl)).max() <= 1e-6
    for rad in (0, 2, 20, 64, 100):
        got = DSCH.get_precision_recall_by_Hamming_Radius(r, rl, q, ql, radius=rad)
        assert abs(got - eval_ref.dsch_radius_precision_ref(r, rl, q, ql, rad)) <= 1e-12
    m = DSCH.mean_average_precision(tq, tr, tql, trl, 5000)
    assert abs(float(m) - eval_ref.maphashing_exact(q, ql, r, rl, 5000)) <= 1e-6
    m = DSCH.mean_average_precision(tq, tr, tql, trl)
    assert abs(float(m) - eval_ref.maphashing_exact(q, ql, r, rl, None)) <= 1e-6
    d = DSCH.calc_hamming_dist(tq[0], tr[:100])
    assert tuple(d.shape) == (1, 100) and np.array_equal(d.numpy(), eval_ref.hamming_ref(q[:1], r[:100]).astype(np.float32))
    # 1-D labels: equality relevance (mean_average_precision's first branch, _utils.py:429-430)
    q1, l1, r1, rl1 = _problem(12, 40, 3000, 32, -1)
    m = DSCH.mean_average_precision(*_t(q1, r1, l1, rl1), 300)
    assert abs(float(m) - eval_ref.maphashing_exact(q1, l1, r1, rl1, 300)) <= 1e-6
    with pytest.raises(NotImplementedError):
        DSCH.mean_average_precision(tq, tr, tql[:, :, None], trl[:, :, None])


def test_pr_rc_hashing_matches_reference_goldens(dsch_golden, tmp_path, monkeypatch):
    from image_retrieval_wavelet_b200.engine import CustomCalculator

    g = dsch_golden
    c = CustomCalculator(k=None, distance_metric="hamming", with_faiss=False)
    for name in g["cases"]:
        q, ql, r, rl = _case(g, name)
        mask = g[f"{name}/not_lone"]
        prec, rec, used = c.pr_rc_hashing_curves(*_t(q, ql, r, rl, mask), chunk=7)
        assert used > 0
        assert np.abs(prec.cpu().numpy() - g[f"{name}/prrc_pr_reference_stable"]).max() <= 1e-6, name
        assert np.abs(rec.cpu().numpy() - g[f"{name}/prrc_rc_reference_stable"]).max() <= 1e-6, name
    monkeypatch.chdir(tmp_path)
    assert c.calculate_pr_rc_hashing(*_t(q, ql, r, rl, mask)) == 0
    import pandas as pd

    df = pd.read_csv(tmp_path / "pr_rc.csv")
    assert np.abs(df["pr"].to_numpy() - g[f"{name}/prrc_pr_reference_stable"]).max() <= 1e-6


def test_curve_kernels_launch():
    from image_retrieval_wavelet_b200 import _cabi
    from image_retrieval_wavelet_b200.engine import DSCH

    q, ql, r, rl = _problem(5, 10, 500, 64, 8)
    before = _cabi.launch_count()
    DSCH.p_topK(*_t(q, r, ql, rl), K=[1, 10])
    assert _cabi.launch_count() - before >= 10      # 4 packs, top-k (hist, scan, walk), cumhits, hist + totals + radius counts
