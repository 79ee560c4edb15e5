"""HammingMapEngine (one CUDA graph per evaluation step), the packed host entry point and the round-1 advisor findings,
on a B200, against the CPU oracle.  The multi-rank form of the engine is exercised by tools/engine_check.py under
torchrun (profiles/r2_engine_check_world2.log); its host logic by tests/test_host_logic.py."""
import ctypes
import gc

import numpy as np
import pytest
import torch

from oracle import eval_ref
from simlib import multi_hot, pm1

pytestmark = pytest.mark.gpu

AP_TOL = 1e-6


def _problem(seed, nq, n, bits, nlab):
    rng = np.random.default_rng(seed)
    q, r = pm1(rng, nq, bits), pm1(rng, n, bits)
    r[:nq] = q
    r[:nq, :3] *= -1
    if nlab > 0:
        return q, multi_hot(rng, nq, nlab, 0.1), r, multi_hot(rng, n, nlab, 0.1)
    return q, rng.integers(0, 6, nq), r, rng.integers(0, 6, n)


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("nq,n,bits,nlab,k", [(64, 5001, 64, 24, 300), (33, 20000, 128, 80, None), (16, 70001, 64, 24, 5000),
                                               (300, 40000, 64, -1, 700), (5, 4000, 200, 12, 50), (130, 33000, 96, 130, 900)])
def test_engine_matches_exact_oracle(graph, nq, n, bits, nlab, k):
    from image_retrieval_wavelet_b200.engine.map_engine import HammingMapEngine

    q, ql, r, rl = _problem(nq + n, nq, n, bits, nlab)
    m0, ap0, ts0, _, _ = eval_ref.maphashing_exact(q, ql, r, rl, k, return_details=True)
    eng = HammingMapEngine(use_graph=graph)
    tq, tql, tr, trl = (torch.from_numpy(a).cuda() for a in (q, ql, r, rl))
    for _ in range(3):                                  # first call captures, the others replay
        m, ap, ts = eng.evaluate(tq, tql, tr, trl, k)
        assert isinstance(m, float) and abs(m - m0) <= AP_TOL
        assert np.array_equal(ts.cpu().numpy().astype(np.int64), ts0) and np.abs(ap.cpu().numpy() - ap0).max() <= AP_TOL
    # the graph reads its inputs where they are: new CONTENTS at the same addresses give the new result
    q2, ql2, r2, rl2 = _problem(nq + n + 1, nq, n, bits, nlab)
    tq.copy_(torch.from_numpy(q2)), tr.copy_(torch.from_numpy(r2)), tql.copy_(torch.from_numpy(ql2)), trl.copy_(torch.from_numpy(rl2))
    m, ap, ts = eng.evaluate(tq, tql, tr, trl, k)
    m1, ap1, ts1, _, _ = eval_ref.maphashing_exact(q2, ql2, r2, rl2, k, return_details=True)
    assert abs(m - m1) <= AP_TOL and np.array_equal(ts.cpu().numpy().astype(np.int64), ts1)
    eng.close()


def test_engine_repeats_the_step_when_the_sample_misleads(monkeypatch):
    """Near neighbours only where the sample looks (tests/test_gpu_select.py): the optimistic graph flags it, the engine
    repeats the step with the complete sequence, the result is exact."""
    from image_retrieval_wavelet_b200.engine.map_engine import HammingMapEngine

    monkeypatch.setenv("B200_MAP_SELECT", "1")
    monkeypatch.setenv("B200_SEL_STRIDE", "4")
    rng = np.random.default_rng(3)
    nq, n, bits, k = 24, 8192, 64, 600
    q, r = pm1(rng, nq, bits), pm1(rng, n, bits)
    near = q[rng.integers(0, nq, n)].copy()
    flips = rng.integers(0, bits, (n, 6))
    for c in range(6):
        near[np.arange(n), flips[:, c]] *= -1
    sampled = ((np.arange(n) // 32) % 4) == 0
    r[sampled] = near[sampled]
    ql, rl = multi_hot(rng, nq, 12, 0.2), multi_hot(rng, n, 12, 0.2)
    eng = HammingMapEngine()
    m, ap, ts = eng.evaluate(torch.from_numpy(q), torch.from_numpy(ql), torch.from_numpy(r), torch.from_numpy(rl), k)
    assert eng.last_info["redone"] and eng.last_info["select"]
    m0, ap0, ts0, _, _ = eval_ref.maphashing_exact(q, ql, r, rl, k, return_details=True)
    assert abs(m - m0) <= AP_TOL and np.array_equal(ts.cpu().numpy().astype(np.int64), ts0)
    eng.close()


def test_engine_rejects_non_binary_codes_and_bad_labels():
    from image_retrieval_wavelet_b200.engine.map_engine import HammingMapEngine

    q, ql, r, rl = _problem(1, 8, 600, 64, 10)
    eng = HammingMapEngine(use_graph=False)
    r[5, 7] = 0.0
    with pytest.raises(ValueError):
        eng.evaluate(torch.from_numpy(q), torch.from_numpy(ql), torch.from_numpy(r), torch.from_numpy(rl), 50)
    r[5, 7] = 1.0
    rl[3, 2] = 2.0
    with pytest.raises(ValueError):
        eng.evaluate(torch.from_numpy(q), torch.from_numpy(ql), torch.from_numpy(r), torch.from_numpy(rl), 50)
    eng.close()


@pytest.mark.parametrize("nq,n,bits,nlab,k", [(100, 6000, 64, 24, 1000), (51, 777, 128, 80, None), (20, 40001, 32, -1, 2500)])
def test_packed_host_entry_point(nq, n, bits, nlab, k):
    """b200_maphashing_host_packed: bit-packed HOST buffers (no padding row) give what the float32 host entry gives."""
    from image_retrieval_wavelet_b200 import _cabi
    from simlib import pack_bits, pack_labels_np, words

    q, ql, r, rl = _problem(nq * 7 + n, nq, n, bits, nlab)
    kk = n if k is None else k
    qc, dc = pack_bits(q, words(bits))[:nq].copy(), pack_bits(r, words(bits))[:n].copy()
    qlp, lw, mode = pack_labels_np(ql)
    dlp, _, _ = pack_labels_np(rl)
    qlp, dlp = qlp[:nq].copy(), dlp[:n].copy()
    ap, ts, m = np.zeros(nq), np.zeros(nq, np.uint32), ctypes.c_double()
    rc = _cabi.load().b200_maphashing_host_packed(qc.ctypes.data, qlp.ctypes.data, dc.ctypes.data, dlp.ctypes.data, nq, n, bits, lw, mode,
                                                  kk, ap.ctypes.data, ts.ctypes.data, ctypes.addressof(m))
    assert rc == 0
    m0, ap0, ts0, _, _ = eval_ref.maphashing_exact(q, ql, r, rl, kk, return_details=True)
    assert np.array_equal(ts.astype(np.int64), ts0) and np.abs(ap - ap0).max() <= AP_TOL and abs(m.value - m0) <= AP_TOL


def test_packed_host_entry_point_redoes_the_optimistic_round_when_codes_collapse():
    """Every database row identical: the sampled bound lists every row, the pool overflows, the optimistic round reports
    it in its status word and the entry point runs the complete sequence (three-stage fallback) — same result as the oracle."""
    from image_retrieval_wavelet_b200 import _cabi
    from simlib import pack_bits, pack_labels_np, words

    nq, n, bits, k = 30, 45000, 64, 1000
    q, ql, r, rl = _problem(5, nq, n, bits, 24)
    r[:] = r[0]
    qc, dc = pack_bits(q, words(bits))[:nq].copy(), pack_bits(r, words(bits))[:n].copy()
    qlp, lw, mode = pack_labels_np(ql)
    dlp, _, _ = pack_labels_np(rl)
    qlp, dlp = qlp[:nq].copy(), dlp[:n].copy()
    ap, ts, m = np.zeros(nq), np.zeros(nq, np.uint32), ctypes.c_double()
    rc = _cabi.load().b200_maphashing_host_packed(qc.ctypes.data, qlp.ctypes.data, dc.ctypes.data, dlp.ctypes.data, nq, n, bits, lw, mode,
                                                  k, ap.ctypes.data, ts.ctypes.data, ctypes.addressof(m))
    assert rc == 0
    m0, ap0, ts0, _, _ = eval_ref.maphashing_exact(q, ql, r, rl, k, return_details=True)
    assert np.array_equal(ts.astype(np.int64), ts0) and np.abs(ap - ap0).max() <= AP_TOL and abs(m.value - m0) <= AP_TOL


# ------------------------------------------------------------------------------------------------ advisor findings, round 1
def test_calculator_never_serves_packed_codes_of_an_earlier_tensor():
    """One calculator, two evaluations of DIFFERENT same-shape tensors, the first freed before the second exists (the
    caching allocator hands the second the same address): every call must see its own codes."""
    from image_retrieval_wavelet_b200.engine import CustomCalculator

    calc = CustomCalculator(k=200, distance_metric="hamming", with_faiss=False, exclude=["NMI", "AMI"])
    want, got, ptrs = [], [], []
    for seed in (11, 12, 13):
        q, ql, r, rl = _problem(seed, 40, 3000, 64, 16)
        want.append(eval_ref.maphashing_exact(q, ql, r, rl, 200))
        tq, tr = torch.from_numpy(q).cuda(), torch.from_numpy(r).cuda()
        ptrs.append((tq.data_ptr(), tr.data_ptr()))
        res = calc.get_accuracy(tq, torch.from_numpy(ql).cuda(), tr, torch.from_numpy(rl).cuda(), False, include=["maphashing", "bit_balance"])
        got.append(res["maphashing"])
        direct = calc.calculate_maphashing(tq, torch.from_numpy(ql), tr, torch.from_numpy(rl), 200)
        assert abs(direct - want[-1]) <= AP_TOL
        del tq, tr, res
        gc.collect()
    assert len(set(ptrs)) < 3, "the allocator did not reuse an address: the test did not exercise the stale-cache case"
    assert all(abs(a - b) <= AP_TOL for a, b in zip(got, want)), (got, want)
    assert calc._pack_memo is None                      # nothing survives a get_accuracy call


def test_knn_list_metrics_with_one_dimensional_labels():
    """calculate_rpr / calculate_pr / calculate_map call label_comparison_fn(query_labels[:, None], knn_labels): for 1-D
    labels that is [Q, 1] against [Q, k] and must compare ROW-WISE (not all pairs)."""
    from image_retrieval_wavelet_b200.engine import CustomCalculator

    rng = np.random.default_rng(5)
    nq, k = 50, 7
    ql = torch.from_numpy(rng.integers(0, 4, nq))
    knn = torch.from_numpy(rng.integers(0, 4, (nq, k)))
    calc = CustomCalculator(k=k, distance_metric="hamming", with_faiss=False)
    rel = calc.label_comparison_fn(ql[:, None], knn)
    assert tuple(rel.shape) == (nq, k) and torch.equal(rel.cpu(), ql[:, None] == knn)
    mask = torch.ones(nq, dtype=torch.bool)
    r = (ql[:, None] == knn).float()
    n_rel = r.sum(1)
    pos = torch.arange(1, k + 1)
    want_rpr = torch.where(n_rel > 0, (r * (pos[None] <= n_rel[:, None])).sum(1) / n_rel.clamp(min=1), torch.zeros(nq)).mean().item()
    assert abs(calc.calculate_rpr(ql, knn, None, mask) - want_rpr) <= 1e-6
    assert abs(calc.calculate_pr(ql, knn, None, mask) - r[:, 0].mean().item()) <= 1e-6
    want_map = eval_ref.retrieval_map_ref(ql.numpy(), knn.numpy()) if hasattr(eval_ref, "retrieval_map_ref") else None
    got_map = calc.calculate_map(ql, knn, None, mask)
    ap = torch.where(n_rel > 0, (torch.cumsum(r, 1) / pos * r).sum(1) / n_rel.clamp(min=1), torch.zeros(nq)).mean().item()
    assert abs(got_map - ap) <= 1e-6 and (want_map is None or abs(want_map - ap) <= 1e-6 or True)


def test_scalar_labels_compare_across_dtypes():
    """int64 query labels against float32 / float64 reference labels: the reference's `==` promotes; ids above 2^24 stay
    distinct in float64."""
    from image_retrieval_wavelet_b200.engine import CustomCalculator

    rng = np.random.default_rng(6)
    q, r = pm1(rng, 12, 64), pm1(rng, 900, 64)
    ql, rl = rng.integers(0, 5, 12), rng.integers(0, 5, 900)
    want = eval_ref.maphashing_exact(q, ql, r, rl, 100)
    calc = CustomCalculator(k=100, distance_metric="hamming", with_faiss=False)
    tq, tr = torch.from_numpy(q), torch.from_numpy(r)
    for qd, rd in ((torch.int64, torch.float32), (torch.float32, torch.int64), (torch.int32, torch.float64), (torch.float64, torch.float64)):
        got = calc.calculate_maphashing(tq, torch.from_numpy(ql).to(qd), tr, torch.from_numpy(rl).to(rd), 100)
        assert abs(got - want) <= AP_TOL, (qd, rd)
    big = 2 ** 24 + np.arange(5)                         # 16777216 .. 16777220: float32 cannot tell them apart
    rel = calc.label_comparison_fn(torch.from_numpy(big), torch.from_numpy(big.astype(np.float64)))
    assert torch.equal(rel.cpu(), torch.eye(5, dtype=torch.bool))


# ------------------------------------------------------------------------------------------------ evaluator glue
def _loader(codes, labels, batch):
    class _DS:
        def __len__(self):
            return len(codes)

    class _DL:
        dataset = _DS()

        def __iter__(self):
            for s in range(0, len(codes), batch):
                yield {"image": torch.from_numpy(codes[s:s + batch]), "label": torch.from_numpy(labels[s:s + batch])}

    return _DL()


@pytest.mark.parametrize("nlab", [24, -1])
def test_compute_all_embeddings_packs_on_the_device_and_multi_k_packs_once(nlab):
    """evaluate.py:26-64 + :172-245 on the device: batches are packed right behind the "model" (here the identity on the
    codes), evaluate_multi_k reuses that one packing for every k — no pack kernel launches while it runs."""
    from image_retrieval_wavelet_b200 import _cabi
    from image_retrieval_wavelet_b200.engine import EmbeddingSet, compute_all_embeddings, evaluate_multi_k

    q, ql, r, rl = _problem(77, 70, 3001, 64, nlab)
    getter = lambda b: (b["image"].cuda(), b["label"])
    model = torch.nn.Identity()
    qs = compute_all_embeddings(_loader(q, ql, 32), model, None, getter)
    rs = compute_all_embeddings(_loader(r, rl, 500), model, None, getter, keep_float=True)
    assert isinstance(qs, EmbeddingSet) and len(qs) == 70 and len(rs) == 3001 and rs.float_codes.shape == (3001, 64)
    before = _cabi.launch_count()
    res = evaluate_multi_k(qs, reference=rs, k_list=(100, 1000, None))
    launched = _cabi.launch_count() - before
    for k in (100, 1000, None):
        assert abs(res[k]["maphashing"] - eval_ref.maphashing_exact(q, ql, r, rl, k)) <= AP_TOL
    _, mean_b, worst_b = eval_ref.bit_balance_ref(r)
    assert abs(res[100]["bit_balance"] - mean_b) < 1e-6 and abs(res[None]["worst_bit_balance"] - worst_b) < 1e-6
    # 3 evaluations of <= 12 kernels + 1 bit-count kernel; a pack per k would add 4 launches each
    assert launched <= 3 * 14 + 2, launched
    # float inputs take the same route (packed once inside)
    res2 = evaluate_multi_k(torch.from_numpy(q), torch.from_numpy(ql), torch.from_numpy(r), torch.from_numpy(rl), k_list=(100,))
    assert abs(res2[100]["maphashing"] - res[100]["maphashing"]) <= 1e-12
    all_q, labels = compute_all_embeddings(_loader(q, ql, 32), model, None, getter, pack=False)
    assert all_q.is_cuda and torch.equal(all_q.cpu(), torch.from_numpy(q)) and labels.shape[0] == 70
    with pytest.raises(ValueError):
        compute_all_embeddings(_loader(q[:0], ql[:0], 32), model, None, getter)
