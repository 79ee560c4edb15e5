"""The kernels' tile programs, executed on the CPU by the simulator (csrc/hostsim.cpp: same __host__ __device__ code
and planners as the CUDA build), against the oracle.  This is what can be checked without a GPU: indexing, halos, tile
planning, counter packing, scan and shard logic.  Device-only behaviour is covered by the -m gpu tests."""
import numpy as np
import pytest

from oracle import c_oracle, eval_ref, filters
from simlib import correlated_codes, multi_hot, pm1, sim_map, sim_swt

SWT_CASES = [
    ((2, 3, 224, 224), "haar", 1, np.uint8),      # BASELINE config C1 shape (small batch)
    ((1, 2, 32, 40), "db2", 2, np.float32),
    ((1, 1, 64, 72), "sym4", 3, np.float32),
    ((1, 1, 518, 518), "haar", 1, np.uint8),      # W % 4 != 0: 64-bit vector path
    ((1, 1, 520, 520), "db4", 3, np.uint8),       # C4 worst halo (49 rows/cols)
    ((1, 2, 48, 36), "bior4.4", 2, np.uint8),     # F = 10, biorthogonal
    ((1, 1, 30, 34), "db2", 1, np.float32),
    ((1, 1, 32, 32), "db7", 2, np.float32),       # F = 14: generic program
    ((1, 1, 16, 16), "db3", 1, np.uint8),
    ((1, 1, 8, 8), "db4", 3, np.float32),         # image smaller than the dilated filter: multiple wraps
    ((1, 1, 64, 64), "haar", 4, np.float32),      # level 4: generic program
    ((3, 1, 136, 200), "coif1", 3, np.uint8),
    ((1, 1, 2, 2), "haar", 1, np.uint8),          # minimum size
]


@pytest.mark.parametrize("shape,name,level,dtype", SWT_CASES)
def test_swt_tile_program(sim, shape, name, level, dtype):
    rng = np.random.default_rng(sum(shape) + level)
    x = rng.integers(0, 256, shape).astype(np.uint8) if dtype == np.uint8 else rng.random(shape, dtype=np.float32)
    lo, hi = filters.filter_bank(name)
    rc, out, plan = sim_swt(sim, x, lo, hi, level)
    assert rc == 0
    ref = c_oracle.swt2(x, lo, hi, level)
    assert not np.isnan(out).any(), "some output pixel was never written"
    for band in range(4):
        tol = 1e-5 * max(np.abs(ref[:, :, band]).max(), 1e-30)
        assert np.abs(out[:, :, band] - ref[:, :, band]).max() <= tol, (band, plan)


@pytest.mark.parametrize("rw", ["0", "1", "2"])
@pytest.mark.parametrize("shape,name,level,dtype", [((1, 2, 70, 518), "haar", 1, np.uint8), ((1, 1, 66, 94), "db4", 1, np.uint8),
                                                     ((2, 1, 40, 36), "db2", 2, np.float32), ((1, 1, 72, 88), "sym4", 3, np.uint8),
                                                     ((1, 1, 64, 80), "haar", 3, np.float32), ((1, 2, 48, 36), "bior4.4", 2, np.uint8),
                                                     ((1, 1, 8, 8), "db4", 3, np.float32), ((1, 1, 130, 518), "db2", 1, np.uint8)])
def test_swt_pass_forms_agree(sim, monkeypatch, rw, shape, name, level, dtype):
    """B200_SWT_RW: 0 two-pass levels (horizontal outputs through shared memory), 1 register-window passes (a window of F
    filtered rows per unit, no intermediate planes), 2 register-window intermediate levels + two-pass last level (the
    default for levels 2-3 under short filters) — same sub-bands, including x / y wrap, W % 4 = 2 and tiny images."""
    monkeypatch.setenv("B200_SWT_RW", rw)
    rng = np.random.default_rng(sum(shape) + level)
    x = rng.integers(0, 256, shape).astype(np.uint8) if dtype == np.uint8 else rng.random(shape, dtype=np.float32)
    lo, hi = filters.filter_bank(name)
    rc, out, plan = sim_swt(sim, x, lo, hi, level)
    assert rc == 0
    ref = c_oracle.swt2(x, lo, hi, level)
    assert not np.isnan(out).any(), "some output pixel was never written"
    for band in range(4):
        assert np.abs(out[:, :, band] - ref[:, :, band]).max() <= 1e-5 * max(np.abs(ref[:, :, band]).max(), 1e-30), (band, plan)


@pytest.mark.parametrize("vs", ["0", "1"])
@pytest.mark.parametrize("rw", ["0", "2"])
@pytest.mark.parametrize("shape,name,level,dtype", [((1, 2, 70, 518), "haar", 1, np.uint8), ((1, 1, 66, 94), "db4", 1, np.uint8),
                                                     ((2, 1, 40, 36), "db2", 2, np.float32), ((1, 1, 72, 88), "sym4", 3, np.uint8),
                                                     ((1, 2, 48, 36), "bior4.4", 2, np.uint8), ((1, 1, 8, 8), "db4", 3, np.float32),
                                                     ((1, 1, 130, 518), "sym4", 1, np.uint8)])
def test_swt_sliding_last_pass_agrees(sim, monkeypatch, vs, rw, shape, name, level, dtype):
    """B200_SWT_VS: 0 blocked last vertical pass (R + F - 1 rows per R outputs, both bands of a half per unit), 1 sliding
    pass (register window of F rows, one band per unit; the default for 8-tap filters at level 1) — same sub-bands for
    tiles that overhang the image bottom, W % 4 = 2 rows and images smaller than one tile."""
    monkeypatch.setenv("B200_SWT_VS", vs)
    monkeypatch.setenv("B200_SWT_RW", rw)
    rng = np.random.default_rng(sum(shape) + level)
    x = rng.integers(0, 256, shape).astype(np.uint8) if dtype == np.uint8 else rng.random(shape, dtype=np.float32)
    lo, hi = filters.filter_bank(name)
    rc, out, plan = sim_swt(sim, x, lo, hi, level)
    assert rc == 0
    ref = c_oracle.swt2(x, lo, hi, level)
    assert not np.isnan(out).any(), "some output pixel was never written"
    for band in range(4):
        assert np.abs(out[:, :, band] - ref[:, :, band]).max() <= 1e-5 * max(np.abs(ref[:, :, band]).max(), 1e-30), (band, plan)


@pytest.mark.parametrize("stage", ["0", "1"])
@pytest.mark.parametrize("shape,name,level", [((1, 2, 70, 518), "haar", 1), ((1, 1, 66, 94), "db4", 1), ((2, 1, 40, 36), "db2", 2),
                                              ((1, 1, 24, 10), "haar", 1)])
def test_swt_uint8_staging_units_agree(sim, monkeypatch, stage, shape, name, level):
    """Both uint8 staging forms (4-pixel units shared with float32, 8-pixel units: three aligned words + funnel shift)
    on rows that start on every byte alignment, with x/y wrap, narrow images and the short last unit of a row."""
    monkeypatch.setenv("B200_SWT_U8STAGE", stage)
    x = np.random.default_rng(len(name) + shape[3]).integers(0, 256, shape).astype(np.uint8)
    lo, hi = filters.filter_bank(name)
    rc, out, plan = sim_swt(sim, x, lo, hi, level)
    assert rc == 0
    ref = c_oracle.swt2(x, lo, hi, level)
    assert not np.isnan(out).any()
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max(), plan


@pytest.mark.parametrize("sms", [1, 16, 148, 1000])
def test_swt_planner_variants_agree(sim, sms):
    """Different SM counts make the planner pick different tiles; the result must not change."""
    x = np.random.default_rng(7).integers(0, 256, (1, 2, 56, 88)).astype(np.uint8)
    lo, hi = filters.filter_bank("db2")
    ref = c_oracle.swt2(x, lo, hi, 2)
    rc, out, plan = sim_swt(sim, x, lo, hi, 2, sms)
    assert rc == 0 and np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max(), plan


def test_u8_conversion_is_bit_identical_to_division(sim):
    """custom_transforms.py:147 divides by 255.0 in float32; the kernels use a 2-instruction FMA form instead."""
    import ctypes

    sim.sim_u8_unit.restype = ctypes.c_float
    got = np.array([sim.sim_u8_unit(ctypes.c_uint(b)) for b in range(256)], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), (np.arange(256, dtype=np.float32) / np.float32(255.0)).view(np.uint32))


def test_swt_rejects_bad_arguments(sim):
    x = np.zeros((1, 1, 6, 8), np.uint8)
    lo, hi = filters.filter_bank("haar")
    assert sim_swt(sim, x, lo, hi, 2)[0] != 0                    # 6 % 4 != 0
    assert sim_swt(sim, x, lo[:1], hi[:1], 1)[0] != 0            # odd filter length


MAP_CASES = [
    # Q, N, bits, labels (-1: 1-D integer labels), k
    (37, 500, 64, 24, 50),
    (37, 500, 64, 24, None),
    (20, 3000, 32, 20, 700),
    (20, 3000, 128, 80, 3000),
    (9, 2000, 96, 130, 100),
    (16, 70000, 64, 24, 66000),      # wide counters (k > 65534)
    (16, 70000, 64, 24, 5000),       # narrow counters, many segments
    (5, 1, 64, 8, 1),
    (5, 2, 64, 8, 5),
    (33, 1000, 200, 200, None),      # 4 code words, 4 label words
    (40, 5000, 48, -1, 300),         # equality labels
    (3, 777, 17, 3, 10),             # odd everything
    (7, 1029, 256, 12, 64),          # B = 256: distances need 9 bits, never stashed
    (6, 333, 254, 5, None),          # largest stashable code width, partial last groups (333 % 32, 333 % 4)
]


@pytest.mark.parametrize("nq,n,bits,nlab,k", MAP_CASES)
@pytest.mark.parametrize("shards,ext,sms,stash", [(1, 0, 148, 1), (1, 1, 148, 0), (3, 0, 148, 1), (8, 0, 4, 1), (2, 0, 148, 0)])
def test_hamming_map_stage_programs(sim, monkeypatch, nq, n, bits, nlab, k, shards, ext, sms, stash):
    """stash = 1: stage B ranks from the (distance, relevance) stash stage A wrote; 0: stage B scores again."""
    monkeypatch.setenv("B200_MAP_STASH", str(stash))
    rng = np.random.default_rng(nq * 1000 + n + bits)
    q, r = pm1(rng, nq, bits), pm1(rng, n, bits)
    if n > nq:
        r[:nq] = q
        r[:nq, :3] *= -1                                 # near-duplicates: short distances exist
    if nlab > 0:
        ql, rl = multi_hot(rng, nq, nlab, 0.1), multi_hot(rng, n, nlab, 0.1)
    else:
        ql, rl = rng.integers(0, 6, nq), rng.integers(0, 6, n)
    m0, ap0, ts0, rank0, dist0 = eval_ref.maphashing_exact(q, ql, r, rl, k, return_details=True)
    m, ap, ts, ri, rd, plan = sim_map(sim, q, ql, r, rl, k, shards, ext, sms, want_rank=True)
    assert plan[3] // 2 == (stash if bits <= 254 else 0), plan                 # the mode under test is the one that ran
    assert np.array_equal(ts.astype(np.int64), ts0), plan                      # hits in the top-k: bit-exact
    assert np.array_equal(ri.astype(np.int64), rank0), plan                    # ranked indices: bit-exact
    assert np.array_equal(rd.astype(np.int64), dist0), plan                    # distances: bit-exact
    assert np.abs(ap - ap0).max() <= 1e-6 and abs(m - m0) <= 1e-6              # AP: fp32 quotients, fp64 sums


@pytest.mark.parametrize("tpq", ["1", "2", "3", "4"])
@pytest.mark.parametrize("nq,n,bits,nlab,k,stash", [(37, 500, 64, 24, 50, 1), (6, 333, 254, 5, None, 0), (20, 3000, 128, 80, 700, 1),
                                                    (3, 777, 17, 3, 10, 1)])
def test_stage_a_threads_per_query(sim, monkeypatch, tpq, nq, n, bits, nlab, k, stash):
    """Stage A with 1-4 threads per query (B200_MAP_TPQ): the threads of a query take alternate 32-row groups of every
    tile — including the partial last group, which exactly one of them owns — and share one counter column."""
    monkeypatch.setenv("B200_MAP_TPQ", tpq)
    monkeypatch.setenv("B200_MAP_STASH", str(stash))
    rng = np.random.default_rng(nq + n + bits)
    q, r = pm1(rng, nq, bits), pm1(rng, n, bits)
    ql, rl = multi_hot(rng, nq, nlab, 0.1), multi_hot(rng, n, nlab, 0.1)
    m0, ap0, ts0, rank0, dist0 = eval_ref.maphashing_exact(q, ql, r, rl, k, return_details=True)
    m, ap, ts, ri, rd, plan = sim_map(sim, q, ql, r, rl, k, 1, 0, 148, want_rank=True)
    assert np.array_equal(ts.astype(np.int64), ts0) and np.array_equal(ri.astype(np.int64), rank0), plan
    assert np.array_equal(rd.astype(np.int64), dist0) and np.abs(ap - ap0).max() <= 1e-6 and abs(m - m0) <= 1e-6


def test_hamming_map_on_reference_goldens(sim, golden):
    """The stage programs against outputs of the real reference code (stable tie order)."""
    for name in golden["cases"]:
        q, r, ql, rl = (golden[f"{name}/{k}"] for k in ("q", "r", "ql", "rl"))
        tk, inc = (int(v) for v in golden[f"{name}/topk"])
        topk = None if tk == -1 else ("max_bin_count" if tk == -2 else tk)
        topk = eval_ref.resolve_topk_ref(topk, rl, bool(inc))
        m, *_ = sim_map(sim, q.astype(np.float32), ql, r.astype(np.float32), rl, topk, 2, 0, 8)
        assert abs(m - float(golden[f"{name}/map_reference_stable"])) <= 1e-6, name


def test_hamming_map_constant_codes_closed_form(sim):
    """MAP-1: one code for everybody => ranking is the index order; AP follows from the labels alone."""
    rng = np.random.default_rng(11)
    ql, rl = multi_hot(rng, 12, 20, 0.15), multi_hot(rng, 150, 20, 0.15)
    q, r = np.ones((12, 64), np.float32), np.ones((150, 64), np.float32)
    m, ap, ts, ri, rd, _ = sim_map(sim, q, ql, r, rl, 40, want_rank=True)
    assert np.array_equal(ri, np.tile(np.arange(40, dtype=np.uint32), (12, 1))) and (rd == 0).all()
    for i in range(12):
        rel = (ql[i] @ rl[:40].T) > 0
        assert abs(ap[i] - eval_ref.ap_from_ranked_relevance(rel)[0]) <= 1e-6


def test_hamming_map_correlated_codes_are_informative(sim):
    rng = np.random.default_rng(12)
    rl = multi_hot(rng, 2000, 24, 0.1)
    ql = multi_hot(rng, 30, 24, 0.1)
    q, r = correlated_codes(rng, ql, rl, 64)
    m, ap, ts, *_ = sim_map(sim, q, ql, r, rl, 500)
    density = ((ql @ rl.T) > 0).mean()
    assert m > density + 0.1
    assert abs(m - eval_ref.maphashing_exact(q, ql, r, rl, 500)) <= 1e-6


def test_wide_plan_keeps_segments_within_16_bit_counters(sim):
    """k > 65534 (wide plan) on a database whose rows all share one code: every row of a segment lands in ONE
    (segment, distance) bucket, so a segment longer than 65534 rows would wrap the 16|16-bit shared counters of stage A
    and of the all-rows stage B.  The planner must cap the segment length in wide mode too."""
    rng = np.random.default_rng(5)
    nq, n, bits = 32, 150000, 64
    ql, rl = multi_hot(rng, nq, 10, 0.2), multi_hot(rng, n, 10, 0.2)
    q, r = np.ones((nq, bits), np.float32), np.ones((n, bits), np.float32)
    m, ap, ts, _, _, plan = sim_map(sim, q, ql, r, rl, None, 1, 0, 1)          # one SM: the planner wants ONE long segment
    assert plan[3] % 2 == 1 and plan[2] <= 65534, plan                         # wide, yet capped
    for i in (0, 7, 31):
        rel = (ql[i] @ rl.T) > 0
        want, hits = eval_ref.ap_from_ranked_relevance(rel)
        assert int(ts[i]) == hits and abs(ap[i] - want) <= 1e-6, (i, plan)
