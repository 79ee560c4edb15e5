"""The HP-SWT oracle against the PyWavelets documentation examples and analytic known answers.

PyWavelets itself is absent, so these — not a run of pywt — are what pins ``oracle.swt_ref`` ("parity unpinned" against
the library proper; see oracle/__init__.py)."""
import numpy as np
import pytest

from oracle import c_oracle, filters, swt_ref

ORTHO = ["haar", "db2", "db3", "db4", "db5", "sym4", "coif1"]


def test_pywt_doc_example_swt_db1_level2():
    """pywt.swt([1..8], 'db1', level=2) from the PyWavelets documentation (SWT-1, SWT-2)."""
    x = np.arange(1, 9, dtype=np.float64)
    lo, hi = filters.filter_bank("db1")
    ca1 = swt_ref.swt_step_1d(x, lo, 1, 0, np.float64)
    cd1 = swt_ref.swt_step_1d(x, hi, 1, 0, np.float64)
    assert np.allclose(ca1, [2.12132034, 3.53553391, 4.94974747, 6.36396103, 7.77817459, 9.19238816, 10.60660172, 6.36396103])
    assert np.allclose(cd1, [-0.70710678] * 7 + [4.94974747])
    assert np.allclose(swt_ref.swt_step_1d(ca1, lo, 2, 0, np.float64), [5, 7, 9, 11, 13, 11, 9, 7])
    assert np.allclose(swt_ref.swt_step_1d(ca1, hi, 2, 0, np.float64), [-2, -2, -2, -2, -2, 2, 6, 2])


@pytest.mark.parametrize("name", ORTHO + ["bior4.4", "bior2.2", "db7"])
@pytest.mark.parametrize("level", [1, 2, 3])
def test_vectorised_step_equals_literal_loop(name, level):
    rng = np.random.default_rng(level)
    lo, hi = filters.filter_bank(name)
    for n in (8, 24):                      # 8 < dilated filter length for long filters: the wrap-around branch
        x = rng.random(n)
        for h in (lo, hi):
            assert np.allclose(swt_ref.swt_step_1d(x, h, level, 0, np.float64), swt_ref.swt_step_1d_loop(x, h, level), atol=1e-12)


@pytest.mark.parametrize("name", ORTHO)
def test_filter_banks_are_orthonormal(name):
    lo, hi = filters.filter_bank(name)
    f = len(lo)
    assert abs(lo.sum() - np.sqrt(2)) < 1e-10 and abs((lo * lo).sum() - 1) < 1e-10 and abs(hi.sum()) < 1e-10
    for m in range(1, f // 2):
        assert abs((lo[2 * m:] * lo[:f - 2 * m]).sum()) < 1e-10
        assert abs((lo[2 * m:] * hi[:f - 2 * m]).sum()) < 1e-10


@pytest.mark.parametrize("name", ["bior1.3", "bior2.2", "bior4.4"])
def test_biorthogonal_banks_are_perfect_reconstruction(name):
    lo, hi = filters.filter_bank(name)
    p = np.convolve(lo, hi * (-1.0) ** np.arange(len(hi)))        # H0(z) H1(-z) must be half-band
    odd, even = np.abs(p[1::2]), np.abs(p[0::2])
    one = odd if odd.max() > even.max() and np.sort(odd)[-2] < 1e-10 else even
    assert abs(np.sort(one)[-1] - 1.0) < 1e-10 and np.sort(one)[-2] < 1e-10


@pytest.mark.parametrize("name", ORTHO)
@pytest.mark.parametrize("level", [1, 2, 3])
def test_constant_image(name, level):
    """SWT-3: orthogonal wavelet, constant c: LL = 2**level * c, details = 0."""
    out = swt_ref.swt2_ref(np.full((16, 24), 0.37), name, level, np.float64)
    assert np.allclose(out[0], (2 ** level) * 0.37, atol=1e-9) and np.allclose(out[1:], 0, atol=1e-9)


def test_haar_ranges_and_zero_mean_details():
    """SWT-4 / studies/results/swt_transform_check_2026-08-12.txt: x in [0,1] => LL in [0,2], details in [-1,1], each
    detail band sums to zero under periodisation."""
    rng = np.random.default_rng(0)
    x = rng.random((32, 48))
    out = swt_ref.swt2_ref(x, "haar", 1, np.float64)
    assert out[0].min() >= 0 and out[0].max() <= 2 and np.abs(out[1:]).max() <= 1
    assert np.abs(out[1:].sum(axis=(1, 2))).max() < 1e-9


@pytest.mark.parametrize("name", ORTHO)
def test_energy_is_quadrupled_at_level_1(name):
    """SWT-5: every undecimated orthogonal 1-D step doubles the energy."""
    x = np.random.default_rng(1).random((16, 32))
    out = swt_ref.swt2_ref(x, name, 1, np.float64)
    assert abs((out ** 2).sum() / (x ** 2).sum() - 4.0) < 1e-9


@pytest.mark.parametrize("name,level", [("haar", 1), ("db2", 2), ("bior4.4", 3)])
def test_circular_shift_equivariance(name, level):
    """SWT-6."""
    x = np.random.default_rng(2).random((32, 40))
    a = swt_ref.swt2_ref(np.roll(x, (5, -3), axis=(0, 1)), name, level, np.float64)
    b = np.roll(swt_ref.swt2_ref(x, name, level, np.float64), (5, -3), axis=(1, 2))
    assert np.allclose(a, b, atol=1e-10)


def test_band_order_is_cA_cH_cV_cD():
    """cH = 'da': detail along axis -2 (rows).  An image that only varies along H has energy in LL and LH only."""
    x = np.tile(np.random.default_rng(3).random((16, 1)), (1, 24))
    out = swt_ref.swt2_ref(x, "haar", 1, np.float64)
    assert np.abs(out[1]).max() > 1e-3 and np.abs(out[2]).max() < 1e-12 and np.abs(out[3]).max() < 1e-12


def test_coarsest_level_only_and_levels_cascade():
    x = np.random.default_rng(4).random((16, 16))
    all_levels = swt_ref.swt2_ref(x, "db2", 3, np.float64, all_levels=True)
    assert len(all_levels) == 3
    assert np.array_equal(all_levels[0], swt_ref.swt2_ref(x, "db2", 3, np.float64))
    assert np.allclose(all_levels[-1], swt_ref.swt2_ref(x, "db2", 1, np.float64))


def test_size_must_divide():
    with pytest.raises(ValueError):
        swt_ref.swt2_ref(np.zeros((6, 8)), "haar", 2)


def test_transform_contract_matches_reference_check_script():
    """studies/verify_swt_transform.py:70-124: [C=3, S=4, H, W], float32, finite, bands distinct."""
    img = np.random.default_rng(5).integers(0, 256, (32, 40, 3), dtype=np.uint8)
    out = swt_ref.swt_transform_ref(img, "haar", 1)
    assert out.shape == (3, 4, 32, 40) and out.dtype == np.float32 and np.isfinite(out).all()
    flat = out.transpose(1, 0, 2, 3).reshape(4, -1)
    norm = flat / np.linalg.norm(flat, axis=1, keepdims=True)
    cos = norm @ norm.T
    assert np.abs(cos[~np.eye(4, dtype=bool)]).mean() < 0.95
    assert swt_ref.fixed_size(518, 518, 3) == (520, 520) and swt_ref.fixed_size(224, 224, 3) == (224, 224)
    raw = swt_ref.raw_stack_ref(img)
    assert raw.shape == (3, 4, 32, 40) and np.array_equal(raw[:, 0], raw[:, 3])


@pytest.mark.parametrize("name,level", [("haar", 1), ("db2", 2), ("sym4", 3), ("bior4.4", 2), ("db7", 1)])
def test_c_oracle_equals_numpy_oracle(name, level):
    x = np.random.default_rng(6).integers(0, 256, (2, 3, 32, 40), dtype=np.uint8)
    lo, hi = filters.filter_bank(name)
    c = c_oracle.swt2(x, lo, hi, level)
    p = swt_ref.swt2_ref(x.astype(np.float32) / np.float32(255), name, level)
    assert np.abs(c - p).max() <= 1e-6
