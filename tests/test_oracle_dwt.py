"""The decimated-DWT oracle (oracle/swt_ref.py: dwt_step_1d, dwt2_ref — a restatement of pywt.dwt / wavedec2 in mode
'symmetric') against PyWavelets' documented examples, an independent convolve-and-downsample formulation and closed forms.
PyWavelets is absent: parity with the package itself is unpinned, as for the SWT."""
import numpy as np

from oracle import filters, swt_ref


def test_pywavelets_documentation_examples():
    x = np.arange(1, 9, dtype=np.float64)
    lo, hi = filters.filter_bank("db1")
    # pywt.dwt([1..8], 'db1') and pywt.wavedec([1..8], 'db1', level=2) (PyWavelets docs)
    assert np.allclose(swt_ref.dwt_step_1d(x, lo, -1, np.float64), [2.12132034, 4.94974747, 7.77817459, 10.60660172])
    assert np.allclose(swt_ref.dwt_step_1d(x, hi, -1, np.float64), [-0.70710678] * 4)
    ca1 = swt_ref.dwt_step_1d(x, lo, -1, np.float64)
    assert np.allclose(swt_ref.dwt_step_1d(ca1, lo, -1, np.float64), [5.0, 13.0])
    assert np.allclose(swt_ref.dwt_step_1d(ca1, hi, -1, np.float64), [-2.0, -2.0])
    # pywt.dwt([1, 2, 3, 4, 5, 6], 'db2', mode='symmetric'): output length (6 + 4 - 1) // 2 = 4
    lo2, hi2 = filters.filter_bank("db2")
    assert swt_ref.dwt_step_1d(np.arange(1.0, 7.0), lo2, -1, np.float64).shape == (4,)


def test_matches_convolve_and_downsample():
    """Independent formulation: pad F - 1 samples by half-sample symmetry, full-overlap convolution, keep odd positions."""
    rng = np.random.default_rng(0)
    for name in ("haar", "db2", "db4", "sym4", "bior4.4", "coif1"):
        lo, hi = filters.filter_bank(name)
        f = len(lo)
        for n in (1, 2, 5, 8, 13, 32, 33):
            x = rng.standard_normal(n)
            pad = x
            while pad.shape[0] < n + 2 * (f - 1):                 # np.pad cannot extend by more than the length at once
                k = min((n + 2 * (f - 1) - pad.shape[0] + 1) // 2, pad.shape[0])
                pad = np.pad(pad, k, mode="symmetric")
            off = (pad.shape[0] - n) // 2 - (f - 1)
            pad = pad[off:off + n + 2 * (f - 1)]
            for h in (lo, hi):
                want = np.convolve(pad, np.asarray(h, np.float64), mode="valid")[1::2][:(n + f - 1) // 2]
                got = swt_ref.dwt_step_1d(x, h, -1, np.float64)
                assert got.shape == want.shape and np.allclose(got, want, atol=1e-12), (name, n)


def test_dwt2_closed_forms():
    rng = np.random.default_rng(1)
    x = rng.random((2, 3, 16, 12)).astype(np.float32)
    out = swt_ref.dwt2_ref(x, "haar", 1)
    a, b, c, d = x[..., 0::2, 0::2], x[..., 0::2, 1::2], x[..., 1::2, 0::2], x[..., 1::2, 1::2]
    assert out.shape == (2, 3, 4, 8, 6)
    assert np.allclose(out[..., 0, :, :], (a + b + c + d) / 2, atol=1e-6)                # cA
    assert np.allclose(np.abs(out[..., 1, :, :]), np.abs(a + b - c - d) / 2, atol=1e-6)  # cH: detail along H
    assert np.allclose(np.abs(out[..., 2, :, :]), np.abs(a - b + c - d) / 2, atol=1e-6)  # cV: detail along W
    assert np.allclose(np.abs(out[..., 3, :, :]), np.abs(a - b - c + d) / 2, atol=1e-6)
    # a constant image: symmetric extension keeps it constant, so cA = 2^L * c and the details vanish
    for name, lv in (("db2", 1), ("sym4", 2), ("db4", 3)):
        out = swt_ref.dwt2_ref(np.full((40, 36), 0.25, np.float32), name, lv, np.float64)
        assert np.allclose(out[0], 0.25 * 2 ** lv, atol=1e-9) and np.abs(out[1:]).max() < 1e-9
    # sizes: H_l = (H_{l-1} + F - 1) // 2
    assert swt_ref.dwt2_ref(np.zeros((32, 32), np.float32), "db4", 2).shape == (4, 13, 13)
    assert swt_ref.dwt2_ref(np.zeros((33, 7), np.float32), "db2", 1).shape == (4, 18, 5)


def test_swt_and_dwt_restatements_share_one_phase_convention():
    """PyWavelets: swt output n is the (periodised) convolution at index n + F/2, dwt output o the (symmetric) one at
    2o + 1 — so away from the borders dwt[o] == swt[2o + 1 - F/2].  The two oracles were restated independently from
    those two formulas; this ties their origins together."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal(64)
    for name in ("haar", "db2", "db4", "sym4", "bior4.4"):
        lo, hi = filters.filter_bank(name)
        f = len(lo)
        for h in (lo, hi):
            s = swt_ref.swt_step_1d(x, h, 1, -1, np.float64)
            d = swt_ref.dwt_step_1d(x, h, -1, np.float64)
            o = np.arange(f, (64 - f) // 2)                      # outputs whose window touches neither border
            assert np.allclose(d[o], s[2 * o + 1 - f // 2], atol=1e-12), name
